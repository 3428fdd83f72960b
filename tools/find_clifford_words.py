"""One-off helper (container only): express each reference Clifford (graphix/_db.py:15-66)
as a word in {h, s} and a phase exp(i*pi*p/4); the output is pasted into graphix_b200/mbqc.py."""
import cmath, itertools, math, sys
import numpy as np
sys.path.insert(0, "/root/repo/tools")
from ref_shim import import_reference
import_reference()
from graphix._db import CLIFFORD

H = np.array([[1, 1], [1, -1]], dtype=complex) / math.sqrt(2)
S = np.array([[1, 0], [0, 1j]], dtype=complex)
out = []
for k, target in enumerate(CLIFFORD):
    found = None
    for L in range(0, 9):
        for word in itertools.product("hs", repeat=L):
            m = np.eye(2, dtype=complex)
            for ch in word:
                m = m @ (H if ch == "h" else S)
            for p in range(8):
                if np.allclose(cmath.exp(1j * math.pi * p / 4) * m, target, atol=1e-12):
                    found = ("".join(word), p); break
            if found: break
        if found: break
    out.append(found)
    print(k, found)
print(tuple(out))
