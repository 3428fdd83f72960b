import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from graphix_b200 import mbqc, B200Statevector
from graphix_b200.sharded import GpuShardEngine, ShardedStatevector
from oracle.oracle import OracleStatevector, outcome_to_operator_matrix
rng = np.random.default_rng(0)
st = mbqc.PlanarState(mbqc.Plane.YZ, 0.37)
for trial in range(12):
    vec = rng.normal(size=3); vec /= np.linalg.norm(vec)
    pm = outcome_to_operator_matrix(vec, int(rng.integers(2)))
    states = [mbqc.BasicStates.PLUS, st, mbqc.BasicStates.MINUS][trial % 3]
    pre = trial % 4
    res = []
    for kind in ("orc", "lib2", "shard"):
        if kind == "orc":
            s = OracleStatevector(nqubit=0, max_qubits=4)
        elif kind == "lib2":
            s = B200Statevector(nqubit=0, max_qubits=4, fusion=2)
        else:
            s = ShardedStatevector(GpuShardEngine(4, device=0))
        s.add_nodes(1, states)
        if pre & 1:
            s.evolve_single(mbqc.Ops.Y, 0)
        if pre & 2:
            s.add_nodes(1, st)
            s.project_qubit(outcome_to_operator_matrix((0.6, 0, 0.8), 0), 1)
        s.project_qubit(pm, 0)
        res.append(complex(s.flatten()[0]))
    print(trial, pre, [np.round(r, 6) for r in res], 'OK' if abs(res[0]-res[1])<1e-12 and abs(res[0]-res[2])<1e-12 else 'MISMATCH')
