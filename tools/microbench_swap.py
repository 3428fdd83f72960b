"""Peer-memory qubit swap rate between two (or more) ranks:

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/microbench_swap.py [local qubits] [reps]

Times ShardedStatevector._swap_in on a |+..+> register (CUDA events around barrier + k_swap_peer + barrier, max over
ranks) for victims on low / middle / top physical bits.  GXSV_SWAP_MODE / GXSV_SWAP_U / GXSV_SWAP_CTAS select the
diagnostic variants of gxsv_swap_peer, GXSV_EXCHANGE=nccl the pack -> ncclSend/Recv -> unpack path."""
from __future__ import annotations

import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main() -> None:
    nloc = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    rank = int(os.environ["RANK"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    from graphix_b200 import mbqc
    from graphix_b200.sharded import GpuShardEngine, ShardedStatevector

    world = dist.get_world_size()
    g = world.bit_length() - 1
    sh = ShardedStatevector(GpuShardEngine(nloc, strict=True), strict=True)
    sh.add_nodes(nloc + g, mbqc.BasicStates.PLUS)
    for i in range(nloc):
        sh.engine.settle(i)
    rows = []
    for label, pick in (("top bit", nloc - 1), ("middle bit", nloc // 2), ("bit 3", 3)):
        sh.victim_priority = lambda pos, pick=pick: 1.0 if sh.q[pos].li == pick else 0.0
        times = []
        for it in range(reps + 1):
            glo = next(o for o in sh.q if o.g is not None)
            # keep evicting the qubit that sits on engine slot `pick`
            sh.victim_priority = lambda pos, pick=pick: 1.0 if sh.q[pos].li == pick else 0.0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            torch.cuda.synchronize()
            e0.record()
            sh._swap_in_impl(glo)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if it:
                times.append(float(t.item()))
        half_bytes = 16 * (1 << (nloc - 1))
        ms = sum(times) / len(times)
        rows.append({"victim": label, "ms": ms, "GBps_per_direction": half_bytes / 1e9 / (ms * 1e-3)})
        if rank == 0:
            print(f"local {nloc} qubits ({16 * (1 << nloc) / 2**30:.0f} GiB), victim on {label:10s}: {ms:8.2f} ms per swap, "
                  f"{rows[-1]['GBps_per_direction']:7.1f} GB/s per direction per GPU  [{sh.transport}]", flush=True)
    n2 = sh.norm2()
    if rank == 0:
        print("norm2", n2, {k: os.environ.get(k) for k in ("GXSV_SWAP_MODE", "GXSV_SWAP_U", "GXSV_SWAP_CTAS", "GXSV_EXCHANGE")})
        os.makedirs("gpurun_out", exist_ok=True)
        tag = "_".join(f"{k[5:].lower()}{v}" for k, v in os.environ.items() if k.startswith("GXSV_SWAP") or k == "GXSV_EXCHANGE")
        with open(f"gpurun_out/microbench_swap_{nloc}_{tag or 'default'}.json", "w") as f:
            json.dump(rows, f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
