"""Run under torch.distributed.run with the gloo backend (CPU): the N>1 plumbing of bench.py.

Each rank builds its pid-hash shard (mode A: one self-contained batch per rank, no data-path
collective), checks ownership with the shard function, aggregates its shard with the CPU oracle and
the ranks then agree on global facts through the same collectives bench.py uses.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_py  # noqa: E402
from parca_agent_b200 import synth  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    w = synth._pid_shard("t", 0x5EED0002, rank, world, 3000, 200, 512, 32, synth.abi.PA_HASH_XXH64X2)
    pids = np.unique(w.hdrs["pid"])
    assert all(synth.xxh64_u32(int(p)) % world == rank for p in pids), "rank owns a foreign pid"
    data, st = oracle_py.run(w)
    assert st["rows"] == w.n and len(data) > 0
    gathered = [None] * world
    dist.all_gather_object(gathered, set(int(p) for p in pids))
    for i in range(world):
        for j in range(i + 1, world):
            assert not (gathered[i] & gathered[j]), "pid shards overlap"
    rows = torch.tensor([st["rows"]], dtype=torch.int64)
    dist.all_reduce(rows, op=dist.ReduceOp.SUM)
    assert int(rows) == 3000 * world
    t = torch.tensor([0.001 * (rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # bench.py: time = max over ranks
    assert abs(float(t) - 0.001 * world) < 1e-12
    dist.barrier()
    if rank == 0:
        print("shard-check ok world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
