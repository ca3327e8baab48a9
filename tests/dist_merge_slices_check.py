"""Run under torch.distributed.run with one GPU per rank: mode B through pa_merge_* over NCCL.

Every rank builds the same workload, keeps its pid-hash shard, and joins the library-level NCCL group (the 128-byte NCCL id
travels through torch.distributed, the job a host agent would do). The merged stream lands in POSIX shared memory that every
rank maps: each rank copies its rows / runs / stream range over its own PCIe link, rank 0 adds dictionaries and metadata and
compares the bytes with the CPU oracle on the stream [shard 0 rows, shard 1 rows, ...].
"""
import ctypes
import os
import sys
from multiprocessing import resource_tracker, shared_memory

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_py  # noqa: E402
from parca_agent_b200 import abi, lib, sharded, synth  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    cases = [synth.edge_workload(seed=33, n=5000, hash_mode=abi.PA_HASH_PROVIDED, external=False),
             synth.edge_workload(seed=34, n=4000, hash_mode=abi.PA_HASH_XXH64X2, external=False),
             synth.config3(n=120_000, u=9_000, p=4_096, npids=96, lsets=6),
             synth.config2(n=400_000, u=20_000, p=32_768)]
    for w in cases:
        idx = sharded.shard_rows(w, world)
        part = w.rows(idx[rank])
        a = lib.from_workload(part, device=local)
        ids = [lib.MergeGroup.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        group = lib.MergeGroup.nccl(a, ids[0], rank, world)
        want, st = oracle_py.run(w.rows(np.concatenate(idx))) if rank == 0 else (None, None)
        shm = None
        for rep in range(2):  # two intervals: adaptive table sizes carry over, the output buffer is reused
            lib.load(a, part)
            a.stage()
            group.process()
            n = group.plan()
            if shm is None:
                names = [None]
                if rank == 0:
                    shm = shared_memory.SharedMemory(create=True, size=max(n, 1))
                    names[0] = shm.name
                dist.broadcast_object_list(names, src=0)
                if rank != 0:
                    shm = shared_memory.SharedMemory(name=names[0])
                    resource_tracker.unregister(shm._name, "shared_memory")  # attached, not owned: Python < 3.13 would unlink it when this process exits
            view = ctypes.c_char.from_buffer(shm.buf)
            res = group.collect(ctypes.addressof(view), n)
            del view
            dist.barrier()
            if rank == 0:
                got = bytes(shm.buf[:n])
                assert got == want, "merged stream differs from the oracle (%s, rep %d): %d vs %d bytes" % (w.name, rep, len(got), len(want))
                assert res.n_rows == w.n and res.n_unique_stacks == st["unique_stacks"] and res.n_locations == st["locations"]
                print("case %s rep %d ok: %d rows, %d stacks, %d bytes, stats %s" % (w.name, rep, res.n_rows, res.n_unique_stacks, n, group.stats()), flush=True)
            dist.barrier()
        group.close()  # unregisters the shared buffer before it is unmapped
        a.close()
        shm.close()
        if rank == 0:
            try:
                shm.unlink()
            except FileNotFoundError:
                pass
    if rank == 0:
        print("merge-slices ok world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
