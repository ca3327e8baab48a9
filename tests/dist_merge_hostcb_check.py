"""Run under torch.distributed.run with the gloo backend: mode B through pa_merge_create_host (collectives on host buffers
over gloo) or, with PA_MERGE_TRANSPORT=shm, pa_merge_create_shm (mailboxes in one page-locked shared-memory segment;
PA_SHM_MAILBOX bytes per rank, small values force multi-round collectives) — one shard aggregator per
PROCESS, the merged stream in POSIX shared memory that every rank maps. All ranks may
share ONE GPU (PA_ONE_GPU=1, the driver's 1-GPU test tier): everything but the NCCL calls themselves is the code the NCCL
group runs (per-process size exchanges, per-process parts of the sliced buffers, validity words completed across processes,
rank 0 adding dictionaries and metadata). Rank 0 compares the bytes with the CPU oracle on [shard 0 rows, shard 1 rows, ...]."""
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
from parca_agent_b200.host_transport import GlooTransport  # noqa: E402


def main():
    local = 0 if os.environ.get("PA_ONE_GPU") == "1" else int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    kind = os.environ.get("PA_MERGE_TRANSPORT", "gloo")
    transport = GlooTransport()
    seg = ["/pa_merge_test_%d" % os.getpid()]
    dist.broadcast_object_list(seg, src=0)
    cases = [synth.edge_workload(seed=33, n=5000, hash_mode=abi.PA_HASH_PROVIDED, external=False),
             synth.edge_workload(seed=34, n=4000, hash_mode=abi.PA_HASH_XXH64X2, external=False),
             synth.config3(n=120_000, u=9_000, p=4_096, npids=96, lsets=6),
             synth.config2(n=300_000, u=20_000, p=32_768)]
    for ci, w in enumerate(cases):
        idx = sharded.shard_rows(w, world)
        if ci == 1:  # an empty shard in the middle of the group
            idx = [idx[0]] + [np.zeros(0, np.int64)] * (world - 2) + [np.concatenate(idx[1:])] if world > 2 else idx
        part = w.rows(idx[rank])
        a = lib.from_workload(part, device=local, frame_id_bytes=4 if ci == 2 else 8)
        if kind == "shm":
            group = lib.MergeGroup.shm(a, "%s_%d" % (seg[0], ci), rank, world, int(os.environ.get("PA_SHM_MAILBOX", "0")))
        else:
            group = lib.MergeGroup.host(a, transport, rank, world)
        want, st = oracle_py.run(w.rows(np.concatenate(idx))) if rank == 0 else (None, None)
        shm = None
        for rep in range(2):
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
                print("case %s rep %d ok: %d rows, %d stacks, %d bytes" % (w.name, rep, res.n_rows, res.n_unique_stacks, n), flush=True)
            dist.barrier()
        group.close()
        a.close()
        shm.close()
        if rank == 0:
            try:
                shm.unlink()
            except FileNotFoundError:
                pass
    assert not transport.errors, transport.errors
    if rank == 0:
        print("merge-hostcb ok world=%d transport=%s" % (world, kind))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
