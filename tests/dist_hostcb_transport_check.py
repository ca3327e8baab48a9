"""Run under torch.distributed.run (gloo, CPU only): the four host collectives of parca_agent_b200.host_transport — the
transport a merge group (pa_merge_create_host) calls back into — checked against their definitions in include/parcaagg.h."""
import ctypes as C
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parca_agent_b200.host_transport import GlooTransport  # noqa: E402


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def u64(v):
    return (C.c_uint64 * len(v))(*v)


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    t = GlooTransport()
    s = t.struct
    # allgather
    send = np.full(7, rank + 1, dtype=np.uint8)
    recv = np.zeros(7 * world, dtype=np.uint8)
    assert s.allgather(None, ptr(send), ptr(recv), 7) == 0
    assert recv.tolist() == sum(([r + 1] * 7 for r in range(world)), [])
    # allgatherv (rank r contributes r*3 bytes; rank 0 contributes nothing)
    counts = [3 * r for r in range(world)]
    displs = [sum(counts[:r]) + 2 * r for r in range(world)]  # gaps between blocks must stay untouched
    send = np.full(max(counts[rank], 1), 10 + rank, dtype=np.uint8)
    recv = np.full(displs[-1] + counts[-1] + 1, 0xEE, dtype=np.uint8)
    assert s.allgatherv(None, ptr(send), ptr(recv), u64(counts), u64(displs)) == 0
    for r in range(world):
        assert recv[displs[r]:displs[r] + counts[r]].tolist() == [10 + r] * counts[r]
    assert recv[-1] == 0xEE
    # alltoallv: rank i sends (i + j + 1) bytes of value 16*i + j to rank j
    sc = [rank + j + 1 for j in range(world)]
    sd = [sum(sc[:j]) for j in range(world)]
    rc = [j + rank + 1 for j in range(world)]
    rd = [sum(rc[:j]) for j in range(world)]
    send = np.concatenate([np.full(sc[j], 16 * rank + j, dtype=np.uint8) for j in range(world)])
    recv = np.zeros(sum(rc), dtype=np.uint8)
    assert s.alltoallv(None, ptr(send), u64(sc), u64(sd), ptr(recv), u64(rc), u64(rd)) == 0
    for j in range(world):
        assert recv[rd[j]:rd[j] + rc[j]].tolist() == [16 * j + rank] * rc[j]
    # allreduce(min) on uint32, including values >= 2^31 and the "unset" marker
    buf = np.array([0xFFFFFFFF, 5 + rank, 0x90000000 + (world - rank), 0xFFFFFFFF if rank else 7], dtype=np.uint32)
    assert s.allreduce_min_u32(None, ptr(buf), len(buf)) == 0
    assert buf.tolist() == [0xFFFFFFFF, 5, 0x90000001, 7]
    assert not t.errors
    dist.barrier()
    if rank == 0:
        print("hostcb-transport ok world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
