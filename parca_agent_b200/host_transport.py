"""Host-side collectives for pa_merge_create_host (include/parcaagg.h): the four callbacks a merge group needs, implemented
with torch.distributed on CPU tensors (gloo). Plumbing only — it lets several processes build one merged record without NCCL
(and lets two processes that share ONE GPU exercise the whole multi-process path); the aggregation stays in the library."""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

_u64p = C.POINTER(C.c_uint64)
ALLGATHER = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64)
ALLGATHERV = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, _u64p, _u64p)
ALLTOALLV = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, _u64p, _u64p, C.c_void_p, _u64p, _u64p)
ALLREDUCE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64)


class PaMergeHostTransport(C.Structure):
    _fields_ = [("user", C.c_void_p), ("allgather", ALLGATHER), ("allgatherv", ALLGATHERV), ("alltoallv", ALLTOALLV), ("allreduce_min_u32", ALLREDUCE)]


def _view(ptr, n):
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(max(int(n), 1),))[:int(n)]


class GlooTransport:
    """Callbacks over the default (or a given) torch.distributed group whose backend handles CPU tensors."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.errors = []
        self.struct = PaMergeHostTransport(None, ALLGATHER(self._allgather), ALLGATHERV(self._allgatherv), ALLTOALLV(self._alltoallv), ALLREDUCE(self._allreduce))

    def _guard(self, fn, *a):
        try:
            fn(*a)
            return 0
        except Exception as e:  # noqa: BLE001 — an exception must not unwind through the C caller
            self.errors.append(repr(e))
            return -1

    def _allgather(self, user, send, recv, nbytes):
        def go():
            t = torch.from_numpy(_view(send, nbytes).copy())
            outs = [torch.empty(int(nbytes), dtype=torch.uint8) for _ in range(self.world)]
            dist.all_gather(outs, t, group=self.group)
            dst = _view(recv, nbytes * self.world)
            for r, o in enumerate(outs):
                dst[r * nbytes:(r + 1) * nbytes] = o.numpy()
        return self._guard(go)

    def _allgatherv(self, user, send, recv, counts, displs):
        def go():
            cnt = [int(counts[r]) for r in range(self.world)]
            dsp = [int(displs[r]) for r in range(self.world)]
            m = max(cnt + [1])
            pad = torch.zeros(m, dtype=torch.uint8)
            if cnt[self.rank]:
                pad[:cnt[self.rank]] = torch.from_numpy(_view(send, cnt[self.rank]).copy())
            outs = [torch.empty(m, dtype=torch.uint8) for _ in range(self.world)]
            dist.all_gather(outs, pad, group=self.group)
            total = max(d + c for d, c in zip(dsp, cnt))
            dst = _view(recv, total)
            for r in range(self.world):
                if cnt[r]:
                    dst[dsp[r]:dsp[r] + cnt[r]] = outs[r][:cnt[r]].numpy()
        return self._guard(go)

    def _alltoallv(self, user, send, scounts, sdispls, recv, rcounts, rdispls):
        def go():
            sc = [int(scounts[r]) for r in range(self.world)]
            sd = [int(sdispls[r]) for r in range(self.world)]
            rc = [int(rcounts[r]) for r in range(self.world)]
            rd = [int(rdispls[r]) for r in range(self.world)]
            src = _view(send, max(d + c for d, c in zip(sd, sc)))
            dst = _view(recv, max(d + c for d, c in zip(rd, rc)))
            reqs, bufs = [], {}
            for r in range(self.world):
                if r == self.rank:
                    if sc[r]:
                        dst[rd[r]:rd[r] + rc[r]] = src[sd[r]:sd[r] + sc[r]]
                    continue
                if rc[r]:
                    bufs[r] = torch.empty(rc[r], dtype=torch.uint8)
                    reqs.append(dist.irecv(bufs[r], src=r, group=self.group))
                if sc[r]:
                    reqs.append(dist.isend(torch.from_numpy(src[sd[r]:sd[r] + sc[r]].copy()), dst=r, group=self.group))
            for q in reqs:
                q.wait()
            for r, b in bufs.items():
                dst[rd[r]:rd[r] + rc[r]] = b.numpy()
        return self._guard(go)

    def _allreduce(self, user, buf, count):
        def go():
            v = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_uint32)), shape=(int(count),))
            t = torch.from_numpy(v.astype(np.int64))  # uint32 values as non-negative int64: MIN is the same
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
            v[:] = t.numpy().astype(np.uint32)
        return self._guard(go)
