"""Multi-GPU "mode B": one merged record batch, bit-identical to what the unsharded stream would give.

Shards are cut by pid hash (`synth.xxh64_u32(pid) % G`, the north-star rule); every shard aggregator hashes and
deduplicates its own rows on its own GPU. What crosses NVLink afterwards is small: 64 bytes per row (the header with the
stack id filled in) plus the frames of each shard's UNIQUE stacks. The merging GPU scatters the rows to their positions in
the global order and runs the provided-id pipeline once, so every dictionary index is assigned in global first-occurrence
order — exactly the reference's rule — without the 8·F bytes per row of frame data ever leaving the shard that hashed it.

torch is used for device buffers, the row scatter and (in `merge_distributed`) the NCCL send/recv: plumbing only; the
aggregation kernels are the library's. No collective is issued for per-shard batches (mode A, the default).
"""
import numpy as np
import torch


def _export(a, frame_base, device):
    """Processed shard aggregator -> (rows [n,64] uint8, frames [nf] int64) on `device`."""
    n, nf = a.shard_sizes()
    hdr = torch.empty((max(n, 1), 64), dtype=torch.uint8, device=device)
    frames = torch.empty(max(nf, 1), dtype=torch.int64, device=device)
    torch.cuda.synchronize(device)
    a.shard_export(frame_base, hdr.data_ptr(), frames.data_ptr())
    return hdr[:n], frames[:nf]


def _stage_merged(merged, hdr_all, frames_all, device, collect=True):
    torch.cuda.synchronize(device)
    merged.stage_device(hdr_all.data_ptr(), hdr_all.shape[0], frames_all.data_ptr(), frames_all.shape[0])
    merged.process()
    return merged.collect() if collect else None


def _stage_parts(merged, parts, n_total, device, collect=True):
    """rows -> global order inside the library's staging kernel (no intermediate merged copy), then the usual pass"""
    torch.cuda.synchronize(device)
    merged.stage_device_parts(parts, n_total)
    merged.process()
    return merged.collect() if collect else None


def merge_local(shards, merged, device=0):
    """Single process, several aggregators on one GPU (tests, or one host feeding one GPU from several rings).

    shards: list of (aggregator that has been stage()d and process()ed, int64 array of the global row index of each of its
    rows, ascending). merged: a PA_HASH_PROVIDED aggregator with the same registrations. Returns merged.collect()."""
    dev = torch.device("cuda", device)
    sizes = [a.shard_sizes() for a, _ in shards]
    n_total, nf_total = sum(s[0] for s in sizes), sum(s[1] for s in sizes)
    base, parts, keep = 0, [], []
    for (a, gidx), (n, nf) in zip(shards, sizes):
        hdr, frames = _export(a, base, dev)
        idx = torch.as_tensor(np.asarray(gidx, dtype=np.int64), device=dev)
        keep.append((hdr, frames, idx))
        parts.append((hdr.data_ptr(), idx.data_ptr(), n, frames.data_ptr(), nf))
        base += nf
    for a, _ in shards:
        a.discard()  # the shard's own record is not needed
    return _stage_parts(merged, parts, n_total, dev)


def merge_distributed(a, gidx, merged=None, dst=0, device=None, collect=True, sizes=None, phases=None):
    """One process per GPU (torch.distributed initialised). Every rank passes its processed shard aggregator and the global
    row indices of its rows; rank `dst` also passes the merging aggregator and gets the merged result, the others get None.
    NCCL: one batched group of point-to-point sends per rank to `dst`. With the gloo backend (CPU tests) the payload is staged
    through host memory; the library calls are the same. collect=False leaves the merged batch processed but uncollected on
    `dst` (benchmarks that time the device-resident step; follow with merged.collect() or merged.discard()).
    sizes: [(rows, frames)] per rank if already known (skips the all_gather). phases: a dict that receives wall-clock
    milliseconds per phase (adds device synchronisations; profiling only)."""
    import time

    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    via_host = dist.get_backend() != "nccl"
    comm_dev = torch.device("cpu") if via_host else dev
    t_last = [time.perf_counter()]

    def mark(name):
        if phases is not None:
            torch.cuda.synchronize(dev)
            now = time.perf_counter()
            phases[name] = phases.get(name, 0.0) + 1e3 * (now - t_last[0])
            t_last[0] = now

    n, nf = a.shard_sizes()
    if sizes is None:
        mine = torch.tensor([n, nf], dtype=torch.int64, device=comm_dev)
        allsz = [torch.zeros(2, dtype=torch.int64, device=comm_dev) for _ in range(world)]
        dist.all_gather(allsz, mine)
        sizes = [(int(t[0]), int(t[1])) for t in allsz]
    assert sizes[rank] == (n, nf)
    base = sum(s[1] for s in sizes[:rank])
    mark("sizes")
    hdr, frames = _export(a, base, dev)
    idx = gidx.to(dev) if torch.is_tensor(gidx) else torch.as_tensor(np.asarray(gidx, dtype=np.int64), device=dev)
    a.discard()
    mark("export")
    if rank != dst:
        payload = [t.to(comm_dev).contiguous() for t in (hdr, idx, frames)]
        if via_host:
            for t in payload:
                dist.send(t, dst)
        else:
            for w_ in dist.batch_isend_irecv([dist.P2POp(dist.isend, t, dst) for t in payload if t.numel()]):
                w_.wait()
        mark("send")
        return None
    n_total, nf_total = sum(s[0] for s in sizes), sum(s[1] for s in sizes)
    recv, ops = {}, []
    for r in range(world):
        if r == dst:
            continue
        rn, rnf = sizes[r]
        recv[r] = (torch.empty((max(rn, 1), 64), dtype=torch.uint8, device=comm_dev)[:rn], torch.empty(max(rn, 1), dtype=torch.int64, device=comm_dev)[:rn],
                   torch.empty(max(rnf, 1), dtype=torch.int64, device=comm_dev)[:rnf])
        if via_host:
            for t in recv[r]:
                dist.recv(t, r)
        else:
            ops += [dist.P2POp(dist.irecv, t, r) for t in recv[r] if t.numel()]
    if ops:
        for w_ in dist.batch_isend_irecv(ops):
            w_.wait()
    mark("recv")
    parts, keep = [], []
    for r in range(world):
        rn, rnf = sizes[r]
        h, i, f = (hdr, idx, frames) if r == dst else tuple(t.to(dev) for t in recv[r])
        keep.append((h, i, f))
        parts.append((h.data_ptr(), i.data_ptr(), rn, f.data_ptr(), rnf))
    out = _stage_parts(merged, parts, n_total, dev, collect)
    mark("scatter_and_merged_pass")
    return out


def shard_rows(w, world):
    """Global row indices of each shard of workload `w` under the pid-hash rule (ascending within a shard)."""
    from . import synth
    pids, inv = np.unique(w.hdrs["pid"], return_inverse=True)
    owner = np.array([synth.xxh64_u32(int(p)) % world for p in pids], dtype=np.int64)[inv]
    return [np.nonzero(owner == r)[0].astype(np.int64) for r in range(world)]
