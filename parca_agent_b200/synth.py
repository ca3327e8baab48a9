"""Deterministic synthetic sample batches for the configs in BASELINE.json.

A Workload is exactly what the reporter's ingest path reads per interval: registered strings,
distinct frames (pa_frame_desc), per-PID labelsets, and N (pa_sample_hdr, frame_id[]) samples.
Generation is numpy-vectorised and seeded (numpy PCG64), so the GPU box regenerates the same
bytes without shipping fixtures. No aggregation logic lives here.
"""
from dataclasses import dataclass, field

import numpy as np

from . import abi


class StringTable:
    """Interns bytes → dense string ids; id 0 is always b""."""

    def __init__(self):
        self.strings = [b""]
        self.index = {b"": 0}

    def sid(self, s):
        if isinstance(s, str):
            s = s.encode()
        i = self.index.get(s)
        if i is None:
            i = len(self.strings)
            self.index[s] = i
            self.strings.append(s)
        return i


@dataclass
class Workload:
    name: str
    strings: list
    frames: np.ndarray            # abi.FRAME_DTYPE
    labelsets: list               # list[list[(name_sid, value_sid)]] sorted by name bytes
    hdrs: np.ndarray              # abi.HDR_DTYPE
    stack_table: np.ndarray = None    # (U, F) uint64 frame ids, uniform-F workloads
    stack_choice: np.ndarray = None   # (N,) index into stack_table
    _frame_ids: np.ndarray = None     # explicit ragged stream (small workloads)
    hash_mode: int = abi.PA_HASH_XXH64X2
    label_flags: int = 0
    samples_per_second: int = 19
    external_labels: list = field(default_factory=list)  # [(name_sid, value_sid)]
    meta: dict = field(default_factory=dict)
    schema: int = abi.PA_SCHEMA_V2    # which sample record the reporter builds (v1 = stacktrace ids only)

    @property
    def n(self):
        return int(self.hdrs.shape[0])

    @property
    def n_frame_ids(self):
        if self._frame_ids is not None:
            return int(self._frame_ids.shape[0])
        return self.n * int(self.stack_table.shape[1])

    @property
    def frame_ids(self):
        if self._frame_ids is None:
            self._frame_ids = np.ascontiguousarray(self.stack_table[self.stack_choice]).reshape(-1)
        return self._frame_ids

    def write_frames(self, out):
        """Materialise the frame-id stream into `out` (uint64 or uint32 [n_frame_ids]) without a temp copy of the stream."""
        if self._frame_ids is not None:
            out[:] = self._frame_ids
        else:
            f = self.stack_table.shape[1]
            table = self.stack_table if out.dtype == self.stack_table.dtype else self.stack_table.astype(out.dtype)
            np.take(table, self.stack_choice, axis=0, out=out.reshape(self.n, f), mode="clip")  # indices are valid by construction; the default mode buffers `out` (45x slower)

    def head(self, n):
        """First n rows as a new workload (frames are a prefix of the stream)."""
        n = min(n, self.n)
        w = Workload(self.name + "[:%d]" % n, self.strings, self.frames, self.labelsets, self.hdrs[:n].copy(),
                     hash_mode=self.hash_mode, label_flags=self.label_flags, samples_per_second=self.samples_per_second,
                     external_labels=list(self.external_labels), meta=dict(self.meta), schema=self.schema)
        if self._frame_ids is None:
            w.stack_table, w.stack_choice = self.stack_table, self.stack_choice[:n]
        else:
            end = int(self.hdrs["frame_off"][n - 1] + self.hdrs["nframes"][n - 1]) if n else 0
            w._frame_ids = self._frame_ids[:end]
        return w

    def rows(self, idx):
        """Arbitrary row subset (in the given order) as a new workload with a repacked frame stream."""
        idx = np.asarray(idx, dtype=np.int64)
        hd = self.hdrs[idx].copy()
        nf = hd["nframes"].astype(np.int64)
        off = np.zeros(len(idx), dtype=np.uint64)
        if len(idx):
            off[1:] = np.cumsum(nf)[:-1]
        w = Workload(self.name + "[subset]", self.strings, self.frames, self.labelsets, hd,
                     hash_mode=self.hash_mode, label_flags=self.label_flags, samples_per_second=self.samples_per_second,
                     external_labels=list(self.external_labels), meta=dict(self.meta), schema=self.schema)
        if self._frame_ids is None:
            w.stack_table, w.stack_choice = self.stack_table, self.stack_choice[idx]
        else:
            src = self._frame_ids
            parts = [src[int(o):int(o) + int(k)] for o, k in zip(self.hdrs["frame_off"][idx], nf)]
            w._frame_ids = np.concatenate(parts) if parts else np.zeros(0, np.uint64)
        hd["frame_off"] = off
        return w


def splitmix64(x):
    with np.errstate(over="ignore"):
        z = np.atleast_1d(np.asarray(x, dtype=np.uint64)) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z if np.ndim(x) else z[0]


def _frame_table(rng, st, P, native, kernel, interp_name):
    """P distinct frames: `native`/`kernel` fractions, the rest interpreted frames of type `interp_name`."""
    fr = np.zeros(P, dtype=abi.FRAME_DTYPE)
    kind = np.full(P, abi.PA_FRAME_INTERP, dtype=np.uint8)
    u = rng.random(P)
    kind[u < native] = abi.PA_FRAME_NATIVE
    kind[(u >= native) & (u < native + kernel)] = abi.PA_FRAME_KERNEL
    fr["kind"] = kind
    t_native, t_kernel, t_interp = st.sid("native"), st.sid("kernel"), st.sid(interp_name)
    fr["type_name_sid"] = np.where(kind == abi.PA_FRAME_NATIVE, t_native, np.where(kind == abi.PA_FRAME_KERNEL, t_kernel, t_interp))
    fr["address_or_lineno"] = rng.integers(0x400000, 0x7FFFFFFFFFFF, P, dtype=np.uint64)
    # executables: 64 binaries; 1/8 have no build id (FileID hex fallback), 1/16 of native frames unknown
    nexe = 64
    exe_file = np.array([st.sid("/usr/bin/app-%02d" % e) for e in range(nexe)], dtype=np.uint32)
    exe_bid = np.array([st.sid("" if e % 8 == 7 else "%040x" % (0xabc0000 + e * 7919)) for e in range(nexe)], dtype=np.uint32)
    e = rng.integers(0, nexe, P)
    known = rng.random(P) >= 1.0 / 16
    is_native = kind == abi.PA_FRAME_NATIVE
    fr["flags"] = np.where(is_native, abi.PA_FRAME_F_MAPPING_FILE | np.where(known, abi.PA_FRAME_F_EXEC_KNOWN, 0), 0).astype(np.uint8)
    fr["exec_file_name_sid"] = np.where(is_native, exe_file[e], 0)
    fr["exec_build_id_sid"] = np.where(is_native, exe_bid[e], 0)
    fr["file_id_hi"] = np.where(is_native, splitmix64(e.astype(np.uint64)), 0)
    fr["file_id_lo"] = np.where(is_native, splitmix64(e.astype(np.uint64) + np.uint64(1000)), 0)
    # symbolised frames: function names shared by ~4 frames each; 1/32 unsymbolised
    sym = ~is_native
    nsym = int(sym.sum())
    nfunc = max(1, nsym // 4)
    kfun = [st.sid("ksym_%d" % i) if i % 3 else st.sid("kernel_function_with_long_name_%06d" % i) for i in range(nfunc)]
    ifun = [st.sid("%s_module.function_name_%07d" % (interp_name, i)) for i in range(nfunc)]
    ifile = [st.sid("/usr/lib/%s/site/mod_%05d.src" % (interp_name, i)) for i in range(max(1, nfunc // 8))]
    fidx = rng.integers(0, nfunc, P)
    unsym = rng.random(P) < 1.0 / 32
    kfun, ifun, ifile = np.asarray(kfun, np.uint32), np.asarray(ifun, np.uint32), np.asarray(ifile, np.uint32)
    fname = np.where(kind == abi.PA_FRAME_KERNEL, kfun[fidx], ifun[fidx])
    fr["function_name_sid"] = np.where(sym & ~unsym, fname, 0)
    fr["source_file_sid"] = np.where((kind == abi.PA_FRAME_INTERP) & ~unsym, ifile[fidx % len(ifile)], 0)
    fr["source_line"] = np.where(sym, rng.integers(1, 5000, P), 0)
    # v1-only attributes, derived without touching the generator so every v2 workload stays byte-identical:
    # a source column on symbolised frames, and a GNU build id + mapping file name on a quarter of the interpreted ones
    ix = np.arange(P, dtype=np.uint64)
    fr["source_column"] = np.where(sym, (splitmix64(ix + np.uint64(77)) % np.uint64(160)).astype(np.uint32), 0)
    libs = np.array([st.sid("/usr/lib/%s/lib%s-ext-%02d.so" % (interp_name, interp_name, i)) for i in range(16)], dtype=np.uint32)
    gnus = np.array([st.sid("%040x" % (0x9e0000 + i * 104729)) for i in range(16)], dtype=np.uint32)
    has_gnu = (kind == abi.PA_FRAME_INTERP) & (ix % np.uint64(4) == 0)
    pick = (splitmix64(ix + np.uint64(991)) % np.uint64(16)).astype(np.int64)
    fr["mapping_file_name_sid"] = np.where(has_gnu, libs[pick], 0)
    fr["gnu_build_id_sid"] = np.where(has_gnu, gnus[pick], 0)
    return fr


def _uniform_batch(name, seed, N, F, U, P, npids, threads_per_pid, ncpu, native, kernel, interp_name,
                   kind=abi.PA_KIND_CPU, zipf=False, labelsets_per_pid=1, hash_mode=abi.PA_HASH_XXH64X2):
    rng = np.random.Generator(np.random.PCG64(seed))
    st = StringTable()
    frames = _frame_table(rng, st, P, native, kernel, interp_name)
    stack_table = rng.integers(0, P, (U, F), dtype=np.uint64)
    if zipf:
        wgt = 1.0 / np.arange(1, U + 1)
        cdf = np.cumsum(wgt / wgt.sum())
        choice = np.searchsorted(cdf, rng.random(N)).astype(np.int64)
        np.minimum(choice, U - 1, out=choice)
        choice = rng.permutation(U)[choice]  # hot stacks are not the low indices
    else:
        choice = rng.integers(0, U, N)
    n_comm, n_node, n_custom = st.sid("comm"), st.sid("node"), st.sid("request_class")
    v_node = st.sid("node-0")
    labelsets = []
    for p in range(npids):
        base = [(n_comm, st.sid("proc-%05d" % p)), (n_node, v_node)]
        if labelsets_per_pid == 1:
            labelsets.append(base)
        else:
            for c in range(labelsets_per_pid):
                labelsets.append(sorted(base + [(n_custom, st.sid("class-%03d" % c))], key=lambda kv: st.strings[kv[0]]))
    thread_comm = np.array([st.sid("worker-%02d" % t) for t in range(threads_per_pid)], dtype=np.uint32)
    pid_idx = rng.integers(0, npids, N)
    thr_idx = rng.integers(0, threads_per_pid, N)
    hd = np.zeros(N, dtype=abi.HDR_DTYPE)
    hd["pid"] = 1000 + pid_idx
    hd["tid"] = 100000 + pid_idx * threads_per_pid + thr_idx
    hd["comm_sid"] = thread_comm[thr_idx]
    hd["labelset_id"] = pid_idx * labelsets_per_pid + (rng.integers(0, labelsets_per_pid, N) if labelsets_per_pid > 1 else 0)
    hd["cpu"] = rng.integers(0, ncpu, N)
    hd["timestamp_ns"] = 1_700_000_000_000_000_000 + np.arange(N, dtype=np.int64) * 52631
    hd["kind"] = kind
    if kind != abi.PA_KIND_CPU:
        hd["value"] = np.exp(rng.uniform(np.log(1e3), np.log(1e7), N)).astype(np.int64)  # log-uniform [1us, 10ms]
    hd["nframes"] = F
    hd["frame_off"] = np.arange(N, dtype=np.uint64) * np.uint64(F)
    hd["hash_hi"] = splitmix64(choice.astype(np.uint64))
    hd["hash_lo"] = splitmix64(choice.astype(np.uint64) ^ np.uint64(0xD1B54A32D192ED03))
    return Workload(name, st.strings, frames, labelsets, hd, stack_table=stack_table, stack_choice=choice, hash_mode=hash_mode,
                    meta={"N": N, "F": F, "U": U, "P": P, "seed": seed})


def config1(hash_mode=abi.PA_HASH_XXH64X2):
    """100k samples, 16-frame stacks, 1k unique stacks (plumbing config)."""
    return _uniform_batch("cfg1_100k_f16_u1k", 0x5EED0001, 100_000, 16, 1_000, 4_096, 64, 4, 8, 0.8, 0.1, "python", hash_mode=hash_mode)


def config2(n=10_000_000, hash_mode=abi.PA_HASH_XXH64X2, u=100_000, p=262_144):
    """Headline: 10M samples, 64-frame stacks, 100k unique stacks."""
    return _uniform_batch("cfg2_10M_f64_u100k" if n == 10_000_000 else "cfg2_scaled_%d" % n, 0x5EED0002, n, 64, u, p, 4_096, 16, 192,
                          0.8, 0.1, "python", hash_mode=hash_mode)


def config3(n=10_000_000, hash_mode=abi.PA_HASH_XXH64X2, u=200_000, p=131_072, npids=5_000, lsets=10):
    """parcagpu-style: Zipf-skewed 32-frame stacks, 50k labelsets, CUDA origin."""
    return _uniform_batch("cfg3_10M_f32_zipf" if n == 10_000_000 else "cfg3_scaled_%d" % n, 0x5EED0003, n, 32, u, p, npids, 8, 192,
                          0.6, 0.0, "cuda", kind=abi.PA_KIND_CUDA, zipf=True, labelsets_per_pid=lsets, hash_mode=hash_mode)


_M64 = (1 << 64) - 1


def xxh64_u32(x, seed=0):
    """XXH64 of one little-endian uint32 (the pid shard key, north_star: shard = xxh64(pid) mod G)."""
    p1, p2, p3, p5 = 11400714785074694791, 14029467366897019727, 1609587929392839161, 2870177450012600261
    h = (seed + p5 + 4) & _M64
    h ^= (x * p1) & _M64
    h = ((((h << 23) | (h >> 41)) & _M64) * p2 + p3) & _M64
    h ^= h >> 33
    h = (h * p2) & _M64
    h ^= h >> 29
    h = (h * p3) & _M64
    h ^= h >> 32
    return h


def _pid_shard(name, seed, rank, world, n, u, p, pids_per_rank, hash_mode):
    """Rows of one rank: the pids with xxh64(pid) % world == rank, `n` samples drawn among them."""
    npids = pids_per_rank * world
    owner = np.array([xxh64_u32(1000 + q) % world for q in range(npids)], dtype=np.int64)
    mine = np.nonzero(owner == rank)[0]
    w = _uniform_batch("%s_shard%d_of_%d" % (name, rank, world), seed + rank, n, 64, u, p, len(mine), 16, 192, 0.8, 0.1, "python",
                       hash_mode=hash_mode)
    local = (w.hdrs["pid"] - 1000).astype(np.int64)
    w.hdrs["pid"] = 1000 + mine[local]  # global pid space: shards are disjoint by construction
    w.meta.update(rank=rank, world=world)
    return w


def config2_shard(rank, world, n=10_000_000, hash_mode=abi.PA_HASH_XXH64X2):
    """Weak-scaling shard of the headline config: every rank aggregates `n` samples of its own pids."""
    return _pid_shard("cfg2", 0x5EED0002, rank, world, n, 100_000, 262_144, 4_096, hash_mode)


def config4_shard(rank, world, n_total=100_000_000, hash_mode=abi.PA_HASH_XXH64X2):
    """Config 4: 100M samples, 1M unique stacks, 65,536 pids sharded by xxh64(pid) mod world."""
    return _pid_shard("cfg4", 0x5EED0004, rank, world, n_total // world, 1_000_000 // world * 2, 1_048_576, 65_536 // world, hash_mode)


_CONFIG4_TABLES = {}


def _config4_tables(U, P, npids, F, threads, world, seed):
    """The tables every rank of a merged batch shares (strings, frames, stack table, labelsets): built once per process."""
    key = (U, P, npids, F, threads, world, seed)
    if key not in _CONFIG4_TABLES:
        _CONFIG4_TABLES.clear()
        rng = np.random.Generator(np.random.PCG64(seed))
        st = StringTable()
        frames = _frame_table(rng, st, P, 0.8, 0.1, "python")
        stack_table = rng.integers(0, P, (U, F), dtype=np.uint64)
        n_comm, n_node, v_node = st.sid("comm"), st.sid("node"), st.sid("node-0")
        labelsets = [[(n_comm, st.sid("proc-%05d" % q)), (n_node, v_node)] for q in range(npids)]
        thread_comm = np.array([st.sid("worker-%02d" % t) for t in range(threads)], dtype=np.uint32)
        owner = np.array([xxh64_u32(1000 + q) % world for q in range(npids)], dtype=np.int64)
        _CONFIG4_TABLES[key] = (st, frames, stack_table, labelsets, thread_comm, owner)
    return _CONFIG4_TABLES[key]


def config4_part(rank, world, rows_per_gpu=12_500_000, stacks_per_gpu=125_000, frames_per_gpu=131_072, pids_per_gpu=8_192,
                 hash_mode=abi.PA_HASH_XXH64X2, seed=0x5EED0004):
    """One rank's ring of a MERGED batch (mode B): BASELINE config 4 at world == 8 (100M samples x 64 frames, 1M unique stacks,
    1,048,576 distinct frames, 65,536 pids sharded by xxh64(pid) mod world), proportionally smaller at other world sizes.
    Strings, frames, labelsets and the stack table are common to every rank (the merge needs identical registrations: ids
    are global); the rows are this rank's pids only. Every rank draws its stacks from the WHOLE stack table, so nearly every
    stack occurs on every GPU — the hardest case for the dictionary merge."""
    U, P, npids, F = stacks_per_gpu * world, frames_per_gpu * world, pids_per_gpu * world, 64
    threads = 16
    st, frames, stack_table, labelsets, thread_comm, owner = _config4_tables(U, P, npids, F, threads, world, seed)
    mine = np.nonzero(owner == rank)[0]
    rr = np.random.Generator(np.random.PCG64(seed + 1 + rank))
    N = rows_per_gpu
    choice = rr.integers(0, U, N)
    pid_idx = mine[rr.integers(0, len(mine), N)]
    thr_idx = rr.integers(0, threads, N)
    hd = np.zeros(N, dtype=abi.HDR_DTYPE)
    hd["pid"] = 1000 + pid_idx
    hd["tid"] = 100000 + pid_idx * threads + thr_idx
    hd["comm_sid"] = thread_comm[thr_idx]
    hd["labelset_id"] = pid_idx
    hd["cpu"] = rr.integers(0, 192, N)
    hd["timestamp_ns"] = 1_700_000_000_000_000_000 + (np.arange(N, dtype=np.int64) * world + rank) * 52631
    hd["kind"] = abi.PA_KIND_CPU
    hd["nframes"] = F
    hd["frame_off"] = np.arange(N, dtype=np.uint64) * np.uint64(F)
    hd["hash_hi"] = splitmix64(choice.astype(np.uint64))
    hd["hash_lo"] = splitmix64(choice.astype(np.uint64) ^ np.uint64(0xD1B54A32D192ED03))
    return Workload("cfg4_part%d_of_%d" % (rank, world), st.strings, frames, labelsets, hd, stack_table=stack_table, stack_choice=choice,
                    hash_mode=hash_mode, meta={"N": N, "F": F, "U": U, "P": P, "seed": seed, "rank": rank, "world": world, "N_total": N * world})


def config5_part(rank, world, rows_per_gpu=47_500_000, hash_mode=abi.PA_HASH_XXH64X2):
    """One rank's share of a BASELINE config-5 window: 19 Hz x 1M threads x 5 s = 95M samples per window over `world` GPUs
    (47.5M rows per GPU at world == 2), 64-frame stacks, 500k unique stacks, 62,500 pids x 16 threads = 1M threads."""
    total_pids = 62_500
    return config4_part(rank, world, rows_per_gpu=rows_per_gpu, stacks_per_gpu=500_000 // world, frames_per_gpu=524_288 // world,
                        pids_per_gpu=total_pids // world, hash_mode=hash_mode, seed=0x5EED0005)


def concat(parts):
    """The stream [part 0's rows, part 1's rows, ...] as one workload (what a merged batch must reproduce). Parts must share
    their tables (config4_part / Workload.rows of one workload)."""
    w0 = parts[0]
    hd = np.concatenate([p.hdrs for p in parts])
    nf = hd["nframes"].astype(np.int64)
    off = np.zeros(len(hd), dtype=np.uint64)
    if len(hd):
        off[1:] = np.cumsum(nf)[:-1]
    hd["frame_off"] = off
    w = Workload("concat", w0.strings, w0.frames, w0.labelsets, hd, hash_mode=w0.hash_mode, label_flags=w0.label_flags,
                 samples_per_second=w0.samples_per_second, external_labels=list(w0.external_labels), meta=dict(w0.meta), schema=w0.schema)
    if all(p._frame_ids is None for p in parts):
        w.stack_table, w.stack_choice = w0.stack_table, np.concatenate([p.stack_choice for p in parts])
    else:
        w._frame_ids = np.concatenate([p.frame_ids for p in parts])
    return w


def ragged(n=2_000_000, u=50_000, p=65_536, max_depth=127, seed=0x5EED00AA, hash_mode=abi.PA_HASH_XXH64X2):
    """Real-world-like stack depths: every unique stack has its own depth in 1..max_depth (not a BASELINE config; used
    to check that the ragged code paths of the hash kernel stay fast and exact)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    base = _uniform_batch("ragged", seed, n, 1, u, p, 256, 8, 64, 0.8, 0.1, "python", hash_mode=hash_mode)
    depth = rng.integers(1, max_depth + 1, u)
    starts = np.zeros(u + 1, dtype=np.int64)
    starts[1:] = np.cumsum(depth)
    pool = rng.integers(0, p, int(starts[-1]), dtype=np.uint64)
    choice = base.stack_choice
    nf = depth[choice]
    off = np.zeros(n, dtype=np.uint64)
    off[1:] = np.cumsum(nf)[:-1]
    idx = np.repeat(starts[choice], nf) + (np.arange(int(nf.sum())) - np.repeat(off.astype(np.int64), nf))
    base.hdrs["nframes"] = nf
    base.hdrs["frame_off"] = off
    return Workload("ragged_%d" % n, base.strings, base.frames, base.labelsets, base.hdrs, _frame_ids=pool[idx], hash_mode=hash_mode,
                    meta={"N": n, "U": u, "P": p, "F": "1..%d" % max_depth})


def edge_workload(seed=7, n=600, hash_mode=abi.PA_HASH_PROVIDED, label_flags=0, external=True):
    """Small adversarial batch: ragged stacks (0..9 frames), every sample kind, every frame kind,
    labelsets with missing names (null back-fill), empty comm (thread_name dropped), hash
    collisions in provided mode (same id, different frames), external labels that are new /
    fully present / partially present, strings longer than 12 bytes and longer than a view block.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    st = StringTable()
    t = {k: st.sid(k) for k in ["native", "kernel", "python", "abort-marker", "ruby"]}
    exe = [(st.sid("/usr/bin/alpha"), st.sid("deadbeef01")), (st.sid("/opt/very/long/path/to/binary-beta"), st.sid("")),
           (st.sid("nvidia.ko"), st.sid("kmodbuild"))]
    big = st.sid("F" * 40000)  # longer than one 32 KiB string-view block
    fnames = [st.sid(s) for s in ["do_syscall_64", "sys_read", "a", "exactly12byt", "thirteen_byte", "mod.fn_" + "x" * 50, "π_unicode_函数"]] + [big]
    files = [st.sid(s) for s in ["a.py", "/srv/app/handlers/request_dispatch.py", ""]]
    descs = []

    def add(kind, **kw):
        d = dict(kind=kind, flags=0, type_name_sid=0, address_or_lineno=0, function_name_sid=0, source_file_sid=0, source_line=0,
                 exec_file_name_sid=0, exec_build_id_sid=0, file_id_hi=0, file_id_lo=0, source_column=0, mapping_file_name_sid=0, gnu_build_id_sid=0)
        d.update(kw)
        descs.append(d)

    for i in range(12):  # native: known w/ build id, known w/o build id (FileID hex), unknown, no mapping
        e = exe[i % 2]
        fl = [3, 3, 1, 0][i % 4]
        add(abi.PA_FRAME_NATIVE, flags=fl, type_name_sid=t["native"], address_or_lineno=0x1000 + 16 * i, exec_file_name_sid=e[0],
            exec_build_id_sid=e[1], file_id_hi=0x0123456789abcdef + i, file_id_lo=0xfedcba9876543210 - i)
    for i in range(10):  # kernel: named/unnamed, module known/unknown
        add(abi.PA_FRAME_KERNEL, flags=(3 if i % 3 == 0 else 0), type_name_sid=t["kernel"], address_or_lineno=0xffffffff81000000 + i,
            function_name_sid=(fnames[i % len(fnames)] if i % 4 else 0), source_line=100 + i, exec_file_name_sid=exe[2][0], source_column=(7 * i) % 5)
    add(abi.PA_FRAME_ABORT, type_name_sid=t["abort-marker"], address_or_lineno=0)
    add(abi.PA_FRAME_ABORT, type_name_sid=t["abort-marker"], address_or_lineno=1)
    add(abi.PA_FRAME_OOMPROF, type_name_sid=t["native"], address_or_lineno=0x77, source_file_sid=files[1], function_name_sid=st.sid("buildid-oom"))
    gnu = [(st.sid("/usr/lib/libpython3.11.so.1.0"), st.sid("gnubuildid-aaaa")), (st.sid("/usr/lib/libruby.so.3.2"), st.sid("gnubuildid-bbbb"))]
    for i in range(14):  # interpreted: python/ruby, with/without function name, empty file path, with/without a GNU build id (v1)
        add(abi.PA_FRAME_INTERP, type_name_sid=t["python" if i % 2 else "ruby"], address_or_lineno=10 + i,
            function_name_sid=(fnames[(i * 3) % len(fnames)] if i % 5 else 0), source_file_sid=files[i % 3], source_line=i, source_column=i % 4,
            mapping_file_name_sid=(gnu[i % 2][0] if i % 3 == 0 else 0), gnu_build_id_sid=(gnu[i % 2][1] if i % 3 == 0 else 0))
    frames = np.zeros(len(descs), dtype=abi.FRAME_DTYPE)
    for i, d in enumerate(descs):
        for k, v in d.items():
            frames[k][i] = v
    P = len(descs)

    names = {k: st.sid(k) for k in ["comm", "node", "job", "cpu", "zone", "thread_name"]}
    vals = [st.sid(v) for v in ["svc-a", "svc-b", "node-0", "batch", "eu-west-1", "override-me", "x" * 300]]
    labelsets = [
        [(names["comm"], vals[0]), (names["node"], vals[2])],
        [(names["comm"], vals[1]), (names["job"], vals[3]), (names["node"], vals[2])],
        [(names["node"], vals[2])],
        [(names["comm"], vals[0]), (names["cpu"], vals[5]), (names["node"], vals[2]), (names["thread_name"], vals[6])],
        [],
        [(names["job"], vals[6]), (names["zone"], vals[4])],
    ]
    comms = [st.sid(c) for c in ["main", "worker", "", "GC Thread#0"]]
    hd = np.zeros(n, dtype=abi.HDR_DTYPE)
    stream = []
    # a pool of stacks, some sharing a provided hash with different frames (collision ⇒ first wins)
    pool = []
    for s in range(40):
        k = int(rng.integers(0, 10))
        fr = rng.integers(0, P, k).astype(np.uint64)
        hid = s if s % 9 else 0  # stacks 0,9,18,... share hash id 0
        pool.append((hid, fr))
    off = 0
    run_ls = 0
    for i in range(n):
        if rng.random() < 0.3:
            run_ls = int(rng.integers(0, len(labelsets)))
        hid, fr = pool[int(rng.integers(0, len(pool)))]
        hd["hash_hi"][i] = splitmix64(np.uint64(hid))
        hd["hash_lo"][i] = splitmix64(np.uint64(hid + 77))
        hd["timestamp_ns"][i] = 1_700_000_000_000_000_000 + i * 1000 - (3000 if i % 50 == 0 else 0)
        kind = int(rng.integers(0, 7)) if rng.random() < 0.2 else (abi.PA_KIND_CPU if (i // 64) % 2 == 0 else abi.PA_KIND_CUDA)
        hd["kind"][i] = kind
        hd["value"][i] = int(rng.integers(-5, 1 << 40))
        hd["pid"][i] = 10 + run_ls
        hd["tid"][i] = int(rng.choice([1, 9, 10, 99, 1234, 4294967295, 65536, 7]))
        hd["comm_sid"][i] = comms[int(rng.integers(0, len(comms)))] if rng.random() < 0.5 else comms[0]
        hd["labelset_id"][i] = run_ls
        hd["cpu"][i] = int(rng.integers(0, 3)) if rng.random() < 0.6 else 2
        hd["nframes"][i] = len(fr)
        hd["frame_off"][i] = off
        stream.append(fr)
        off += len(fr)
    frame_ids = np.concatenate(stream) if stream else np.zeros(0, np.uint64)
    ext = []
    if external:
        ext = [(st.sid("cluster"), st.sid("prod")), (names["node"], st.sid("node-ext")), (names["zone"], st.sid("zone-ext"))]
    return Workload("edge_seed%d" % seed, st.strings, frames, labelsets, hd, _frame_ids=frame_ids, hash_mode=hash_mode,
                    label_flags=label_flags, samples_per_second=19, external_labels=ext, meta={"seed": seed})
