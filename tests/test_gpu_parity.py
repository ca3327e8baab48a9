"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Bar: bit-exact Arrow IPC stream bytes. On a mismatch the decoded physical components are diffed
(tests/pyref.py) so the failure names the first differing buffer element.
"""
import numpy as np
import pyarrow as pa
import pytest

import kat_workloads as kw
import pyref
from parca_agent_b200 import abi, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from parca_agent_b200 import lib
    lib.lib()
    return lib


def assert_same(oracle, gpu, w, **kw_run):
    want, st = oracle.run(w)
    got, r = gpu.run(w, **kw_run)
    assert r.n_rows == st["rows"]
    if got != want:
        d = None
        if want and got:
            ex = pyref.extract_v1 if getattr(w, "schema", 0) == abi.PA_SCHEMA_V1 else pyref.extract
            xa = ex(pa.ipc.open_stream(want).read_all())
            xb = ex(pa.ipc.open_stream(got).read_all())
            d = pyref.diff(xa, xb)
        raise AssertionError("IPC bytes differ (len %d vs %d); first logical difference: %s" % (len(want), len(got), d))
    if getattr(w, "schema", 0) == abi.PA_SCHEMA_V1:
        assert r.n_unique_stacks == st["unique_stacks"]
    else:
        assert (r.n_unique_stacks, r.n_locations, r.n_functions, r.n_location_indices) == (
            st["unique_stacks"], st["locations"], st["functions"], st["location_indices"])
    return r


@pytest.mark.parametrize("name", ["stack_dedup", "writer_basic", "multiple_frame_types", "func_dedup_in_stack", "null_lines"])
def test_reference_unit_test_inputs(oracle, gpu, name):
    assert_same(oracle, gpu, getattr(kw, name)())


@pytest.mark.parametrize("flags", [0, 1, 2, 4, 7])
def test_labels_for_tid_inputs(oracle, gpu, flags):
    w, cpus = kw.labels_cpu_sequence(flags)
    assert_same(oracle, gpu, w)


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
@pytest.mark.parametrize("mode", [abi.PA_HASH_PROVIDED, abi.PA_HASH_XXH64X2])
@pytest.mark.parametrize("external", [False, True])
def test_edge_batches(oracle, gpu, seed, mode, external):
    assert_same(oracle, gpu, synth.edge_workload(seed=seed, hash_mode=mode, external=external))


@pytest.mark.parametrize("flags", [1, 2, 4, 7, 3])
def test_edge_label_flags(oracle, gpu, flags):
    assert_same(oracle, gpu, synth.edge_workload(seed=11, label_flags=flags, external=True))


def test_edge_large_and_chunked(oracle, gpu):
    w = synth.edge_workload(seed=21, n=20000, hash_mode=abi.PA_HASH_XXH64X2)
    assert_same(oracle, gpu, w)
    assert_same(oracle, gpu, w, chunk_samples=777)  # many H2D chunks, copy/compute overlap path


def test_empty_batch(gpu):
    w = synth.edge_workload(seed=1)
    a = gpu.from_workload(w)
    r = a.flush()
    assert r.n_rows == 0 and r.ipc_len == 0
    a.close()


def test_config1_full(oracle, gpu):
    r = assert_same(oracle, gpu, synth.config1())
    assert r.n_unique_stacks == 1000 and r.gpu_launches > 0


def test_config1_provided_hash(oracle, gpu):
    assert_same(oracle, gpu, synth.config1(hash_mode=abi.PA_HASH_PROVIDED))


def test_config2_scaled(oracle, gpu):
    assert_same(oracle, gpu, synth.config2(n=400_000, u=20_000, p=32_768), chunk_samples=65536)


def test_config3_scaled_zipf(oracle, gpu):
    assert_same(oracle, gpu, synth.config3(n=300_000, u=30_000, p=16_384, npids=500, lsets=10))


def test_config2_full(oracle, gpu):
    """BASELINE.json config 2 at its full size (10M samples x 64 frames, 100k stacks): the IPC stream of one flush through
    the C ABI equals the oracle's byte for byte (the same comparison bench.py repeats on its timed end-to-end flush)."""
    w = synth.config2()
    r = assert_same(oracle, gpu, w, chunk_samples=1 << 20)
    assert r.n_rows == 10_000_000 and r.n_unique_stacks == 100_000


def test_config3_full(oracle, gpu):
    """BASELINE.json config 3 at its full size (10M samples, Zipf-skewed 32-frame stacks, 50k labelsets, CUDA origin)."""
    w = synth.config3()
    r = assert_same(oracle, gpu, w, chunk_samples=1 << 20)
    assert r.n_rows == 10_000_000 and r.n_unique_stacks == len(np.unique(w.stack_choice))


@pytest.mark.parametrize("mode", [abi.PA_HASH_PROVIDED, abi.PA_HASH_XXH64X2])
def test_narrow_frame_id_ring(oracle, gpu, mode):
    """pa_agg_config.frame_id_bytes = 4: the ring and the HBM frame stream hold uint32 ids (half the bytes); every output byte
    is the same, including ragged stacks at odd offsets, empty stacks and stacks deeper than 64 frames"""
    for w in (synth.edge_workload(seed=51, n=6000, hash_mode=mode), synth.config1(hash_mode=mode).head(50_000), wide_label_workload(12, n=800, seed=3)):
        w.hash_mode = mode
        assert_same(oracle, gpu, w, frame_id_bytes=4)
        assert_same(oracle, gpu, w, frame_id_bytes=4, chunk_samples=777)
    assert_same(oracle, gpu, as_v1(synth.edge_workload(seed=52, n=3000, hash_mode=mode)), frame_id_bytes=4)


def test_early_copy_out_of_trailing_columns(oracle, gpu, monkeypatch):
    """A multi-chunk flush copies stacktrace_id / value / timestamp to the host while the ids are still uploading, at positions
    counted from the END of the output buffer of the previous flush. Intervals of different sizes and shapes through ONE
    aggregator: bigger than the buffer (re-allocation drops the early copies), smaller (the stream starts in the middle of the
    buffer), mixed sample kinds (many kind runs: too many -> ordinary path), LZ4 bodies, a narrow ring."""
    monkeypatch.setenv("PA_EARLY_D2H_ALWAYS", "1")
    base = synth.config1()
    edge = synth.edge_workload(seed=61, n=9000, hash_mode=abi.PA_HASH_XXH64X2, external=False)
    for kw in ({}, {"frame_id_bytes": 4}, {"ipc_compression": abi.PA_IPC_LZ4_FRAME}):
        a = gpu.from_workload(base, chunk_samples=4096, **kw)
        for n in ((20_000, 20_000, 60_000, 5_000, 33_333, 1) if not kw else (10_000, 30_000, 4_000)):
            w = base.head(n)
            want, st = oracle.run(w)
            gpu.load(a, w)
            res = a.flush()
            got = res.ipc_bytes()
            if kw.get("ipc_compression"):
                assert pa.ipc.open_stream(got).read_all().equals(pa.ipc.open_stream(want).read_all())
            else:
                assert got == want, "interval of %d rows differs" % n
            assert res.n_rows == st["rows"]
        a.close()
    a = gpu.from_workload(base, chunk_samples=4096)   # a batch dropped after its early copies were started must not leak into the next one
    w = base.head(30_000)
    gpu.load(a, w)
    a.flush()
    gpu.load(a, base.head(25_000))
    a.stage(); a.process(); a.discard()
    gpu.load(a, w)
    assert a.flush().ipc_bytes() == oracle.run(w)[0]
    a.close()
    a = gpu.from_workload(edge, chunk_samples=1000)   # every sample kind, alternating: hundreds of kind runs
    want, _ = oracle.run(edge)
    for _ in range(3):
        gpu.load(a, edge)
        assert a.flush().ipc_bytes() == want
    a.close()
    w = synth.config1().head(30_000)
    w.hdrs["kind"][:] = np.arange(w.n) % 2 * abi.PA_KIND_OFFCPU  # 30k kind runs: far beyond what the early path sizes
    a = gpu.from_workload(w, chunk_samples=5000)
    want, _ = oracle.run(w)
    for _ in range(2):
        gpu.load(a, w)
        assert a.flush().ipc_bytes() == want
    a.close()


def test_single_ring_flag(oracle, gpu):
    """PA_CFG_SINGLE_RING: one ring buffer; ingest is refused while a staged batch holds it, and works again after collect"""
    w = synth.config1().head(20_000)
    want, _ = oracle.run(w)
    a = gpu.from_workload(w, flags=abi.PA_CFG_SINGLE_RING)
    for _ in range(3):
        gpu.load(a, w)
        assert a.flush().ipc_bytes() == want
    gpu.load(a, w)
    a.stage()
    with pytest.raises(gpu.PaError) as e:
        a.acquire(1, 0)
    assert e.value.code == -28  # PA_ENOSPC
    a.process()
    assert a.collect().ipc_bytes() == want
    gpu.load(a, w)
    assert a.flush().ipc_bytes() == want
    a.close()


def test_stack_ids_match_xxh64(oracle, gpu):
    w = synth.edge_workload(seed=5, hash_mode=abi.PA_HASH_XXH64X2)
    a = gpu.from_workload(w)
    gpu.load(a, w)
    a.stage()
    a.process()
    ids = a.debug_stack_ids(w.n)
    fr = w.frame_ids
    for r in range(0, w.n, 7):
        h = w.hdrs[r]
        data = fr[int(h["frame_off"]):int(h["frame_off"]) + int(h["nframes"])].astype("<u8").tobytes()
        want = oracle.xxh64(data, 0).to_bytes(8, "big") + oracle.xxh64(data, abi.PA_XXH_SEED_LO).to_bytes(8, "big")
        assert ids[r].tobytes() == want
    res = a.collect()
    counts = a.debug_stack_counts(res.n_unique_stacks)
    assert int(counts.sum()) == w.n  # every sample counted exactly once (the per-stack count side table)
    a.close()


@pytest.mark.parametrize("make", [lambda: synth.edge_workload(seed=9, n=5000), lambda: synth.config3(n=120_000, u=5_000, p=4_096, npids=200, lsets=6),
                                  lambda: synth.config1(hash_mode=abi.PA_HASH_PROVIDED)])
def test_pair_counts_side_table(gpu, make):
    """(labelset, stack) -> count, first-occurrence order: the hash-and-count view of a batch (not part of the reference's
    record, which keeps one row per sample). Checked against a plain dictionary walk over the rows."""
    w = make()
    a = gpu.from_workload(w)
    gpu.load(a, w)
    a.stage()
    a.process()
    ids = a.debug_stack_ids(w.n)
    stack_ord, pairs = {}, {}
    for r in range(w.n):
        sid = ids[r].tobytes()
        o = stack_ord.setdefault(sid, len(stack_ord))
        key = (int(w.hdrs["labelset_id"][r]), o)
        pairs[key] = pairs.get(key, 0) + 1
    ls, st, ct, n = a.debug_pair_counts(w.n)
    assert n == len(pairs) and len(ls) == n
    assert list(zip(ls.tolist(), st.tolist())) == list(pairs.keys()) and ct.tolist() == list(pairs.values())
    assert int(ct.sum()) == w.n
    ls2, st2, ct2, n2 = a.debug_pair_counts(3)          # a short buffer reports the total and fills what fits
    assert n2 == n and ls2.tolist() == ls[:3].tolist() and ct2.tolist() == ct[:3].tolist()
    res = a.collect()
    per_stack = a.debug_stack_counts(res.n_unique_stacks)
    assert np.array_equal(np.bincount(st, weights=ct, minlength=res.n_unique_stacks).astype(np.uint32), per_stack)
    a.close()


def test_staged_pipeline_equals_flush_and_repeats(oracle, gpu):
    """stage/process/collect == flush; process() is repeatable; the ring double-buffers across flushes."""
    w = synth.config1().head(30000)
    want, _ = oracle.run(w)
    a = gpu.from_workload(w, max_samples=40000, max_frames=40000 * 16)
    for _ in range(3):
        gpu.load(a, w)
        a.stage()
        a.process()
        a.process()
        assert a.collect().ipc_bytes() == want
        gpu.load(a, w)
        assert a.flush().ipc_bytes() == want
    ms, launches = a.kernel_ms("total")
    assert ms > 0 and launches > 0
    a.close()


def test_two_batches_different_content(oracle, gpu):
    """Dictionaries are rebuilt per interval: batch 2 must not see batch 1's state."""
    w = synth.edge_workload(seed=31, n=3000)
    a = gpu.from_workload(w, max_samples=4000, max_frames=40000)
    for part in (w.head(1000), w.rows(np.arange(1000, 3000))):
        want, _ = oracle.run(part)
        gpu.load(a, part)
        assert a.flush().ipc_bytes() == want
    a.close()


def test_bad_ids_are_reported(gpu):
    w = synth.edge_workload(seed=1, n=50)
    w.hdrs["labelset_id"][3] = 10_000
    a = gpu.from_workload(w)
    gpu.load(a, w)
    with pytest.raises(gpu.PaError):
        a.flush()
    a.close()


@pytest.mark.parametrize("variant", ["direct", "staged", "wide", "widenp", "widepf", "widepf3", "bulk", "bulk6x2", "tma", "tma12x4", "tma24x2", "tma12x2r", "tma13x2r", "tma8x3r", "tmag13x2", "tmag9x3", "tmag6x4"])
def test_hash_kernel_variants(oracle, gpu, variant, monkeypatch):
    """Both hash kernels (direct global loads / cp.async-staged) must produce identical bytes, including
    ragged stacks, odd frame offsets (8-byte staging path) and stacks deeper than one staging slot."""
    monkeypatch.setenv("PA_HASH_VARIANT", variant)
    assert_same(oracle, gpu, synth.edge_workload(seed=41, n=5000, hash_mode=abi.PA_HASH_XXH64X2))
    assert_same(oracle, gpu, synth.config1())
    # 100-frame stacks exceed the 64-frame staging slot: the kernel must fall back per batch
    rng = np.random.Generator(np.random.PCG64(5))
    w = synth.config1().head(4000)
    deep = rng.integers(0, 4096, (4000, 100), dtype=np.uint64)
    w.stack_table, w.stack_choice = None, None
    w._frame_ids = deep.reshape(-1)
    w.hdrs["nframes"] = 100
    w.hdrs["frame_off"] = np.arange(4000, dtype=np.uint64) * np.uint64(100)
    w.hdrs["nframes"][::7] = 63  # mixed depths inside one warp batch (frames stay where they are)
    assert_same(oracle, gpu, w)
    assert_same(oracle, gpu, w, chunk_samples=333)  # tiles that end in the middle of a warp, several launches
    if variant in ("wide", "widepf", "tma", "tmag13x2"):  # wide / widepf also read a narrow (uint32) frame stream (tma hands those to wide)
        assert_same(oracle, gpu, w, frame_id_bytes=4)
        assert_same(oracle, gpu, synth.edge_workload(seed=43, n=5000, hash_mode=abi.PA_HASH_XXH64X2), frame_id_bytes=4, chunk_samples=1000)
        assert_same(oracle, gpu, synth.ragged(n=60_000, u=3_000, p=4_096))


def test_large_batch_properties(oracle, gpu):
    """Size-independent properties on a batch too large for the oracle to be the only check
    (2M samples x 64 frames): unique counts, ListView consistency, ids == XXH64x2 of the frames on a
    sample of rows, run-end monotonicity, dictionary uniqueness, and an oracle comparison on a prefix."""
    w = synth.config2(n=2_000_000, u=50_000, p=65_536)
    a = gpu.from_workload(w, chunk_samples=1 << 18)
    gpu.load(a, w)
    res = a.flush()
    assert res.n_rows == w.n and res.n_unique_stacks == len(np.unique(w.stack_choice))
    t = pa.ipc.open_stream(pa.py_buffer(res.ipc)).read_all()
    assert t.num_rows == w.n
    st = t.column("stacktrace").chunk(0)
    off, size = st.offsets.to_numpy(), st.sizes.to_numpy()
    assert (size == 64).all() and off.min() == 0 and off.max() + 64 == len(st.values) == res.n_location_indices
    # rows with the same stack share (offset, size); first occurrences appear in increasing offset order
    first = {}
    for r in range(0, w.n, 997):
        first.setdefault(int(w.stack_choice[r]), int(off[r]))
        assert first[int(w.stack_choice[r])] == int(off[r])
    uniq_off = off[np.sort(np.unique(w.stack_choice, return_index=True)[1])]
    assert (np.diff(uniq_off) == 64).all()
    ids = t.column("stacktrace_id").chunk(0).storage
    fr = w.stack_table
    for r in range(0, w.n, 100_003):
        data = fr[w.stack_choice[r]].astype("<u8").tobytes()
        want = oracle.xxh64(data, 0).to_bytes(8, "big") + oracle.xxh64(data, abi.PA_XXH_SEED_LO).to_bytes(8, "big")
        assert ids[r].as_py() == want
    labels = t.column("labels").chunk(0)
    for i in range(labels.type.num_fields):
        col = labels.field(i)
        re_ = col.run_ends.to_numpy()
        assert (np.diff(re_) > 0).all() and re_[-1] == w.n
        d = col.values.dictionary.to_pylist()
        assert len(d) == len(set(d))
    assert t.column("timestamp").chunk(0).cast(pa.int64()).to_numpy().tolist()[:3] == w.hdrs["timestamp_ns"][:3].tolist()
    a.close()
    sub = w.head(50_000)
    assert gpu.run(sub)[0] == oracle.run(sub)[0]




def test_table_overflow_retry(gpu):
    """Adaptive table sizing: batch 2 has far more unique stacks and thread ids than batch 1 predicted, so
    the first attempt overflows both hash tables and process() must redo the batch with larger ones."""
    F = 2
    rng = np.random.Generator(np.random.PCG64(9))
    base = synth.config1(hash_mode=abi.PA_HASH_PROVIDED)

    def batch(n, unique):
        hd = np.zeros(n, dtype=abi.HDR_DTYPE)
        sid = np.arange(n, dtype=np.uint64) if unique else rng.integers(0, 10, n).astype(np.uint64)
        hd["hash_hi"] = synth.splitmix64(sid)
        hd["hash_lo"] = synth.splitmix64(sid + np.uint64(12345))
        hd["timestamp_ns"] = np.arange(n)
        hd["tid"] = np.arange(n) if unique else 7
        hd["pid"] = 1
        hd["comm_sid"] = 0
        hd["labelset_id"] = 0
        hd["cpu"] = 1
        hd["nframes"] = F
        hd["frame_off"] = np.arange(n, dtype=np.uint64) * np.uint64(F)
        fr = (np.repeat(sid, F) * np.uint64(7) + np.tile(np.arange(F, dtype=np.uint64), n)) % np.uint64(4096)
        return synth.Workload("ovf", base.strings, base.frames, base.labelsets, hd, _frame_ids=fr, hash_mode=abi.PA_HASH_PROVIDED)

    w1, w2 = batch(1_200_000, False), batch(2_600_000, True)
    a = gpu.from_workload(w2)
    gpu.load(a, w1)
    r1 = a.flush()
    assert r1.n_rows == w1.n and r1.n_unique_stacks == 10
    gpu.load(a, w2)
    r2 = a.flush()
    assert r2.n_rows == w2.n and r2.n_unique_stacks == w2.n and r2.n_location_indices == w2.n * F
    t = pa.ipc.open_stream(pa.py_buffer(r2.ipc)).read_all()
    tid = t.column("labels").chunk(0).field("thread_id")
    assert len(tid.values.dictionary) == w2.n and tid.run_ends.to_numpy()[-1] == w2.n
    assert tid.values.dictionary[12345].as_py() == "12345"  # first-occurrence order == row order here
    st = t.column("stacktrace").chunk(0)
    assert (st.offsets.to_numpy() == np.arange(w2.n) * F).all()
    a.close()


def test_randomized_sweep(oracle, gpu):
    """40 random small batches: random size, hash mode, label flags and external labels."""
    rng = np.random.Generator(np.random.PCG64(2024))
    for i in range(40):
        n = int(rng.integers(1, 2500))
        mode = abi.PA_HASH_PROVIDED if rng.random() < 0.5 else abi.PA_HASH_XXH64X2
        w = synth.edge_workload(seed=1000 + i, n=n, hash_mode=mode, label_flags=int(rng.integers(0, 8)), external=bool(rng.random() < 0.5))
        assert_same(oracle, gpu, w, chunk_samples=int(rng.choice([0, 97, 512])))


def test_registration_concurrent_with_flush(oracle, gpu):
    """ParcaReporter registers strings / frames / labelsets from producer threads while the report ticker flushes
    (reporter.cpp: FlushOnce runs without the ingest lock). The pass must only ever use the registration state it snapshotted
    under the lock: no out-of-range table reads, every flushed batch still byte-exact."""
    import threading
    w = synth.edge_workload(seed=61, n=4000, hash_mode=abi.PA_HASH_XXH64X2, external=False)
    want, _ = oracle.run(w)
    a = gpu.from_workload(w)
    stop, errors = threading.Event(), []

    names = a.register_strings([b"late_label_a", b"late_label_b"])  # a fixed set of label NAMES (columns), ever new values

    def registrar():
        i = 0
        try:
            while not stop.is_set():
                first = a.register_strings([("late-%d-%d" % (i, k)).encode() for k in range(8)])
                fr = np.zeros(4, dtype=abi.FRAME_DTYPE)
                fr["kind"] = abi.PA_FRAME_INTERP
                fr["type_name_sid"] = first
                fr["function_name_sid"] = first + 1
                fr["source_file_sid"] = first + 2
                fr["address_or_lineno"] = np.arange(4) + i
                a.register_frames(fr)
                a.register_labelsets([[(names, first + 4)], [(names, first + 5), (names + 1, first + 7)]])
                i += 1
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    t = threading.Thread(target=registrar)
    t.start()
    try:
        for _ in range(40):
            gpu.load(a, w)
            assert a.flush().ipc_bytes() == want  # the late registrations are not referenced by any row: the record is unchanged
    finally:
        stop.set()
        t.join()
    a.close()
    assert not errors, errors


def test_two_aggregators_concurrently(oracle, gpu):
    """Two instances on one GPU, flushed from two host threads at once (the streaming bench mode)."""
    import threading
    ws = [synth.config1().head(30000), synth.edge_workload(seed=77, n=20000, hash_mode=abi.PA_HASH_XXH64X2)]
    want = [oracle.run(w)[0] for w in ws]
    aggs = [gpu.from_workload(w) for w in ws]
    errors = []

    def worker(i):
        try:
            for _ in range(6):
                gpu.load(aggs[i], ws[i])
                if aggs[i].flush().ipc_bytes() != want[i]:
                    errors.append("instance %d produced different bytes" % i)
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for a in aggs:
        a.close()
    assert not errors, errors


# ---- v1 schema sample record (the reference's default schema) -----------------------------------
def as_v1(w):
    w.schema = abi.PA_SCHEMA_V1
    return w


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("mode", [abi.PA_HASH_PROVIDED, abi.PA_HASH_XXH64X2])
@pytest.mark.parametrize("external", [False, True])
def test_v1_edge_batches(oracle, gpu, seed, mode, external):
    assert_same(oracle, gpu, as_v1(synth.edge_workload(seed=seed, hash_mode=mode, external=external)))


def test_v1_config1_and_flags(oracle, gpu):
    assert_same(oracle, gpu, as_v1(synth.config1()))
    assert_same(oracle, gpu, as_v1(synth.config1(hash_mode=abi.PA_HASH_PROVIDED)), chunk_samples=4096)
    for flags in (1, 2, 4, 7):
        assert_same(oracle, gpu, as_v1(synth.edge_workload(seed=12, label_flags=flags)))
    for name in ("stack_dedup", "writer_basic", "multiple_frame_types"):
        assert_same(oracle, gpu, as_v1(getattr(kw, name)()))


def test_v1_scaled_and_repeated(oracle, gpu):
    w = as_v1(synth.config3(n=200_000, u=20_000, p=16_384, npids=300, lsets=8))
    want, _ = oracle.run(w)
    a = gpu.from_workload(w)
    for _ in range(2):
        gpu.load(a, w)
        assert a.flush().ipc_bytes() == want
    a.close()


def wide_label_workload(n_names, n=1500, seed=5):
    """Many distinct label names (sorted byte-wise incl. non-ASCII), sparse presence, deep stacks with repeated frames."""
    rng = np.random.Generator(np.random.PCG64(seed))
    base = synth.edge_workload(seed=seed, n=n, hash_mode=abi.PA_HASH_XXH64X2, external=False)
    st = synth.StringTable()
    st.strings, st.index = list(base.strings), {s: i for i, s in enumerate(base.strings)}
    names = ["lbl_%02d" % i for i in range(n_names - 3)] + ["zzé中", "Aardvark", "a"]
    name_sids = [st.sid(x) for x in names]
    vals = [st.sid("v%d" % i) for i in range(17)] + [st.sid("x" * 70)]
    labelsets = []
    for _ in range(24):
        chosen = sorted(rng.choice(len(names), size=int(rng.integers(0, len(names) + 1)), replace=False), key=lambda i: st.strings[name_sids[i]])
        labelsets.append([(name_sids[i], vals[int(rng.integers(0, len(vals)))]) for i in chosen])
    hd = base.hdrs.copy()
    hd["labelset_id"] = rng.integers(0, len(labelsets), n)
    # deep stacks (up to 200 frames, repeated frame ids inside one stack)
    P = len(base.frames)
    stream, off = [], 0
    for i in range(n):
        k = int(rng.choice([0, 1, 3, 64, 65, 129, 200]))
        fr = rng.integers(0, min(P, 7), k).astype(np.uint64)
        hd["nframes"][i], hd["frame_off"][i] = k, off
        stream.append(fr)
        off += k
    w = synth.Workload("wide_labels_%d" % n_names, st.strings, base.frames, labelsets, hd, _frame_ids=np.concatenate(stream),
                       hash_mode=abi.PA_HASH_XXH64X2, samples_per_second=97)
    return w


@pytest.mark.parametrize("n_names", [5, 35])
def test_many_label_columns_and_deep_stacks(oracle, gpu, n_names):
    w = wide_label_workload(n_names)
    assert_same(oracle, gpu, w)
    w.schema = abi.PA_SCHEMA_V1
    assert_same(oracle, gpu, w)


def test_many_label_names(oracle, gpu):
    """round 1 capped a batch at 38 distinct label names; k8s pod labels go well beyond that"""
    assert_same(oracle, gpu, wide_label_workload(45, n=400))
    assert_same(oracle, gpu, wide_label_workload(200, n=300, seed=9))


def test_too_many_label_names_is_an_error(gpu):
    w = wide_label_workload(300, n=50)
    a = gpu.from_workload(w)
    gpu.load(a, w)
    with pytest.raises(gpu.PaError) as e:
        a.flush()
    assert e.value.code == -34  # PA_ERANGE (kernel-parameter tables hold 246 label columns)
    a.close()


def test_ring_full_and_invalid_inputs(gpu):
    w = synth.edge_workload(seed=2, n=100)
    a = gpu.from_workload(w, max_samples=100)
    gpu.load(a, w)
    with pytest.raises(gpu.PaError) as e:
        a.acquire(1, 0)
    assert e.value.code == -28  # PA_ENOSPC: the caller must flush
    assert a.flush().n_rows == 100
    bad = synth.edge_workload(seed=2, n=100)
    bad.hdrs["cpu"][5] = 70000
    gpu.load(a, bad)
    with pytest.raises(gpu.PaError):
        a.flush()
    bad = synth.edge_workload(seed=2, n=100)
    bad.hdrs["kind"][7] = 9
    gpu.load(a, bad)
    with pytest.raises(gpu.PaError):
        a.flush()
    gpu.load(a, w)  # the aggregator stays usable after rejected batches
    assert a.flush().n_rows == 100
    a.close()


def test_adaptive_tables_across_flushes(oracle, gpu):
    """Batches above 2^20 rows switch the stack / thread-id tables to capacities predicted from the previous
    interval; flush the same 1.2M-row batch three times (worst-case, then adaptive sizes) and compare each
    result with the oracle, in both schemas."""
    for schema in (abi.PA_SCHEMA_V2, abi.PA_SCHEMA_V1):
        w = synth.config2(n=1_200_000, u=30_000, p=32_768)
        w.schema = schema
        want, _ = oracle.run(w)
        a = gpu.from_workload(w, chunk_samples=1 << 18)
        for i in range(3):
            gpu.load(a, w)
            assert a.flush().ipc_bytes() == want, "flush %d differs (schema %d)" % (i, schema)
        a.close()


def test_ragged_depths(oracle, gpu):
    """Every stack has its own depth (1..127 frames): remainder tiers of the hash kernel, odd offsets, all variants."""
    w = synth.ragged(n=150_000, u=4_000, p=8_192)
    assert_same(oracle, gpu, w)


# ---- v1 stacktrace record (buildStacktraceRecord + the device-resident store of known stacks) ----
def missing_ids(k):
    return [bytes([0xEE, i]) * 8 for i in range(k)]


def assert_same_stacktraces(o, a, ids):
    """o: oracle.Oracle, a: Aggregator, both fed the same intervals; ids: list of 16-byte ids."""
    want, nloc = o.stacktraces(b"".join(ids))
    r = a.stacktraces(b"".join(ids))
    got = r.ipc_bytes()
    if got != want:
        d = pyref.diff(pyref.extract_stacktraces(want), pyref.extract_stacktraces(got))
        raise AssertionError("stacktrace record bytes differ (len %d vs %d); first logical difference: %s" % (len(want), len(got), d))
    assert (r.n_rows, r.n_locations) == (len(ids), nloc)
    return r


def feed_both(oracle, gpu, w, o=None, a=None, **kw_agg):
    """One v1 interval through both sides; returns (o, a, this interval's unique ids in first-occurrence order)."""
    as_v1(w)
    o = o or oracle.Oracle(w, stack_cache_entries=kw_agg.get("stack_cache_entries", 0))
    a = a or gpu.from_workload(w, **kw_agg)
    o.ingest(w.hdrs, w.frame_ids)
    want, st = o.flush()
    gpu.load(a, w)
    r = a.flush()
    assert r.ipc_bytes() == want
    ids = a.last_stack_ids(r.n_unique_stacks)
    dict_ids = pa.ipc.open_stream(want).read_all().column("stacktrace_id").chunk(0).values.dictionary.to_pylist() if want else []
    assert [bytes(x) for x in ids] == dict_ids  # == the ids the reference walks at parca_reporter.go:1307-1328
    return o, a, dict_ids


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("mode", [abi.PA_HASH_PROVIDED, abi.PA_HASH_XXH64X2])
def test_stacktrace_record_edge(oracle, gpu, seed, mode):
    o, a, ids = feed_both(oracle, gpu, synth.edge_workload(seed=seed, hash_mode=mode))
    assert_same_stacktraces(o, a, ids)                                   # offline mode: every new stack of the interval
    mixed = list(ids)
    for i, m in enumerate(missing_ids(4)):
        mixed.insert((5 * i) % (len(mixed) + 1), m)
    assert_same_stacktraces(o, a, mixed + ids[:3])                       # unknown ids and repeats, as a server may ask
    assert_same_stacktraces(o, a, list(reversed(ids)))
    a.close(); o.close()


def test_stacktrace_record_degenerate(oracle, gpu):
    w = synth.edge_workload(seed=5)
    o, a, ids = feed_both(oracle, gpu, w)
    assert_same_stacktraces(o, a, [])                                    # an empty record is still written (:1332)
    assert_same_stacktraces(o, a, missing_ids(1))
    assert_same_stacktraces(o, a, missing_ids(70))
    known = pyref.known_stacks(w)
    empties = [k for k, v in known.items() if len(v) == 0]
    assert empties
    assert_same_stacktraces(o, a, empties)                               # zero locations, null list entries only
    assert_same_stacktraces(o, a, empties + ids[:1] + empties)
    a.close(); o.close()


def test_stacktrace_record_configs(oracle, gpu):
    for w in (synth.config1(), synth.config1(hash_mode=abi.PA_HASH_PROVIDED).head(20_000), synth.ragged(n=30_000, u=2_000, p=4_096)):
        o, a, ids = feed_both(oracle, gpu, w)
        r = assert_same_stacktraces(o, a, ids)
        assert r.n_locations > len(ids) and r.gpu_launches > 0
        assert_same_stacktraces(o, a, ids[::7] + missing_ids(2))
        a.close(); o.close()


def test_stacktrace_store_persists_across_intervals(oracle, gpu):
    """The store is the `stacks` LRU: first occurrence wins (also across intervals), later intervals only add."""
    w = synth.edge_workload(seed=41, n=3000)
    o, a, ids1 = feed_both(oracle, gpu, w.head(1200), max_samples=4000, max_frames=40000)
    _, _, ids2 = feed_both(oracle, gpu, w.rows(np.arange(1200, 3000)), o=o, a=a)
    assert set(ids1) & set(ids2)
    both = ids1 + [i for i in ids2 if i not in set(ids1)]
    assert_same_stacktraces(o, a, both + missing_ids(2))
    _, _, ids3 = feed_both(oracle, gpu, w.head(0), o=o, a=a)               # an empty interval changes nothing
    assert ids3 == []
    assert_same_stacktraces(o, a, both)
    a.close(); o.close()


def test_stacktrace_store_is_an_lru(oracle, gpu):
    """`stacks` is lru.SyncedLRU (parca_reporter.go:106): `Get` per sample and per requested id moves an entry to the front, `Add`
    evicts the entry at the back. The store must answer every request like the oracle's LRU of the same capacity: evicted
    stacks come back as "missing stacktrace" rows (:1556-1573), stacks refreshed by a later sample or by a request survive,
    an evicted stack that is sampled again is known again, a batch with more distinct stacks than the cache keeps its last C."""
    big = synth.config2(n=12_000, u=10_000, p=8_192)                      # 64-frame stacks, ~3.6k distinct per 4.5k rows
    w1, w2 = big.head(4500), big.rows(np.arange(6000, 10_500))
    o, a, ids1 = feed_both(oracle, gpu, w1, max_samples=8000, max_frames=8000 * 64, stack_cache_entries=4096)
    assert 2048 < len(ids1) <= 4096
    assert_same_stacktraces(o, a, ids1[:300])                             # these 300 are now the most recently used
    _, _, ids2 = feed_both(oracle, gpu, w2, o=o, a=a)
    assert len(set(ids1) | set(ids2)) > 4096                              # something had to go
    assert_same_stacktraces(o, a, ids1 + ids2 + missing_ids(1))
    a.close(); o.close()
    # many small intervals against a tiny cache: evictions every interval, revivals, compactions of the slot table and the arena
    w = synth.config2(n=9000, u=2500, p=4096)
    o = a = None
    seen = []
    rng = np.random.Generator(np.random.PCG64(9))
    for k in range(18):  # 2500 distinct stacks against 1536 usable slots: the table is compacted along the way
        part = w.rows(np.sort(rng.choice(w.n, 900 if k == 7 else 220, replace=False)))  # interval 7 alone holds more stacks than the cache
        o, a, ids = feed_both(oracle, gpu, part, o=o, a=a, max_samples=4000, max_frames=400_000, stack_cache_entries=300)
        seen += [i for i in ids if i not in set(seen)]
        probe = [seen[int(j)] for j in rng.choice(len(seen), min(len(seen), 120), replace=False)]
        assert_same_stacktraces(o, a, probe + missing_ids(1) + probe[:5])  # the request itself refreshes what it finds
    assert_same_stacktraces(o, a, seen)
    assert a.kernel_ms("store_compactions")[1] >= 1 and a.kernel_ms("store_evictions")[1] >= 10
    a.close(); o.close()
    # one interval with more distinct stacks than the cache holds: the last C by last occurrence stay
    o, a, ids = feed_both(oracle, gpu, w1, max_samples=8000, max_frames=8000 * 64, stack_cache_entries=1000)
    assert len(ids) > 1000
    assert_same_stacktraces(o, a, ids)
    _, _, ids_b = feed_both(oracle, gpu, w2.head(300), o=o, a=a)
    assert_same_stacktraces(o, a, ids + ids_b)
    a.close(); o.close()


def test_stacktrace_store_out_of_frame_space_starts_over(oracle, gpu):
    """The frame arena (pa_agg_config.stack_cache_frames) is this library's own limit. Out of it: evicted entries are compacted
    away first; if the live ones alone do not fit, the store is cleared and refilled from the current batch, so it then
    behaves like a reporter that has only seen that batch — older stacks come back as "missing stacktrace" rows."""
    big = synth.config2(n=12_000, u=10_000, p=8_192)
    w1, w2 = big.head(4500), big.rows(np.arange(6000, 10_500))
    o, a, ids1 = feed_both(oracle, gpu, w1, max_samples=8000, max_frames=8000 * 64, stack_cache_entries=1 << 16, stack_cache_frames=4096 * 64)
    assert 2048 < len(ids1) <= 4096
    assert_same_stacktraces(o, a, ids1)
    _, _, ids2 = feed_both(oracle, gpu, w2, o=o, a=a)
    assert len(set(ids1) | set(ids2)) > 4096
    o2 = oracle.Oracle(as_v1(w2))                                         # a reporter that only ever saw the second batch
    o2.ingest(w2.hdrs, w2.frame_ids)
    o2.flush()
    assert_same_stacktraces(o2, a, ids2 + ids1[:50] + missing_ids(1))
    a.close(); o.close(); o2.close()


def test_stacktrace_custom_unknown_type_and_errors(oracle, gpu):
    w = as_v1(synth.edge_workload(seed=8))
    sid = len(w.strings)
    w.strings = list(w.strings) + [b"unknown-frame-type-from-libpf"]
    w.unknown_frame_type_sid = sid
    o, a, ids = feed_both(oracle, gpu, w, unknown_frame_type_sid=sid)
    assert_same_stacktraces(o, a, missing_ids(2) + ids[:5])
    gpu.load(a, w)
    a.stage()
    with pytest.raises(gpu.PaError) as e:                                  # the staged batch still owns the scratch arena
        a.stacktraces(b"".join(ids[:1]))
    assert e.value.code == -22
    a.process(); a.collect()
    a.close(); o.close()
    w2 = synth.edge_workload(seed=8)                                      # v2 aggregators keep no store
    a2 = gpu.from_workload(w2)
    with pytest.raises(gpu.PaError) as e:
        a2.stacktraces(b"\x00" * 16)
    assert e.value.code == -22
    a2.close()


# ---- multi-GPU mode B: one merged batch from pid-hash shards (here: several shard aggregators on one GPU) ----------
def merged_from_shards(gpu, w, world, schema):
    from parca_agent_b200 import sharded
    rows = sharded.shard_rows(w, world)
    shards = []
    for idx in rows:
        part = w.rows(idx)
        part.schema = abi.PA_SCHEMA_V2
        a = gpu.from_workload(part)
        gpu.load(a, part)
        a.stage()
        a.process()
        shards.append((a, idx))
    m = gpu.Aggregator(hash_mode=abi.PA_HASH_PROVIDED, label_flags=w.label_flags, samples_per_second=w.samples_per_second,
                       external_labels=w.external_labels, max_samples=max(w.n, 1), max_frames=max(w.n_frame_ids, 1), schema=schema)
    assert m.register_strings(w.strings[1:]) == 1
    m.register_frames(w.frames)
    m.register_labelsets(w.labelsets)
    res = sharded.merge_local(shards, m)
    data = res.ipc_bytes()
    for a, _ in shards:
        a.close()
    return data, res, m, [len(i) for i in rows]


@pytest.mark.parametrize("schema", [abi.PA_SCHEMA_V2, abi.PA_SCHEMA_V1])
@pytest.mark.parametrize("world", [2, 3])
def test_mode_b_merged_batch_equals_unsharded_oracle(oracle, gpu, schema, world):
    """Shards hash and deduplicate on their own; the merged batch must be byte-for-byte what the reference path gives on
    the unsharded stream (global first-occurrence order of every dictionary, run ends across shard boundaries, stacks
    seen by several shards, provided-id collisions whose first occurrence sits on another shard)."""
    cases = [synth.edge_workload(seed=21, n=4000, hash_mode=abi.PA_HASH_PROVIDED), synth.edge_workload(seed=22, n=4000, hash_mode=abi.PA_HASH_XXH64X2),
             synth.config3(n=60_000, u=4_000, p=4_096, npids=64, lsets=6), synth.ragged(n=20_000, u=1_500, p=2_048)]
    for w in cases:
        w.schema = schema
        want, st = oracle.run(w)
        got, res, m, sizes = merged_from_shards(gpu, w, world, schema)
        assert min(sizes) > 0 and sum(sizes) == w.n
        if got != want:
            ex = pyref.extract_v1 if schema == abi.PA_SCHEMA_V1 else pyref.extract
            d = pyref.diff(ex(pa.ipc.open_stream(want).read_all()), ex(pa.ipc.open_stream(got).read_all()))
            raise AssertionError("%s world=%d: merged IPC differs (len %d vs %d): %s" % (w.name, world, len(want), len(got), d))
        assert res.n_rows == w.n and res.n_unique_stacks == st["unique_stacks"]
        m.close()


def test_mode_b_device_staging_rules(gpu):
    w = synth.edge_workload(seed=3, n=300)
    a = gpu.from_workload(w)
    with pytest.raises(gpu.PaError):
        a.shard_sizes()                                                   # nothing processed yet
    gpu.load(a, w)
    a.stage(); a.process()
    n, nf = a.shard_sizes()
    assert n == w.n and 0 < nf <= w.n_frame_ids
    with pytest.raises(gpu.PaError) as e:
        a.stage_device(0, 0, 0, 0)                                        # a staged batch is still pending
    assert e.value.code == -22
    a.collect()
    a.stage_device(0, 0, 0, 0)                                            # an empty device batch behaves like an empty interval
    a.process()
    assert a.collect().n_rows == 0
    a.close()
    v1 = as_v1(synth.edge_workload(seed=3, n=300))
    b = gpu.from_workload(v1)
    gpu.load(b, v1)
    b.stage(); b.process()
    with pytest.raises(gpu.PaError) as e:
        b.shard_export(0, 1, 1)                                           # v1 aggregators do not gather unique stacks
    assert e.value.code == -22
    b.collect(); b.close()


def test_concurrent_producers_and_flush_on_one_aggregator(gpu):
    """reporter.Reporter's threading contract: ReportTraceEvent is called from many goroutines while the report
    ticker flushes; the writer is swapped under the ingest lock (parca_reporter.go:1745-1748). Four producer threads submit
    while a fifth flushes: every row must come out exactly once, batches must decode, and each producer's rows must keep
    their order inside every batch (row order == lock acquisition order)."""
    import threading
    import time
    w = synth.edge_workload(seed=77, n=24_000, hash_mode=abi.PA_HASH_XXH64X2)
    w.hdrs["timestamp_ns"] = np.arange(w.n)                       # unique row tag
    a = gpu.from_workload(w, max_samples=3_000, max_frames=40_000)    # small ring: producers hit PA_ENOSPC and retry
    nprod, stop, batches, errors = 4, threading.Event(), [], []
    fo = w.hdrs["frame_off"].astype(np.int64)
    nf = w.hdrs["nframes"].astype(np.int64)
    frames = w.frame_ids

    def producer(p):
        try:
            rows = np.arange(p, w.n, nprod)
            for s in range(0, len(rows), 50):
                part = rows[s:s + 50]
                hd = w.hdrs[part].copy()
                fr = np.concatenate([frames[fo[r]:fo[r] + nf[r]] for r in part]) if len(part) else np.zeros(0, np.uint64)
                while True:
                    try:
                        a.submit(hd, fr)
                        break
                    except gpu.PaError as e:
                        if e.code != -28:
                            raise
                        time.sleep(0.001)                         # ring full: wait for the flusher
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    def flusher():
        try:
            while not stop.is_set():
                r = a.flush()
                if r.n_rows:
                    batches.append(r.ipc_bytes())
                time.sleep(0.002)
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    ts = [threading.Thread(target=producer, args=(p,)) for p in range(nprod)]
    fl = threading.Thread(target=flusher)
    fl.start()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    stop.set()
    fl.join()
    r = a.flush()
    if r.n_rows:
        batches.append(r.ipc_bytes())
    a.close()
    assert not errors, errors
    assert len(batches) > 3
    seen = []
    for b in batches:
        t = pa.ipc.open_stream(b).read_all()
        tags = t.column("timestamp").cast(pa.int64()).to_numpy()  # nanoseconds since epoch == the row tag
        for p in range(nprod):                                    # per-producer order inside the batch
            mine = tags[tags % nprod == p]
            assert np.all(np.diff(mine) > 0)
        seen.append(tags)
    allt = np.concatenate(seen)
    assert len(allt) == w.n and np.array_equal(np.sort(allt), np.arange(w.n))
