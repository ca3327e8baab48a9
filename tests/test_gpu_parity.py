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
            xa = pyref.extract(pa.ipc.open_stream(want).read_all())
            xb = pyref.extract(pa.ipc.open_stream(got).read_all())
            d = pyref.diff(xa, xb)
        raise AssertionError("IPC bytes differ (len %d vs %d); first logical difference: %s" % (len(want), len(got), d))
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
