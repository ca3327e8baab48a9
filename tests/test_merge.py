"""Mode B with an O(unique keys) exchange (pa_merge_*): several shard aggregators build ONE record batch that must equal,
byte for byte, the oracle's record for the stream [shard 0's rows, shard 1's rows, ...].

The in-process group (all shards on one device) runs the same kernels and the same host code as the NCCL group; only the
transport differs, so the driver's 1-GPU tier exercises the whole merge. The NCCL transport itself is covered by
tests/dist_merge_slices_check.py (torchrun, one GPU per rank; run here when the box has >= 2 GPUs)."""
import os
import subprocess
import sys

import numpy as np
import pyarrow as pa
import pytest

import pyref
from parca_agent_b200 import abi, sharded, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gpu():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from parca_agent_b200 import lib
    lib.lib()
    return lib


def merged_vs_oracle(oracle, gpu, w, idx_lists, staged=False, repeat=1, chunk_samples=0, **kw):
    """Shards = the given row subsets of w (in that order); returns the merged Result after comparing with the oracle."""
    parts = [w.rows(ix) for ix in idx_lists]
    order = np.concatenate([np.asarray(ix, dtype=np.int64) for ix in idx_lists]) if idx_lists else np.zeros(0, np.int64)
    ref = w.rows(order)
    want, st = oracle.run(ref)
    aggs = [gpu.from_workload(p, chunk_samples=chunk_samples, **kw) for p in parts]
    g = gpu.MergeGroup.local(aggs)
    res = None
    for _ in range(repeat):
        for a, p in zip(aggs, parts):
            gpu.load(a, p)
        if staged:
            for a in aggs:
                a.stage()
            g.process()
            n = g.plan()
            res = g.collect()
            assert res.ipc_len == n
        else:
            res = g.flush()
        got = res.ipc_bytes()
        res.ipc = np.frombuffer(got, dtype=np.uint8)  # the library buffer goes away with the group: keep a copy
        if got != want:
            d = None
            if want and got:
                d = pyref.diff(pyref.extract(pa.ipc.open_stream(want).read_all()), pyref.extract(pa.ipc.open_stream(got).read_all()))
            raise AssertionError("merged IPC bytes differ (len %d vs %d); first logical difference: %s" % (len(want), len(got), d))
        assert res.n_rows == st["rows"]
        if st["rows"]:
            assert (res.n_unique_stacks, res.n_locations, res.n_functions, res.n_location_indices) == (
                st["unique_stacks"], st["locations"], st["functions"], st["location_indices"])
    g.close()
    for a in aggs:
        a.close()
    return res


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_config1_pid_sharded(oracle, gpu, world):
    w = synth.config1().head(40_000)
    merged_vs_oracle(oracle, gpu, w, sharded.shard_rows(w, world))


@pytest.mark.parametrize("mode", [abi.PA_HASH_PROVIDED, abi.PA_HASH_XXH64X2])
@pytest.mark.parametrize("seed", [3, 4])
def test_edge_batches(oracle, gpu, mode, seed):
    """ragged stacks, every sample kind / frame kind, null label runs (validity bits that meet inside a byte at shard borders),
    provided ids that collide (first occurrence across ALL shards wins), the all-zero id"""
    w = synth.edge_workload(seed=seed, n=3000, hash_mode=mode, external=False)
    for world in (2, 5):
        merged_vs_oracle(oracle, gpu, w, sharded.shard_rows(w, world))


def test_external_label_new_column(oracle, gpu):
    w = synth.edge_workload(seed=8, n=1500, external=False)
    st = {s: i for i, s in enumerate(w.strings)}
    w.strings = list(w.strings) + [b"cluster", b"prod"]
    w.external_labels = [(len(w.strings) - 2, len(w.strings) - 1)]
    assert b"cluster" not in st
    merged_vs_oracle(oracle, gpu, w, sharded.shard_rows(w, 3))


def test_runs_merge_across_shard_borders(oracle, gpu):
    """contiguous row blocks as shards: cpu / thread / labelset runs that straddle a border must come out as ONE run, tiny
    shards (1-3 rows, whole shard inside one run) and empty shards included"""
    w = synth.edge_workload(seed=12, n=2000, external=False)
    w.hdrs["cpu"][:] = np.repeat(np.arange(2000 // 50), 50)[:2000] % 3   # long constant stretches
    w.hdrs["tid"][:] = 7
    cuts = [0, 100, 101, 103, 103, 650, 651, 1999, 2000]
    idx = [np.arange(cuts[i], cuts[i + 1]) for i in range(len(cuts) - 1)]
    res = merged_vs_oracle(oracle, gpu, w, idx)
    t = pa.ipc.open_stream(pa.py_buffer(res.ipc)).read_all()
    tid = t.column("labels").chunk(0).field("thread_id")
    assert len(tid.run_ends) == 1 and tid.run_ends[0].as_py() == 2000


def test_narrow_ring_members(oracle, gpu):
    w = synth.edge_workload(seed=14, n=4000, hash_mode=abi.PA_HASH_XXH64X2, external=False)
    merged_vs_oracle(oracle, gpu, w, sharded.shard_rows(w, 3), frame_id_bytes=4)


def test_config3_zipf_many_labelsets(oracle, gpu):
    w = synth.config3(n=150_000, u=12_000, p=8_192, npids=300, lsets=8)
    merged_vs_oracle(oracle, gpu, w, sharded.shard_rows(w, 4), chunk_samples=20_000)


def test_config2_scaled_staged_and_repeated(oracle, gpu):
    """stage / process / plan / collect; two intervals back to back (adaptive table sizes carry over)"""
    w = synth.config2(n=300_000, u=15_000, p=32_768)
    merged_vs_oracle(oracle, gpu, w, sharded.shard_rows(w, 4), staged=True, repeat=2)


def test_empty_and_all_empty(oracle, gpu):
    w = synth.config1().head(5000)
    idx = sharded.shard_rows(w, 2)
    merged_vs_oracle(oracle, gpu, w, [idx[0], np.zeros(0, np.int64), idx[1]])
    parts = [w.rows(np.zeros(0, np.int64)) for _ in range(2)]
    aggs = [gpu.from_workload(p, max_samples=16, max_frames=64) for p in parts]
    g = gpu.MergeGroup.local(aggs)
    r = g.flush()
    assert r.n_rows == 0 and r.ipc_len == 0
    g.close()
    for a in aggs:
        a.close()


def test_mismatched_registrations_are_rejected(gpu):
    w1, w2 = synth.config1().head(100), synth.edge_workload(seed=1, n=100, external=False)
    aggs = [gpu.from_workload(w1), gpu.from_workload(w2)]
    g = gpu.MergeGroup.local(aggs)
    for a, w in zip(aggs, (w1, w2)):
        gpu.load(a, w)
    with pytest.raises(gpu.PaError):
        g.flush()
    g.close()
    for a in aggs:
        a.close()


def test_single_aggregator_still_works_after_merge(oracle, gpu):
    """a member can go back to per-shard batches (mode A) between merged intervals"""
    w = synth.config1().head(20_000)
    idx = sharded.shard_rows(w, 2)
    parts = [w.rows(ix) for ix in idx]
    aggs = [gpu.from_workload(p) for p in parts]
    g = gpu.MergeGroup.local(aggs)
    for a, p in zip(aggs, parts):
        gpu.load(a, p)
    assert g.flush().ipc_bytes() == oracle.run(w.rows(np.concatenate(idx)))[0]
    for a, p in zip(aggs, parts):
        gpu.load(a, p)
        assert a.flush().ipc_bytes() == oracle.run(p)[0]
    g.close()
    for a in aggs:
        a.close()


@pytest.mark.parametrize("world,transport,mailbox", [(2, "gloo", 0), (3, "gloo", 0), (2, "shm", 0), (3, "shm", 0), (3, "shm", 8192)])
def test_multi_process_group_over_host_transport(world, transport, mailbox):
    """one shard per PROCESS, all processes on this one GPU, the stream in shared memory: the exchange over gloo callbacks on host
    buffers (pa_merge_create_host) or through page-locked shared-memory mailboxes (pa_merge_create_shm; an 8 KiB mailbox makes
    every collective take many rounds)"""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", PA_ONE_GPU="1", PA_MERGE_TRANSPORT=transport, PA_SHM_MAILBOX=str(mailbox))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", str(29533 + world + (10 if transport == "shm" else 0) + (5 if mailbox else 0)), os.path.join(ROOT, "tests", "dist_merge_hostcb_check.py")], capture_output=True, text=True, env=env, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "merge-hostcb ok world=%d transport=%s" % (world, transport) in p.stdout


def test_nccl_group_on_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (tests/dist_merge_slices_check.py under torchrun)")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29531", os.path.join(ROOT, "tests", "dist_merge_slices_check.py")], capture_output=True, text=True, env=env, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "merge-slices ok world=2" in p.stdout
