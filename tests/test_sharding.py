"""Multi-rank path (mode A): pid-hash sharding, per-rank batches, max-over-ranks timing.
CPU: gloo, world size 2. GPU: each shard through the CUDA path equals the oracle on that sub-stream."""
import os
import subprocess
import sys

import numpy as np
import pytest

from parca_agent_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_function_matches_xxh64(oracle):
    for pid in [0, 1, 1000, 65535, 4242424, 2**32 - 1]:
        assert synth.xxh64_u32(pid) == oracle.xxh64(int(pid).to_bytes(4, "little"), 0)


def test_shards_partition_the_pid_space():
    owners = {}
    for rank in range(4):
        w = synth._pid_shard("t", 1, rank, 4, 500, 50, 128, 16, synth.abi.PA_HASH_PROVIDED)
        for p in np.unique(w.hdrs["pid"]):
            assert owners.setdefault(int(p), rank) == rank
            assert synth.xxh64_u32(int(p)) % 4 == rank


def test_world_size_2_gloo():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29513", os.path.join(ROOT, "tests", "dist_shard_check.py")], capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "shard-check ok world=2" in p.stdout


def test_host_transport_collectives_world_3_gloo():
    """the host collectives a multi-process merge group calls back into (pa_merge_create_host), against their definitions"""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "3", "--master-addr", "127.0.0.1",
                        "--master-port", "29519", os.path.join(ROOT, "tests", "dist_hostcb_transport_check.py")], capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "hostcb-transport ok world=3" in p.stdout


@pytest.mark.gpu
def test_each_shard_matches_oracle_on_gpu(oracle):
    from parca_agent_b200 import lib
    for rank in range(2):
        w = synth._pid_shard("t", 0x5EED0002, rank, 2, 50_000, 2_000, 4_096, 64, synth.abi.PA_HASH_XXH64X2)
        assert lib.run(w)[0] == oracle.run(w)[0]


@pytest.mark.gpu
def test_mode_b_distributed_merge_two_ranks_one_gpu():
    """sharded.merge_distributed with two ranks sharing the GPU (gloo, payload staged through host memory): the same
    orchestration the NCCL path uses, runnable on a 1-GPU box. `PA_DIST_BACKEND=nccl` + one GPU per rank is the real thing."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", PA_DIST_BACKEND="gloo")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.join(ROOT, "tests", "dist_merge_check.py")], capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "merge-check ok world=2 backend=gloo" in p.stdout


def test_shard_rows_partition_and_keep_order():
    """sharded.shard_rows: every row lands on exactly one shard, a pid never straddles shards, row order inside a shard is
    the global order (what makes the merged batch's first-occurrence ranks come out right)."""
    from parca_agent_b200 import sharded
    w = synth.config3(n=20_000, u=500, p=1_024, npids=200, lsets=3)
    for world in (1, 2, 3, 8):
        rows = sharded.shard_rows(w, world)
        assert len(rows) == world
        allr = np.concatenate(rows)
        assert len(allr) == w.n and np.array_equal(np.sort(allr), np.arange(w.n))
        for r, idx in enumerate(rows):
            assert np.all(np.diff(idx) > 0)
            assert all(synth.xxh64_u32(int(p)) % world == r for p in np.unique(w.hdrs["pid"][idx]))
