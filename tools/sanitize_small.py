"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck are 10-100x slower than a normal run): the kernels and host paths
added in round 2 — early copy-out during a multi-chunk flush, the v1 LRU store (evictions, revivals, compaction, over-full batch),
the pipelined / bulk-copy-staged hash variants, the narrow ring, a 3-shard merged record.

usage: PA_EARLY_D2H_ALWAYS=1 compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle_py  # noqa: E402
from parca_agent_b200 import abi, lib, sharded, synth  # noqa: E402


def main():
    oracle_py.build()
    # early copy-out: several intervals through one aggregator, 7 chunks each
    base = synth.edge_workload(seed=3, n=3000, hash_mode=abi.PA_HASH_XXH64X2, external=False)
    for kw in ({}, {"frame_id_bytes": 4}):
        a = lib.from_workload(base, chunk_samples=450, **kw)
        for n in (3000, 3000, 1200, 2999):
            w = base.head(n)
            lib.load(a, w)
            assert a.flush().ipc_bytes() == oracle_py.run(w)[0]
        a.close()
    # hash kernel variants (each reads PA_HASH_VARIANT at create)
    rag = synth.ragged(n=3000, u=300, p=1024)
    for v in ("wide", "widepf", "tma", "tma24x2", "tma13x2r", "tmag13x2"):
        os.environ["PA_HASH_VARIANT"] = v
        for w in (base, rag):
            got, _ = lib.run(w, chunk_samples=700)
            assert got == oracle_py.run(w)[0], v
    os.environ.pop("PA_HASH_VARIANT")
    # v1 LRU store
    big = synth.config2(n=3000, u=800, p=1024)
    big.schema = abi.PA_SCHEMA_V1
    o = oracle_py.Oracle(big, stack_cache_entries=100)
    a = lib.from_workload(big, max_samples=3000, max_frames=3000 * 64, stack_cache_entries=100)
    rng = np.random.Generator(np.random.PCG64(1))
    seen = []
    for k in range(14):
        part = big.rows(np.sort(rng.choice(big.n, 300 if k == 5 else 70, replace=False)))
        part.schema = abi.PA_SCHEMA_V1
        o.ingest(part.hdrs, part.frame_ids)
        want, _ = o.flush()
        lib.load(a, part)
        r = a.flush()
        assert r.ipc_bytes() == want
        ids = [bytes(x) for x in a.last_stack_ids(r.n_unique_stacks)]
        seen += [i for i in ids if i not in set(seen)]
        probe = b"".join(seen[:: max(1, len(seen) // 60)])
        assert a.stacktraces(probe).ipc_bytes() == o.stacktraces(probe)[0]
    assert a.kernel_ms("store_compactions")[1] >= 1
    a.close(); o.close()
    # merged record from 3 shards on one device
    w = synth.edge_workload(seed=4, n=2500, hash_mode=abi.PA_HASH_XXH64X2, external=False)
    idx = sharded.shard_rows(w, 3)
    parts = [w.rows(ix) for ix in idx]
    aggs = [lib.from_workload(p) for p in parts]
    g = lib.MergeGroup.local(aggs)
    for x, p in zip(aggs, parts):
        lib.load(x, p)
    assert g.flush().ipc_bytes() == oracle_py.run(w.rows(np.concatenate(idx)))[0]
    g.close()
    for x in aggs:
        x.close()
    print("sanitize_small ok")


if __name__ == "__main__":
    main()
