"""One-off differential stress run on a GPU box (not collected by pytest): many random small batches through every mode —
v2 and v1 sample records, the v1 stacktrace record with shuffled / unknown / repeated ids, multi-interval stores with tiny LRU
capacities, random chunking (PA_EARLY_D2H_ALWAYS=1 sends every multi-chunk flush after the first through the early copy-out).

usage: python tests/stress_gpu.py [--cases 300] [--seed 1]   (prints a one-line summary; exits non-zero on the first mismatch)
"""
import argparse
import os
import sys

import numpy as np
import pyarrow as pa

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle_py  # noqa: E402
from parca_agent_b200 import abi, lib, synth  # noqa: E402


def one_case(rng, i):
    n = int(rng.integers(1, 4000))
    mode = abi.PA_HASH_PROVIDED if rng.random() < 0.5 else abi.PA_HASH_XXH64X2
    w = synth.edge_workload(seed=50_000 + i, n=n, hash_mode=mode, label_flags=int(rng.integers(0, 8)), external=bool(rng.random() < 0.5))
    v1 = rng.random() < 0.6
    w.schema = abi.PA_SCHEMA_V1 if v1 else abi.PA_SCHEMA_V2
    chunk = int(rng.choice([0, 61, 97, 512, 4096]))
    cuts = sorted(set([0, n] + [int(x) for x in rng.integers(0, n + 1, int(rng.integers(0, 3)))]))
    # small `stacks` LRU capacities: evictions, revivals, over-full intervals. Not with provided ids: edge workloads make different
    # stacks share an id on purpose, and which of them an entry holds after an eviction inside an interval is the one documented
    # deviation of the store
    cap = int(rng.choice([0, 0, 5, 12, 25, 40])) if v1 and mode == abi.PA_HASH_XXH64X2 else 0
    o = oracle_py.Oracle(w, stack_cache_entries=cap)
    a = lib.from_workload(w, chunk_samples=chunk, stack_cache_entries=cap)
    seen = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        part = w.rows(np.arange(lo, hi))
        part.schema = w.schema
        o.ingest(part.hdrs, part.frame_ids)
        want, _ = o.flush()
        lib.load(a, part)
        r = a.flush()
        if r.ipc_bytes() != want:
            return "sample record differs (case %d rows %d..%d v1=%s mode=%d chunk=%d)" % (i, lo, hi, v1, mode, chunk)
        if v1 and want:
            ids = pa.ipc.open_stream(want).read_all().column("stacktrace_id").chunk(0).values.dictionary.to_pylist()
            seen += [x for x in ids if x not in set(seen)]
            req = list(seen)
            rng.shuffle(req)
            req = req[: int(rng.integers(0, len(req) + 1))]
            for k in range(int(rng.integers(0, 4))):
                req.insert(int(rng.integers(0, len(req) + 1)), bytes(rng.integers(0, 256, 16, dtype=np.uint8)))
            if req and rng.random() < 0.3:
                req += req[:2]
            blob = b"".join(req)
            ws, nloc = o.stacktraces(blob)
            rs = a.stacktraces(blob)
            if rs.ipc_bytes() != ws or rs.n_locations != nloc:
                return "stacktrace record differs (case %d, %d ids, cache capacity %d, hash mode %d, rows %d..%d of %d, chunk %d)" % (i, len(req), cap, mode, lo, hi, n, chunk)
    a.close()
    o.close()
    return None


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=300)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    rng = np.random.Generator(np.random.PCG64(args.seed))
    for i in range(args.cases):
        err = one_case(rng, i)
        if err:
            print("MISMATCH:", err)
            sys.exit(1)
    print("stress ok: %d cases, seed %d" % (args.cases, args.seed))
