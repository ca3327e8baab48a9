"""Generates tests/golden/oracle_v1_sha256.json: SHA-256 of the oracle's v1-schema output for fixed workloads — the
sample record (reporter/arrow.go:274-316) and the stacktrace record (buildStacktraceRecord, parca_reporter.go:1545-1739)
for that interval's stacks plus two ids nobody has seen. Run in the build container: `python tests/golden/make_v1_golden.py`.

Like oracle_ipc_sha256.json these digests freeze the oracle, they are not Go-produced bytes ("parity unpinned").
"""
import hashlib
import json
import os
import sys

import pyarrow as pa

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle_py  # noqa: E402
from parca_agent_b200 import abi, synth  # noqa: E402

MISSING = [bytes([0xEE, i]) * 8 for i in range(2)]


def cases():
    for seed in (1, 2):
        for mode, mname in ((abi.PA_HASH_PROVIDED, "provided"), (abi.PA_HASH_XXH64X2, "xxh64x2")):
            yield "edge:seed%d:%s" % (seed, mname), synth.edge_workload(seed=seed, hash_mode=mode)
    yield "config1:head3000", synth.config1().head(3000)
    yield "ragged:small", synth.ragged(n=4000, u=300, p=1024)


def request_ids(sample_ipc):
    """The interval's unique stack ids in dictionary order, with the two unknown ids spliced in."""
    ids = pa.ipc.open_stream(sample_ipc).read_all().column("stacktrace_id").chunk(0).values.dictionary.to_pylist()
    return [MISSING[0]] + ids[: len(ids) // 2] + [MISSING[1]] + ids[len(ids) // 2:]


def run_oracle(w):
    w.schema = abi.PA_SCHEMA_V1
    o = oracle_py.Oracle(w)
    o.ingest(w.hdrs, w.frame_ids)
    sample, st = o.flush()
    ids = request_ids(sample)
    stack, nloc = o.stacktraces(b"".join(ids))
    o.close()
    return sample, stack, ids, nloc


if __name__ == "__main__":
    out = {}
    for name, w in cases():
        sample, stack, ids, nloc = run_oracle(w)
        out[name] = {"sample_sha256": hashlib.sha256(sample).hexdigest(), "sample_bytes": len(sample), "stacktraces_sha256": hashlib.sha256(stack).hexdigest(),
                     "stacktraces_bytes": len(stack), "ids": len(ids), "locations": nloc}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_v1_sha256.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", len(out), "digests")
