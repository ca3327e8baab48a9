"""Generates tests/golden/oracle_ipc_sha256.json: SHA-256 of the oracle's IPC stream for fixed
workloads. Run in the build container: `python tests/golden/make_ipc_golden.py`.

These digests freeze the oracle's output (any change to oracle/ or the workload generators shows
up as a digest change that must be justified). They are NOT Go-produced bytes: the reference cannot
run here (see DESIGN.md "parity unpinned"); the logical pin is tests/test_oracle_pyarrow.py.
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kat_workloads as kw  # noqa: E402
from oracle import oracle_py  # noqa: E402
from parca_agent_b200 import abi, synth  # noqa: E402


def cases():
    for name in ["stack_dedup", "writer_basic", "multiple_frame_types", "func_dedup_in_stack", "null_lines"]:
        yield "kat:" + name, getattr(kw, name)()
    for seed in (1, 2, 3):
        for mode, mname in ((abi.PA_HASH_PROVIDED, "provided"), (abi.PA_HASH_XXH64X2, "xxh64x2")):
            yield "edge:seed%d:%s" % (seed, mname), synth.edge_workload(seed=seed, hash_mode=mode)
    yield "config1:head3000", synth.config1().head(3000)
    yield "config1:full", synth.config1()
    yield "config3:scaled", synth.config3(n=2000, u=300, p=512, npids=20, lsets=5)


if __name__ == "__main__":
    out = {}
    for name, w in cases():
        data, st = oracle_py.run(w)
        out[name] = {"sha256": hashlib.sha256(data).hexdigest(), "bytes": len(data), "rows": st["rows"], "unique_stacks": st["unique_stacks"],
                     "locations": st["locations"], "functions": st["functions"]}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_ipc_sha256.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", len(out), "digests")
