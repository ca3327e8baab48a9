"""Committed golden digests of the oracle's IPC output (tests/golden/oracle_ipc_sha256.json)."""
import hashlib
import importlib.util
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_ipc_golden", os.path.join(HERE, "golden", "make_ipc_golden.py"))
gold = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gold)
GOLD = json.load(open(os.path.join(HERE, "golden", "oracle_ipc_sha256.json")))


def test_oracle_matches_golden_digests(oracle):
    seen = set()
    for name, w in gold.cases():
        data, st = oracle.run(w)
        g = GOLD[name]
        assert (len(data), st["rows"], st["unique_stacks"], st["locations"], st["functions"]) == (
            g["bytes"], g["rows"], g["unique_stacks"], g["locations"], g["functions"]), name
        assert hashlib.sha256(data).hexdigest() == g["sha256"], name
        seen.add(name)
    assert seen == set(GOLD)


@pytest.mark.gpu
def test_cuda_path_matches_golden_digests():
    from parca_agent_b200 import lib
    for name, w in gold.cases():
        data, _ = lib.run(w)
        assert hashlib.sha256(data).hexdigest() == GOLD[name]["sha256"], name


# ---- v1 schema: sample record + stacktrace record ----------------------------------------------------
spec1 = importlib.util.spec_from_file_location("make_v1_golden", os.path.join(HERE, "golden", "make_v1_golden.py"))
gold1 = importlib.util.module_from_spec(spec1)
spec1.loader.exec_module(gold1)
GOLD1 = json.load(open(os.path.join(HERE, "golden", "oracle_v1_sha256.json")))


def test_oracle_matches_v1_golden_digests(oracle):
    seen = set()
    for name, w in gold1.cases():
        sample, stack, ids, nloc = gold1.run_oracle(w)
        g = GOLD1[name]
        assert (len(sample), len(stack), len(ids), nloc) == (g["sample_bytes"], g["stacktraces_bytes"], g["ids"], g["locations"]), name
        assert hashlib.sha256(sample).hexdigest() == g["sample_sha256"] and hashlib.sha256(stack).hexdigest() == g["stacktraces_sha256"], name
        seen.add(name)
    assert seen == set(GOLD1)


@pytest.mark.gpu
def test_cuda_path_matches_v1_golden_digests():
    from parca_agent_b200 import abi, lib
    for name, w in gold1.cases():
        w.schema = abi.PA_SCHEMA_V1
        a = lib.from_workload(w)
        lib.load(a, w)
        sample = a.flush().ipc_bytes()
        g = GOLD1[name]
        assert hashlib.sha256(sample).hexdigest() == g["sample_sha256"], name
        r = a.stacktraces(b"".join(gold1.request_ids(sample)))
        assert hashlib.sha256(r.ipc_bytes()).hexdigest() == g["stacktraces_sha256"] and r.n_locations == g["locations"], name
        a.close()
