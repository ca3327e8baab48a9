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
