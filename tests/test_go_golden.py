"""Go-produced golden bytes (tools/golden/): consumed when present. Without them parity stays pinned to the oracle only —
the reference is Go and no Go toolchain exists in this image."""
import glob
import json
import os

import pytest

import kat_workloads as kw
from parca_agent_b200 import padata

GO_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "go")
FILES = sorted(glob.glob(os.path.join(GO_DIR, "*.arrows")))


def cases():
    types = None
    p = os.path.join(GO_DIR, "frame_types.json")
    if os.path.exists(p):
        types = json.load(open(p))
    return kw.go_pin_cases(types)


def test_inputs_exist_for_every_go_case(oracle):
    """the workloads the Go recipe mirrors are valid inputs (this part runs with or without Go files)"""
    for name, w in cases().items():
        data, st = oracle.run(w)
        assert st["rows"] == w.n and data


@pytest.mark.skipif(not FILES, reason="no Go-produced golden files under tests/golden/go (see tools/golden/README.md)")
def test_oracle_reproduces_go_bytes(oracle):
    c = cases()
    for f in FILES:
        name = os.path.splitext(os.path.basename(f))[0]
        want = open(f, "rb").read()
        got, _ = oracle.run(c[name])
        assert got == want, "oracle differs from the Go-produced stream for case %s" % name
        log = os.path.join(GO_DIR, name + ".padata")
        if os.path.exists(log):
            assert padata.read(open(log, "rb").read())[1] == [want]


@pytest.mark.gpu
@pytest.mark.skipif(not FILES, reason="no Go-produced golden files under tests/golden/go")
def test_cuda_path_reproduces_go_bytes():
    from parca_agent_b200 import lib
    c = cases()
    for f in FILES:
        name = os.path.splitext(os.path.basename(f))[0]
        assert lib.run(c[name])[0] == open(f, "rb").read(), name
