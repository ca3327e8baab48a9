"""The C++ host layer (parca::ParcaReporter, parca_agent_b200/csrc/reporter.hpp) — the mirror of the
reference's reporter.Reporter implementation above the C ABI. CPU part runs with a recording sink;
the GPU part drives the real library and is compared bit-exactly with the oracle."""
import os
import subprocess

import pytest

import kat_workloads as kw

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "test_reporter")


def build_bin():
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "test_reporter.cpp")
    lib = os.path.join(ROOT, "parca_agent_b200", "libparcaagg.so")
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < max(os.path.getmtime(src), os.path.getmtime(lib)):
        subprocess.run(["g++", "-O1", "-std=c++17", "-o", BIN, src, "-L" + os.path.join(ROOT, "parca_agent_b200"), "-lparcaagg",
                        "-Wl,-rpath," + os.path.join(ROOT, "parca_agent_b200"), "-lpthread"], check=True)
    return BIN


def test_host_layer_logic():
    out = subprocess.run([build_bin()], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == "ok"


@pytest.mark.gpu
def test_reporter_through_c_abi_matches_oracle(oracle, tmp_path):
    """arrow_v2_test.go:257-316 (native + kernel rows) fed through ParcaReporter::ReportTraceEvent."""
    out_file = str(tmp_path / "stream.arrows")
    p = subprocess.run([build_bin(), "--gpu", out_file], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    want, _ = oracle.run(kw.multiple_frame_types())
    assert open(out_file, "rb").read() == want
