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


@pytest.mark.gpu
def test_reporter_v1_offline_log_matches_oracle(oracle, tmp_path):
    """v1 schema + offline mode (parca_reporter.go:1262-1349): every interval logs its sample record followed by the
    stacktrace record of the stacks the log has not seen yet — an empty one when there are none."""
    import pyarrow as pa

    from parca_agent_b200 import abi, padata
    out_file = str(tmp_path / "log.padata")
    p = subprocess.run([build_bin(), "--gpu-v1", out_file], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    version, batches = padata.read(open(out_file, "rb").read())
    assert version == 0 and len(batches) == 4
    w = kw.multiple_frame_types()
    w.schema = abi.PA_SCHEMA_V1
    o = oracle.Oracle(w)
    want = []
    for interval in range(2):
        o.ingest(w.hdrs, w.frame_ids)
        sample, _ = o.flush()
        ids = pa.ipc.open_stream(sample).read_all().column("stacktrace_id").chunk(0).values.dictionary.to_pylist()
        want += [sample, o.stacktraces(b"".join(ids if interval == 0 else []))[0]]
    assert [bytes(b) for b in batches] == want
