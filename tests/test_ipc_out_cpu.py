"""CPU tier: the PRODUCT's IPC stream planner (parca_agent_b200/csrc/ipc_out.hpp) on host-resident buffers.

tests/cpp/test_ipc_out.cpp builds one toy record twice — as pa::Node trees for the product planner and with the oracle's
array model + writer — and requires identical bytes; here the result is also decoded and validated with pyarrow. (On the
GPU tier the same planner is exercised with device buffers by every parity test.)"""
import os
import subprocess

import pyarrow as pa

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_ipc_planner_matches_oracle_writer_and_decodes(tmp_path):
    src = os.path.join(ROOT, "tests", "cpp", "test_ipc_out.cpp")
    exe, out = str(tmp_path / "test_ipc_out"), str(tmp_path / "toy.arrows")
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", exe, src], check=True)
    p = subprocess.run([exe, out], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    batches = list(pa.ipc.open_stream(open(out, "rb").read()))
    assert len(batches) == 1
    t = batches[0]
    t.validate(full=True)
    rows = t.to_pylist()
    assert [r["value"] for r in rows] == [10, -20, None, 40, 50]
    assert [r["label"] for r in rows] == [b"hello", b"hello", b"abc", b"abc", b"abc"]
    assert [r["is_complete"] for r in rows] == [True, False, True, False, True]
    assert rows[3]["locations"] is None and rows[1]["locations"] == []
    assert rows[0]["locations"][1]["function"] == {"name": "a_function_name_longer_than_twelve_bytes", "file": "a.py", "start_line": 0}
    assert [r["view"] for r in rows] == [[0], [1, 1], [], [0, 1, 1], [0]]
    assert t.schema.metadata == {b"parca_write_schema_version": b"test"}
    assert t.schema.field("stacktrace_id").type.extension_name == "arrow.uuid"
