"""Known-answer tests carried over from the reference's own unit tests, run against the CPU oracle.

Reference: reporter/arrow_v2_test.go (6 tests), reporter/parca_reporter_test.go (maybeFixTruncation,
labelsForTID) of parca-dev/parca-agent; XXH64 vectors from python-xxhash (tests/golden/xxh64_kat.json).
"""
import json
import os

import numpy as np
import pyarrow as pa
import pytest

import kat_workloads as kw
import pyref

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def decode(data):
    return pa.ipc.open_stream(data).read_all()


def test_xxh64_known_answers(oracle):
    kat = json.load(open(os.path.join(GOLD, "xxh64_kat.json")))
    for v in kat["words"]:
        data = np.asarray(v["words"], dtype="<u8").tobytes()
        assert oracle.xxh64(data, 0) == v["seed0"]
        assert oracle.xxh64(data, 0x9E3779B97F4A7C15) == v["seedlo"]
        assert pyref.xxh64_words(v["words"], 0) == v["seed0"]
        assert pyref.xxh64_words(v["words"], 0x9E3779B97F4A7C15) == v["seedlo"]
    for v in kat["bytes"]:
        data = bytes.fromhex(v["hex"])
        assert oracle.xxh64(data, 0) == v["seed0"]
        assert oracle.xxh64(data, 7) == v["seed7"]
    # SURVEY §8c anchors
    assert oracle.xxh64(b"", 0) == 0xEF46DB3751D8E999
    assert oracle.xxh64(b"abc", 0) == 0x44BC2CF5AD770999


def test_function_dict_dedup(oracle):  # arrow_v2_test.go:13-47
    L = oracle.lib()
    b = L.orc_funcdict_new()
    assert L.orc_funcdict_append(b, b"main", b"main.go", 10) == 0 and L.orc_funcdict_len(b) == 1
    assert L.orc_funcdict_append(b, b"main", b"main.go", 10) == 0 and L.orc_funcdict_len(b) == 1
    assert L.orc_funcdict_append(b, b"helper", b"util.go", 5) == 1 and L.orc_funcdict_len(b) == 2
    assert L.orc_funcdict_append(b, b"main", b"main.go", 10) == 0 and L.orc_funcdict_len(b) == 2
    L.orc_funcdict_free(b)


def test_stack_dedup_counts(oracle):  # arrow_v2_test.go:144-204
    w = kw.stack_dedup()
    for k, (rows, uniq) in enumerate([(1, 1), (2, 1), (3, 2), (4, 2)], start=1):
        _, st = oracle.run(w.head(k))
        assert (st["rows"], st["unique_stacks"]) == (rows, uniq)
    data, st = oracle.run(w)
    t = decode(data)
    assert t.num_rows == 4 and len(t.column("stacktrace").chunk(0)) == 4
    x = pyref.extract(t)
    assert x["stacktrace"]["offsets"] == [0, 0, 2, 0] and x["stacktrace"]["sizes"] == [2, 2, 1, 2]
    assert x["stacktrace"]["indices"] == [0, 1, 2]
    assert x["stacktrace_id"][0] == (1).to_bytes(8, "big") + (2).to_bytes(8, "big")


def test_writer_basic_schema(oracle):  # arrow_v2_test.go:206-255
    data, st = oracle.run(kw.writer_basic())
    t = decode(data)
    assert t.num_rows == 1
    assert t.schema.metadata[b"parca_write_schema_version"] == b"v2"
    assert t.schema.equals(pyref.expected_schema(["pod", "service"]), check_metadata=True)
    t.validate(full=True)
    row = t.to_pylist()[0]
    assert row["labels"] == {"pod": "pod-1", "service": "my-service"}
    assert row["value"] == 1 and row["period"] == 10**9 // 19 and row["duration"] == 10**9
    assert row["stacktrace"][0]["mapping_file"] == "/usr/bin/test" and row["stacktrace"][0]["mapping_build_id"] == "abc123"


def test_multiple_frame_types(oracle):  # arrow_v2_test.go:257-316
    data, st = oracle.run(kw.multiple_frame_types())
    t = decode(data)
    t.validate(full=True)
    assert t.num_rows == 2
    rows = t.to_pylist()
    assert rows[0]["stacktrace"][0]["frame_type"] == "native" and rows[0]["stacktrace"][0]["lines"] is None
    k = rows[1]["stacktrace"][0]
    assert k["mapping_file"] == "[kernel.kallsyms]" and k["mapping_build_id"] is None
    assert k["lines"] == [{"line": 100, "column": 0, "function": {"system_name": "do_syscall_64", "filename": "vmlinux", "start_line": 0}}]


def test_function_dedup_in_stacktrace(oracle):  # arrow_v2_test.go:318-364
    data, st = oracle.run(kw.func_dedup_in_stack())
    assert st["locations"] == 3 and st["functions"] == 2
    t = decode(data)
    assert t.num_rows == 1 and len(t.to_pylist()[0]["stacktrace"]) == 3


def test_null_lines_for_unsymbolized(oracle):  # arrow_v2_test.go:366-411
    data, _ = oracle.run(kw.null_lines())
    t = decode(data)
    lines = t.column("stacktrace").chunk(0).values.dictionary.field("lines")
    assert not lines[0].is_valid, "native frame (no lines) must have null lines"
    assert lines[1].is_valid, "kernel frame (has lines) must have non-null lines"


def test_maybe_fix_truncation(oracle):  # parca_reporter_test.go:18-41
    chinese = "Go（又稱Golang[4]）是Google開發的一种静态强类型、編譯型、并发型，并具有垃圾回收功能的编程语言。".encode()
    chinese2 = "Linux是一种自由和开放源码的类Unix操作系统。".encode()
    cases = [(b"ASCII string", b"ASCII string", True), (chinese[0:4], None, False), (chinese[0:48], chinese[0:47], True),
             (chinese2[0:48], chinese2[0:48], True), (chinese2, chinese2, True)]
    for s, want, ok in cases:
        got, gok = oracle.fix_truncation(s, 48)
        assert gok == ok and got == want


def test_labels_for_tid_cpu_not_stale(oracle):  # parca_reporter_test.go:64-98
    w, cpus = kw.labels_cpu_sequence()
    t = decode(oracle.run(w)[0])
    rows = t.to_pylist()
    assert [r["labels"]["cpu"] for r in rows] == [str(c) for c in cpus]
    assert all(r["labels"]["thread_id"] == "4243" and r["labels"]["thread_name"] == "myprocess" and r["labels"]["node"] == "test-node" for r in rows)


@pytest.mark.parametrize("flags,present", [(0, {"cpu", "thread_id", "thread_name", "node"}), (1, {"thread_id", "thread_name", "node"}),
                                           (2, {"cpu", "thread_name", "node"}), (4, {"cpu", "thread_id", "node"}), (7, {"node"})])
def test_labels_disable_flags(oracle, flags, present):  # parca_reporter_test.go:100-150
    w, _ = kw.labels_cpu_sequence(flags)
    t = decode(oracle.run(w)[0])
    assert set(t.to_pylist()[0]["labels"].keys()) == present
    assert t.to_pylist()[0]["labels"]["node"] == "test-node"
