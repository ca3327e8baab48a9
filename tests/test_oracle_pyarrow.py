"""Pins the C++ oracle's logical + physical output with an independent decode (pyarrow 24) and an
independent literal Python transcription (tests/pyref.py) on small and adversarial batches."""
import numpy as np
import pyarrow as pa
import pytest

import ipc_inspect
import pyref
from parca_agent_b200 import abi, synth


def check(oracle, w, validate=True):
    data, st = oracle.run(w)
    t = pa.ipc.open_stream(data).read_all()
    got = pyref.extract(t)
    want = pyref.reference_record(w)
    d = pyref.diff(want, got)
    assert d is None, d
    assert t.schema.equals(pyref.expected_schema(list(want["labels"].keys())), check_metadata=True)
    assert st["rows"] == want["rows"] == w.n
    assert st["locations"] == len(want["stacktrace"]["loc"]["address"])
    assert st["functions"] == len(want["stacktrace"]["loc"]["lines"]["func"]["start_line"])
    assert st["location_indices"] == len(want["stacktrace"]["indices"])
    if validate:
        t.validate(full=True)
    return data, t


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("mode", [abi.PA_HASH_PROVIDED, abi.PA_HASH_XXH64X2])
def test_edge_batches(oracle, seed, mode):
    # external labels colliding with fully-present labels yield a zero-length run (arrow_v2.go:562),
    # which pyarrow's full validation rejects — that is the reference's behaviour, so skip validate there
    check(oracle, synth.edge_workload(seed=seed, hash_mode=mode), validate=False)
    check(oracle, synth.edge_workload(seed=seed, hash_mode=mode, external=False), validate=True)


@pytest.mark.parametrize("flags", [1, 2, 4, 7, 3])
def test_edge_label_flags(oracle, flags):
    check(oracle, synth.edge_workload(seed=11, label_flags=flags, external=False))


def test_config1_prefix(oracle):
    w = synth.config1().head(3000)
    check(oracle, w)


def test_config3_style_prefix(oracle):
    w = synth.config3(n=2000, u=300, p=512, npids=20, lsets=5)
    check(oracle, w)


def test_empty_batch_is_skipped(oracle):
    w = synth.edge_workload(seed=1).head(0)
    data, st = oracle.run(w)
    assert data == b"" and st["rows"] == 0  # reporter/parca_reporter.go:1775-1778


def test_ipc_layout(oracle):
    """Framing facts the product writer must share: dictionary ids pre-order, batches inner-first."""
    w = synth.edge_workload(seed=5, external=False)
    data, _ = oracle.run(w)
    msgs = ipc_inspect.messages(data)
    assert msgs[0]["header"] == "Schema" and msgs[-1]["header"] == "EOS" and msgs[-2]["header"] == "RecordBatch"
    ids = ipc_inspect.dict_ids(msgs[0]["fields"])
    nl = len([f for f in msgs[0]["fields"][0]["children"]])
    assert [i for _, i in ids] == list(range(nl + 6))
    order = [m["id"] for m in msgs if m["header"] == "DictionaryBatch"]
    assert order == list(range(nl)) + [nl + 1, nl + 2, nl + 3, nl + 5, nl + 4, nl]
    for m in msgs[:-1]:
        assert m["body_at"] % 8 == 0 and m["bodyLength"] % 8 == 0
        if "batch" in m:
            assert all(off % 8 == 0 for off, _ in m["batch"]["buffers"])
    # stacktrace_id is the arrow.uuid extension over fixed_size_binary[16]
    f = msgs[0]["fields"][2]
    assert f["type"] == "FixedSizeBinary" and f["byteWidth"] == 16 and ("ARROW:extension:name", "arrow.uuid") in f["metadata"]


def test_string_view_blocks(oracle):
    """system_name > 12 bytes goes to 32 KiB view blocks; a 40 kB name gets a block of its own."""
    w = synth.edge_workload(seed=4, external=False)
    data, _ = oracle.run(w)
    msgs = ipc_inspect.messages(data)
    nl = len(msgs[0]["fields"][0]["children"])
    fb = [m for m in msgs if m["header"] == "DictionaryBatch" and m["id"] == nl + 4][0]
    assert fb["batch"]["variadic"][0] >= 1


# ---- v1 schema (the reference's default, --remote-store-use-v2-schema=false) --------------------
def check_v1(oracle, w):
    w.schema = abi.PA_SCHEMA_V1
    data, st = oracle.run(w)
    t = pa.ipc.open_stream(data).read_all()
    got = pyref.extract_v1(t)
    want = pyref.reference_record_v1(w)
    d = pyref.diff(want, got)
    assert d is None, d
    assert t.schema.equals(pyref.expected_schema_v1(list(want["labels"].keys())), check_metadata=True)
    assert st["rows"] == w.n and st["unique_stacks"] == len(want["stacktrace_id"]["dict"])
    return t


@pytest.mark.parametrize("seed", [1, 2])
@pytest.mark.parametrize("mode", [abi.PA_HASH_PROVIDED, abi.PA_HASH_XXH64X2])
def test_v1_edge_batches(oracle, seed, mode):
    check_v1(oracle, synth.edge_workload(seed=seed, hash_mode=mode))
    t = check_v1(oracle, synth.edge_workload(seed=seed, hash_mode=mode, external=False))
    t.validate(full=True)


def test_v1_config1_prefix(oracle):
    t = check_v1(oracle, synth.config1().head(3000))
    t.validate(full=True)
    row = t.slice(0, 1).to_pylist()[0]
    assert row["period"] == 10**9 // 19 and row["duration"] == 10**9 and row["producer"] == b"parca_agent" and row["temporality"] == b"delta"
    assert len(row["stacktrace_id"]) == 16


# ---- v1 stacktrace record (buildStacktraceRecord, parca_reporter.go:1545-1739) -------------------
def stacktrace_ids(w, known, extra_missing=3):
    """The ids a server would ask for: every stack of the batch in first-occurrence order, plus ids nobody has seen."""
    ids = list(known.keys())
    for i in range(extra_missing):
        ids.insert((i * 7) % (len(ids) + 1), bytes([0xEE, i]) * 8)
    return ids


def check_stacktraces(oracle, w, ids=None):
    w.schema = abi.PA_SCHEMA_V1
    o = oracle.Oracle(w)
    o.ingest(w.hdrs, w.frame_ids)
    o.flush()
    known = pyref.known_stacks(w)
    ids = stacktrace_ids(w, known) if ids is None else ids
    data, nloc = o.stacktraces(b"".join(ids))
    o.close()
    t = pa.ipc.open_stream(data).read_all()
    want = pyref.reference_stacktraces(w, ids, known)
    got = pyref.extract_stacktraces(data)
    d = pyref.diff(want, got)
    assert d is None, d
    assert t.schema.equals(pyref.expected_schema_stacktraces(), check_metadata=True)
    assert nloc == len(want["locations"]["address"])
    return t, want


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("mode", [abi.PA_HASH_PROVIDED, abi.PA_HASH_XXH64X2])
def test_stacktrace_record_edge(oracle, seed, mode):
    t, want = check_stacktraces(oracle, synth.edge_workload(seed=seed, hash_mode=mode))
    t.validate(full=True)
    assert not all(want["is_complete"]) and any(want["is_complete"])
    assert b"missing stacktrace" in want["locations"]["lines"]["function_name"]["dict"]
    assert False in want["locations"]["valid"]  # an empty stack is a null list entry (:1576-1580)


def test_stacktrace_record_config1(oracle):
    t, want = check_stacktraces(oracle, synth.config1().head(400))
    t.validate(full=True)
    row = t.slice(1, 1).to_pylist()[0]
    assert len(row["stacktrace_id"]) == 16 and len(row["locations"]) == 16
    assert row["locations"][0]["mapping_start"] == 0 and row["locations"][0]["mapping_limit"] == 0


def test_stacktrace_record_degenerate(oracle):
    w = synth.edge_workload(seed=5)
    check_stacktraces(oracle, w, ids=[])                      # no new stacks this interval: an empty record is still written (:1332)
    check_stacktraces(oracle, w, ids=[b"\x01" * 16])           # only a missing stack
    known = pyref.known_stacks(w)
    empties = [k for k, v in known.items() if len(v) == 0]
    assert empties
    check_stacktraces(oracle, w, ids=empties)                 # only empty stacks: zero locations, null list entries


@pytest.mark.parametrize("seed", list(range(10, 22)))
def test_stacktrace_record_random_batches(oracle, seed):
    """More seeds of the adversarial generator, random request orders: the oracle's buildStacktraceRecord against the literal
    Python transcription (the only pin this record has — the reference ships no test for it)."""
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(seed))
    mode = abi.PA_HASH_PROVIDED if seed % 2 else abi.PA_HASH_XXH64X2
    w = synth.edge_workload(seed=seed, n=int(rng.integers(50, 1500)), hash_mode=mode)
    known = pyref.known_stacks(w)
    ids = list(known.keys())
    rng.shuffle(ids)
    ids = ids[: int(rng.integers(0, len(ids) + 1))] + [bytes(rng.integers(0, 256, 16, dtype=np.uint8)) for _ in range(int(rng.integers(0, 3)))]
    rng.shuffle(ids)
    check_stacktraces(oracle, w, ids=ids)


@pytest.mark.parametrize("seed", list(range(30, 42)))
def test_sample_records_random_batches(oracle, seed):
    """More seeds, sizes, label-flag combinations and external-label choices for both sample records: the C++ oracle against
    the literal Python transcription (logical pin; the reference has no golden bytes for this path)."""
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(seed))
    mode = abi.PA_HASH_PROVIDED if seed % 2 else abi.PA_HASH_XXH64X2
    kw = dict(seed=seed, n=int(rng.integers(1, 1200)), hash_mode=mode, label_flags=int(rng.integers(0, 8)), external=bool(rng.random() < 0.5))
    check(oracle, synth.edge_workload(**kw), validate=False)
    check_v1(oracle, synth.edge_workload(**kw))
