"""Optional LZ4_FRAME body compression (the reference's gRPC framing, ipc.WithLZ4(), parca_reporter.go:1851).

Not a bit-exact mode — the Go side compresses with pierrec/lz4, which is not available here — so the check is: the stream
is valid Arrow IPC with BodyCompression set, and decodes to exactly the record of the uncompressed stream. Host-only code
(parca_agent_b200/csrc/ipc_lz4.hpp through pa_ipc_compress_lz4), so the core of it runs on the CPU tier."""
import pyarrow as pa
import pytest

import ipc_inspect
from parca_agent_b200 import abi, lib, synth


def roundtrip(plain):
    packed = lib.compress_lz4(plain)
    a, b = list(pa.ipc.open_stream(plain)), list(pa.ipc.open_stream(packed))
    assert len(a) == len(b) == 1
    assert a[0].schema.equals(b[0].schema, check_metadata=True)
    assert a[0].equals(b[0])
    a[0].validate(full=True)
    b[0].validate(full=True)
    return packed


def test_compressed_stream_decodes_to_the_same_record(oracle):
    for w in (synth.edge_workload(seed=1, external=False), synth.config1().head(4_000)):  # pyarrow's equals() on REE columns is slow
        for schema in (abi.PA_SCHEMA_V2, abi.PA_SCHEMA_V1):
            w.schema = schema
            plain, _ = oracle.run(w)
            packed = roundtrip(plain)
            if w.n > 3_000:
                assert len(packed) < 0.8 * len(plain)
    # the v1 stacktrace record (List<Struct<... List<Struct>>>, REE, Bool) as well
    w = synth.config1().head(3_000)
    w.schema = abi.PA_SCHEMA_V1
    o = oracle.Oracle(w)
    o.ingest(w.hdrs, w.frame_ids)
    sample, _ = o.flush()
    ids = pa.ipc.open_stream(sample).read_all().column("stacktrace_id").chunk(0).values.dictionary.to_pylist()
    st, _ = o.stacktraces(b"".join(ids))
    roundtrip(st)
    o.close()


def test_compressed_layout(oracle):
    """Every record/dictionary batch carries BodyCompression; non-empty buffers start with their int64 uncompressed length
    (or -1 = stored raw) followed by an LZ4 frame; the schema message is untouched."""
    import struct
    plain, _ = oracle.run(synth.config1().head(5_000))
    packed = lib.compress_lz4(plain)
    mp, mc = ipc_inspect.messages(plain), ipc_inspect.messages(packed)
    assert [m["header"] for m in mp] == [m["header"] for m in mc] and mp[0]["header"] == "Schema" and mp[-1]["header"] == "EOS"
    assert plain[:mp[0]["body_at"]] == packed[:mc[0]["body_at"]]  # schema message verbatim
    saw_frame = saw_raw = False
    for m, c in zip(mp[1:-1], mc[1:-1]):
        assert c["batch"]["compression"] and not m["batch"]["compression"]
        assert len(m["batch"]["buffers"]) == len(c["batch"]["buffers"]) and m["batch"]["nodes"] == c["batch"]["nodes"]
        assert m["batch"].get("variadic") == c["batch"].get("variadic") and m["batch"]["length"] == c["batch"]["length"]
        for (_, n), (off, cn) in zip(m["batch"]["buffers"], c["batch"]["buffers"]):
            if n == 0:
                assert cn == 0
                continue
            at = c["body_at"] + off
            prefix = struct.unpack_from("<q", packed, at)[0]
            if prefix == -1:
                saw_raw = True
                assert cn == n + 8
            else:
                saw_frame = True
                assert prefix == n and packed[at + 8:at + 12] == bytes.fromhex("04224d18")
    assert saw_frame and saw_raw


def test_compress_rejects_garbage():
    with pytest.raises(lib.PaError):
        lib.compress_lz4(b"\x00" * 64)
    with pytest.raises(lib.PaError):
        lib.compress_lz4(b"")


@pytest.mark.gpu
def test_aggregator_lz4_mode_matches_plain_mode():
    for schema in (abi.PA_SCHEMA_V2, abi.PA_SCHEMA_V1):
        w = synth.config3(n=50_000, u=3_000, p=4_096, npids=64, lsets=6)
        w.schema = schema
        plain, _ = lib.run(w)
        a = lib.from_workload(w, ipc_compression=abi.PA_IPC_LZ4_FRAME)
        lib.load(a, w)
        r = a.flush()
        packed = r.ipc_bytes()
        # byte-identical to the standalone compressor, whose output the CPU tests show to be record-identical
        assert len(packed) < len(plain) and packed == lib.compress_lz4(plain)
        t = list(pa.ipc.open_stream(packed))[0]
        assert t.num_rows == w.n and t.column("value").to_pylist() == list(pa.ipc.open_stream(plain))[0].column("value").to_pylist()
        if schema == abi.PA_SCHEMA_V1:
            ids = a.last_stack_ids(r.n_unique_stacks).tobytes()
            st = a.stacktraces(ids).ipc_bytes()
            assert list(pa.ipc.open_stream(st))[0].num_rows == r.n_unique_stacks
        a.close()


def test_compress_survives_corrupted_streams(oracle):
    """pa_ipc_compress_lz4 is an exported entry point: truncations and bit flips of a valid stream must give an error code
    or a (possibly meaningless) stream, never a crash — its flatbuffer reader checks every offset against the metadata."""
    import numpy as np
    plain, _ = oracle.run(synth.edge_workload(seed=2, n=300))
    rng = np.random.Generator(np.random.PCG64(5))
    meta_end = ipc_inspect.messages(plain)[1]["body_at"]  # schema + first dictionary batch metadata: where the offsets live
    outcomes = {"ok": 0, "err": 0}
    for k in range(300):
        b = bytearray(plain)
        if k % 3 == 0:
            b = b[: int(rng.integers(0, len(b)))]
        else:
            for _ in range(int(rng.integers(1, 4))):
                pos = int(rng.integers(0, meta_end if k % 3 == 1 else len(b)))
                b[pos] ^= 1 << int(rng.integers(0, 8))
        try:
            lib.compress_lz4(bytes(b))
            outcomes["ok"] += 1
        except lib.PaError:
            outcomes["err"] += 1
    assert outcomes["err"] > 50 and outcomes["ok"] + outcomes["err"] == 300
