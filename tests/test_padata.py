"""Offline log framing round trip (reporter/parca_reporter.go:1102-1116, :1807-1831)."""
import io

import pyarrow as pa
import pytest

from parca_agent_b200 import padata, synth


def test_header_and_counter_layout():
    f = io.BytesIO()
    w = padata.Writer(f)
    assert f.getvalue() == bytes([0xA6, 0xE7, 0xCC, 0xCA, 0, 0, 0, 0])
    w.append(b"\x01\x02\x03")
    w.append(b"\x09\x08")
    assert f.getvalue() == bytes([0xA6, 0xE7, 0xCC, 0xCA, 0, 0, 0, 2, 0, 0, 0, 3, 1, 2, 3, 0, 0, 0, 2, 9, 8])
    assert padata.read(f.getvalue()) == (0, [b"\x01\x02\x03", b"\x09\x08"])
    with pytest.raises(ValueError):
        padata.read(b"nope....")
    with pytest.raises(ValueError):
        padata.read(f.getvalue()[:-1])


def test_round_trip_of_oracle_batches(oracle):
    f = io.BytesIO()
    w = padata.Writer(f)
    batches = [oracle.run(synth.edge_workload(seed=s, external=False))[0] for s in (1, 2)]
    for b in batches:
        w.append(b)
    version, got = padata.read(f.getvalue())
    assert version == 0 and got == batches
    for b in got:
        assert pa.ipc.open_stream(b).read_all().num_rows == 600


@pytest.mark.gpu
def test_offline_log_of_gpu_flushes(oracle):
    from parca_agent_b200 import lib
    w = synth.edge_workload(seed=3, n=900)
    a = lib.from_workload(w)
    f = io.BytesIO()
    log = padata.Writer(f)
    parts = [w.head(400), w.rows(range(400, 900))]
    for p in parts:
        lib.load(a, p)
        log.append(a.flush().ipc_bytes())
    a.close()
    _, got = padata.read(f.getvalue())
    assert got == [oracle.run(p)[0] for p in parts]
