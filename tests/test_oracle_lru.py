"""CPU tier: the oracle's model of `stacks` (lru.SyncedLRU[libpf.TraceHash, libpf.Frames], reporter/parca_reporter.go:106, :224-227,
:1555) against an independent OrderedDict model. The LRU library (github.com/elastic/go-freelru) is not vendored with the reference,
so what is pinned here is the textbook behaviour its README states: Get refreshes, Add evicts the least recently used entry."""
import collections

import numpy as np
import pyarrow as pa

from parca_agent_b200 import abi, synth


class ModelLru:
    def __init__(self, cap):
        self.cap, self.d = cap, collections.OrderedDict()

    def get(self, k):
        if k not in self.d:
            return False
        self.d.move_to_end(k)
        return True

    def add(self, k):
        if len(self.d) >= self.cap:
            self.d.popitem(last=False)
        self.d[k] = True


def missing_flags(ipc):
    """per requested id: True if the record answers it with the "missing stacktrace" placeholder (:1556-1573)"""
    b = list(pa.ipc.open_stream(ipc))[0]
    ll = b.column(b.schema.get_field_index("locations"))
    loc = ll.values
    lines = loc.field(loc.type.get_field_index("lines"))
    fn = lines.values.field(lines.values.type.get_field_index("function_name"))
    names = fn.dictionary.to_pylist()
    idx = fn.indices.to_pylist()
    out = []
    lo, li = ll.offsets.to_pylist(), lines.offsets.to_pylist()
    for i in range(b.num_rows):
        first_loc = lo[i]
        n_loc = lo[i + 1] - lo[i]
        if n_loc != 1 or li[first_loc + 1] - li[first_loc] != 1:
            out.append(False)
            continue
        out.append(names[idx[li[first_loc]]] in (b"missing stacktrace", "missing stacktrace"))
    return out


def test_oracle_stack_cache_is_an_lru(oracle):
    w = synth.config2(n=6000, u=1500, p=4096)  # ~1500 distinct 64-frame stacks
    w.schema = abi.PA_SCHEMA_V1
    cap = 120
    o = oracle.Oracle(w, stack_cache_entries=cap)
    model = ModelLru(cap)
    rng = np.random.Generator(np.random.PCG64(3))
    seen = []
    for k in range(10):
        part = w.rows(np.sort(rng.choice(w.n, 300 if k % 4 == 3 else 70, replace=False)))  # every fourth interval alone overflows the cache
        part.schema = abi.PA_SCHEMA_V1
        o.ingest(part.hdrs, part.frame_ids)
        ipc, _ = o.flush()
        t = pa.ipc.open_stream(ipc).read_all()
        col = t.column("stacktrace_id").chunk(0)
        # run-end encoded: one value per run of equal ids; repeating an access to the same key does not change an LRU's order
        per_row = [bytes(x) for x in col.values.dictionary.take(col.values.indices).to_pylist()]
        assert col.run_ends[-1].as_py() == t.num_rows
        for i, sid in enumerate(per_row):  # ReportTraceEvent, sample by sample (:224-227); samples without frames are hashed like any other
            if not model.get(sid):
                model.add(sid)
        for sid in per_row:
            if sid not in seen:
                seen.append(sid)
        probe = [seen[int(j)] for j in rng.choice(len(seen), min(len(seen), 80), replace=False)] + [b"\xEE\x01" * 8]
        want = [not model.get(sid) for sid in probe]  # buildStacktraceRecord: Get per requested id, in order (:1555)
        got = missing_flags(o.stacktraces(b"".join(probe))[0])
        assert got == want, "interval %d" % k
    assert len(seen) > cap  # the scenario did evict
    o.close()
