"""Literal, dict-and-list Python transcription of the reference V2 path — TEST INFRASTRUCTURE.

A second, independent restatement (pure-Python loops, small inputs only) used to pin the C++
oracle: tests decode the oracle's IPC bytes with pyarrow, pull the physical components out with
`extract()`, and compare them with `reference_record()` below. Citations are into
parca-dev/parca-agent (reporter/parca_reporter.go, reporter/arrow_v2.go, reporter/arrow.go).
"""
import pyarrow as pa

from parca_agent_b200 import abi

KIND_TABLE = {  # reporter/parca_reporter.go:338-363: (value_from_hdr, duration, period|None=1e9/sps, delta, 5 strings)
    abi.PA_KIND_CPU: (False, 10**9, None, True, "parca_agent", "samples", "count", "cpu", "nanoseconds"),
    abi.PA_KIND_OFFCPU: (True, 10**9, 0, True, "parca_agent", "wallclock", "nanoseconds", "samples", "count"),
    abi.PA_KIND_CUDA: (True, 10**9, 1, True, "parca_agent", "cuda", "nanoseconds", "cuda", "nanoseconds"),
    abi.PA_KIND_MEM_INUSE_OBJECTS: (True, 0, 512 * 1024, False, "memory", "inuse_objects", "count", "space", "bytes"),
    abi.PA_KIND_MEM_INUSE_SPACE: (True, 0, 512 * 1024, False, "memory", "inuse_space", "bytes", "space", "bytes"),
    abi.PA_KIND_MEM_ALLOC_OBJECTS: (True, 0, 512 * 1024, False, "memory", "alloc_objects", "count", "space", "bytes"),
    abi.PA_KIND_MEM_ALLOC_SPACE: (True, 0, 512 * 1024, False, "memory", "alloc_space", "bytes", "space", "bytes"),
}

_P1, _P2, _P3, _P4, _P5 = 11400714785074694791, 14029467366897019727, 1609587929392839161, 9650029242287828579, 2870177450012600261
_M = (1 << 64) - 1


def _rotl(x, r):
    return ((x << r) | (x >> (64 - r))) & _M


def _round(acc, inp):
    return (_rotl((acc + inp * _P2) & _M, 31) * _P1) & _M


def xxh64_words(words, seed):
    """XXH64 over little-endian uint64 words (length is always a multiple of 8 here)."""
    n = len(words) * 8
    i = 0
    if n >= 32:
        v = [(seed + _P1 + _P2) & _M, (seed + _P2) & _M, seed & _M, (seed - _P1) & _M]
        while i + 4 <= len(words):
            for k in range(4):
                v[k] = _round(v[k], words[i + k])
            i += 4
        h = (_rotl(v[0], 1) + _rotl(v[1], 7) + _rotl(v[2], 12) + _rotl(v[3], 18)) & _M
        for k in range(4):
            h = ((h ^ _round(0, v[k])) * _P1 + _P4) & _M
    else:
        h = (seed + _P5) & _M
    h = (h + n) & _M
    while i < len(words):
        h ^= _round(0, words[i])
        h = (_rotl(h, 27) * _P1 + _P4) & _M
        i += 1
    h ^= h >> 33
    h = (h * _P2) & _M
    h ^= h >> 29
    h = (h * _P3) & _M
    h ^= h >> 32
    return h


class _Ree:
    """array.RunEndEncodedBuilder bookkeeping: run ends are flushed lazily (finishRun)."""

    def __init__(self):
        self.length = 0
        self.run_ends = []

    def finish(self):
        if self.length:
            self.run_ends.append(self.length)

    def append(self, n):
        self.finish()
        self.length += n

    def cont(self, n):
        self.length += n


class _DictRee:  # reporter/arrow.go:83-139
    def __init__(self):
        self.ree = _Ree()
        self.idx = []      # None = null
        self.memo = {}
        self.dict = []

    def null(self):
        self.ree.finish()
        self.idx.append(None)
        self.ree.length += 1

    def ensure(self, n):
        while self.ree.length < n:
            self.null()

    def _bd_append(self, v):
        if v not in self.memo:
            self.memo[v] = len(self.dict)
            self.dict.append(v)
        self.idx.append(self.memo[v])

    def append(self, v):
        if self.idx and self.idx[-1] is not None and self.dict[self.idx[-1]] == v:
            self.ree.cont(1)
            return
        self.ree.append(1)
        self._bd_append(v)

    def out(self):
        self.ree.finish()
        return {"run_ends": self.ree.run_ends, "indices": self.idx, "dict": self.dict}


class _ValRee:  # reporter/arrow.go:23-70, :153-207 (string / int64 / uint64 values)
    def __init__(self, nullable_check=True):
        self.ree = _Ree()
        self.vals = []
        self.nullable_check = nullable_check

    def null(self):
        self.ree.finish()
        self.vals.append(None)
        self.ree.length += 1

    def append(self, v):
        if self.vals and self.vals[-1] is not None and self.vals[-1] == v:
            self.ree.cont(1)
            return
        self.ree.append(1)
        self.vals.append(v)

    def out(self):
        self.ree.finish()
        return {"run_ends": self.ree.run_ends, "values": self.vals}


class _Dict:  # array.BinaryDictionaryBuilder
    def __init__(self):
        self.idx, self.memo, self.dict = [], {}, []

    def append(self, v):
        if v not in self.memo:
            self.memo[v] = len(self.dict)
            self.dict.append(v)
        self.idx.append(self.memo[v])

    def null(self):
        self.idx.append(None)

    def out(self):
        return {"indices": self.idx, "dict": self.dict}


def labels_for_tid(w, h, S):
    """reporter/parca_reporter.go:568-632."""
    lb = {S(n): S(v) for n, v in w.labelsets[int(h["labelset_id"])]}
    if w.label_flags & 7 != 7:
        def set_(name, v):
            if v == b"":
                lb.pop(name, None)
            else:
                lb[name] = v
        if not w.label_flags & abi.PA_LABEL_DISABLE_CPU:
            set_(b"cpu", str(int(h["cpu"])).encode())
        if not w.label_flags & abi.PA_LABEL_DISABLE_THREAD_ID:
            set_(b"thread_id", str(int(h["tid"])).encode())
        if not w.label_flags & abi.PA_LABEL_DISABLE_THREAD_COMM:
            set_(b"thread_name", S(int(h["comm_sid"])))
    return sorted(lb.items())


def reference_record(w):
    """Returns the physical components of the record the reference would build for workload `w`."""
    S = lambda sid: w.strings[int(sid)]  # noqa: E731
    labels = {}
    st_index = {}
    offsets, sizes, indices = [], [], []
    loc_index = {}
    loc = {"address": [], "frame_type": _Dict(), "mapping_file": _Dict(), "mapping_build_id": _Dict(),
           "line_offsets": [], "line": [], "column": [], "func_indices": []}
    func_index = {}
    func = {"system_name": [], "filename": _Dict(), "start_line": []}
    ids, ts, value = [], [], []
    producer, sample_type, sample_unit, period_type, period_unit, temporality = (_ValRee() for _ in range(6))
    period, duration = _ValRee(), _ValRee()
    frame_ids = w.frame_ids

    def append_function(sys, fname, start):  # arrow_v2.go:186-208
        key = (sys, fname, start)
        if key in func_index:
            return func_index[key]
        i = len(func_index)
        func_index[key] = i
        func["system_name"].append(sys if sys != b"" else None)
        if fname == b"":
            func["filename"].null()
        else:
            func["filename"].append(fname)
        func["start_line"].append(start)
        return i

    def line(n, fn):
        loc["line"].append(n)
        loc["column"].append(0)
        loc["func_indices"].append(fn)

    def append_location(fid):  # parca_reporter.go:418-555
        if fid in loc_index:
            return loc_index[fid]
        i = len(loc_index)
        loc_index[fid] = i
        f = w.frames[fid]
        loc["line_offsets"].append(len(loc["line"]))
        loc["address"].append(int(f["address_or_lineno"]))
        kind = int(f["kind"])
        exists = (int(f["flags"]) & 3) == 3
        tname = S(f["type_name_sid"])
        if kind == abi.PA_FRAME_ABORT:
            loc["frame_type"].append(tname)
            loc["mapping_file"].append(b"agent-internal-error-frame")
            loc["mapping_build_id"].null()
            line(0, append_function(b"aborted", b"", 0))
        elif kind == abi.PA_FRAME_NATIVE:
            loc["frame_type"].append(tname)
            if exists:
                loc["mapping_file"].append(S(f["exec_file_name_sid"]))
                bid = S(f["exec_build_id_sid"])
                loc["mapping_build_id"].append(bid if bid != b"" else b"%016x%016x" % (int(f["file_id_hi"]), int(f["file_id_lo"])))
            else:
                loc["mapping_file"].append(b"UNKNOWN")
                loc["mapping_build_id"].null()
        elif kind == abi.PA_FRAME_KERNEL:
            loc["frame_type"].append(tname)
            loc["mapping_file"].append(b"[kernel.kallsyms]")
            loc["mapping_build_id"].null()
            module = S(f["exec_file_name_sid"]) if exists else b"vmlinux"
            fn = S(f["function_name_sid"])
            if fn != b"":
                sym, ln = fn, int(f["source_line"])
            else:
                sym, ln = b"UNKNOWN", 0
            line(ln, append_function(sym, module, 0))
        elif kind == abi.PA_FRAME_OOMPROF:
            loc["frame_type"].append(tname)
            loc["mapping_file"].append(S(f["source_file_sid"]))
            loc["mapping_build_id"].append(S(f["function_name_sid"]))
        else:
            loc["frame_type"].append(tname)
            loc["mapping_file"].append(tname)
            loc["mapping_build_id"].null()
            fn = S(f["function_name_sid"])
            if fn != b"":
                name, path, ln = fn, S(f["source_file_sid"]), int(f["source_line"])
            else:
                name, path, ln = b"UNREPORTED", b"UNREPORTED", 0
            if path == b"":
                path = b"UNKNOWN"
            line(ln, append_function(name, path, 0))
        return i

    for r in range(w.n):
        h = w.hdrs[r]
        fr = [int(x) for x in frame_ids[int(h["frame_off"]):int(h["frame_off"]) + int(h["nframes"])]]
        if w.hash_mode == abi.PA_HASH_XXH64X2:
            key = (xxh64_words(fr, 0), xxh64_words(fr, abi.PA_XXH_SEED_LO))
        else:
            key = (int(h["hash_hi"]), int(h["hash_lo"]))
        nrows = len(value)
        for name, v in labels_for_tid(w, h, S):  # writeSampleV2 :376-378 + Label() arrow_v2.go:543-552
            b = labels.setdefault(name, _DictRee())
            b.ensure(nrows)
            b.append(v)
        if key in st_index:  # AppendStacktrace arrow_v2.go:288-322
            o, s = st_index[key]
        else:
            o = len(indices)
            for fid in fr:
                indices.append(append_location(fid))
            s = len(fr)
            st_index[key] = (o, s)
        offsets.append(o)
        sizes.append(s)
        ids.append(key[0].to_bytes(8, "big") + key[1].to_bytes(8, "big"))
        from_hdr, dur, per, delta, prod, stype, sunit, ptype, punit = KIND_TABLE[int(h["kind"])]
        ts.append(int(h["timestamp_ns"]))
        value.append(int(h["value"]) if from_hdr else 1)
        sample_type.append(stype)
        sample_unit.append(sunit)
        period_type.append(ptype)
        period_unit.append(punit)
        producer.append(prod)
        duration.append(dur)
        period.append(10**9 // w.samples_per_second if per is None else per)
        if delta:
            temporality.append("delta")
        else:
            temporality.null()

    rows = len(value)
    for n_sid, v_sid in w.external_labels:  # LabelAll arrow_v2.go:555-564
        b = labels.setdefault(S(n_sid), _DictRee())
        b.ree.append(rows - b.ree.length)
        b._bd_append(S(v_sid))
    for b in labels.values():
        b.ensure(rows)
    nloc = len(loc["address"])
    nlines = len(loc["line"])
    lo = loc["line_offsets"]
    line_sizes = [(lo[i + 1] - lo[i]) if i < nloc - 1 else nlines - lo[i] for i in range(nloc)]
    return {
        "rows": rows,
        "labels": {k.decode(): v.out() for k, v in sorted(labels.items())},
        "stacktrace": {
            "offsets": offsets, "sizes": sizes, "indices": indices,
            "loc": {
                "address": loc["address"], "frame_type": loc["frame_type"].out(), "mapping_file": loc["mapping_file"].out(),
                "mapping_build_id": loc["mapping_build_id"].out(),
                "lines": {"offsets": lo, "sizes": line_sizes, "valid": [s > 0 for s in line_sizes],
                          "line": loc["line"], "column": loc["column"], "func_indices": loc["func_indices"],
                          "func": {"system_name": func["system_name"], "filename": func["filename"].out(), "start_line": func["start_line"]}},
            },
        },
        "stacktrace_id": ids, "value": value,
        "producer": producer.out(), "sample_type": sample_type.out(), "sample_unit": sample_unit.out(),
        "period_type": period_type.out(), "period_unit": period_unit.out(), "temporality": temporality.out(),
        "period": period.out(), "duration": duration.out(), "timestamp": ts,
    }


# ---- pulling the same components out of a decoded pyarrow batch ------------------------------
def _b(x):
    return None if x is None else (x.encode() if isinstance(x, str) else bytes(x))


def _dict_out(arr):
    return {"indices": arr.indices.to_pylist(), "dict": [_b(v) for v in arr.dictionary.to_pylist()]}


def _ree_out(arr, strings=True):
    vals = arr.values.to_pylist()
    return {"run_ends": arr.run_ends.to_pylist(), "values": vals}


def extract(batch):
    """Physical components of a decoded V2 record batch (pyarrow RecordBatch or 1-chunk Table)."""
    if isinstance(batch, pa.Table):
        batch = batch.combine_chunks().to_batches()[0] if batch.num_rows else None
    col = {f.name: batch.column(i) for i, f in enumerate(batch.schema)}
    labels = {}
    larr = col["labels"]
    for i, f in enumerate(larr.type):
        a = larr.field(i)
        d = _dict_out(a.values)
        labels[f.name] = {"run_ends": a.run_ends.to_pylist(), "indices": d["indices"], "dict": d["dict"]}
    st = col["stacktrace"]
    locd = st.values
    ls = locd.dictionary
    lines = ls.field("lines")
    lstruct = lines.values
    fd = lstruct.field("function")
    fs = fd.dictionary
    nloc = len(ls)
    out = {
        "rows": batch.num_rows,
        "labels": labels,
        "stacktrace": {
            "offsets": st.offsets.to_pylist(), "sizes": st.sizes.to_pylist(), "indices": locd.indices.to_pylist(),
            "loc": {
                "address": ls.field("address").to_pylist(),
                "frame_type": _dict_out(ls.field("frame_type")), "mapping_file": _dict_out(ls.field("mapping_file")),
                "mapping_build_id": _dict_out(ls.field("mapping_build_id")),
                "lines": {"offsets": lines.offsets.to_pylist(), "sizes": lines.sizes.to_pylist(),
                          "valid": [lines[i].is_valid for i in range(nloc)],
                          "line": lstruct.field("line").to_pylist(), "column": lstruct.field("column").to_pylist(),
                          "func_indices": fd.indices.to_pylist(),
                          "func": {"system_name": [_b(v) for v in fs.field("system_name").to_pylist()],
                                   "filename": _dict_out(fs.field("filename")), "start_line": fs.field("start_line").to_pylist()}},
            },
        },
        "stacktrace_id": [bytes(v) for v in col["stacktrace_id"].storage.to_pylist()],
        "value": col["value"].to_pylist(),
        "timestamp": col["timestamp"].cast(pa.int64()).to_pylist(),
    }
    for name in ("producer", "sample_type", "sample_unit", "period_type", "period_unit", "temporality", "period", "duration"):
        out[name] = _ree_out(col[name])
    return out


def expected_schema(label_names):
    """The v2 sample schema (reporter/arrow_v2.go:35-160, :581-609) built independently in pyarrow."""
    u32 = pa.uint32()
    func = pa.struct([pa.field("system_name", pa.string_view(), True), pa.field("filename", pa.dictionary(u32, pa.string()), True),
                      pa.field("start_line", pa.uint64(), False)])
    line = pa.struct([pa.field("line", pa.uint64(), False), pa.field("column", pa.uint64(), False),
                      pa.field("function", pa.dictionary(u32, func), False)])
    loc = pa.struct([pa.field("address", pa.uint64(), False), pa.field("frame_type", pa.dictionary(u32, pa.string()), True),
                     pa.field("mapping_file", pa.dictionary(u32, pa.string()), True),
                     pa.field("mapping_build_id", pa.dictionary(u32, pa.string()), True), pa.field("lines", pa.list_view(line), True)])
    lab = pa.run_end_encoded(pa.int32(), pa.dictionary(u32, pa.string()))
    rs = pa.run_end_encoded(pa.int32(), pa.string())
    return pa.schema([
        pa.field("labels", pa.struct([pa.field(n, lab, True) for n in label_names]), False),
        pa.field("stacktrace", pa.list_view(pa.dictionary(u32, loc)), True),
        pa.field("stacktrace_id", pa.uuid(), False),
        pa.field("value", pa.int64(), False),
        pa.field("producer", rs, False), pa.field("sample_type", rs, False), pa.field("sample_unit", rs, False),
        pa.field("period_type", rs, False), pa.field("period_unit", rs, False), pa.field("temporality", rs, True),
        pa.field("period", pa.run_end_encoded(pa.int32(), pa.int64()), False),
        pa.field("duration", pa.run_end_encoded(pa.int32(), pa.uint64()), False),
        pa.field("timestamp", pa.timestamp("ns", "UTC"), False),
    ], metadata={"parca_write_schema_version": "v2"})


def diff(a, b, path=""):
    """First difference between two extracted structures (None when equal) — readable failures."""
    if isinstance(a, dict) and isinstance(b, dict):
        if a.keys() != b.keys():
            return "%s: keys %s != %s" % (path, sorted(a), sorted(b))
        for k in a:
            d = diff(a[k], b[k], path + "/" + str(k))
            if d:
                return d
        return None
    if isinstance(a, list) and isinstance(b, list):
        if len(a) != len(b):
            return "%s: len %d != %d" % (path, len(a), len(b))
        for i, (x, y) in enumerate(zip(a, b)):
            if x != y:
                return "%s[%d]: %r != %r" % (path, i, x if not isinstance(x, bytes) else x[:40], y if not isinstance(y, bytes) else y[:40])
        return None
    return None if a == b else "%s: %r != %r" % (path, a, b)


# ================================================================================================
# v1 schema (reporter/parca_reporter.go:246-328 + reporter/arrow.go:260-332, :484-521)
V1_KIND_TABLE = {  # (value_from_hdr, duration, period|None=1e9/sps, delta, 5 strings); v1 keeps period=1e9/Hz for off-CPU and CUDA
    abi.PA_KIND_CPU: (False, 10**9, None, True, "parca_agent", "samples", "count", "cpu", "nanoseconds"),
    abi.PA_KIND_OFFCPU: (True, 10**9, None, True, "parca_agent", "wallclock", "nanoseconds", "samples", "count"),
    abi.PA_KIND_CUDA: (True, 10**9, None, True, "parca_agent", "cuda", "nanoseconds", "cuda", "nanoseconds"),
    abi.PA_KIND_MEM_INUSE_OBJECTS: (True, 0, 512 * 1024, False, "memory", "inuse_objects", "count", "space", "bytes"),
    abi.PA_KIND_MEM_INUSE_SPACE: (True, 0, 512 * 1024, False, "memory", "inuse_space", "bytes", "space", "bytes"),
    abi.PA_KIND_MEM_ALLOC_OBJECTS: (True, 0, 512 * 1024, False, "memory", "alloc_objects", "count", "space", "bytes"),
    abi.PA_KIND_MEM_ALLOC_SPACE: (True, 0, 512 * 1024, False, "memory", "alloc_space", "bytes", "space", "bytes"),
}


def reference_record_v1(w):
    S = lambda sid: w.strings[int(sid)]  # noqa: E731
    labels = {}
    stid = _DictRee()
    names = ["producer", "sample_type", "sample_unit", "period_type", "period_unit", "temporality"]
    cols = {n: _DictRee() for n in names}
    period, duration, timestamp = _ValRee(), _ValRee(), _ValRee()
    value = []
    frame_ids = w.frame_ids
    for r in range(w.n):
        h = w.hdrs[r]
        if w.hash_mode == abi.PA_HASH_XXH64X2:
            fr = [int(x) for x in frame_ids[int(h["frame_off"]):int(h["frame_off"]) + int(h["nframes"])]]
            key = (xxh64_words(fr, 0), xxh64_words(fr, abi.PA_XXH_SEED_LO))
        else:
            key = (int(h["hash_hi"]), int(h["hash_lo"]))
        nrows = len(value)
        for name, v in labels_for_tid(w, h, S):
            b = labels.setdefault(name, _DictRee())
            b.ensure(nrows)
            b.append(v)
        from_hdr, dur, per, delta, prod, stype, sunit, ptype, punit = V1_KIND_TABLE[int(h["kind"])]
        stid.append(key[0].to_bytes(8, "big") + key[1].to_bytes(8, "big"))
        timestamp.append(int(h["timestamp_ns"]))
        value.append(int(h["value"]) if from_hdr else 1)
        cols["sample_type"].append(stype.encode())
        cols["sample_unit"].append(sunit.encode())
        cols["period_type"].append(ptype.encode())
        cols["period_unit"].append(punit.encode())
        cols["producer"].append(prod.encode())
        duration.append(dur)
        period.append(10**9 // w.samples_per_second if per is None else per)
        if delta:
            cols["temporality"].append(b"delta")
        else:
            cols["temporality"].null()
    rows = len(value)
    for n_sid, v_sid in w.external_labels:
        b = labels.setdefault(S(n_sid), _DictRee())
        b.ree.append(rows - b.ree.length)
        b._bd_append(S(v_sid))
    for b in labels.values():
        b.ensure(rows)
    out = {"rows": rows, "labels": {k.decode(): v.out() for k, v in sorted(labels.items())}, "stacktrace_id": stid.out(), "value": value,
           "period": period.out(), "duration": duration.out(), "timestamp": timestamp.out()}
    for n in names:
        out[n] = cols[n].out()
    return out


def extract_v1(batch):
    if isinstance(batch, pa.Table):
        batch = batch.combine_chunks().to_batches()[0]
    col = {f.name: batch.column(i) for i, f in enumerate(batch.schema)}

    def dict_ree(a):
        d = _dict_out(a.values)
        return {"run_ends": a.run_ends.to_pylist(), "indices": d["indices"], "dict": d["dict"]}

    out = {"rows": batch.num_rows, "labels": {}, "value": col["value"].to_pylist()}
    for name, a in col.items():
        if name.startswith("labels."):
            out["labels"][name[len("labels."):]] = dict_ree(a)
    for n in ("stacktrace_id", "producer", "sample_type", "sample_unit", "period_type", "period_unit", "temporality"):
        out[n] = dict_ree(col[n])
    for n in ("period", "duration", "timestamp"):
        out[n] = _ree_out(col[n])
    return out


def expected_schema_v1(label_names):
    lab = pa.run_end_encoded(pa.int32(), pa.dictionary(pa.uint32(), pa.binary()))
    ri = pa.run_end_encoded(pa.int32(), pa.int64())
    fields = [pa.field("labels." + n, lab, True) for n in label_names]
    fields += [pa.field("stacktrace_id", lab, False), pa.field("value", pa.int64(), False)]
    fields += [pa.field(n, lab, False) for n in ("producer", "sample_type", "sample_unit", "period_type", "period_unit", "temporality")]
    fields += [pa.field("period", ri, False), pa.field("duration", ri, False), pa.field("timestamp", ri, False)]
    return pa.schema(fields, metadata={"parca_write_schema_version": "v1"})


# ---- v1 stacktrace record: buildStacktraceRecord (reporter/parca_reporter.go:1545-1739) ----------
def known_stacks(w, known=None):
    """The `stacks` LRU after ingesting w (:224-227): hash bytes -> frame ids of the FIRST occurrence."""
    known = {} if known is None else known
    frame_ids = w.frame_ids
    for r in range(w.n):
        h = w.hdrs[r]
        fr = [int(x) for x in frame_ids[int(h["frame_off"]):int(h["frame_off"]) + int(h["nframes"])]]
        if w.hash_mode == abi.PA_HASH_XXH64X2:
            key = (xxh64_words(fr, 0), xxh64_words(fr, abi.PA_XXH_SEED_LO))
        else:
            key = (int(h["hash_hi"]), int(h["hash_lo"]))
        kb = key[0].to_bytes(8, "big") + key[1].to_bytes(8, "big")
        if kb not in known:
            known[kb] = fr
    return known


class _List:  # array.ListBuilder
    def __init__(self):
        self.offsets, self.valid = [], []

    def append(self, v, child_len):
        self.offsets.append(child_len)
        self.valid.append(bool(v))

    def out(self, child_len):
        return {"offsets": self.offsets + [child_len], "valid": self.valid}


def reference_stacktraces(w, ids, known, unknown_type=b"unknown"):
    S = lambda sid: w.strings[int(sid)]  # noqa: E731
    is_complete, loc_list, lines = [], _List(), _List()
    address, line_no, column, start_line = [], [], [], []
    frame_type, mapping_file, mapping_build_id, fn_filename = _DictRee(), _DictRee(), _DictRee(), _DictRee()
    fn_name, fn_sys = _Dict(), _Dict()

    def line(n, col, fn, filename):
        lines.append(True, len(line_no))
        line_no.append(n)
        column.append(col)
        fn_name.append(fn)
        fn_sys.append(b"")
        if filename is None:
            fn_filename.null()
        else:
            fn_filename.append(filename)
        start_line.append(0)

    for sid in ids:
        complete = True
        if sid not in known:  # :1556-1573
            loc_list.append(True, len(address))
            address.append(0)
            frame_type.append(unknown_type)
            mapping_file.null()
            mapping_build_id.null()
            line(0, 0, b"missing stacktrace", None)
            is_complete.append(False)
            continue
        trace = known[sid]
        loc_list.append(len(trace) != 0, len(address))
        for fid in trace:
            f = w.frames[fid]
            address.append(int(f["address_or_lineno"]))
            kind = int(f["kind"])
            exists = (int(f["flags"]) & 3) == 3
            if kind == abi.PA_FRAME_ABORT:
                frame_type.append(S(f["type_name_sid"]))
                mapping_file.append(b"agent-internal-error-frame")
                mapping_build_id.null()
                line(0, 0, b"aborted", None)
                continue
            if kind == abi.PA_FRAME_NATIVE:
                frame_type.append(S(f["type_name_sid"]))
                if exists:
                    mapping_file.append(S(f["exec_file_name_sid"]))
                    bid = S(f["exec_build_id_sid"])
                    mapping_build_id.append(bid if bid else b"%016x%016x" % (int(f["file_id_hi"]), int(f["file_id_lo"])))
                else:
                    mapping_file.append(b"UNKNOWN")
                    mapping_build_id.null()
                    complete = False
                lines.append(False, len(line_no))
            elif kind == abi.PA_FRAME_KERNEL:
                frame_type.append(S(f["type_name_sid"]))
                module = S(f["exec_file_name_sid"]) if exists else b"vmlinux"
                if S(f["function_name_sid"]):
                    symbol, n = S(f["function_name_sid"]), int(f["source_line"])
                else:
                    symbol, n = b"UNKNOWN", 0
                    complete = False
                mapping_build_id.null()
                line(n, int(f["source_column"]), symbol, module)
                mapping_file.append(b"[kernel.kallsyms]")
            elif kind == abi.PA_FRAME_OOMPROF:
                frame_type.append(S(f["type_name_sid"]))
                mapping_file.append(S(f["source_file_sid"]))
                mapping_build_id.append(S(f["function_name_sid"]))
                lines.append(False, len(line_no))
                complete = False
            else:
                frame_type.append(S(f["type_name_sid"]))
                if S(f["function_name_sid"]):
                    name, path, n = S(f["function_name_sid"]), S(f["source_file_sid"]), int(f["source_line"])
                else:
                    name, path, n = b"UNREPORTED", b"UNREPORTED", 0
                    complete = False
                if not path:
                    path = b"UNKNOWN"
                if S(f["gnu_build_id_sid"]):
                    mapping_file.append(S(f["mapping_file_name_sid"]))
                    mapping_build_id.append(S(f["gnu_build_id_sid"]))
                else:
                    mapping_file.append(S(f["type_name_sid"]))
                    mapping_build_id.null()
                line(n, int(f["source_column"]), name, path)
        is_complete.append(complete)
    nloc = len(address)
    zero_ree = {"run_ends": [nloc] if nloc else [], "values": [0]}  # AppendN(0, numMappings), arrow.go:231-238
    return {
        "rows": len(ids), "stacktrace_id": list(ids), "is_complete": is_complete,
        "locations": dict(loc_list.out(nloc), address=address, frame_type=frame_type.out(), mapping_start=zero_ree, mapping_limit=zero_ree,
                          mapping_offset=zero_ree, mapping_file=mapping_file.out(), mapping_build_id=mapping_build_id.out(),
                          lines=dict(lines.out(len(line_no)), line=line_no, column=column, function_name=fn_name.out(),
                                     function_system_name=fn_sys.out(), function_filename=fn_filename.out(), function_start_line=start_line)),
    }


def extract_stacktraces(batch):
    if isinstance(batch, (bytes, bytearray)):  # the stream holds exactly one record batch, possibly with zero rows
        batch = list(pa.ipc.open_stream(batch))[0]
    col = {f.name: batch.column(i) for i, f in enumerate(batch.schema)}

    def dict_ree(a):
        d = _dict_out(a.values)
        return {"run_ends": a.run_ends.to_pylist(), "indices": d["indices"], "dict": d["dict"]}

    def lst(a):
        return {"offsets": a.offsets.to_pylist(), "valid": a.is_valid().to_pylist()}

    ll = col["locations"]
    loc = ll.values
    f = {loc.type.field(i).name: loc.field(i) for i in range(loc.type.num_fields)}
    lines = f["lines"]
    ln = lines.values
    g = {ln.type.field(i).name: ln.field(i) for i in range(ln.type.num_fields)}
    return {
        "rows": batch.num_rows, "stacktrace_id": col["stacktrace_id"].to_pylist(), "is_complete": col["is_complete"].to_pylist(),
        "locations": dict(lst(ll), address=f["address"].to_pylist(), frame_type=dict_ree(f["frame_type"]), mapping_start=_ree_out(f["mapping_start"]),
                          mapping_limit=_ree_out(f["mapping_limit"]), mapping_offset=_ree_out(f["mapping_offset"]),
                          mapping_file=dict_ree(f["mapping_file"]), mapping_build_id=dict_ree(f["mapping_build_id"]),
                          lines=dict(lst(lines), line=g["line"].to_pylist(), column=g["column"].to_pylist(), function_name=_dict_out(g["function_name"]),
                                     function_system_name=_dict_out(g["function_system_name"]), function_filename=dict_ree(g["function_filename"]),
                                     function_start_line=g["function_start_line"].to_pylist())),
    }


def expected_schema_stacktraces():
    db = pa.dictionary(pa.uint32(), pa.binary())
    rd = pa.run_end_encoded(pa.int32(), db)
    ru = pa.run_end_encoded(pa.int32(), pa.uint64())
    line = pa.struct([pa.field("line", pa.int64(), False), pa.field("column", pa.uint64(), False), pa.field("function_name", db, False),
                      pa.field("function_system_name", db, False), pa.field("function_filename", rd, False),
                      pa.field("function_start_line", pa.int64(), False)])
    loc = pa.struct([pa.field("address", pa.uint64(), False), pa.field("frame_type", rd, False), pa.field("mapping_start", ru, False),
                     pa.field("mapping_limit", ru, False), pa.field("mapping_offset", ru, False), pa.field("mapping_file", rd, False),
                     pa.field("mapping_build_id", rd, False), pa.field("lines", pa.list_(pa.field("item", line, True)), False)])
    return pa.schema([pa.field("stacktrace_id", pa.binary(), False), pa.field("is_complete", pa.bool_(), False),
                      pa.field("locations", pa.list_(pa.field("item", loc, True)), False)], metadata={"parca_write_schema_version": "v1"})
