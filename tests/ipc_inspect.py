"""Minimal Arrow IPC stream inspector (no flatbuffers package needed).

Test infrastructure: walks the encapsulated-message framing and decodes the
flatbuffer metadata (Message / Schema / Field / RecordBatch / DictionaryBatch,
per Arrow format/Message.fbs + Schema.fbs) into plain dicts so tests can check
node/buffer layouts, dictionary ids and dictionary-batch order.
"""
import struct

TYPE_NAMES = {0: "NONE", 1: "Null", 2: "Int", 3: "FloatingPoint", 4: "Binary", 5: "Utf8", 6: "Bool",
              7: "Decimal", 8: "Date", 9: "Time", 10: "Timestamp", 11: "Interval", 12: "List",
              13: "Struct_", 14: "Union", 15: "FixedSizeBinary", 16: "FixedSizeList", 17: "Map",
              18: "Duration", 19: "LargeBinary", 20: "LargeUtf8", 21: "LargeList", 22: "RunEndEncoded",
              23: "BinaryView", 24: "Utf8View", 25: "ListView", 26: "LargeListView"}
HEADER_NAMES = {0: "NONE", 1: "Schema", 2: "DictionaryBatch", 3: "RecordBatch"}


class Table:
    def __init__(self, buf, pos):
        self.buf, self.pos = buf, pos
        self.vt = pos - struct.unpack_from("<i", buf, pos)[0]
        self.vtsize = struct.unpack_from("<H", buf, self.vt)[0]

    def _off(self, slot):
        o = 4 + 2 * slot
        if o >= self.vtsize:
            return 0
        return struct.unpack_from("<H", self.buf, self.vt + o)[0]

    def scalar(self, slot, fmt, default=0):
        o = self._off(slot)
        return struct.unpack_from("<" + fmt, self.buf, self.pos + o)[0] if o else default

    def present(self, slot):
        return self._off(slot) != 0

    def _indirect(self, slot):
        o = self._off(slot)
        if not o:
            return None
        p = self.pos + o
        return p + struct.unpack_from("<I", self.buf, p)[0]

    def table(self, slot):
        p = self._indirect(slot)
        return Table(self.buf, p) if p is not None else None

    def string(self, slot):
        p = self._indirect(slot)
        if p is None:
            return None
        n = struct.unpack_from("<I", self.buf, p)[0]
        return bytes(self.buf[p + 4:p + 4 + n]).decode()

    def vec_tables(self, slot):
        p = self._indirect(slot)
        if p is None:
            return None
        n = struct.unpack_from("<I", self.buf, p)[0]
        out = []
        for i in range(n):
            e = p + 4 + 4 * i
            out.append(Table(self.buf, e + struct.unpack_from("<I", self.buf, e)[0]))
        return out

    def vec_structs(self, slot, fmt):
        p = self._indirect(slot)
        if p is None:
            return None
        n = struct.unpack_from("<I", self.buf, p)[0]
        sz = struct.calcsize("<" + fmt)
        return [struct.unpack_from("<" + fmt, self.buf, p + 4 + sz * i) for i in range(n)]


def _kv(tabs):
    return None if tabs is None else [(t.string(0), t.string(1)) for t in tabs]


def _type(t, tt):
    name = TYPE_NAMES.get(tt, str(tt))
    d = {"type": name}
    if t is None:
        return d
    if name == "Int":
        d["bitWidth"] = t.scalar(0, "i"); d["signed"] = bool(t.scalar(1, "B"))
    elif name == "Timestamp":
        d["unit"] = t.scalar(0, "h"); d["tz"] = t.string(1)
    elif name == "FixedSizeBinary":
        d["byteWidth"] = t.scalar(0, "i")
    return d


def _field(f):
    d = {"name": f.string(0), "nullable": bool(f.scalar(1, "B"))}
    d.update(_type(f.table(3), f.scalar(2, "B")))
    de = f.table(4)
    if de is not None:
        it = de.table(1)
        d["dict"] = {"id": de.scalar(0, "q"),
                     "index": None if it is None else (it.scalar(0, "i"), bool(it.scalar(1, "B"))),
                     "ordered": bool(de.scalar(2, "B"))}
    ch = f.vec_tables(5)
    d["children"] = [_field(c) for c in (ch or [])]
    md = _kv(f.vec_tables(6))
    if md is not None:
        d["metadata"] = md
    return d


def _record_batch(rb):
    d = {"length": rb.scalar(0, "q"),
         "nodes": rb.vec_structs(1, "qq") or [],
         "buffers": rb.vec_structs(2, "qq") or [],
         "compression": rb.present(3)}
    v = rb.vec_structs(4, "q")
    if v is not None:
        d["variadic"] = [x[0] for x in v]
    return d


def messages(data):
    """Yield dicts for every encapsulated message in an IPC stream."""
    data = memoryview(data)
    pos = 0
    out = []
    while pos < len(data):
        cont, mlen = struct.unpack_from("<Ii", data, pos)
        assert cont == 0xFFFFFFFF, "missing continuation marker at %d" % pos
        pos += 8
        if mlen == 0:
            out.append({"header": "EOS", "at": pos - 8})
            break
        fb = data[pos:pos + mlen]
        root = Table(fb, struct.unpack_from("<I", fb, 0)[0])
        msg = {"at": pos - 8, "meta_len": mlen, "version": root.scalar(0, "h"),
               "header": HEADER_NAMES.get(root.scalar(1, "B")), "bodyLength": root.scalar(3, "q")}
        h = root.table(2)
        if msg["header"] == "Schema":
            msg["endianness"] = h.scalar(0, "h")
            msg["fields"] = [_field(f) for f in h.vec_tables(1)]
            msg["metadata"] = _kv(h.vec_tables(2))
        elif msg["header"] == "DictionaryBatch":
            msg["id"] = h.scalar(0, "q")
            msg["isDelta"] = bool(h.scalar(2, "B"))
            msg["batch"] = _record_batch(h.table(1))
        elif msg["header"] == "RecordBatch":
            msg["batch"] = _record_batch(h)
        pos += mlen
        msg["body_at"] = pos
        pos += msg["bodyLength"]
        out.append(msg)
    return out


def dict_ids(fields, out=None, path=""):
    """Flatten (path, dict id) pairs in schema pre-order."""
    out = [] if out is None else out
    for f in fields:
        p = path + "/" + f["name"]
        if "dict" in f:
            out.append((p, f["dict"]["id"]))
        dict_ids(f["children"], out, p)
    return out


if __name__ == "__main__":
    import sys, pprint
    for m in messages(open(sys.argv[1], "rb").read()):
        if m["header"] == "Schema":
            pprint.pprint(dict_ids(m["fields"]))
            pprint.pprint(m["fields"], width=160, compact=True)
            print("schema metadata", m["metadata"])
        else:
            pprint.pprint(m, width=160, compact=True)
