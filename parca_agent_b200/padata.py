"""Offline-mode `.padata` log framing (host side; no compute).

Reference: setupOfflineModeLog (reporter/parca_reporter.go:1102-1116) writes the 8-byte header
`A6 E7 CC CA | u16 BE format version (0) | u16 BE number of batches`; logDataForOfflineModeV2
(:1807-1831) appends `u32 BE size || IPC stream` per interval and then patches the batch counter at
offset 6. The reader side mirrors uploader/log_uploader.go:486-520 (magic check, counter, sized batches).
The C++ twin is parca::OfflineLog (parca_agent_b200/csrc/reporter.hpp).
"""
import struct

MAGIC = bytes([0xA6, 0xE7, 0xCC, 0xCA])


class Writer:
    def __init__(self, fileobj, version=0):
        self.f, self.n = fileobj, 0
        self.f.write(MAGIC + struct.pack(">HH", version, 0))

    def append(self, ipc_bytes):
        """One interval's IPC stream; the counter is updated only after the payload is written (fsync order in :1819-1831)."""
        self.f.write(struct.pack(">I", len(ipc_bytes)))
        self.f.write(ipc_bytes)
        self.f.flush()
        self.n += 1
        pos = self.f.tell()
        self.f.seek(6)
        self.f.write(struct.pack(">H", self.n & 0xFFFF))
        self.f.seek(pos)


def read(data):
    """bytes of a .padata file -> (version, [ipc stream bytes per batch]). Raises ValueError on a bad file."""
    if len(data) < 8 or data[:4] != MAGIC:
        raise ValueError("not a .padata file (bad magic)")
    version, nbatches = struct.unpack(">HH", data[4:8])
    out, pos = [], 8
    for _ in range(nbatches):
        if pos + 4 > len(data):
            raise ValueError("truncated batch header")
        (sz,) = struct.unpack(">I", data[pos:pos + 4])
        pos += 4
        if pos + sz > len(data):
            raise ValueError("truncated batch")
        out.append(bytes(data[pos:pos + sz]))
        pos += sz
    return version, out
