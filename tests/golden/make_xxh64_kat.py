"""Generates tests/golden/xxh64_kat.json with the python-xxhash package (xxhash 3.7.0 here).

Run in the build container: `python tests/golden/make_xxh64_kat.py`. The vectors pin the oracle's
XXH64 restatement (oracle/xxh64.h) and, through it, the CUDA hash kernel.
"""
import json
import os
import struct

import numpy as np
import xxhash

SEED_LO = 0x9E3779B97F4A7C15
rng = np.random.Generator(np.random.PCG64(1234))
vec = []
for nwords in [0, 1, 2, 3, 4, 5, 7, 8, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 128, 255]:
    words = rng.integers(0, 2**63, nwords, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, nwords, dtype=np.uint64)
    data = words.astype("<u8").tobytes()
    vec.append({"words": [int(w) for w in words], "seed0": xxhash.xxh64(data, seed=0).intdigest(),
                "seedlo": xxhash.xxh64(data, seed=SEED_LO).intdigest()})
byte_vec = []
for s in [b"", b"abc", b"a", b"0123456789ab", struct.pack("<64Q", *range(64)), bytes(range(37)), bytes(range(101))]:
    byte_vec.append({"hex": s.hex(), "seed0": xxhash.xxh64(s, seed=0).intdigest(), "seed7": xxhash.xxh64(s, seed=7).intdigest()})
out = {"generator": "xxhash " + xxhash.VERSION, "words": vec, "bytes": byte_vec}
with open(os.path.join(os.path.dirname(__file__), "xxh64_kat.json"), "w") as f:
    json.dump(out, f, indent=0)
print("wrote", len(vec), "+", len(byte_vec), "vectors")
