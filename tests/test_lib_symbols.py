"""CPU-side checks of the product library: it loads, exports every symbol include/parcaagg.h
declares, and its host-only helpers (no GPU involved) restate the reference functions."""
import ctypes
import json
import os
import re
import subprocess

import numpy as np

from parca_agent_b200 import abi, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "parcaagg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pa_[a-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol():
    L = lib.lib()
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), "missing export " + n
    assert sorted(lib.EXPORTS) == names
    assert L.pa_agg_abi_version() == abi.PA_ABI_VERSION


def test_struct_sizes_match_header(tmp_path):
    """The numpy/ctypes mirrors against what a C compiler makes of include/parcaagg.h (sizes and a few offsets)."""
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "parcaagg.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(pa_sample_hdr), sizeof(pa_frame_desc), sizeof(pa_agg_config), sizeof(pa_agg_result),"
                   "offsetof(pa_frame_desc, file_id_hi), offsetof(pa_frame_desc, gnu_build_id_sid), offsetof(pa_agg_config, stack_cache_entries));return 0;}\n")
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert got == [abi.HDR_DTYPE.itemsize, abi.FRAME_DTYPE.itemsize, ctypes.sizeof(abi.PaAggConfig), ctypes.sizeof(abi.PaAggResult),
                   abi.FRAME_DTYPE.fields["file_id_hi"][1], abi.FRAME_DTYPE.fields["gnu_build_id_sid"][1], abi.PaAggConfig.stack_cache_entries.offset]
    assert got[:2] == [64, 64]


def test_host_xxh64_known_answers():
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "xxh64_kat.json")))
    for v in kat["words"]:
        data = np.asarray(v["words"], dtype="<u8").tobytes()
        assert lib.xxh64(data, 0) == v["seed0"] and lib.xxh64(data, abi.PA_XXH_SEED_LO) == v["seedlo"]
    for v in kat["bytes"]:
        data = bytes.fromhex(v["hex"])
        assert lib.xxh64(data, 0) == v["seed0"] and lib.xxh64(data, 7) == v["seed7"]


def test_host_fix_truncation():  # reporter/parca_reporter_test.go:18-41
    chinese = "Go（又稱Golang[4]）是Google開發的一种静态强类型、編譯型、并发型，并具有垃圾回收功能的编程语言。".encode()
    chinese2 = "Linux是一种自由和开放源码的类Unix操作系统。".encode()
    cases = [(b"ASCII string", b"ASCII string", True), (chinese[0:4], None, False), (chinese[0:48], chinese[0:47], True),
             (chinese2[0:48], chinese2[0:48], True), (chinese2, chinese2, True)]
    for s, want, ok in cases:
        assert lib.fix_truncation(s, 48) == (want, ok)


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        return
    try:
        lib.Aggregator(max_samples=16)
    except lib.PaError as e:
        assert e.code == -19  # PA_ENODEV: no silent CPU fallback
    else:
        raise AssertionError("Aggregator must not construct without a CUDA device")
