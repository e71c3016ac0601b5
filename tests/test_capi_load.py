"""The C-ABI library: builds in-tree, loads, exports every symbol include/lorab200.h
declares, and fails loudly without a GPU (no CPU fallback).  No compute calls here.
"""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    hdr = (ROOT / "include" / "lorab200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(lorab200_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_are_exported(lib):
    L = ctypes.CDLL(str(lib.LIB_PATH))
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/lorab200.h but not exported"
    # and the python binding declares exactly the exported set
    assert sorted(lib.EXPORTS) == syms


def test_abi_version_and_strerror(lib):
    L = lib.load_library()
    assert L.lorab200_abi_version() == 3      # 2: rx state / packet structs carry the raw peak and total; lorab200_rx_host; 3: preamble search
    assert L.lorab200_strerror(0) == b"ok"
    assert L.lorab200_strerror(-1) == b"invalid argument"
    assert L.lorab200_table_len(7, lib.TABLE_UP) == 128
    assert L.lorab200_table_len(12, lib.TABLE_FINE) == 4096 * 128
    assert L.lorab200_table_len(6, lib.TABLE_UP) == 0


def test_built_for_sm100a_with_tma_and_packed_fp32(lib):
    """The shipped binary is sm_100a SASS containing the TMA bulk copy and packed FP32 ops."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "-sass", str(lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "UBLKCP" in out, "cp.async.bulk (TMA) missing from SASS"
    assert "FMUL2" in out and "FADD2" in out and "FFMA2" in out, "packed f32x2 ops missing from SASS"
    assert "SYNCS" in out, "mbarrier ops missing from SASS"


def test_create_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lib.Lorab200Error) as e:
        lib.Demod(7)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_bad_arguments_rejected_before_touching_cuda(lib):
    L = lib.load_library()
    ctx = ctypes.c_void_p()
    assert L.lorab200_create(ctypes.byref(ctx), 0, 6, 0) == -1   # sf out of range
    assert L.lorab200_create(ctypes.byref(ctx), 0, 13, 0) == -1
    assert L.lorab200_create(None, 0, 7, 0) == -1
    assert L.lorab200_demod_dev(None, None, 0, 0, 0, None, None, None, None, None) == -1
    L.lorab200_destroy(None)  # must be a no-op


def test_header_is_plain_c(tmp_path):
    """include/lorab200.h must be consumable by a C compiler (the drop-in boundary is a C ABI)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not on PATH")
    src = tmp_path / "use.c"
    src.write_text(
        '#include "lorab200.h"\n'
        "int use(void) {\n"
        "    lorab200_ctx *c = 0; lorab200_result r[4]; lorab200_fine f = {0, 0.0f};\n"
        "    float iq[2 * 128 * 4];\n"
        "    if (lorab200_create(&c, 0, 7, LORAB200_FLAG_DEFAULT) != LORAB200_OK) return 1;\n"
        "    int st = lorab200_demod_host(c, iq, 4, 128, LORAB200_TABLE_UP, &f, r, 0, 0);\n"
        "    lorab200_destroy(c);\n"
        "    return st + (int)sizeof(lorab200_result) - 16;\n"
        "}\n")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only",
                        f"-I{ROOT / 'include'}", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_size_helpers_match_the_reference_blocks(lib):
    """lorab200_encode_symbols / lorab200_mod_samples are pure host functions (no context, no GPU):
    they must agree with what the unmodified reference encoder and modulator produce."""
    import ctypes as C
    import numpy as np
    import loopback
    if not loopback.available():
        pytest.skip("oracle/_ref/libloraloop.so not built")
    L = lib.load_library()
    rng = np.random.default_rng(0)
    for sf, rdd, ex, crc in ((7, 4, 1, 1), (9, 1, 1, 0), (12, 3, 0, 1), (10, 0, 0, 0)):
        cfg = (C.c_int32 * 5)(0, rdd, ex, crc, 1)
        for ln in (1, 7, 20, 100, 255):
            ref = loopback.encode_cfg((sf, 0, rdd, ex, crc, 1), bytes(rng.integers(0, 256, ln, dtype=np.uint8)))
            assert L.lorab200_encode_symbols(sf, C.cast(cfg, C.c_void_p), ln) == ref.size
    for sf, nsy, pad in ((7, 12, 3), (10, 40, 1), (12, 8, 0)):
        ref = loopback.modulate(sf, np.zeros(nsy, np.uint16), padding=pad)
        assert L.lorab200_mod_samples(sf, 1, nsy, pad) == ref.size
    assert L.lorab200_mod_samples(6, 1, 4, 1) == 0 and L.lorab200_mod_samples(7, 0, 4, 1) == 0
