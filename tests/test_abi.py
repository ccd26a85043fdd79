"""The drop-in boundary: libeccbatch.so loads and exports exactly what include/eccbatch.h declares.

CPU only — no compute calls (there is no CPU fallback to call).
"""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "eccbatch.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ecb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_documented_surface():
    fns = header_functions()
    for must in ("ecb_init", "ecb_destroy", "ecb_ed25519_mul_base", "ecb_ed25519_mul", "ecb_x25519", "ecb_x448", "ecb_wei_mul",
                 "ecb_wei_mul_base", "ecb_ecdsa_verify_hashed", "ecb_ed25519_verify_prehashed", "ecb_ed25519_mul_base_dev"):
        assert must in fns


def test_library_exports_every_declared_symbol():
    from eccoxide_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        pytest.fail("libeccbatch.so not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_functions():
        assert hasattr(lib, name), "missing export %s" % name


def test_python_binding_covers_the_header_exactly():
    from eccoxide_b200 import _lib

    assert sorted(_lib.SIGNATURES) == header_functions()
    _lib.load()


def test_rust_ffi_lists_every_exported_symbol():
    """bindings/rust/src/ffi.rs is generated from the header (tools/gen_rust_ffi.py): one extern item per exported
    function, the error codes and the curve ids — the committed file must be what the generator prints."""
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_ffi.py")], capture_output=True, text=True, check=True).stdout
    assert out == open(os.path.join(ROOT, "bindings", "rust", "src", "ffi.rs")).read()
    assert sorted(re.findall(r"pub fn (ecb_[a-z0-9_]+)\(", out)) == header_functions()


def test_no_cuda_device_means_error_not_fallback():
    """Without a GPU every entry point must fail loudly (ECB_ERR_CUDA); nothing is computed on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from eccoxide_b200 import Context, EccBatchError

    with pytest.raises(EccBatchError):
        Context()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "eccoxide_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inc")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.replace("test oracle", "").lower() or f in (), "%s mentions the oracle" % f


def test_pipeline_chunk_schedule_covers_every_slice_exactly():
    """Host logic of the pipelined entry points (run_sharded, eccbatch.cu): the chunk boundaries of a
    device's slice are strictly increasing from lo to hi, no chunk exceeds the configured size, long
    batches start and end on ramped (halved) chunks, short ones are still cut for overlap."""
    import numpy as np

    from eccoxide_b200 import _lib

    lib = _lib.load()
    C = 189440

    def plan(lo, hi, chunk=C, ramp=2):
        cnt = lib.ecb_debug_chunk_plan(lo, hi, chunk, ramp, None, 0)
        assert cnt >= 2
        b = np.zeros(cnt, dtype=np.uint64)
        assert lib.ecb_debug_chunk_plan(lo, hi, chunk, ramp, b.ctypes.data_as(ctypes.c_void_p), cnt) == cnt
        return b.astype(np.int64)

    for lo, hi in ((0, 1), (0, 1000), (7, 7 + 65536), (0, C), (0, C + 1), (0, 473600), (0, 473601), (5, 5 + (1 << 20)), (0, 569097), (0, 1 << 24)):
        for ramp in range(5):
            for chunk in (1000, C):
                b = plan(lo, hi, chunk, ramp)
                sizes = np.diff(b)
                assert b[0] == lo and b[-1] == hi and (sizes > 0).all() and sizes.max() <= chunk, (lo, hi, ramp, chunk)
    sizes = np.diff(plan(0, 1 << 20))
    assert list(sizes[:2]) == [C // 4, C // 2] and list(sizes[-2:]) == [C // 2, C // 4]
    assert len(np.diff(plan(0, 1 << 16))) == 4          # short batch: four chunks in flight
    assert len(np.diff(plan(0, 1 << 16, ramp=0))) == 1  # ramp 0: plain equal chunks
    assert len(np.diff(plan(0, 1000))) == 1
    assert lib.ecb_debug_chunk_plan(10, 5, C, 2, None, 0) == -1
