"""The C-ABI library builds, loads and exports every symbol include/hsv.h declares (no compute, CPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from megatts2_hierspeechpp_b200 import _lib, build
    build.build()
    return _lib.load()


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "hsv.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hsv_[A-Za-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    from megatts2_hierspeechpp_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in hsv.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_helpers(lib):
    assert lib.hsv_version() == 100
    assert lib.hsv_blk16_rows(1) == 64 + 512
    assert lib.hsv_blk16_rows(512) == 64 + 512
    assert lib.hsv_blk16_rows(513) == 64 + 1024
    from megatts2_hierspeechpp_b200 import ops
    for L in (1, 127, 128, 129, 2000, 160000):
        assert ops.blk16_rows(L) == lib.hsv_blk16_rows(L)


def test_argument_errors_are_reported(lib):
    # argument validation happens before any CUDA call, so it is testable without a GPU
    rc = lib.hsv_act1d_snakebeta(None, None, None, None, 1, 8, 16, 0, 1.0, None)
    assert rc == -1 and b"null" in lib.hsv_last_error()
    one = ctypes.c_void_p(16)
    rc = lib.hsv_act1d_snakebeta(one, one, one, one, 1, 12, 16, 1, 1.0, None)
    assert rc == -1 and b"C % 16" in lib.hsv_last_error()
    rc = lib.hsv_conv1d_umma(one, one, None, None, one, None, 0, 1.0, 1, 24, 32, 100, 3, 1, 32, None)
    assert rc == -1 and b"Cin" in lib.hsv_last_error()
    rc = lib.hsv_conv1d_umma(one, one, None, None, one, None, 0, 1.0, 1, 32, 32, 100, 11, 7, 32, None)
    assert rc == -1 and b"halo" in lib.hsv_last_error()
    rc = lib.hsv_conv_transpose1d_direct(one, one, None, None, one, 1, 8, 8, 10, 7, 4, None)
    assert rc == -1


def test_sass_uses_blackwell_tensor_path():
    """cuobjdump of the built library shows tcgen05 MMA (UTC*MMA), TMEM loads and bulk TMA copies."""
    import shutil
    import subprocess
    from megatts2_hierspeechpp_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert re.search(r"UTC\w*MMA", sass), "no tcgen05.mma in SASS"
    assert "LDTM" in sass, "no tcgen05.ld in SASS"
    assert "UBLKCP" in sass, "no cp.async.bulk in SASS"


def test_ops_refuse_cpu_tensors():
    import torch
    from megatts2_hierspeechpp_b200 import ops
    x = torch.zeros(1, 8, 16)
    a = torch.zeros(8)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.act1d(x, a, a)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.conv1d_direct(x, torch.zeros(4, 8, 3), None, pad=1)
