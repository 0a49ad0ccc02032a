"""CPU-side checks of the drop-in boundary: the C-ABI shared library builds for sm_100a, loads
without a GPU, exports every symbol include/libmidaspom_cuda.h declares, and refuses to run
(loudly, no CPU fallback) when no B200 is present."""
import ctypes as C
import re
from pathlib import Path

import pytest

import midaspom_b200 as mb
from midaspom_b200 import build as mbuild

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "libmidaspom_cuda.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mp_[a-z_0-9]+)\s*\(", text)))


def test_header_and_python_mirror_agree():
    assert declared_symbols() == sorted(mb.ABI_SYMBOLS)


def test_library_builds_and_exports_every_symbol():
    lib_path = mbuild.build()
    assert lib_path.exists()
    lib = C.CDLL(str(lib_path))
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    lib.mp_version.restype = C.c_char_p
    assert b"sm_100a" in lib.mp_version()


def test_library_carries_sm100a_sass_only():
    import shutil, subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", str(mbuild.build())], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu():
    lib = mb.load_library()
    if lib.mp_device_count() > 0:
        pytest.skip("a GPU is present; the refusal path is for CPU-only hosts")
    with pytest.raises(mb.MpError, match="no CUDA device|no CPU fallback"):
        mb.Engine(8, 7)


def test_bad_arguments_are_reported_not_fatal():
    lib = mb.load_library()
    cfg = mb.MpConfig(0, 1, 0, 0, 0, 0, 0, 0, 1, 0.5)
    h = C.c_void_p()
    assert lib.mp_create(C.byref(cfg), C.byref(h)) == -1          # MP_ERR_ARG before any CUDA call
    assert b"n_patches" in lib.mp_last_error(None)
    assert lib.mp_destroy(None) == 0
