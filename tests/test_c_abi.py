"""The C-ABI library loads without a GPU and exports every symbol include/orca_b200.h declares.
No compute calls here (CPU box)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "orca_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(orca_[a-z_0-9]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from collision_avoidance_b200 import build
    build.build()
    from collision_avoidance_b200 import _lib
    return _lib.load()


def test_header_and_binding_agree():
    from collision_avoidance_b200 import _lib
    assert _declared_symbols() == sorted(_lib.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in _declared_symbols():
        assert hasattr(lib, name), name


def test_abi_version_and_struct_layout(lib):
    from collision_avoidance_b200 import _lib
    assert lib.orca_abi_version() == 2
    assert ctypes.sizeof(_lib.OrcaParams) == 28
    # the library validates struct_size; the Python mirror must be 8-byte aligned and match C
    assert ctypes.sizeof(_lib.OrcaEnvStepArgs) % 8 == 0


def test_argument_errors_surface_as_exceptions(lib):
    from collision_avoidance_b200 import _lib
    h = ctypes.c_void_p()
    p = _lib.OrcaParams(1 / 60., 5.0, 10, 1.5, 1.5, 0.5, 1.0)
    with pytest.raises(ValueError):
        _lib.check(lib.orca_create(ctypes.byref(p), 0, 0, 16, ctypes.byref(h)))  # num_envs = 0
    p2 = _lib.OrcaParams(1 / 60., 5.0, 40, 1.5, 1.5, 0.5, 1.0)
    with pytest.raises(NotImplementedError):
        _lib.check(lib.orca_create(ctypes.byref(p2), 0, 4, 16, ctypes.byref(h)))  # k > 16
    with pytest.raises(ValueError):
        _lib.check(lib.orca_step(None, None, None, None, None))


def test_missing_library_fails_loudly(monkeypatch):
    from collision_avoidance_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/liborca_b200.so")
    with pytest.raises(_lib.OrcaLibraryError):
        _lib.load()


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "collision_avoidance_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "rvo2_oracle" not in src, f


def test_library_has_no_viaddmnmx_with_a_uniform_operand():
    """ptxas 12.9.86 (sm_100a) fuses min(uniform - x, y) into `VIADDMNMX R, R, UR, R` without a
    negate bit, i.e. min(uniform + x, y) (tools/probes/viaddmnmx_probe.cu; it made the observation
    kernel's last chunk run past the batch).  The shipped library must not contain that form."""
    import re
    import shutil
    import subprocess
    from collision_avoidance_b200 import build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump) or not os.path.exists(build.LIB_PATH):
        pytest.skip("cuobjdump or the built library is not here")
    sass = subprocess.run([cuobjdump, "-sass", build.LIB_PATH], capture_output=True, text=True, check=True).stdout
    bad = [l.strip() for l in sass.splitlines() if re.search(r"VIADDMNMX(\.\w+)* R\d+, R\d+, UR\d+", l)]
    assert not bad, bad[:3]
