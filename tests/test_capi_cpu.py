"""CPU-side checks of the drop-in boundary: libhdgpu.so loads, exports every symbol that
include/hyperdeal_b200.h declares, and refuses to compute without a CUDA device."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, has_gpu


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "hyperdeal_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hd_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from hyperdeal_b200 import api, build

    build.build()
    L = ctypes.CDLL(api.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(L, name), name
    # the Python binding covers the same surface
    assert sorted(api.EXPORTS) == declared


def test_version_and_error_string():
    from hyperdeal_b200 import api

    L = api.lib()
    assert L.hd_version() >= 100
    assert isinstance(L.hd_last_error(), bytes)


@pytest.mark.skipif(has_gpu(), reason="checks the no-device behaviour")
def test_no_cpu_fallback():
    from hyperdeal_b200 import api

    with pytest.raises(api.HdError):
        api.Context(0)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under hyperdeal_b200/ may reference it."""
    pkg = os.path.join(ROOT, "hyperdeal_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".hpp", ".cc", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "hd_oracle" not in src and "import oracle" not in src and "from oracle" not in src, f


def test_header_is_plain_c(tmp_path):
    """the drop-in boundary is a C ABI: the header must compile as C99 without any C++ or CUDA type in it"""
    import subprocess

    src = tmp_path / "t.c"
    src.write_text('#include "hyperdeal_b200.h"\nint main(void) { hd_mesh_desc d; hd_halo_send s; (void)d; (void)s; return hd_version() < 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
