"""The C-ABI boundary without a GPU: libsofa_b200.so loads, exports every function include/sofa_b200.h declares,
the ctypes binding covers all of them, and the product fails loudly (no CPU fallback) when there is no CUDA
device or no library.  No compute entry point is called here."""
import ctypes
import os
import re

import pytest

from helpers import REPO
from sofa_b200 import _lib

HEADER = os.path.join(REPO, "include", "sofa_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sofa_b200_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    assert len(names) >= 40
    for must in ("sofa_b200_create", "sofa_b200_upload", "sofa_b200_parse", "sofa_b200_emit_csv", "sofa_b200_emit_json",
                 "sofa_b200_gpu_profile", "sofa_b200_window_reduce", "sofa_b200_side_preprocess", "sofa_b200_net_profile",
                 "sofa_b200_nvsmi_profile"):
        assert must in names


def test_library_exports_every_declared_symbol():
    lib = _lib.load_lib()
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert missing == []
    assert lib.sofa_b200_abi_version() == 2


def test_ctypes_binding_covers_the_header():
    assert sorted(set(_lib.EXPORTS)) == declared_functions()


def test_no_device_means_a_loud_error_not_a_fallback():
    lib = _lib.load_lib()
    ctx = ctypes.c_void_p()
    rc = lib.sofa_b200_create(ctypes.byref(ctx), 0)
    if rc == 0:                                   # a GPU is present (GPU box): nothing to check here
        lib.sofa_b200_destroy(ctx)
        pytest.skip("CUDA device present")
    assert rc < 0 and not ctx.value
    assert lib.sofa_b200_last_error()            # a message is available
    from sofa_b200.engine import CudaEngine
    with pytest.raises(_lib.SofaB200Error):
        CudaEngine(0)


def test_missing_library_raises(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", os.path.join(REPO, "sofa_b200", "no_such_lib.so"))
    with pytest.raises(ImportError):
        _lib.load_lib()


def test_product_modules_do_not_import_the_oracle():
    for root, _, files in os.walk(os.path.join(REPO, "sofa_b200")):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(root, fn)).read()
                assert "oracle" not in re.sub(r"#.*", "", src).replace("Oracle", ""), fn
    assert "oracle" not in open(os.path.join(REPO, "bin", "sofa")).read()
