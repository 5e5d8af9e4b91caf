"""The C-ABI library loads without a GPU and exports exactly what include/vpca.h declares (no compute calls here)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _header_functions():
    text = (ROOT / "include" / "vpca.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vpca_[a-z_0-9]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as entry
    from spark_examples_b200 import native
    if not native.library_path().exists():
        entry.build()
    return native.load_library()


def test_header_and_binding_agree(lib):
    from spark_examples_b200 import native
    declared = _header_functions()
    assert declared == sorted(native.EXPORTED_SYMBOLS)
    for sym in declared:
        assert hasattr(lib, sym), f"libvpca.so does not export {sym}"


def test_header_is_plain_c():
    text = (ROOT / "include" / "vpca.h").read_text()
    assert 'extern "C"' in text
    assert text.count("VariantsPca.scala") >= 8           # every entry point cites the reference lines it replaces
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)     # comments stripped
    for forbidden in ("torch", "at::", "std::", "#include <string>", "class ", "template"):
        assert forbidden not in code


def test_struct_layout_matches_header(lib):
    from spark_examples_b200 import native
    assert ctypes.sizeof(native.VpcaConfig) == 72
    assert native.VpcaConfig.stream.offset == 48 and native.VpcaConfig.d_gram.offset == 56
    assert native.VpcaConfig.gram_band_row0.offset == 64 and native.VpcaConfig.gram_band_rows.offset == 68
    assert ctypes.sizeof(native.VpcaStats) == 64
    assert lib.vpca_version() == 2


def test_create_fails_loudly_without_gpu(lib):
    """No CPU fallback: on a box without a CUDA device vpca_create must fail with a CUDA error, not succeed."""
    import torch
    from spark_examples_b200 import native
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(native.VpcaError) as ei:
        native.NativePca(16)
    assert ei.value.code == native.VPCA_ERR_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_bad_config_rejected(lib):
    from spark_examples_b200 import native
    cfg = native.VpcaConfig(struct_size=12, n_samples=8)
    h = ctypes.c_void_p()
    assert lib.vpca_create(ctypes.byref(cfg), ctypes.byref(h)) == native.VPCA_ERR_BAD_ARG
    assert b"struct_size" in lib.vpca_last_error(None)
    assert lib.vpca_create(None, ctypes.byref(h)) == native.VPCA_ERR_BAD_ARG
    assert lib.vpca_destroy(None) == native.VPCA_OK


def test_product_never_imports_the_oracle():
    """tier rule: only tests/, smoke() and bench.py's CPU legs may touch oracle/."""
    for path in (ROOT / "spark_examples_b200").rglob("*"):
        if path.suffix in (".py", ".cu", ".cuh", ".h") and path.is_file():
            text = path.read_text(errors="ignore")
            assert "oracle" not in text.replace("oracle regenerates", "").replace("CPU oracle", ""), path
