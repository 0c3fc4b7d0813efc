"""CPU: the C-ABI library builds, loads and exports every symbol include/wdpm_b200.h declares.
No compute call is made (there is no GPU here and no CPU path in the library)."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "wdpm_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wdpm_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("wdpm_create", "wdpm_upload", "wdpm_run_block", "wdpm_download_water", "wdpm_destroy", "wdpm_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol(cuda_lib):
    for name in declared_symbols():
        assert hasattr(cuda_lib, name), f"{name} declared in include/wdpm_b200.h but not exported"


def test_abi_version_and_struct_size(cuda_lib):
    from wdpm_b200 import solver
    assert cuda_lib.wdpm_abi_version() == 1
    assert C.sizeof(solver._Config) == 88
    assert C.sizeof(solver._BlockResult) == 48


def test_fused_variant_table(cuda_lib):
    from wdpm_b200 import solver
    for dtype, esize in ((solver.F64, 8), (solver.F32, 4)):
        v = 1
        while True:
            info = solver.fused_variant_info(v, dtype)
            if info is None:
                break
            assert info["strip_cols"] % 12 == 0 and 0 < info["strip_cols"] < info["window_cols"]
            assert info["smem_bytes"] <= 227 * 1024, (v, dtype, info)
            v += 1
        assert v > 3


def test_no_cpu_fallback(cuda_lib):
    """Without a CUDA device create() must fail loudly rather than compute on the host."""
    import torch
    from wdpm_b200 import ADD, Solver, WdpmError
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(WdpmError) as ei:
        Solver(8, 8, -9999.0, ADD)
    assert ei.value.code == -2


def test_product_does_not_import_oracle():
    for path in list((ROOT / "wdpm_b200").rglob("*.py")) + list((ROOT / "wdpm_b200").rglob("*.c*")) + list((ROOT / "wdpm_b200").rglob("*.h")):
        text = path.read_text(errors="ignore")
        assert "pyoracle" not in text and "wdpm_oracle" not in text and "oracle/" not in text.replace("oracle/ ", ""), path
