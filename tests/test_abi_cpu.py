"""The C-ABI library loads, exports every symbol include/smx.h declares, and refuses to run
without a CUDA device (no CPU fallback)."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_header_symbols_are_exported():
    from multiplexnanopore_b200 import _lib
    header = (ROOT / "include" / "smx.h").read_text()
    declared = set(re.findall(r"\b(smx_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/smx.h but not exported by libsmx.so"
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    assert lib.smx_version() >= 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from multiplexnanopore_b200 import Engine, EngineError
    with pytest.raises(EngineError, match="no CUDA device|no CPU fallback|failed"):
        Engine(0)


def test_product_never_imports_oracle():
    for py in (ROOT / "multiplexnanopore_b200").rglob("*.py"):
        text = py.read_text()
        assert "oracle" not in re.sub(r"#.*", "", text).replace("no oracle", ""), f"{py} mentions the oracle"
