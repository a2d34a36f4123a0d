"""CPU-side checks of the boundary: the library loads and exports every symbol include/mfsdbg.h declares; without
a GPU every compute entry point refuses loudly (no CPU fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "mfsdbg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mfsdbg_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from mitoflex_b200 import lib
    assert _declared() == sorted(lib.SYMBOLS)


def test_library_exports_every_symbol():
    from mitoflex_b200 import lib
    L = lib.load()
    for name in _declared():
        assert hasattr(L, name), name
    assert L.mfsdbg_version() == 100


def test_no_cpu_fallback_without_device(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mitoflex_b200 import lib
    assert lib.device_count() == 0
    with pytest.raises(lib.MfsdbgError) as ei:
        lib.Context(0)
    assert ei.value.code == lib.ENODEV
    with pytest.raises(lib.MfsdbgError) as ei:
        lib.count(k=21, min_count=2, read_lib_file=str(tmp_path / "reads.lib"), output_prefix=str(tmp_path / "out"))
    assert ei.value.code == lib.ENODEV
    with pytest.raises(lib.MfsdbgError) as ei:
        lib.count(k=21, min_count=2)
    assert ei.value.code == lib.EINVAL
