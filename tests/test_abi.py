"""The C-ABI shared library loads and exports every symbol include/ttam.h declares (no compute here)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def built():
    from two_tower_augmented_with_adaptive_mimic_mechanism_b200 import _lib
    _lib.build()
    return _lib


def _declared():
    text = (ROOT / "include" / "ttam.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ttam_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(built):
    names = _declared()
    assert len(names) >= 25
    handle = ctypes.CDLL(str(built.LIB_PATH))
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, f"declared in ttam.h but not exported: {missing}"


def test_bindings_cover_the_header(built):
    assert sorted(built.SIGNATURES) == _declared()
    lib = built.lib()
    assert lib.ttam_version() >= 100
    assert isinstance(lib.ttam_last_error(), bytes)


def test_missing_library_fails_loudly(built, monkeypatch, tmp_path):
    monkeypatch.setattr(built, "LIB_PATH", tmp_path / "nope.so")
    monkeypatch.setattr(built, "_LIB", None)
    with pytest.raises(built.TtamError, match="no CPU fallback"):
        built.lib()


def test_library_is_sm100a_only(built):
    import subprocess
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", str(built.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_product_never_imports_oracle():
    pkg = ROOT / "two_tower_augmented_with_adaptive_mimic_mechanism_b200"
    for f in pkg.glob("*.py"):
        assert not re.search(r"^\s*(from|import)\s+oracle\b", f.read_text(), flags=re.M), f


def test_ctypes_structs_match_the_header(built, tmp_path):
    """The ctypes mirrors in _lib.py (TowerDesc, TowerBufs, TowerGrads, TensorList) have the size and the field offsets a C
    compiler gives the structs of include/ttam.h (gcc, plain C: the header is the ABI contract)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "layout.c"
    src.write_text('''
#include <stdio.h>
#include <stddef.h>
#include "ttam.h"
int main(void) {
  printf("%zu %zu %zu %zu\\n", sizeof(ttam_tower_desc), sizeof(ttam_tower_bufs), sizeof(ttam_tower_grads), sizeof(ttam_tensor_list));
  printf("%zu %zu %zu %zu %zu\\n", offsetof(ttam_tower_desc, X), offsetof(ttam_tower_desc, dropout_p), offsetof(ttam_tower_desc, seed),
         offsetof(ttam_tower_desc, state), offsetof(ttam_tower_bufs, q));
  printf("%zu %zu %zu\\n", offsetof(ttam_tower_grads, dW1), offsetof(ttam_tower_grads, accumulate), offsetof(ttam_tower_grads, phase));
  printf("%zu %zu %zu\\n", offsetof(ttam_tensor_list, p), offsetof(ttam_tensor_list, v), offsetof(ttam_tensor_list, numel));
  return 0;
}
''')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    rows = [[int(x) for x in line.split()] for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()]
    L = built
    assert rows[0] == [ctypes.sizeof(L.TowerDesc), ctypes.sizeof(L.TowerBufs), ctypes.sizeof(L.TowerGrads), ctypes.sizeof(L.TensorList)]
    assert rows[1] == [L.TowerDesc.X.offset, L.TowerDesc.dropout_p.offset, L.TowerDesc.seed.offset, L.TowerDesc.state.offset, L.TowerBufs.q.offset]
    assert rows[2] == [L.TowerGrads.dW1.offset, L.TowerGrads.accumulate.offset, L.TowerGrads.phase.offset]
    assert rows[3] == [L.TensorList.p.offset, L.TensorList.v.offset, L.TensorList.numel.offset]
