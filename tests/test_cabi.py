"""The C-ABI shared library: loads without a GPU and exports every symbol include/fw25.h declares."""

import ctypes
import re
from pathlib import Path

from fullwave25_b200 import engine, mapgen

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "fw25.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fw25_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_entry_points():
    syms = declared_symbols()
    for must in ("fw25_run", "fw25_create", "fw25_destroy", "fw25_sweep_u", "fw25_sweep_p",
                 "fw25_inject", "fw25_record", "fw25_last_error"):
        assert must in syms


def test_library_loads_and_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(str(built_lib))
    for s in declared_symbols():
        assert hasattr(lib, s), f"libfw25.so does not export {s}"


def test_python_binding_covers_the_header(built_lib):
    assert set(engine.EXPORTS) | set(mapgen.SIGS) == set(declared_symbols())
    assert engine.lib().fw25_abi_version() == 6
    assert engine.lib().fw25_pitch(1241) == 1248


def test_struct_layout_matches_header(built_lib):
    # 8 int32 + 2 float, 15 pointers, 3 x (int32 + pad + pointers), 3 int32 + pad, 4 pointers, aniso, out_box
    assert ctypes.sizeof(engine.CProblem) == 8 * 4 + 2 * 4 + 15 * 8 + (8 + 16) + (8 + 8) + (8 + 8) + 16 + 32 + 8 + 8
    assert ctypes.sizeof(engine.CAniso) == 2 * (3 + 6 + 6) * 8
    assert ctypes.sizeof(engine.CSlab) == 16
    assert ctypes.sizeof(engine.CStats) == 8 * 8 + 2 * 4
    # fw25_medium: 8 int32, 2 double, 3 + 3 + 10 + 2 + 3 pointers, 2 int32, 4 double, 1 pointer, 2 int32
    assert ctypes.sizeof(mapgen.CMedium) == 8 * 4 + 2 * 8 + 21 * 8 + 2 * 4 + 4 * 8 + 8 + 4 * 4
