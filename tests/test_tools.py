"""CPU checks of the design tools that back statements in DESIGN.md."""
import importlib.util
import os


def _load(name):
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", name)
    spec = importlib.util.spec_from_file_location(name[:-3], path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_tile_major_swizzle_is_conflict_free(capsys):
    _load("swizzle_check.py").main()          # asserts internally
    assert capsys.readouterr().out.startswith("OK")


def test_swizzle_checker_detects_the_unswizzled_layout():
    m = _load("swizzle_check.py")
    # without the XOR, the eight state rows of a DMMA operand load share their banks
    addr = [(l >> 2) * 32 + (l & 3) for l in range(16)]
    assert m.half_warp_conflicts(addr)
