"""The table-driven exponential of brdf_model.cuh (BG_EXP_TABLE): the committed constants are the generator's, and the
algorithm -- emulated with exact rational arithmetic, one rounding per FMA as on the device -- stays within 2.2e-16
(2 ulp) of a 60-digit exponential over the fast path's range |y| < 700."""
import importlib.util
import os
import re
import struct

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _generator():
    spec = importlib.util.spec_from_file_location("make_exp_table", os.path.join(ROOT, "profiles", "tools", "make_exp_table.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _committed():
    return open(os.path.join(ROOT, "brdf_b200", "csrc", "exp_table.inc")).read()


def test_committed_table_is_the_generators():
    g, inc = _generator(), _committed()
    words = [int(w, 16) for w in re.findall(r"0x([0-9a-f]{16})ull", inc)]
    assert len(words) == g.N == 64
    for i, w in enumerate(words):
        # stored with (i << 14) taken off the high word: adding it back gives the correctly rounded 2^(i/64)
        value = struct.unpack("<d", struct.pack("<Q", w + (i << (14 + 32))))[0]
        assert value == g.TAB[i], i
        assert 1.0 <= value < 2.0
    red = re.search(r"kExpTabRed\[4\] = \{([^}]*)\}", inc).group(1).split(",")
    assert [float.fromhex(v.strip()) for v in red] == [g.INV, g.MAGIC, -g.HI, -g.LO]
    poly = re.search(r"kExpTabPoly\[4\] = \{([^}]*)\}", inc).group(1).split(",")
    assert [float.fromhex(v.strip()) for v in poly] == g.POLY


def test_reduction_constant_is_exact_for_every_index():
    """j * hi must be exact in double for every j the fast path can see (|y| < 700 -> |j| < 2^17): hi has 33 significant bits."""
    g = _generator()
    m, e = g.HI.hex().split("p")
    mant = int(m.replace("0x1.", ""), 16)
    assert mant & ((1 << 20) - 1) == 0          # 52 - 20 = 32 fraction bits + the leading one
    assert 700.0 * g.INV < 2 ** 17


def test_emulated_algorithm_stays_within_two_ulp():
    g = _generator()
    assert g.check(1500) < 2.2e-16
    assert g.exp_tab(0.0) == 1.0                # pow(t, 0) == 1 exactly on the fast path
