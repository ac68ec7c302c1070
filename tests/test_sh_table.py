"""The SH coefficient table: shipped data (reference generator's literals) vs first-principles derivation,
and the closed-form known answers of SURVEY.md section 8c."""
import hashlib
import math

import numpy as np

from range_b200 import sh_table


def test_known_answers(sh_entries):
    e = sh_entries
    assert e[(0, 0)] == (1.0, {0: 0.886226925452758})
    assert e[(1, 0)] == (1.0, {1: 1.53499006191973})
    assert e[(1, 1)] == (0.48860251190292, {0: 1.0})
    assert e[(2, 0)] == (1.0, {2: 2.97249547320451, 0: -0.990831824401503})
    assert e[(39, 39)] == (1.064079376195, {0: 1.0})
    assert e[(5, 3)] == (0.00931882475114763, {2: 472.5, 0: -52.5})
    # the m = 0 convention: sqrt((2l+1)/4*pi), i.e. pi x the orthonormal value (generate_ylms.py:29)
    assert abs(e[(0, 0)][1][0] - math.sqrt(math.pi / 4)) < 1e-15


def test_table_is_pinned():
    """sha256 of the shipped literals (regenerate with tools/make_sh_table.py if the generator changes)"""
    z = np.load(sh_table.DATA_FILE)
    h = hashlib.sha256()
    for k in ("l", "am", "pref", "off", "power", "coef"):
        h.update(np.ascontiguousarray(z[k]).tobytes())
    assert len(z["pref"]) == 820 and len(z["coef"]) == 5950
    assert h.hexdigest() == PINNED


def test_derivation_matches_generator_to_the_15th_digit(sh_entries):
    derived = sh_table.derive_entries(40)
    exact, total = 0, 0
    for key, (p, cs) in sh_entries.items():
        pd, cd = derived[key]
        assert set(cs) == set(cd)
        for k in cs:
            a, b = p * cs[k], pd * cd[k]
            total += 1
            exact += a == b
            assert abs(a - b) <= 2.5e-14 * abs(a), (key, k, a, b)   # <= 2 units of the 15th digit
    assert exact / total > 0.95


def test_flat_table_layout():
    t = sh_table.build_table(40)
    assert t["pref"].shape == (820,) and t["off"][-1] == len(t["coef"])
    # entry order is am-major; Horner in c^2 from the top power
    i = t["index"][(5, 3)]
    assert list(t["coef"][t["off"][i]:t["off"][i + 1]]) == [472.5, -52.5] and t["par"][i] == 0
    i = t["index"][(4, 0)]
    assert list(t["coef"][t["off"][i]:t["off"][i + 1]]) == [11.6317283965674, -9.97005291134353, 0.997005291134353]
    i = t["index"][(3, 2)]
    assert t["par"][i] == 1 and t["pref"][i] == 1.44530572132028
    # a smaller L is a prefix-compatible subset
    t10 = sh_table.build_table(10)
    assert t10["pref"].shape == (55,)


PINNED = "1be50abeccf00a83347e2d2c915b9011bd60e8325c65ad216771807c9bacb6b8"
