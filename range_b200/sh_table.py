"""Coefficient table of the SatCLIP 'analytic' real spherical harmonics, derived from first principles.

The reference evaluates feature ``l*l + l + m`` (l < L, -l <= m <= l) through sympy-generated closed
polynomials (generator: range/location_models/satclip/positional_encoding/
spherical_harmonics_generate_ylms.py:19-42, consumer: .../spherical_harmonics.py:27-42).  Every one of
those functions has the shape

    Y_lm(theta, phi) = p_lm * (1 - c^2)^(|m|/2) * Q_lm(c) * T_m(phi),      c = cos(theta)

with ``T_0 = 1, T_{m>0} = cos(m phi), T_{m<0} = sin(|m| phi)`` and ``Q_lm`` the |m|-th derivative of the
Legendre polynomial P_l.  The generator prints every number with 15 significant digits
(``.evalf()``), and for l >= 26 that truncation - not fp64 rounding - dominates the reference's
deviation from the exact harmonics, so the reference's *function* is defined by these 15-digit
doubles.  This module re-derives them with exact rational arithmetic (it does NOT read the reference):

  * m != 0:  p = sqrt(2) * sqrt((2l+1)/(4 pi) * (l-|m|)!/(l+|m|)!)           (generate_ylms.py:23-35;
             the generator's explicit (-1)^m cancels sympy's Condon-Shortley phase), coefficients of
             Q are the exact rationals of d^m P_l / dx^m, each rounded to 15 digits;  when Q is a
             monomial its coefficient is folded into p before rounding (sympy flattens the product).
  * m == 0:  the generator's ``sqrt((2*l + 1) / 4 * pi)`` (operator precedence: pi in the numerator)
             is a 15-digit float that sympy distributes over P_l's coefficients, i.e. each printed
             coefficient is round15(round15(N_l) * round15(r_k)).

sympy's ``evalf`` is not always correctly rounded, so ~1.4 % of the literals derived here differ from
the generator's output by one unit in the 15th digit.  Because that digit matters (a 1e-15 relative
change of a 1e14-sized coefficient moves Y_39^0 by ~0.1 near the poles), the *product* table for
L <= 40 is the data file ``range_b200/data/sh_analytic_L40.npz`` - the numbers printed by the
reference's own generator, parsed by tools/make_sh_table.py - and this derivation is (a) the
documentation of what those numbers are, (b) the cross-check in tests/test_sh_table.py (every literal
within 2 units of the 15th digit), (c) the fallback for L > 40 (flagged ``derived``).
"""
import os
from fractions import Fraction
from math import comb, factorial

import mpmath
import numpy as np

_DPS = 15


def _round15(x):
    """What sympy's ``Float(x, 15)`` -> ``str`` -> ``float`` round trip yields for an mpmath number x:
    round to a 53-bit mantissa, print 15 significant digits, parse."""
    with mpmath.workprec(53):
        f = +mpmath.mpf(x)                       # round to 53 bits
        s = mpmath.nstr(f, _DPS, strip_zeros=False)
    return float(s)


def _legendre_coeffs(l):
    """exact coefficients {power: Fraction} of P_l(x)"""
    out = {}
    for k in range(l // 2 + 1):
        out[l - 2 * k] = Fraction((-1) ** k * comb(l, k) * comb(2 * l - 2 * k, l), 2 ** l)
    return out


def _derivative(coeffs, m):
    out = {}
    for p, a in coeffs.items():
        if p >= m:
            out[p - m] = a * (factorial(p) // factorial(p - m))
    return out


def _frac_mpf(fr):
    return mpmath.mpf(fr.numerator) / mpmath.mpf(fr.denominator)


def ylm_entry(l, am):
    """(prefactor, {power: coeff}) of the (l, |m|=am) function, as the 15-digit doubles the reference uses."""
    with mpmath.workprec(400):
        q = _derivative(_legendre_coeffs(l), am)
        if am == 0:
            # sqrt((2*l + 1) / 4 * pi): python float (2l+1)/4 (exact in binary), sympy pi
            n15 = _round15(mpmath.sqrt(mpmath.mpf(2 * l + 1) / 4 * mpmath.pi))
            coeffs = {}
            for p, a in q.items():
                a15 = _round15(_frac_mpf(a))
                coeffs[p] = _round15(mpmath.mpf(n15) * mpmath.mpf(a15))
            return 1.0, coeffs
        norm = mpmath.sqrt(2) * mpmath.sqrt(
            mpmath.mpf(2 * l + 1) / (4 * mpmath.pi)
            * mpmath.mpf(factorial(l - am)) / mpmath.mpf(factorial(l + am)))
        if len(q) == 1:
            (p, a), = q.items()
            return _round15(norm * _frac_mpf(a)), {p: 1.0}
        return _round15(norm), {p: _round15(_frac_mpf(a)) for p, a in q.items()}


def derive_entries(L=40):
    return {(l, am): ylm_entry(l, am) for am in range(L) for l in range(am, L)}


DATA_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "sh_analytic_L40.npz")


def load_entries(L=40):
    """entries {(l, am): (pref, {power: coeff})} for l < L.  L <= 40: the reference generator's own
    literals (shipped data file); L > 40: first-principles derivation."""
    if L <= 40:
        z = np.load(DATA_FILE)
        ls, ams, pref, off, powers, coefs = (z[k] for k in ("l", "am", "pref", "off", "power", "coef"))
        out = {}
        for i in range(len(ls)):
            if ls[i] < L:
                out[(int(ls[i]), int(ams[i]))] = (float(pref[i]), {int(k): float(a) for k, a in
                                                  zip(powers[off[i]:off[i + 1]], coefs[off[i]:off[i + 1]])})
        return out
    return derive_entries(L)


def build_table(L=40, entries=None):
    """Flat device table.

    Returns dict with
      pref   (L*(L+1)/2,)  fp64   prefactor of entry e(l,am), entries ordered am-major: for am: for l>=am
      off    (n_entries+1,) int32 start of each entry's coefficients in ``coef``
      coef   (n_coef,)     fp64   Horner coefficients in c^2, highest power first
      par    (n_entries,)  int32  parity of the polynomial (1 -> multiply by c once)
    Entry e evaluates  Q(c) = c^par * Horner(c^2; coef[off[e]:off[e+1]]).
    """
    if entries is None:
        entries = load_entries(L)
    pref, off, coef, par, index = [], [0], [], [], {}
    for am in range(L):
        for l in range(am, L):
            p, cs = entries[(l, am)]
            index[(l, am)] = len(pref)
            pref.append(p)
            parity = (l - am) & 1
            powers = sorted(cs, reverse=True)
            assert all((k & 1) == parity for k in powers)
            # dense in c^2 from the top power down to `parity`
            top = powers[0]
            for k in range(top, parity - 1, -2):
                coef.append(cs.get(k, 0.0))
            par.append(parity)
            off.append(len(coef))
    return dict(pref=np.asarray(pref, np.float64), off=np.asarray(off, np.int32),
                coef=np.asarray(coef, np.float64), par=np.asarray(par, np.int32), index=index, L=L)


def closed_form_norms(L):
    """Normalisation factors of harmonics_calculation='closed-form' (spherical_harmonics_closed_form.py:28-40), computed
    in Python floats exactly as the reference does: SH_renormalization(l, m) = sqrt((2l+1) (l-m)! / (4 pi (l+m)!)),
    times sqrt(2) for m != 0.  |m|-major order (for am in range(L): for l in range(am, L)), float64."""
    import math
    out = []
    for am in range(L):
        for l in range(am, L):
            renorm = math.sqrt((2.0 * l + 1.0) * math.factorial(l - am) / (4 * math.pi * math.factorial(l + am)))
            out.append(renorm if am == 0 else math.sqrt(2.0) * renorm)
    return np.asarray(out, np.float64)
