"""Parse the sympy-generated `spherical_harmonics_ylm.py` (output of the reference's own
generator, range/location_models/satclip/positional_encoding/spherical_harmonics_generate_ylms.py:37-42)
into a numeric table {(l,m): (prefactor, exponent_of_(1-c^2), {power: coeff}, trig)}.

Used only to build golden fixtures (tests/golden/) and to validate range_b200/sh_table.py.
"""
import re, sys, ast

FUNC = re.compile(r"def Yl(\d+)_m(_minus_)?(\d+)\(theta, phi\):\n    return (.*)")

def _split_top(expr, seps):
    """split expr on top-level (paren depth 0) occurrences of any char in seps; keeps the separator
    as prefix of the following chunk when it is +/-"""
    out, depth, cur = [], 0, ""
    i = 0
    while i < len(expr):
        ch = expr[i]
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if depth == 0 and ch in seps:
            # '**' is not a '*' separator; 'e+39' / 'e-31' are not +/- separators
            if ch == "*" and (expr[i:i+2] == "**" or (i > 0 and expr[i-1] == "*")):
                cur += ch; i += 1; continue
            if ch in "+-" and i > 0 and expr[i-1] in "eE" and expr[i-2].isdigit():
                cur += ch; i += 1; continue
            if ch in "+-":
                if cur.strip():
                    out.append(cur.strip())
                cur = ch
            else:
                out.append(cur.strip()); cur = ""
        else:
            cur += ch
        i += 1
    if cur.strip():
        out.append(cur.strip())
    return out

def parse_poly(expr):
    """expr: sum of terms  [+-] a*cos(theta)**k | a*cos(theta) | a  -> {k: a}"""
    coeffs = {}
    for term in _split_top(expr, "+-"):
        t = term.replace(" ", "")
        sign = 1.0
        if t[0] == "-":
            sign, t = -1.0, t[1:]
        elif t[0] == "+":
            t = t[1:]
        factors = _split_top(t, "*")
        a, k = 1.0, 0
        for f in factors:
            if f.startswith("cos(theta)"):
                k += int(f[len("cos(theta)**"):]) if "**" in f else 1
            else:
                a *= float(f)
        assert k not in coeffs
        coeffs[k] = sign * a
    return coeffs

def parse_body(body):
    """returns (prefactor, exponent, {k: a}, trig) with trig in (None, ('cos', m), ('sin', m))"""
    body = body.strip()
    top_terms = _split_top(body, "+-")
    if len(top_terms) > 1 or "phi" not in body and "(1.0" not in body:
        # bare polynomial in cos(theta) (m == 0) or a constant / p*c
        return 1.0, 0.0, parse_poly(body), None
    factors = _split_top(body, "*")
    pref, expo, poly, trig, extra_c = 1.0, 0.0, None, None, 0
    for f in factors:
        if f.startswith("(1.0 - cos(theta)**2)"):
            rest = f[len("(1.0 - cos(theta)**2)"):]
            expo = float(rest[2:]) if rest.startswith("**") else 1.0
        elif f.startswith("cos(theta)"):
            extra_c += int(f[len("cos(theta)**"):]) if "**" in f else 1
        elif f.startswith("cos(") or f.startswith("sin("):
            arg = f[4:-1]
            m = 1 if arg == "phi" else int(arg[:-len("*phi")])
            trig = (f[:3], m)
        elif f.startswith("("):
            poly = parse_poly(f[1:-1])
        else:
            pref *= float(f)
    if poly is None:
        poly = {extra_c: 1.0}
    else:
        assert extra_c == 0
    return pref, expo, poly, trig

def parse_file(path):
    src = open(path).read()
    table = {}
    for mt in FUNC.finditer(src):
        l = int(mt.group(1)); m = int(mt.group(3)) * (-1 if mt.group(2) else 1)
        table[(l, m)] = parse_body(mt.group(4))
    return table

if __name__ == "__main__":
    tab = parse_file(sys.argv[1])
    print(len(tab), "functions")
    nlit = sum(len(v[2]) + 1 for v in tab.values())
    print("literals", nlit)
    for key in [(0,0),(1,-1),(2,1),(5,3),(5,-3),(4,0),(3,2),(3,3),(39,39),(1,0)]:
        print(key, tab[key])
