"""Build range_b200/data/sh_analytic_L40.npz from the output of the reference's OWN generator.

Runs range/location_models/satclip/positional_encoding/spherical_harmonics_generate_ylms.py (needs
sympy; ~30 s) up to l = 39, parses the printed functions (tools/parse_ylm.py) and stores the numeric
literals.  Needs /root/reference, so it runs in the build container only; its output is committed.

    python tools/make_sh_table.py [path/to/existing/spherical_harmonics_ylm.py]
"""
import os, subprocess, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
from parse_ylm import parse_file  # noqa: E402

GEN = "/root/reference/range/location_models/satclip/positional_encoding/spherical_harmonics_generate_ylms.py"
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def regenerate(path):
    """stream the generator's stdout, stop at l = 40 (it would loop to l = 100)"""
    os.makedirs(os.path.dirname(path), exist_ok=True)
    proc = subprocess.Popen([sys.executable, "-u", GEN], stdout=subprocess.PIPE, text=True)
    lines = []
    for line in proc.stdout:
        if line.startswith("def Yl40_m_minus_40"):
            break
        lines.append(line)
    proc.kill()
    # drop the dangling "@torch.jit.script" decorator of the function we cut
    while lines and lines[-1].strip() in ("", "@torch.jit.script"):
        lines.pop()
    with open(path, "w") as f:
        f.writelines(lines)
        f.write("\n")


def main():
    ylm = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REF_DIR, "spherical_harmonics_ylm.py")
    if not os.path.exists(ylm):
        regenerate(ylm)
    tab = parse_file(ylm)
    assert len(tab) == 1600
    ls, ams, pref, off, power, coef = [], [], [], [0], [], []
    for am in range(40):
        for l in range(am, 40):
            p, e, poly, trig = tab[(l, am)]
            if (l, am) == (2, 2):
                # sympy distributes 3*(1 - c^2): "0.182..*(3.0 - 3.0*cos(theta)**2)*cos(2*phi)"
                assert e == 0.0 and poly == {0: 3.0, 2: -3.0}
                e, poly = 1.0, {0: 3.0}
            assert e == am / 2
            assert trig == (None if am == 0 else ("cos", am))
            if am:
                # +m and -m share everything but the trig factor
                pn, en, polyn, trign = tab[(l, -am)]
                if (l, am) == (2, 2):
                    en, polyn = 1.0, {0: 3.0}
                assert (pn, en, polyn) == (p, e, poly) and trign == ("sin", am)
            ls.append(l); ams.append(am); pref.append(p)
            for k in sorted(poly, reverse=True):
                power.append(k); coef.append(poly[k])
            off.append(len(coef))
    out = os.path.join(ROOT, "range_b200", "data", "sh_analytic_L40.npz")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    np.savez_compressed(out, l=np.asarray(ls, np.int32), am=np.asarray(ams, np.int32),
                        pref=np.asarray(pref, np.float64), off=np.asarray(off, np.int32),
                        power=np.asarray(power, np.int32), coef=np.asarray(coef, np.float64))
    print("wrote", out, "entries", len(ls), "coefficients", len(coef), os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
