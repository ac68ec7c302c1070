"""CPU oracle for the RANGE / RANGE+ embedding hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU (torch fp64/fp32 + numpy, exactly where the reference uses each), the
algorithm of mvrl/RANGE's `load_model(...)` / `model(locs)` path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import it; the
product (`range_b200/`) never does and fails loudly without its CUDA extension.

Pinned: the reference is pure Python, so it was imported unmodified in the build container
(tests/golden/make_golden.py: regenerated `spherical_harmonics_ylm.py` with the reference's own
generator, stubbed the absent third-party imports, fabricated a random-init SatCLIP-L40 checkpoint and a
synthetic DB) and this restatement agrees with it to 2.3e-16 (location columns) / fp32 rounding
(retrieved columns); the reference's inputs/outputs are committed as tests/golden/*.npz and
tests/test_oracle.py re-checks the oracle against them on every run.  The reference ships no tests or
golden vectors of its own.

Reference lines followed (paths relative to /root/reference/range):
  * spherical harmonics  location_models/satclip/positional_encoding/spherical_harmonics.py:27-42 and the
    generated closed forms (generator spherical_harmonics_generate_ylms.py:19-42)
  * SIREN                location_models/satclip/location_encoder.py:98-151
  * DB preparation       range.py:78-100 and utils/utils.py:11-16
  * forward              range.py:206-242
"""
import math

import numpy as np
import torch

SEM_TEMP = {"RANGE": 15.0, "RANGE+": 12.0}   # range.py:103,108
GEO_TEMP = 40.0                              # range.py:109


# ------------------------------------------------------------------------------------------------
# spherical harmonics (spherical_harmonics.py:27-42 + generated Yl{l}_m{m})
# ------------------------------------------------------------------------------------------------
def sh_analytic(lonlat, L, entries):
    """lonlat (N,2) float64 tensor (lon, lat) degrees -> (N, L*L) float64.

    `entries[(l, am)] = (pref, {power: coeff})` are the generated closed forms (range_b200.sh_table
    .load_entries).  Evaluated the way the generated code is: every power through `**`, terms summed left
    to right from the highest power, factors multiplied left to right
    ``pref * (1 - c**2)**(am/2) * POLY * trig``; m == 0 functions are bare polynomials.
    """
    lonlat = torch.as_tensor(lonlat, dtype=torch.float64)
    lon, lat = lonlat[:, 0], lonlat[:, 1]
    phi = torch.deg2rad(lon + 180)          # spherical_harmonics.py:31
    theta = torch.deg2rad(lat + 90)         # :32
    c = torch.cos(theta)
    one_minus = 1.0 - c ** 2
    Y = torch.empty(lonlat.shape[0], L * L, dtype=torch.float64)
    for l in range(L):
        for am in range(l + 1):
            pref, poly = entries[(l, am)]
            acc = None
            for k in sorted(poly, reverse=True):
                if k == 0:
                    term = torch.full_like(c, poly[k])
                elif k == 1:
                    term = poly[k] * c
                else:
                    term = poly[k] * c ** k
                acc = term if acc is None else acc + term
            if am == 0:
                Y[:, l * l + l] = acc
                continue
            base = pref * one_minus ** (am / 2.0)
            legendre = base * acc
            Y[:, l * l + l + am] = legendre * torch.cos(am * phi)
            Y[:, l * l + l - am] = legendre * torch.sin(am * phi)
    return Y


def closed_form_norms(L):
    """(sqrt(2) *) SH_renormalization(l, m) exactly as spherical_harmonics_closed_form.py:28-40 computes them in Python
    floats; returned |m|-major: for am in range(L): for l in range(am, L)"""
    out = []
    for am in range(L):
        for l in range(am, L):
            renorm = math.sqrt((2.0 * l + 1.0) * math.factorial(l - am) / (4 * math.pi * math.factorial(l + am)))
            out.append(renorm if am == 0 else math.sqrt(2.0) * renorm)
    return np.asarray(out, np.float64)


def sh_closed_form(lonlat, L):
    """harmonics_calculation='closed-form': spherical_harmonics.py:27-42 with spherical_harmonics_closed_form.py:8-40
    (unnormalised associated-Legendre recurrence WITH Condon-Shortley phase, orthonormal m = 0), same operation order."""
    lonlat = torch.as_tensor(lonlat, dtype=torch.float64)
    phi = torch.deg2rad(lonlat[:, 0] + 180)
    theta = torch.deg2rad(lonlat[:, 1] + 90)
    x = torch.cos(theta)
    Y = torch.empty(lonlat.shape[0], L * L, dtype=torch.float64)

    def legendre(l, m):                                      # closed_form.py:8-26
        pmm = torch.ones_like(x)
        if m > 0:
            somx2 = torch.sqrt((1 - x) * (1 + x))
            fact = 1.0
            for _ in range(1, m + 1):
                pmm = pmm * (-fact) * somx2
                fact += 2.0
        if l == m:
            return pmm
        pmmp1 = x * (2.0 * m + 1.0) * pmm
        if l == m + 1:
            return pmmp1
        pll = torch.zeros_like(x)
        for ll in range(m + 2, l + 1):
            pll = ((2.0 * ll - 1.0) * x * pmmp1 - (ll + m - 1.0) * pmm) / (ll - m)
            pmm = pmmp1
            pmmp1 = pll
        return pll

    def renorm(l, m):                                        # :28-30
        return math.sqrt((2.0 * l + 1.0) * math.factorial(l - m) / (4 * math.pi * math.factorial(l + m)))

    for l in range(L):
        for m in range(-l, l + 1):
            if m == 0:
                y = renorm(l, 0) * legendre(l, 0)
            elif m > 0:
                y = math.sqrt(2.0) * renorm(l, m) * torch.cos(m * phi) * legendre(l, m)
            else:
                y = math.sqrt(2.0) * renorm(l, -m) * torch.sin(-m * phi) * legendre(l, -m)
            Y[:, l * l + l + m] = y
    return Y


def sh_exact(lonlat, L):
    """Mathematically exact harmonics in the SAME convention (stable recurrence, fp64): the accuracy
    yardstick that shows how far the reference's 15-digit polynomials are from the true functions."""
    lonlat = np.asarray(lonlat, np.float64)
    phi = np.deg2rad(lonlat[:, 0] + 180)
    theta = np.deg2rad(lonlat[:, 1] + 90)
    c, s = np.cos(theta), np.sin(theta)
    N = len(c)
    Y = np.empty((N, L * L))
    # fully normalised recurrence without Condon-Shortley phase
    pmm = np.full(N, math.sqrt(1.0 / (4 * math.pi)))
    for m in range(L):
        if m > 0:
            pmm = pmm * s * math.sqrt((2 * m + 1) / (2 * m))
        p_prev, p_cur = None, pmm
        for l in range(m, L):
            if l == m:
                p = pmm
            elif l == m + 1:
                p = math.sqrt(2 * m + 3) * c * pmm
            else:
                a = math.sqrt((4 * l * l - 1) / (l * l - m * m))
                b = math.sqrt(((l - 1) ** 2 - m * m) / (4 * (l - 1) ** 2 - 1))
                p = a * (c * p_cur - b * p_prev)
            if l > m:
                p_prev, p_cur = p_cur, p
            else:
                p_prev, p_cur = None, p
            if m == 0:
                Y[:, l * l + l] = p * math.pi        # sqrt((2l+1)/4*pi) == pi * sqrt((2l+1)/(4 pi))
            else:
                Y[:, l * l + l + m] = math.sqrt(2) * p * np.cos(m * phi)
                Y[:, l * l + l - m] = math.sqrt(2) * p * np.sin(m * phi)
    return Y


# ------------------------------------------------------------------------------------------------
# SIREN (location_encoder.py:98-151): sin(30 (Y W0^T + b0)) -> sin(h W1^T + b1) ... -> last linear
# ------------------------------------------------------------------------------------------------
def siren(Y, weights, w0_initial=30.0, w0=1.0):
    """weights: list of (W, b) float64 tensors; all but the last are followed by sin(w0 * .)"""
    x = torch.as_tensor(Y, dtype=torch.float64)
    n = len(weights)
    for i, (W, b) in enumerate(weights):
        x = torch.nn.functional.linear(x, W, b)                 # location_encoder.py:147
        if i < n - 1:
            x = torch.sin((w0_initial if i == 0 else w0) * x)   # :119
    return x


def rad_to_cart(locations):
    """utils/utils.py:11-16 (numpy, dtype-preserving)"""
    x = np.cos(locations[:, 1]) * np.cos(locations[:, 0])
    y = np.cos(locations[:, 1]) * np.sin(locations[:, 0])
    z = np.sin(locations[:, 1])
    return np.stack([x, y, z], axis=1)


def prepare_db(locs, satclip_embeddings, image_embeddings):
    """range.py:78-100.  Returns K (M,256) fp32 row-normalised, V (M,1024) fp32, xyz (M,3) fp32."""
    db_locs = locs.astype(np.float32)                                             # :79
    K = satclip_embeddings.astype(np.float32)                                     # :85
    K = K / np.linalg.norm(K, ord=2, axis=1, keepdims=True)                       # :89
    V = image_embeddings.astype(np.float32)                                       # :90
    xyz = rad_to_cart(db_locs * math.pi / 180)                                    # :93-95 (fp32)
    return K, V, xyz


class RangeOracle:
    """Restatement of range.py's LocationEncoder for 'RANGE' / 'RANGE+'."""

    def __init__(self, model_name, weights, entries, db, L=40, beta=0.5, exact=False, harmonics="analytic"):
        if model_name not in SEM_TEMP:
            raise ValueError("Unimplemented RANGE model")                        # range.py:114
        self.model_name, self.L, self.entries = model_name, L, entries
        self.harmonics = harmonics          # 'analytic' (generated closed forms) or 'closed-form' (recurrence)
        self.weights = [(torch.as_tensor(W, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64))
                        for W, b in weights]
        K, V, xyz = prepare_db(db["locs"], db["satclip_embeddings"], db["image_embeddings"])
        self.exact = exact          # fp64 retrieval: accuracy yardstick, not the reference's arithmetic
        dt = torch.float64 if exact else torch.float32
        self.K, self.V, self.xyz = (torch.tensor(a).to(dt) for a in (K, V, xyz))
        self.temp, self.geo_temp, self.beta = SEM_TEMP[model_name], GEO_TEMP, beta
        self.location_feature_dim = 1024 + 256                                    # :86

    def encode(self, coords):
        Y = (sh_analytic(coords, self.L, self.entries) if getattr(self, "harmonics", "analytic") == "analytic"
             else sh_closed_form(coords, self.L))
        e = siren(Y, self.weights)
        return e / e.norm(p=2, dim=-1, keepdim=True)                              # :212

    @torch.no_grad()
    def __call__(self, coords):
        coords = torch.as_tensor(coords, dtype=torch.float64)
        q = self.encode(coords)
        dt = self.K.dtype
        sim = q.to(dt) @ self.K.t()                                               # :213
        sim = torch.softmax(sim * self.temp, dim=-1)                              # :215
        hi = sim @ self.V                                                         # :217
        if self.model_name == "RANGE":
            return np.concatenate((hi.numpy(), q.numpy()), axis=1)               # :222
        xyz_q = torch.tensor(rad_to_cart(coords.numpy() * math.pi / 180))        # :225-229 (fp64)
        ang = xyz_q.to(dt) @ self.xyz.t()                                         # :231
        ang = torch.softmax(ang * self.geo_temp, dim=-1)                          # :234
        ahi = ang @ self.V                                                        # :236
        out = (1 - self.beta) * ahi + self.beta * hi                              # :238
        return np.concatenate((out.numpy(), q.numpy()), axis=1)                  # :240


# ------------------------------------------------------------------------------------------------
# synthetic inputs: seeded, no network
# ------------------------------------------------------------------------------------------------
def siren_init(L=40, H=512, n_hidden=2, out_dim=256, seed=0):
    """random-init SIREN with the reference's own init rule (location_encoder.py:137-144)"""
    g = torch.Generator().manual_seed(seed)
    dims = [L * L] + [H] * n_hidden + [out_dim]
    ws = []
    for i in range(len(dims) - 1):
        din, dout = dims[i], dims[i + 1]
        std = (1.0 / din) if i == 0 else math.sqrt(6.0 / din) / 1.0
        W = (torch.rand(dout, din, generator=g, dtype=torch.float64) * 2 - 1) * std
        b = (torch.rand(dout, generator=g, dtype=torch.float64) * 2 - 1) * std
        ws.append((W, b))
    return ws


def area_uniform(n, rng):
    lon = rng.uniform(-180, 180, n)
    lat = np.degrees(np.arcsin(rng.uniform(-1, 1, n)))
    return np.stack([lon, lat], 1)


def synthetic_db(M, seed=0, kind="iid", encoder=None):
    """`iid`: N(0,1) keys/values.  `structured`: keys near the encoder's own embedding of the location,
    values a noisy linear image of the key with non-zero mean (peaky softmax)."""
    rng = np.random.default_rng(seed)
    locs = area_uniform(M, rng)
    if kind == "iid":
        K = rng.standard_normal((M, 256))
        V = rng.standard_normal((M, 1024))
    else:
        E = encoder(torch.tensor(locs)).numpy()
        K = E + 0.1 * rng.standard_normal((M, 256))
        R = rng.standard_normal((256, 1024))
        V = K @ R / 16 + 0.5 + 0.2 * rng.standard_normal((M, 1024))
    return dict(locs=locs, satclip_embeddings=K, image_embeddings=V)
