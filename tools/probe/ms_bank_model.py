"""CPU model of the shared-memory wavefronts of the gathers in the ray-march kernels (no GPU needed).

The L1/shared data pipe serves a 128-bit warp load as four quarter warps of 8 lanes; a quarter warp
costs as many wavefronts as the largest number of DIFFERENT 16-byte vectors it reads from one bank
group (8 groups of 4 banks), lanes reading the same vector share it. With that rule and the
geometry of the Earth bench workload (256 x 128 x 32 scattering table, 8 x 32 rows, 51 samples per
ray) the model gives 5.23 wavefronts per gather load for the round-1 layout of
multiple_scattering_rows_kernel; ncu measured 1.27 conflict wavefronts per gather load on top of 4
(profiles/r2_ncu_full_census.csv), i.e. the same. It was then used to pick the layout of
kernel_raymarch.cu (parity split + odd slab pitch + lane predicate) before spending GPU time:

    python tools/probe/ms_bank_model.py            # multiple scattering: layouts side by side
    python tools/probe/ms_bank_model.py --single   # single scattering (sun lookups)

Numbers are wavefronts per 128-bit gather instruction (4 = no conflict and no broadcast).
"""
import math
import sys

import numpy as np

BOTTOM, TOP = 6360.0, 6420.0
H = math.sqrt(TOP * TOP - BOTTOM * BOTTOM)
R_N, MU_N, MUS_N, NU_N, NS, T_W = 32, 128, 32, 8, 50, 256
MU_S_MIN = math.cos(102 / 180 * math.pi)


def dist_top(r, mu):
    return np.maximum(-r * mu + np.sqrt(np.maximum(r * r * (mu * mu - 1) + TOP * TOP, 0)), 0)


def dist_bottom(r, mu):
    return np.maximum(-r * mu - np.sqrt(np.maximum(r * r * (mu * mu - 1) + BOTTOM * BOTTOM, 0)), 0)


D_MIN, D_MAX = TOP - BOTTOM, H
A = (dist_top(BOTTOM, MU_S_MIN) - D_MIN) / (D_MAX - D_MIN)


def mu_s_columns():
    x = np.arange(MUS_N) / (MUS_N - 1)
    a = (A - x * A) / (1 + x * A)
    d = D_MIN + np.minimum(a, A) * (D_MAX - D_MIN)
    return np.where(d == 0, 1.0, np.clip((H * H - d * d) / (2 * BOTTOM * d), -1, 1))


MUS_COL = mu_s_columns()
NUS = np.arange(NU_N) / (NU_N - 1) * 2 - 1


def x_mu_s(mu_s):
    """mu_s -> texel coordinate of the scattering table's mu_s axis."""
    d = dist_top(BOTTOM, mu_s)
    a = (d - D_MIN) / (D_MAX - D_MIN)
    return np.maximum(1 - a / A, 0) / (1 + a) * (MUS_N - 1)


def row_setup(k, j):
    """Ray of block (layer k, mu row j) and the nu taps of its 8 x 32 texels."""
    rho = H * k / (R_N - 1)
    r = math.sqrt(rho * rho + BOTTOM * BOTTOM)
    u = (j + 0.5) / MU_N
    half = MU_N / 2
    if j < MU_N // 2:
        x = ((1 - 2 * u) - 0.5 / half) / (1 - 1 / half)
        d = (r - BOTTOM) + (rho - (r - BOTTOM)) * x
        mu = -1.0 if d == 0 else min(max(-(rho * rho + d * d) / (2 * r * d), -1), 1)
        d_end = float(dist_bottom(r, mu))
    else:
        x = ((2 * u - 1) - 0.5 / half) / (1 - 1 / half)
        d = (TOP - r) + (rho + H - (TOP - r)) * x
        mu = 1.0 if d == 0 else min(max((H * H - rho * rho - d * d) / (2 * r * d), -1), 1)
        d_end = float(dist_top(r, mu))
    s = np.sqrt((1 - mu * mu) * (1 - MUS_COL * MUS_COL))
    nu = np.clip(NUS[:, None] * np.ones(MUS_N)[None, :], (mu * MUS_COL - s)[None, :], (mu * MUS_COL + s)[None, :])
    xn = (nu + 1) / 2 * (NU_N - 1)
    i0 = np.minimum(np.floor(xn), NU_N - 1).astype(int)
    w = xn - i0
    i1 = np.minimum(i0 + 1, NU_N - 1)
    on = (np.abs(w) < 1e-7) | (np.abs(w - 1) < 1e-7) | (i0 == i1)
    slab_a = np.where(on & (np.abs(w - 1) < 1e-7), i1, i0)
    return r, mu, d_end, nu, slab_a, i1, on


def wavefronts(pos, active=None):
    total = 0
    for q in range(4):
        t = pos[q * 8:(q + 1) * 8]
        if active is not None:
            t = t[active[q * 8:(q + 1) * 8]]
        if len(t):
            total += np.bincount(np.unique(t) % 8, minlength=8).max()
    return total


def on_slab_first(on):
    ids = np.arange(256)
    o = on.reshape(-1)
    return np.concatenate([ids[o], ids[~o]])


def multiple_scattering(rows, parity_split, slab_pitch, predicate):
    """Gathers of multiple_scattering_rows_kernel: texel pair (i0, i0 + 1) of one slab (lanes on a slab)
    or of two slabs; texels of a slab at positions slab * slab_pitch + ... inside a plane."""
    n_inst = n_wave = 0
    for k, j in rows:
        r, mu, d_end, nu, slab_a, slab_b, on = row_setup(k, j)
        if d_end <= 0:
            continue
        order = on_slab_first(on)
        inu, imus = order // MUS_N, order % MUS_N
        t_nu, t_mus, t_on = nu[inu, imus], MUS_COL[imus], on[inu, imus]
        s_a, s_b = slab_a[inu, imus], slab_b[inu, imus]
        for i in range(NS + 1):
            d = d_end * i / NS
            r_i = min(max(math.sqrt(d * d + 2 * r * mu * d + r * r), BOTTOM), TOP)
            xs = np.clip(x_mu_s(np.clip((r * t_mus + d * t_nu) / r_i, -1, 1)), 0, MUS_N - 1)
            i0 = np.minimum(np.floor(xs), MUS_N - 2).astype(int)
            if parity_split:
                taps = ((i0 + 1) >> 1, 100000 * 8 + (i0 >> 1))   # even texel, odd texel (second half of the plane)
            else:
                taps = (i0, i0 + 1)
            for w in range(8):
                sl = slice(w * 32, (w + 1) * 32)
                for t in taps:
                    n_inst += 1
                    n_wave += wavefronts(s_a[sl] * slab_pitch + t[sl])
                if not t_on[sl].all():
                    act = ~t_on[sl] if predicate else None
                    for t in taps:
                        n_inst += 1
                        n_wave += wavefronts(s_b[sl] * slab_pitch + t[sl], act)
    return n_wave / n_inst, n_wave


def single_scattering(rows, nu_lanes, parity_split):
    """Sun lookups of single_scattering_kernel: texel pair of the staged transmittance row at r_i."""
    n_inst = n_wave = 0
    tid = np.arange(256)
    x = (tid % NU_N) * MUS_N + tid // NU_N if nu_lanes else tid
    inu, imus = x // MUS_N, x % MUS_N
    for k, j in rows:
        r, mu, d_end, nu, _, _, _ = row_setup(k, j)
        if d_end <= 0:
            continue
        t_nu, t_mus = nu[inu, imus], MUS_COL[imus]
        for i in range(NS + 1):
            d = d_end * i / NS
            r_i = min(max(math.sqrt(d * d + 2 * r * mu * d + r * r), BOTTOM), TOP)
            rho_i = math.sqrt(max(r_i * r_i - BOTTOM * BOTTOM, 0))
            p = np.clip(r * t_mus + d * t_nu, -r_i, r_i)
            q = TOP * TOP - r_i * r_i
            s = np.sqrt(p * p + q)
            with np.errstate(invalid="ignore", divide="ignore"):
                dt = np.where(p > 0, q / (p + s), s - p)
            xt = np.clip((dt - (TOP - r_i)) / (rho_i + H - (TOP - r_i)) * (T_W - 1), 0, T_W - 1)
            i0 = np.minimum(np.floor(xt), T_W - 2).astype(int)
            taps = ((i0 + 1) >> 1, 100000 * 8 + (i0 >> 1)) if parity_split else (i0, i0 + 1)
            for w in range(8):
                sl = slice(w * 32, (w + 1) * 32)
                for t in taps:
                    n_inst += 1
                    n_wave += wavefronts(t[sl])
    return n_wave / n_inst, n_wave


if __name__ == "__main__":
    rows = [(k, j) for k in range(1, 32, 5) for j in range(3, 128, 9)]
    if "--single" in sys.argv:
        for name, args in (("columns on lanes, natural order", (False, False)),
                           ("nu on lanes, natural order (round 2)", (True, False)),
                           ("nu on lanes, parity split", (True, True))):
            print(f"{name:44s} {single_scattering(rows, *args)[0]:.3f}")
    else:
        base = None
        for name, args in (("natural order, slab pitch 32 (round 1)", (False, 32, False)),
                           ("parity split, slab pitch 16", (True, 16, False)),
                           ("parity split, slab pitch 16, lane predicate", (True, 16, True)),
                           ("parity split, slab pitch 17, lane predicate", (True, 17, True)),
                           ("parity split, slab pitch 18, lane predicate", (True, 18, True))):
            per, total = multiple_scattering(rows, *args)
            base = base or total
            print(f"{name:48s} {per:.3f} per load, {total / base:.3f} of the round-1 wavefronts")
