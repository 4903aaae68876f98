"""The algebra behind the scattering-density kernel (csrc/kernel_density.cu), checked on the CPU in
fp64 against the oracle's literal restatement of ComputeScatteringDensity (functions.glsl:1163-1260).

The kernel does not do the reference's 4-D lookups. It relies on three facts:
  1. every lookup of the pass happens at the output texel's own r and mu_s, so the 4-D fetch
     degenerates to a bilinear fetch in (mu, nu) at a fixed (layer k, column i_mu_s), and the mu
     footprint depends on (k, theta) only;
  2. the nu interpolation is L(x) = V[0] + sum_s (V[s+1] - V[s]) sat(x - s), x = (nu1 + 1)(NU - 1)/2:
     linear in the table values, so the per-direction work is a set of channel-independent weights;
  3. the cosine of the sun zenith angle at the ground point is affine in nu1:
     (r mu_s + d_ground nu1) / bottom, so the ground irradiance is a piecewise-linear ramp sum too.
  4. omega_s is a unit vector, so over the 32 azimuths of one polar direction x stays inside
     x0 +- A, x0 = (mu_s cos(theta) + 1)(NU - 1)/2, A = sin(theta) sqrt(1 - mu_s^2)(NU - 1)/2 -- an
     interval shared by every texel of a (layer, mu_s column) block. Ramps below it are identically 1
     and telescope into the base value V[lo]; ramps above it are identically 0: the kernel stages the
     rows rebased at lo = floor(x0 - A) and sweeps only the live ramps.
This test rebuilds one texel of the density table from those statements, in plain numpy, and
compares it with the oracle (which loops over directions and calls the literal GetScattering /
GetIrradiance / GetTransmittance). No GPU, no kernel code: it pins the formulation, not the port.
"""
import math

import numpy as np
import pytest

import precomputed_atmospheric_scattering_b200 as pas

SIZES = dict(t_w=64, t_h=16, r=8, mu=32, mu_s=8, nu=8, e_w=16, e_h=4)


@pytest.fixture(scope="module")
def world(orc):
    spec = pas.small_planet()
    cp = pas.channel_params(spec, [680.0, 550.0, 440.0])
    o = orc.Oracle(cp, orc.Sizes(**SIZES))
    T = o.transmittance()
    dE1 = o.direct_irradiance(T)
    dR, dM = o.single_scattering(T)
    dJ2 = o.scattering_density(T, dR, dM, dR, dE1, 2)
    dE2 = o.indirect_irradiance(dR, dM, dR, 1)
    dS2, _ = o.multiple_scattering(T, dJ2)
    return o, cp, T, dE1, dR, dM, dE2, dS2


def sat(x):
    return np.clip(x, 0.0, 1.0)


def ramp_interp(V, x):
    """Statement 2: V [channels, NU] at uniform knots, x in knot units (any real: clamped like the
    reference's texel fetch)."""
    D = np.diff(V, axis=1)
    return V[:, 0] + (D * sat(x - np.arange(V.shape[1] - 1))[None, :]).sum(axis=1)


def tap(x, n):
    i = math.floor(x)
    return min(max(i, 0), n - 1), min(max(i + 1, 0), n - 1), x - i


def density_texel(o, cp, T, tab_R, tab_M, tab_S, dE, order, k, j, i_nu, i_mu_s):
    R, MU, MUS, NU, EW = SIZES["r"], SIZES["mu"], SIZES["mu_s"], SIZES["nu"], SIZES["e_w"]
    x, y, z = i_nu * MUS + i_mu_s + 0.5, j + 0.5, k + 0.5
    r, mu, mu_s, nu, _ = o.rmumusnu_from_frag_coord(x, y, z)
    bottom, g = cp.bottom_radius, cp.mie_phase_function_g
    pR = lambda c: 3.0 / (16.0 * math.pi) * (1.0 + c * c)
    kM = 3.0 / (8.0 * math.pi) * (1.0 - g * g) / (2.0 + g * g)
    pM = lambda c: kM * (1.0 + c * c) / (1.0 + g * g - 2.0 * g * c) ** 1.5
    # omega = (sqrt(1 - mu^2), 0, mu), omega_s = (sx, sy, mu_s) (functions.glsl:1181-1185)
    wx = math.sqrt(1.0 - mu * mu)
    sx = 0.0 if wx == 0.0 else (nu - mu * mu_s) / wx
    sy = math.sqrt(max(1.0 - sx * sx - mu_s * mu_s, 0.0))
    beta_R = np.asarray(cp.rayleigh_scattering) * o.profile_density(0, r - bottom)
    beta_M = np.asarray(cp.mie_scattering) * o.profile_density(1, r - bottom)
    albedo = np.asarray(cp.ground_albedo)
    shape = (len(albedo), R, MU, NU, MUS)
    A_R, A_M, A_S = (t.reshape(shape) for t in (tab_R, tab_M, tab_S))
    E0 = dE[:, 0, :]                                    # statement 3 needs row 0 (r = bottom) only
    out = np.zeros(len(albedo))
    dtheta = dphi = math.pi / 16
    for l in range(16):
        theta = (l + 0.5) * dtheta
        ct, st = math.cos(theta), math.sin(theta)
        hit = bool(o.ray_intersects_ground(r, ct))
        # statement 1: mu footprint of (r, cos theta) -- shared by every texel of the layer
        u = o.scattering_uvwz_from_rmumusnu(r, ct, mu_s, 0.0, hit)
        j0, j1, wj = tap(u[2] * MU - 0.5, MU)
        row = lambda A: (1.0 - wj) * A[:, k, j0, :, i_mu_s] + wj * A[:, k, j1, :, i_mu_s]   # [c, NU]
        V_R, V_M, V_S = row(A_R), row(A_M), row(A_S)
        if hit:
            d_g = o.distance_to_bottom(r, ct)
            t_ground = np.array(o.get_transmittance(T, r, ct, d_g, True))
        for m in range(32):
            phi = (m + 0.5) * dphi
            wi = (math.cos(phi) * st, math.sin(phi) * st, ct)
            nu1 = sx * wi[0] + sy * wi[1] + mu_s * wi[2]
            nu2 = wx * wi[0] + mu * wi[2]
            xk = (nu1 + 1.0) * 0.5 * (NU - 1)
            if order == 2:
                incident = ramp_interp(V_R, xk) * pR(nu1) + ramp_interp(V_M, xk) * pM(nu1)
            else:
                incident = ramp_interp(V_S, xk)
            if hit:
                cos_ground = (r * mu_s + d_g * nu1) / bottom                       # statement 3
                xe = (cos_ground * 0.5 + 0.5) * (EW - 1)
                incident = incident + t_ground * albedo / math.pi * ramp_interp(E0, xe)
            out += incident * (beta_R * pR(nu2) + beta_M * pM(nu2)) * (dtheta * dphi * st)
    want = np.array(o.scattering_density_point(T, tab_R, tab_M, tab_S, dE, r, mu, mu_s, nu, order))
    return out, want


@pytest.mark.parametrize("order", [2, 3])
@pytest.mark.parametrize("texel", [(0, 3, 1, 2), (3, 20, 6, 5), (7, 31, 0, 0), (5, 15, 7, 7), (1, 16, 4, 3)])
def test_degenerate_lookups_and_ramp_sums_reproduce_the_literal_integral(world, order, texel):
    o, cp, T, dE1, dR, dM, dE2, dS2 = world
    dE = dE1 if order == 2 else dE2
    got, want = density_texel(o, cp, T, dR, dM, dS2, dE, order, *texel)
    # the literal 4-D fetch touches the neighbouring layer / column with weights of ~1e-12 (the r and
    # mu_s of a texel centre map back to texel coordinates up to rounding): nothing else differs
    assert np.allclose(got, want, rtol=1e-8, atol=1e-300), (got, want)


def test_ramp_sum_is_the_clamped_linear_interpolation():
    rng = np.random.default_rng(1)
    V = rng.uniform(0.1, 2.0, size=(3, 8))
    for x in np.concatenate([rng.uniform(-1.0, 8.0, 200), np.arange(0.0, 8.0)]):
        xc = min(max(x, 0.0), 7.0)
        i = min(int(math.floor(xc)), 6)
        want = V[:, i] + (xc - i) * (V[:, i + 1] - V[:, i])
        assert np.allclose(ramp_interp(V, x), want, rtol=1e-13)


def window_of(mu_s, theta, nu_n, margin=1e-3):
    """Statement 4, as csrc/kernel_density.cuh computes it per (block, direction): first knot and the
    number of live ramps."""
    sc = 0.5 * (nu_n - 1)
    x0 = (mu_s * math.cos(theta) + 1.0) * sc
    a = sc * math.sin(theta) * math.sqrt(max(1.0 - mu_s * mu_s, 0.0)) + margin
    lo = min(max(math.floor(x0 - a), 0), nu_n - 2)
    hi = min(max(math.ceil(x0 + a), lo + 1), nu_n - 1)
    return lo, hi


def test_ramp_window_is_block_uniform_and_the_rebased_sum_is_exact(world):
    o = world[0]
    R, MU, MUS, NU = SIZES["r"], SIZES["mu"], SIZES["mu_s"], SIZES["nu"]
    rng = np.random.default_rng(3)
    V = rng.uniform(0.1, 2.0, size=(3, NU))
    D = np.diff(V, axis=1)
    live_counts = []
    for k in (0, 3, R - 1):
        for i_mu_s in range(MUS):
            mu_s_col = None
            for l in range(16):
                theta = (l + 0.5) * math.pi / 16
                ct, st = math.cos(theta), math.sin(theta)
                xs = []
                for j in range(0, MU, 3):
                    for i_nu in range(NU):
                        r, mu, mu_s, nu, _ = o.rmumusnu_from_frag_coord(i_nu * MUS + i_mu_s + 0.5, j + 0.5, k + 0.5)
                        mu_s_col = mu_s if mu_s_col is None else mu_s_col
                        assert mu_s == mu_s_col                      # one mu_s per block
                        wx = math.sqrt(1.0 - mu * mu)
                        sx = 0.0 if wx == 0.0 else (nu - mu * mu_s) / wx
                        sy = math.sqrt(max(1.0 - sx * sx - mu_s * mu_s, 0.0))
                        for m in range(32):
                            phi = (m + 0.5) * math.pi / 16
                            nu1 = sx * math.cos(phi) * st + sy * math.sin(phi) * st + mu_s * ct
                            xs.append((nu1 + 1.0) * 0.5 * (NU - 1))
                xs = np.asarray(xs)
                lo, hi = window_of(mu_s_col, theta, NU)
                live_counts.append(hi - lo)
                # every x of the block lies in [lo, hi] up to the clamping of the lookup itself
                xc = np.clip(xs, 0.0, NU - 1.0)
                assert xc.min() >= lo - 1e-9 and xc.max() <= hi + 1e-9, (k, i_mu_s, l, xs.min(), xs.max(), lo, hi)
                # rebased ramp sum: base V[lo] + live ramps only == the full ramp sum
                for x in xs[::97]:
                    full = ramp_interp(V, x)
                    reb = V[:, lo] + (D[:, lo:hi] * sat(x - lo - np.arange(hi - lo))[None, :]).sum(axis=1)
                    assert np.allclose(reb, full, rtol=1e-13), (x, lo, hi)
    assert 1 <= min(live_counts) and max(live_counts) == NU - 1
    assert np.mean(live_counts) < NU - 1.5       # the window does cut work (5.0 of 7 at the reference's sizes)
