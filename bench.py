"""Headline benchmark: the full LUT precompute (atmosphere::Model::Init equivalent) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 1..5]

Default workload = BASELINE.json configs[1] (`--config 2`): Earth/demo atmosphere
(atmosphere/demo/demo.cc:188-284), precomputed-luminance mode with 15 wavelengths (5 RGB batches in
the reference's terms), 4 scattering orders, the reference's table sizes (T 256x64, E 64x16,
S 256x128x32), combined scattering textures, half-precision final 3-D tables -- the demo's own
settings. One "step" = one complete precompute of all tables. Metric (BASELINE.json): wall-clock ms
of that precompute; lower is better. The other BASELINE configs are `--config 1` (RGB), `3` (10
orders), `4` (4x tables per dimension) and `5` (64 atmospheres + 1080p renders); the driver's line
is the default.

  value / ms_per_step  device time of one Init, CUDA events on the library's own stream (start of
                       the first kernel to the end of the last), mean over K steps, max over ranks.
                       Nothing is resident beforehand except the model parameters; every table is
                       recomputed each step.
  e2e                  the same job through the reference-facing API with HOST buffers: construct
                       the Model from host arrays (the 19 constructor arguments; parameters reach the
                       GPU as kernel arguments), Init(4), the three product tables land in registered
                       host memory, destroy. Wall clock around the synchronous calls.
  parity               outside the timed region every rank checks the tables it has just timed
                       against the committed outputs of the unmodified reference (tests/golden/):
                       all 4096 rows of S through per-row digests, E and T texel by texel. The
                       process exits non-zero when a rank is out of tolerance.
  roofline             dominant kernel: EXECUTED fp32 flops per launch (SASS opcode census of the ncu
                       capture of the same kernels, profiles/ncu_hot_kernels.json) / live CUDA-event
                       time / the FP32 FMA peak measured on this device. This path is FP32-pipe or
                       L1-data-pipe bound, not HBM or tensor bound; the
                       canonical-work figure of SURVEY 8(d) is kept beside it as `canonical_frac`.
  cpu_baseline         the UNMODIFIED reference CPU model (oracle/_ref) on the host cores, bounded
                       row sample, extrapolated to the full job.

--impl reference times that CPU reference alone (rank 0 only under torchrun).
Multi-GPU (torchrun, one rank per GPU): r-slab sharding inside the library; the density kernel stores
its slab to every rank through the NVLS multicast address of a symmetric arena (fallbacks: unicast
CUDA IPC peer mappings, NCCL), flag barriers between orders; strong scaling (the job is fixed, N GPUs
share it).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

X4 = dict(transmittance_width=1024, transmittance_height=256, scattering_r=128, scattering_mu=512,
          scattering_mu_s=128, scattering_nu=8, irradiance_width=64, irradiance_height=16)
SIZES_TEXT = "T 256x64, E 64x16, S 256x128x32"
# BASELINE.json configs[0..4]
CONFIGS = {
    1: dict(wavelengths=3, orders=4, half=True, sizes=None,
            metric="lut_precompute_ms_4_orders_rgb",
            workload=f"earth-demo atmosphere, RGB (3 wavelengths), 4 scattering orders, {SIZES_TEXT}, "
                     "combined textures, half-precision final tables"),
    2: dict(wavelengths=15, orders=4, half=True, sizes=None,
            metric="lut_precompute_ms_4_orders_15_wavelengths",
            workload="earth-demo atmosphere, 15 precomputed wavelengths (5 RGB batches), 4 scattering orders, "
                     f"{SIZES_TEXT}, combined textures, half-precision final tables"),
    3: dict(wavelengths=15, orders=10, half=True, sizes=None,
            metric="lut_precompute_ms_10_orders_15_wavelengths",
            workload=f"earth-demo atmosphere, 15 precomputed wavelengths, 10 scattering orders, {SIZES_TEXT}, "
                     "combined textures, half-precision final tables"),
    4: dict(wavelengths=15, orders=4, half=False, sizes=X4,
            metric="lut_precompute_ms_4_orders_15_wavelengths_4x_tables",
            workload="earth-demo atmosphere, 15 precomputed wavelengths, 4 scattering orders, 4x tables per "
                     "dimension: T 1024x256, E 64x16, S 1024x512x128 ((nu, mu_s) = (8, 128)), combined "
                     "textures, fp32 final tables"),
    5: dict(wavelengths=3, orders=4, half=True, sizes=None,
            metric="ensemble_64_atmospheres_precompute_and_1080p_sky_radiance_ms",
            workload="64 earth atmospheres (4 turbidity x 4 ozone x 4 albedo, seeded sweep), RGB, 4 scattering "
                     f"orders each, {SIZES_TEXT}, 16 precomputations in flight, then one 1920x1080 GetSkyRadiance "
                     "image per atmosphere"),
}


def config_dict(cfg_id: int, world_size: int) -> dict:
    """The `config` object of the JSON line: identical in the B200 arm and in the reference arm."""
    c = CONFIGS[cfg_id]
    return {"workload": c["workload"], "baseline_config": cfg_id, "orders": c["orders"],
            "wavelengths": c["wavelengths"],
            "parallelism": "1 GPU" if world_size == 1 else (
                f"r-slabs over {world_size} GPUs, " + ("NCCL all-gather / all-reduce per order"
                if os.environ.get("PAS_EXCHANGE") == "nccl" else
                "density slabs stored to all ranks by the kernel " +
                ("into CUDA IPC peer mappings (unicast)" if os.environ.get("PAS_EXCHANGE") == "peer" else
                 "through the NVLS multicast address of a symmetric arena") + ", flag barriers")),
            "l2": "no flush between steps: every step recomputes and rewrites all tables "
                  "(5 x 60 MiB intermediates + products > 126 MB L2), nothing is reused across steps"}


# ---- canonical algorithmic work ---------------------------------
def canonical_work(C=15, K=4, n_t=256 * 64, n_e=64 * 16, n_s=256 * 128 * 32):
    """FP32 flops (FMA = 2) and MUFU ops per pass in the minimal formulation; per-launch figures."""
    w_t, w_1, w_d, w_i, w_m = n_t * 501, n_s * 51, n_s * 512, n_e * 1024, n_s * 51
    flops = {
        "transmittance": w_t * 38 + n_t * 6 * C,
        "single_scattering": w_1 * (80 + 22 * C) + n_s * 14 * C,
        "scattering_density_2": w_d * (47 + 25 * C),
        "scattering_density_n": w_d * (36 + 16 * C),
        "indirect_irradiance_2": w_i * (60 + 65 * C),
        "indirect_irradiance_n": w_i * (60 + 32 * C),
        "multiple_scattering": w_m * (80 + 41 * C) + n_s * 14 * C,
    }
    mufu = {
        "transmittance": w_t * 4 + n_t * C,
        "single_scattering": w_1 * (11 + C),
        "scattering_density_2": w_d * 2,
        "scattering_density_n": w_d * 1,
        "indirect_irradiance_2": w_i * 6,
        "indirect_irradiance_n": w_i * 6,
        "multiple_scattering": w_m * (9 + C),
    }
    total_f = (flops["transmittance"] + flops["single_scattering"] + flops["scattering_density_2"] +
               flops["indirect_irradiance_2"] + (K - 2) * (flops["scattering_density_n"] +
                                                           flops["indirect_irradiance_n"]) +
               (K - 1) * flops["multiple_scattering"])
    total_u = (mufu["transmittance"] + mufu["single_scattering"] + mufu["scattering_density_2"] +
               mufu["indirect_irradiance_2"] + (K - 2) * (mufu["scattering_density_n"] +
                                                          mufu["indirect_irradiance_n"]) +
               (K - 1) * mufu["multiple_scattering"])
    return flops, mufu, total_f, total_u


def phase_key(name: str) -> str:
    """Timing name ('scattering_density_3') -> canonical_work key."""
    for base in ("scattering_density", "indirect_irradiance"):
        if name.startswith(base):
            return base + ("_2" if name.endswith("_2") else "_n")
    for base in ("multiple_scattering", "single_scattering", "transmittance"):
        if name.startswith(base):
            return base
    return name


# ---- clocks ---------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons of one GPU every 5 ms while the timed region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.005)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ---- the reference CPU arm ------------------------------------------------------------------------
def _odd_stride(total_rows: int, want_rows: int) -> int:
    """A stride that visits ~want_rows rows and is odd (so it walks through every mu row j, the row
    index being k * 128 + j)."""
    s = max(1, total_rows // max(1, want_rows))
    return s if s % 2 == 1 else s + 1


class ReferenceCpu:
    """The unmodified reference CPU model (atmosphere/reference, compiled in place into
    oracle/_ref/libpas_ref.so) on a bounded sample of the workload: every pass of
    atmosphere/reference/model.cc:140-237 runs on a strided subset of its texel rows and the
    measured time is scaled by rows / rows sampled. The reference always computes its 47 spectral
    lanes in fp64 whatever the number of wavelengths asked for; the bench channels sit in lanes
    0..C-1. Its table sizes are compile-time constants (atmosphere/constants.h:47-61), so only the
    configs with the reference's sizes can run."""

    def __init__(self, cfg_id: int = 2):
        from oracle import ref
        import precomputed_atmospheric_scattering_b200 as pas
        c = CONFIGS[cfg_id]
        if c["sizes"] is not None or cfg_id == 5:
            raise RuntimeError("the reference CPU model has compile-time table sizes and no batch API: "
                               "configs 4 and 5 have no CPU arm")
        if not ref.available():
            raise RuntimeError("oracle/_ref/libpas_ref.so is missing (build() compiles it where "
                               "/root/reference exists; the prebuilt file travels to the GPU box)")
        self.threads = os.cpu_count() or 1
        self.orders, self.wavelengths = c["orders"], c["wavelengths"]
        spec = pas.earth(c["wavelengths"], half_precision=c["half"])
        cp = pas.channel_params(spec, pas.precomputed_wavelengths(c["wavelengths"]))
        self.model = ref.RefModel(cp, nthreads=self.threads)
        rows3 = 32 * 128
        self.stride_ray = _odd_stride(rows3, 16 * self.threads)
        self.stride_density = _odd_stride(rows3, 4 * self.threads)
        self.rows3 = rows3

    def sample(self) -> str:
        return (f"every pass on a strided row subset, scaled to the full tables: single/multiple "
                f"scattering 1 row in {self.stride_ray}, scattering density 1 row in "
                f"{self.stride_density} (of {self.rows3} rows of 256 texels), 2-D tables in full; "
                f"47 fp64 lanes (the reference's fixed spectrum width), {self.threads} threads")

    def step(self):
        """One bounded sample; returns (estimated full-job seconds, seconds actually spent)."""
        m = self.model
        est = spent = 0.0

        def run(name, order=0, stride=1):
            nonlocal est, spent
            t = m.phase(name, order, stride)
            rows = self.rows3
            done = len(range(0, rows, stride))
            spent += t
            est += t * (rows / done if stride > 1 else 1.0)

        run("transmittance")
        run("direct_irradiance")
        run("single_scattering", 0, self.stride_ray)
        for order in range(2, self.orders + 1):
            run("scattering_density", order, self.stride_density)
            run("indirect_irradiance", order)
            run("multiple_scattering", order, self.stride_ray)
        return est, spent

    def baseline(self, ms: float, spent_s: float) -> dict:
        """`cpu_baseline` object for an estimate of `ms` per full job."""
        return {"value": round(ms, 1), "unit": "ms", "cores": self.threads, "kind": "reference",
                "sample": self.sample(), "extrapolated": True,
                "measured_seconds_per_sample": round(spent_s, 2),
                "value_scaled_to_bench_lanes_of_47": round(ms * self.wavelengths / 47.0, 1)}


def run_reference(args, rank):
    if rank != 0:
        return
    c = CONFIGS[args.config]
    try:
        cpu = ReferenceCpu(args.config)
    except Exception as e:  # pragma: no cover
        print(json.dumps({"impl": "reference", "unavailable": str(e)}))
        return
    for _ in range(args.warmup):
        cpu.step()
    est, spent = [], []
    for _ in range(args.steps):
        e, s = cpu.step()
        est.append(e)
        spent.append(s)
    ms = 1e3 * sum(est) / len(est)
    line = {
        "impl": "reference", "metric": c["metric"], "value": ms, "unit": "ms", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.config, args.gpus),
        # the value is an ESTIMATE of the full job: each step runs a strided row sample and scales it
        "extrapolated": True,
        "measured_ms_per_step": 1e3 * sum(spent) / len(spent),
        # the reference computes 47 fp64 lanes whatever the wavelength count: like for like with the
        # C channels of the B200 arm
        "value_scaled_to_bench_lanes_of_47": ms * c["wavelengths"] / 47.0,
        "cpu_baseline": cpu.baseline(ms, sum(spent) / len(spent)),
        "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---- parity of what was timed ----------------------------------------------------------------------
def parity_check(cfg_id, S, E, T, L, half):
    """Tables of the timed model against the committed reference outputs (tests/golden/). Config 2:
    the luminance product through tests.parity.check_bench_product. Config 1 (radiance mode): every
    row of S against the per-channel reference digests, E / T texel by texel."""
    import numpy as np
    from tests import parity
    two, _, _ = parity.load_golden()
    rows = parity.load_rows()
    if cfg_id == 2:
        return parity.check_bench_product(S, E, T, L, two, rows, half_precision=half)
    lanes = slice(15, 18)
    S = np.asarray(S, dtype=np.float64)
    m = parity.digest_metrics(np.moveaxis(S[..., :3], -1, 0), rows, "scattering", lanes, floor=1e-3)
    a = parity.digest_metrics(np.moveaxis(S[..., 3:], -1, 0), rows, "delta_mie", slice(15, 16), floor=1e-3)
    e = parity.error_metrics(np.moveaxis(np.asarray(E)[..., :3], -1, 0), two["irradiance"][lanes])
    t = parity.error_metrics(np.moveaxis(np.asarray(T)[..., :3], -1, 0), two["transmittance"][lanes])
    tol_sum, tol_max = (parity.HALF_TOL_SUM, parity.HALF_TOL_TEXEL) if half else (parity.REL_TOL, parity.REL_TOL)
    s_sum, s_max = max(m["sum"], m["wsum"], a["sum"], a["wsum"]), max(m["max"], a["max"])
    out = {"scattering_row_sums": s_sum, "scattering_row_sums_tol": tol_sum, "scattering_row_max": s_max,
           "scattering_row_max_tol": tol_max, "irradiance": e["max_floor"], "transmittance": t["max_floor"],
           "tol": parity.REL_TOL, "n_texels": m["texels"] + a["texels"] + int(two["irradiance"][lanes].size) +
           int(two["transmittance"][lanes].size), "nan": m["nan"] + a["nan"] + e["nan"] + t["nan"]}
    out["max_floor"] = parity.REL_TOL * max(e["max_floor"] / parity.REL_TOL, t["max_floor"] / parity.REL_TOL,
                                            s_sum / tol_sum, s_max / tol_max)
    out["ok"] = bool(out["nan"] == 0 and out["max_floor"] <= parity.REL_TOL)
    return out


# ---- the B200 arm ---------------------------------------------------------------------------------
def run_b200(args, rank, world_size, local_rank):
    import numpy as np
    import torch

    import precomputed_atmospheric_scattering_b200 as pas
    from precomputed_atmospheric_scattering_b200 import world

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the precompute has no CPU fallback")
    torch.cuda.set_device(local_rank)
    distributed = world_size > 1
    if distributed:
        import torch.distributed as dist
        # keep stdout to the one JSON line: NCCL prints "NCCL version ..." on stdout at the VERSION
        # level and honours NCCL_DEBUG_FILE only above it
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    if args.config == 5:
        return run_ensemble(args, rank, world_size, local_rank, barrier)
    cfg = CONFIGS[args.config]
    ORDERS, C = cfg["orders"], cfg["wavelengths"]
    spec = pas.earth(C, half_precision=cfg["half"], combine_scattering_textures=True)
    kw = {"sizes": cfg["sizes"]} if cfg["sizes"] else {}
    if distributed and cfg["sizes"] and cfg["sizes"]["scattering_r"] % world_size:
        raise SystemExit("scattering_r must be divisible by the number of GPUs")

    def new_model():
        m = pas.Model.from_spec(spec, device=local_rank, **kw)
        if distributed:
            world.attach(m)
        return m

    model = new_model()
    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        model.Init(ORDERS)
    peaks = pas.measure_device_peaks(local_rank) if rank == 0 else None

    # ---- timed region: K complete precomputes, tables rebuilt from the parameters each time ----
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    device_ms, phases, launches = [], {}, 0
    for _ in range(args.steps):
        model.Init(ORDERS)
        tm = model.last_timings()
        device_ms.append(sum(tm.values()))
        for k, v in tm.items():
            phases[k] = phases.get(k, 0.0) + v / args.steps
        launches += model.last_launch_count()
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0) / args.steps
    clocks = sampler.finish()
    ms_per_step = world.max_over_ranks(sum(device_ms) / len(device_ms))
    wall_ms = world.max_over_ranks(wall_ms)
    if os.environ.get("PAS_BENCH_RANK_PHASES"):
        # per-rank phase timings (which slab is the slow one): one JSON line per rank on stderr
        print(json.dumps({"rank": rank, "phases_ms": {k: round(v, 4) for k, v in phases.items()}}), file=sys.stderr, flush=True)

    # ---- parity of the tables just timed, on every rank, outside the timed region ---------------------
    parity = None
    if args.config in (1, 2):
        parity = parity_check(args.config, model.texture(pas.TEXTURE_SCATTERING), model.irradiance,
                              model.transmittance, model.luminance_matrix(), cfg["half"])
        parity["max_floor"] = world.max_over_ranks(parity["max_floor"])
        parity["ranks_checked"] = world_size
        parity["ranks_ok"] = int(round(world_size - world.sum_over_ranks(0.0 if parity["ok"] else 1.0)))
        parity["ok"] = parity["ranks_ok"] == world_size
        parity["against"] = ("tests/golden/earth18_rows.npz + earth18_2d.npz: outputs of the unmodified reference "
                             "CPU model (oracle/run_reference.py); every rank checks the complete tables it holds")

    # ---- e2e: host arrays -> Model -> Init -> host tables, every step -------------------------------
    info = {w: model.texture_info(w) for w in (pas.TEXTURE_TRANSMITTANCE, pas.TEXTURE_SCATTERING,
                                               pas.TEXTURE_IRRADIANCE)}
    host, shared = {}, None
    own_layers = distributed and os.environ.get("PAS_EXCHANGE", "symm") != "nccl" and \
        os.environ.get("PAS_HOST_TABLES", "shared") == "shared"
    if own_layers:
        # N > 1: ONE set of page-locked host tables shared by the ranks (POSIX shared memory); every rank
        # copies the layers it computed into it, Init returns once every rank's part is there
        shared = world.shared_host_tables(model, list(info))
        host = dict(shared.arrays)
    else:
        for w, i in info.items():
            shape = ((i.depth,) if i.depth > 1 else ()) + (i.height, i.width, 4)
            dtype = np.float16 if i.bytes_per_channel == 2 else np.float32
            t = torch.empty(shape, dtype=torch.float16 if dtype == np.float16 else torch.float32).pin_memory()
            host[w] = t.numpy()
    d2h = sum(a.nbytes for a in host.values())
    # parameters handed to the library per step: 7 spectra + wavelengths (48 doubles each), layers, scalars
    h2d = 8 * len(spec.wavelengths) * 8 + 5 * 5 * 8 + 12 * 8
    model.close()

    def e2e_step():
        # the public API a caller that uploads the tables uses: host buffers registered up front,
        # Init fills them (each table is copied out as soon as it is final) and returns when done
        m = new_model()
        m.set_host_outputs(transmittance=host[pas.TEXTURE_TRANSMITTANCE], scattering=host[pas.TEXTURE_SCATTERING],
                           irradiance=host[pas.TEXTURE_IRRADIANCE])
        if own_layers:
            m.set_host_output_mode(True)
        m.Init(ORDERS)
        L = m.luminance_matrix()
        m.close()
        return L

    for _ in range(3):
        e2e_step()
    barrier()
    if not own_layers or rank == 0:
        for a in host.values():
            a[...] = 0
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        L = e2e_step()
    barrier()
    e2e_ms = world.max_over_ranks(1e3 * (time.perf_counter() - t0) / args.steps)
    if parity is not None:
        # the host tables the e2e path delivered
        p2 = parity_check(args.config, host[pas.TEXTURE_SCATTERING].astype(np.float32),
                          host[pas.TEXTURE_IRRADIANCE], host[pas.TEXTURE_TRANSMITTANCE], L, cfg["half"])
        parity["e2e_host_tables_max_floor"] = world.max_over_ranks(p2["max_floor"])
        ok2 = int(round(world_size - world.sum_over_ranks(0.0 if p2["ok"] else 1.0))) == world_size
        parity["ok"] = bool(parity["ok"] and ok2)

    if shared is not None:
        host = {}
        shared.close()
    if rank != 0:
        return 0 if (parity is None or parity["ok"]) else 3
    # ---- roofline of the dominant kernel ------------------------------------------------------------
    sz = cfg["sizes"] or {}
    n_t = sz.get("transmittance_width", 256) * sz.get("transmittance_height", 64)
    n_s = (sz.get("scattering_nu", 8) * sz.get("scattering_mu_s", 32) * sz.get("scattering_mu", 128) *
           sz.get("scattering_r", 32))
    flops, mufu, total_f, total_u = canonical_work(C=C, K=ORDERS, n_t=n_t, n_s=n_s)
    # ncu evidence of the same kernels (profiles/ncu_hot_kernels.json, made by tools/ncu_to_json.py from
    # one `ncu --set full` capture of the config-2 workload on B200): executed thread instructions per
    # SASS opcode -> executed fp32 flops per launch, DRAM bytes per launch = `traffic`, L1 wavefronts
    ncu = None
    tpath = os.path.join(ROOT, "profiles", "ncu_hot_kernels.json")
    if os.path.exists(tpath) and args.config == 2:
        ncu = json.load(open(tpath))
    sm_hz = 1e6 * (clocks["sm_mhz"] or peaks.get("sm_mhz") or 1965.0)
    kernels = {}
    # per-rank figures: a rank computes 1/N of every 3-D pass (its r-slab); rank 0's timings
    share = {k: (1.0 if k == "transmittance" else 1.0 / world_size) for k in flops}
    for name, ms in phases.items():
        key = phase_key(name)
        if key not in flops or ms <= 0:
            continue
        f, u = flops[key] * share[key], mufu[key] * share[key]
        k = {"ms": round(ms, 4), "canonical_gflop": round(f / 1e9, 2),
             "canonical_frac_fp32_peak": round(f / (ms * 1e-3) / 1e12 / peaks["fp32_tflops"], 4),
             "canonical_frac_mufu_peak": round(u / (ms * 1e-3) / 1e9 / peaks["mufu_gops"], 4)}
        e = ncu["passes"].get(key) if ncu else None
        if e and "fp32_flop_executed" in e:
            ef, eu = e["fp32_flop_executed"] * share[key], e["mufu_executed"] * share[key]
            k["executed_gflop"] = round(ef / 1e9, 2)
            k["tflops"] = round(ef / (ms * 1e-3) / 1e12, 2)
            k["frac_fp32_peak"] = round(ef / (ms * 1e-3) / 1e12 / peaks["fp32_tflops"], 4)
            k["frac_mufu_peak"] = round(eu / (ms * 1e-3) / 1e9 / peaks["mufu_gops"], 4)
            if e.get("l1_wavefronts_per_sm"):
                # one 128-byte wavefront per cycle per SM is the L1 / shared-memory data-pipe peak
                k["l1_wavefront_frac"] = round(e["l1_wavefronts_per_sm"] * share[key] / (ms * 1e-3 * sm_hz), 4)
            k["ncu"] = {m: e[m] for m in ("traffic_bytes", "fma_pipe_cycles_active_pct", "issue_active_pct",
                                          "l1_data_pipe_wavefronts_pct", "dram_throughput_pct") if m in e}
        kernels[name] = k
    dom = max(kernels, key=lambda k: kernels[k]["ms"])
    kd = kernels[dom]
    executed = "frac_fp32_peak" in kd
    exec_total = None
    if executed and all("executed_gflop" in k for n, k in kernels.items() if phase_key(n) != "transmittance"):
        exec_total = sum(k.get("executed_gflop", 0.0) for k in kernels.values()) * world_size
    roofline = {
        "kernel": dom, "bound": "fp32",
        "achieved": kd["tflops"] if executed else round(kd["canonical_gflop"] / kd["ms"], 2),
        "peak": round(peaks["fp32_tflops"], 2), "unit": "TFLOP/s",
        "frac": kd["frac_fp32_peak"] if executed else kd["canonical_frac_fp32_peak"],
        "canonical_frac": kd["canonical_frac_fp32_peak"],
        "traffic": (ncu["passes"].get(phase_key(dom), {}).get("traffic_bytes") if ncu else None),
        "peak_source": "measured on this device by pas_measure_device_peaks (FMA microbenchmark); "
                       "MEASURED_PEAKS.json has no FP32 figure",
        "achieved_is": ("EXECUTED fp32 flops per launch (FFMA x2, FFMA2 x4, FMUL, FMUL2 x2, FADD, FADD2 x2 thread "
                        "instructions of the ncu capture of the same kernel) / live CUDA-event time of the pass; "
                        "canonical_frac = the minimal-formulation count of SURVEY.md 8(d) over the same time "
                        "(it exceeds 1 where the kernel needs fewer flops than that formulation)") if executed else
                       "canonical FP32 flops of SURVEY.md 8(d) / live CUDA-event time (no ncu census for this config)",
        "ncu_source": ncu["source"] if ncu else None,
        "mufu_peak_gops": round(peaks["mufu_gops"], 1),
        "whole_job": {"canonical_gflop": round(total_f / 1e9, 1), "canonical_mufu_gop": round(total_u / 1e9, 2),
                      "executed_gflop": round(exec_total, 1) if exec_total else None,
                      "frac_fp32_peak": round(exec_total * 1e9 / (ms_per_step * 1e-3) / 1e12 /
                                              (world_size * peaks["fp32_tflops"]), 4) if exec_total else None,
                      "canonical_frac_fp32_peak": round(total_f / (ms_per_step * 1e-3) / 1e12 / (world_size * peaks["fp32_tflops"]), 4),
                      "canonical_frac_mufu_peak": round(total_u / (ms_per_step * 1e-3) / 1e9 / (world_size * peaks["mufu_gops"]), 4)},
        "kernels": kernels,
    }
    line = {
        "metric": cfg["metric"], "value": round(ms_per_step, 4), "unit": "ms", "n_gpus": world_size,
        "steps": args.steps, "warmup": warmup, "ms_per_step": round(ms_per_step, 4),
        "wall_ms_per_step": round(wall_ms, 4), "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.config, world_size),
        "e2e": {"value": round(e2e_ms, 4), "unit": "ms", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "what": "Model(host arrays) + Init with T, E copied into registered pinned host buffers as they become final and S written there by the last multiple-scattering pass itself + destroy"
                        + ("; one set of host tables shared by the ranks, every rank copies the layers it computed" if own_layers else "")},
        "gpu_launches": launches, "clocks": clocks, "parity": parity, "roofline": roofline,
        "phases_ms": {k: round(v, 4) for k, v in phases.items()},
    }
    if world_size == 1 and not args.no_cpu_baseline:
        try:
            cpu = ReferenceCpu(args.config)
            cpu.step()
            budget, est, spent = 20.0, [], 0.0
            while spent < budget and len(est) < 8:
                e, s = cpu.step()
                est.append(e)
                spent += s
            ms = 1e3 * sum(est) / len(est)
            line["cpu_baseline"] = cpu.baseline(ms, spent / len(est))
        except Exception as e:  # pragma: no cover
            line["cpu_baseline"] = {"value": None, "unit": "ms", "cores": os.cpu_count(), "kind": "reference",
                                    "sample": f"unavailable: {e}"}
    print(json.dumps(line))
    return 0 if (parity is None or parity["ok"]) else 3


def run_ensemble(args, rank, world_size, local_rank, barrier):
    """BASELINE config 5: 64 atmospheres precomputed with 16 in flight (pas_model_init_async), then, per
    atmosphere, one 1920x1080 image of sky radiance through the product's render-time lookup
    (pas_model_get_sky_radiance = GetSkyRadiance of functions.glsl:1705-1769, one query per pixel, view
    rays of the reference's test camera, inputs and outputs resident on the device). Multi-GPU: the
    atmospheres are dealt round-robin to the ranks (independent models, no exchange)."""
    import numpy as np
    import torch
    import precomputed_atmospheric_scattering_b200 as pas
    from precomputed_atmospheric_scattering_b200 import ensemble, world
    from tests import scene
    cfg = CONFIGS[5]
    specs = ensemble.sweep(num_precomputed_wavelengths=3, half_precision=True)[rank::world_size]
    W, H = 1920, 1080
    view = scene.model_test_view(65.0, 90.0, False, width=W, height=H, sun_angular_radius=specs[0].sun_angular_radius)
    # view rays of every pixel (reference/model_test.cc:690-711), camera relative to the planet centre
    M = np.asarray(view.model_from_clip, dtype=np.float64).reshape(3, 3)
    x = 2.0 * (np.arange(W) + 0.5) / W - 1.0
    y = 1.0 - 2.0 * (np.arange(H) + 0.5) / H
    rays = np.stack(np.broadcast_arrays(x[None, :], y[:, None], 1.0), axis=-1) @ M.T
    rays /= np.linalg.norm(rays, axis=-1, keepdims=True)
    dev = torch.device("cuda", local_rank)
    n = W * H
    d_rays = torch.from_numpy(rays.reshape(n, 3)).to(dev)
    d_cam = (torch.tensor(view.camera, dtype=torch.float64) - torch.tensor(view.earth_center, dtype=torch.float64)).to(dev).repeat(n, 1)
    d_sun = torch.tensor(view.sun_direction, dtype=torch.float64, device=dev).repeat(n, 1)
    d_L = torch.empty((n, 3), dtype=torch.float32, device=dev)
    d_T = torch.empty((n, 3), dtype=torch.float32, device=dev)
    host_image = torch.empty((n, 3), dtype=torch.float32).pin_memory()

    def step():
        t0 = time.perf_counter()
        models = ensemble.precompute(specs, cfg["orders"], device=local_rank, max_in_flight=16)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        render_ms = 0.0
        for m in models:
            m.sky_radiance_device(n, d_cam.data_ptr(), d_rays.data_ptr(), 0, d_sun.data_ptr(), d_L.data_ptr(), d_T.data_ptr())
            render_ms += m.last_render_ms()
            host_image.copy_(d_L)       # every image goes back to the host
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        launches = sum(m.last_launch_count() + 1 for m in models)
        for m in models:
            m.close()
        return 1e3 * (t1 - t0), 1e3 * (t2 - t1), render_ms, launches

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    pre, ren, dev_ren, launches = [], [], [], 0
    steps = max(1, min(args.steps, 5))
    for _ in range(steps):
        a, b, c, nl = step()
        pre.append(a), ren.append(b), dev_ren.append(c)
        launches += nl
    barrier()
    clocks = sampler.finish()
    pre_ms = world.max_over_ranks(sum(pre) / steps)
    ren_ms = world.max_over_ranks(sum(ren) / steps)
    finite = bool(torch.isfinite(host_image).all() and (host_image >= 0).all() and host_image.max() > 0)
    if rank != 0:
        return 0
    total = pre_ms + ren_ms
    print(json.dumps({
        "metric": cfg["metric"], "value": round(total, 3), "unit": "ms", "n_gpus": world_size, "steps": steps,
        "warmup": max(1, min(args.warmup, 2)), "ms_per_step": round(total, 3), "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(5, world_size),
        "precompute_ms_64_atmospheres": round(pre_ms, 3), "atmospheres_per_second": round(64e3 / pre_ms, 1),
        "render_ms_64_x_1080p_wall_with_readback": round(ren_ms, 3),
        "render_kernel_ms_per_1080p_image": round(sum(dev_ren) / steps / len(specs), 4),
        "e2e": {"value": round(total, 3), "unit": "ms", "h2d_bytes_per_step": 64 * 3368,
                "d2h_bytes_per_step": 64 * n * 12,
                "what": "host wall clock: 64 x Model(host arrays) + InitAsync/Wait, then 64 x 1080p GetSkyRadiance images read back to pinned host memory"},
        "gpu_launches": launches, "clocks": clocks, "parity": None, "images_finite": finite,
        "parity_note": "tests/test_gpu_configs.py checks this batch against blocking Init and the oracle, and renders of the "
                       "reference's test scene with these tables against the oracle renderer",
    }))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--wavelengths", type=int, default=0,
                    help="override the wavelength count of the config (scaling studies: --config 4 --wavelengths 3)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.wavelengths and args.wavelengths != CONFIGS[args.config]["wavelengths"]:
        c = CONFIGS[args.config]
        c["workload"] = c["workload"].replace(f"{c['wavelengths']} precomputed wavelengths", f"{args.wavelengths} precomputed wavelengths")
        c["metric"] = c["metric"].replace(f"{c['wavelengths']}_wavelengths", f"{args.wavelengths}_wavelengths")
        c["wavelengths"] = args.wavelengths
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world_size != args.gpus:
        if args.gpus > 1 and world_size == 1:
            raise SystemExit(f"--gpus {args.gpus}: launch with python -m torch.distributed.run "
                             f"--nproc-per-node {args.gpus} (one rank per GPU)")
    rc = run_b200(args, rank, world_size, local_rank)
    if world_size > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rc:
        sys.exit(rc)


if __name__ == "__main__":
    main()
