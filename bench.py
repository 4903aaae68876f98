"""Headline benchmark: the full LUT precompute (atmosphere::Model::Init equivalent) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): Earth/demo atmosphere (atmosphere/demo/demo.cc:188-284),
precomputed-luminance mode with 15 wavelengths (5 RGB batches in the reference's terms), 4 scattering
orders, the reference's table sizes (T 256x64, E 64x16, S 256x128x32), combined scattering textures,
half-precision final 3-D tables -- the demo's own settings. One "step" = one complete precompute of
all tables. Metric (BASELINE.json): wall-clock ms of that precompute; lower is better.

  value / ms_per_step  device time of one Init, CUDA events on the library's own stream (start of
                       the first kernel to the end of the last), mean over K steps, max over ranks.
                       Nothing is resident beforehand except the model parameters; every table is
                       recomputed each step.
  e2e                  the same job through the reference-facing API with HOST buffers: construct
                       the Model from host arrays (the 19 constructor arguments; parameters reach the
                       GPU as kernel arguments), Init(4), copy the three product tables back to host
                       memory, destroy. Wall clock around the synchronous calls.
  roofline             dominant kernel against the measured FP32 FMA peak of this device (this path is
                       FP32/SFU bound, not HBM or tensor bound: SURVEY.md section 8d).
  cpu_baseline         the UNMODIFIED reference CPU model (oracle/_ref) on the host cores, bounded
                       row sample, extrapolated to the full job.

--impl reference times that CPU reference alone (rank 0 only under torchrun).
Multi-GPU (torchrun, one rank per GPU): r-slab sharding inside the library with NCCL all-gathers
between orders; strong scaling (the job is fixed, N GPUs share it).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ORDERS = 4
WAVELENGTHS = 15
METRIC = "lut_precompute_ms_4_orders_15_wavelengths"
WORKLOAD = ("earth-demo atmosphere, 15 precomputed wavelengths (5 RGB batches), 4 scattering orders, "
            "T 256x64, E 64x16, S 256x128x32, combined textures, half-precision final tables")


# ---- canonical algorithmic work ---------------------------------
def canonical_work(C=WAVELENGTHS, K=ORDERS, n_t=256 * 64, n_e=64 * 16, n_s=256 * 128 * 32):
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
    lanes in fp64 whatever the number of wavelengths asked for; the 15 bench channels sit in lanes
    0..14."""

    def __init__(self):
        from oracle import ref
        import precomputed_atmospheric_scattering_b200 as pas
        if not ref.available():
            raise RuntimeError("oracle/_ref/libpas_ref.so is missing (build() compiles it where "
                               "/root/reference exists; the prebuilt file travels to the GPU box)")
        self.threads = os.cpu_count() or 1
        spec = pas.earth(WAVELENGTHS, half_precision=True)
        cp = pas.channel_params(spec, pas.precomputed_wavelengths(WAVELENGTHS))
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
        for order in range(2, ORDERS + 1):
            run("scattering_density", order, self.stride_density)
            run("indirect_irradiance", order)
            run("multiple_scattering", order, self.stride_ray)
        return est, spent


def run_reference(args, rank):
    if rank != 0:
        return
    try:
        cpu = ReferenceCpu()
    except Exception as e:  # pragma: no cover
        print(json.dumps({"impl": "reference", "unavailable": str(e)}))
        return
    for _ in range(args.warmup):
        cpu.step()
    est = []
    for _ in range(args.steps):
        e, _ = cpu.step()
        est.append(e)
    ms = 1e3 * sum(est) / len(est)
    line = {
        "impl": "reference", "metric": METRIC, "value": ms, "unit": "ms", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "orders": ORDERS, "wavelengths": WAVELENGTHS},
        "cpu_baseline": {"value": ms, "unit": "ms", "cores": cpu.threads, "kind": "reference",
                         "sample": cpu.sample(),
                         "value_scaled_to_15_of_47_lanes": ms * WAVELENGTHS / 47.0},
        "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


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

    spec = pas.earth(WAVELENGTHS, half_precision=True, combine_scattering_textures=True)

    def new_model():
        m = pas.Model.from_spec(spec, device=local_rank)
        if distributed:
            world.attach(m)
        return m

    model = new_model()
    for _ in range(max(args.warmup, 3)):
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

    # ---- e2e: host arrays -> Model -> Init -> host tables, every step -------------------------------
    info = {w: model.texture_info(w) for w in (pas.TEXTURE_TRANSMITTANCE, pas.TEXTURE_SCATTERING,
                                               pas.TEXTURE_IRRADIANCE)}
    host = {}
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
        m.Init(ORDERS)
        m.close()

    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_ms = world.max_over_ranks(1e3 * (time.perf_counter() - t0) / args.steps)

    if rank != 0:
        return
    # ---- roofline of the dominant kernel ------------------------------------------------------------
    flops, mufu, total_f, total_u = canonical_work()
    kernels = {}
    # per-rank figures: a rank computes 1/N of every 3-D pass (its r-slab); rank 0's timings
    share = {k: (1.0 if k == "transmittance" else 1.0 / world_size) for k in flops}
    for name, ms in phases.items():
        key = phase_key(name)
        if key in flops and ms > 0:
            f, u = flops[key] * share[key], mufu[key] * share[key]
            kernels[name] = {"ms": round(ms, 4), "canonical_gflop": round(f / 1e9, 2),
                             "tflops": round(f / (ms * 1e-3) / 1e12, 2),
                             "frac_fp32_peak": round(f / (ms * 1e-3) / 1e12 / peaks["fp32_tflops"], 4),
                             "frac_mufu_peak": round(u / (ms * 1e-3) / 1e9 / peaks["mufu_gops"], 4)}
    dom = max(kernels, key=lambda k: kernels[k]["ms"])
    # ncu evidence of the same kernels (profiles/ncu_hot_kernels.json, made by tools/ncu_to_json.py
    # from one `ncu --set full` capture on B200): DRAM bytes per launch = `traffic`, and the executed
    # pipe utilisations, which is what bounds these kernels (the canonical flop count above credits
    # the minimal formulation; the density kernel executes fewer instructions than that)
    traffic, ncu = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_hot_kernels.json")
    if os.path.exists(tpath):
        ncu = json.load(open(tpath))
        traffic = ncu["passes"].get(phase_key(dom), {}).get("traffic_bytes")
        for name, k in kernels.items():
            e = ncu["passes"].get(phase_key(name))
            if e:
                k["ncu"] = {m: e[m] for m in ("traffic_bytes", "fma_pipe_cycles_active_pct", "issue_active_pct",
                                              "l1_data_pipe_wavefronts_pct", "dram_throughput_pct")}
    roofline = {
        "kernel": dom, "bound": "fp32", "achieved": kernels[dom]["tflops"], "peak": round(peaks["fp32_tflops"], 2),
        "unit": "TFLOP/s", "frac": kernels[dom]["frac_fp32_peak"], "traffic": traffic,
        "peak_source": "measured on this device by pas_measure_device_peaks (FMA microbenchmark); "
                       "MEASURED_PEAKS.json has no FP32 figure",
        "achieved_is": "canonical FP32 flops of SURVEY.md 8(d) / live CUDA-event time of the pass; > peak "
                       "means the kernel needs fewer flops than the canonical formulation counts",
        "ncu_source": ncu["source"] if ncu else None,
        "mufu_peak_gops": round(peaks["mufu_gops"], 1),
        "whole_job": {"canonical_gflop": round(total_f / 1e9, 1), "canonical_mufu_gop": round(total_u / 1e9, 2),
                      "frac_fp32_peak": round(total_f / (ms_per_step * 1e-3) / 1e12 / (world_size * peaks["fp32_tflops"]), 4),
                      "frac_mufu_peak": round(total_u / (ms_per_step * 1e-3) / 1e9 / (world_size * peaks["mufu_gops"]), 4)},
        "kernels": kernels,
    }
    line = {
        "metric": METRIC, "value": round(ms_per_step, 4), "unit": "ms", "n_gpus": world_size,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4),
        "wall_ms_per_step": round(wall_ms, 4), "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "orders": ORDERS, "wavelengths": WAVELENGTHS,
                   "parallelism": "1 GPU" if world_size == 1 else (
                       f"r-slabs over {world_size} GPUs, " + ("NCCL all-gather / all-reduce per order"
                       if os.environ.get("PAS_EXCHANGE") == "nccl" else
                       "density slabs stored to all ranks by the kernel over NVLink peer memory, flag barriers")),
                   "l2": "no flush between steps: every step recomputes and rewrites all tables "
                         "(5 x 60 MiB intermediates + products > 126 MB L2), nothing is reused across steps"},
        "e2e": {"value": round(e2e_ms, 4), "unit": "ms", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "what": "Model(host arrays) + Init(4) with T, S, E copied into registered pinned host buffers as they become final + destroy"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
    }
    if world_size == 1 and not args.no_cpu_baseline:
        try:
            cpu = ReferenceCpu()
            cpu.step()
            budget, est, spent = 20.0, [], 0.0
            while spent < budget and len(est) < 8:
                e, s = cpu.step()
                est.append(e)
                spent += s
            ms = 1e3 * sum(est) / len(est)
            line["cpu_baseline"] = {"value": round(ms, 1), "unit": "ms", "cores": cpu.threads,
                                    "kind": "reference", "sample": cpu.sample(),
                                    "value_scaled_to_15_of_47_lanes": round(ms * WAVELENGTHS / 47.0, 1)}
        except Exception as e:  # pragma: no cover
            line["cpu_baseline"] = {"value": None, "unit": "ms", "cores": os.cpu_count(), "kind": "reference",
                                    "sample": f"unavailable: {e}"}
    print(json.dumps(line))
    if distributed:
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world_size != args.gpus:
        if args.gpus > 1 and world_size == 1:
            raise SystemExit(f"--gpus {args.gpus}: launch with python -m torch.distributed.run "
                             f"--nproc-per-node {args.gpus} (one rank per GPU)")
    run_b200(args, rank, world_size, local_rank)
    if world_size > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
