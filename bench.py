#!/usr/bin/env python
"""bench.py -- log-prob + gradient evaluations per second, one process per GPU.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

Headline (``metric`` / ``value`` / ``e2e`` / ``roofline``): the C2 workload -- BASELINE.json configs[1]: EPL+shear lens,
SersicEllipse lens light + source, 60x60, supersample 2, 13x13 PSF, batch 4096 per GPU.  A "step" is one pass of the hot
path over one batch: ``ForwardProbModel.log_prob`` and its gradient w.r.t. ``z`` for ``bs`` samples (bijector + prior,
ray-shooting + light, PSF conv + pool, chi^2 likelihood, and the hand adjoint of all of it).  ``value`` times it with ``z``
resident in HBM; ``e2e`` times the same call through the C ABI's host-buffer entry point (``gl_logprob_grad_host``: pinned
host ``z`` in, ``logp`` / ``red_chi2`` / ``dz`` out, copies inside the timed region).  Samples are independent, so N GPUs run
N shards with no data-path collective ("weak" scaling); timing is CUDA events, max over ranks.

Extra keys, timed in the same process (skip with --headline-only):
  ``c4``      BASELINE.json configs[3]: cluster model (NFW + 30-member dPIE scaling relation + shear), 200x200, ss=2,
              batch 1024 per GPU, log-prob + gradient;
  ``c3``      configs[2]: EPL+shear / Shapelets(n_max=10) through ``lstsq_simulate`` (BackwardProbModel), batch 2048 per GPU;
  ``c5_svi`` / ``c5_hmc``  configs[4]: the VI + HMC drivers (``ModellingSequence``) on the cluster model, GLOBAL batch 16384
              split over the N ranks (strong scaling), with the share of a step spent in the NCCL all-reduce.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "logprob+grad evals/sec (60x60, ss=2)"
UNIT = "evals/s"
BS_PER_GPU = 4096
WORKLOAD = "C2: EPL+shear / SersicEllipse x2, 60x60, ss=2, PSF 13x13, MAP logprob+grad (BASELINE.json configs[1])"
C5_GLOBAL_BATCH = 16384
PARAMS_PER_SAMPLE = 22


def make_config(world):
    """`config` of the JSON line: identical for the CUDA arm and the reference arm (same workload, same batch)."""
    return {"workload": WORKLOAD, "global_batch": BS_PER_GPU * world, "batch_per_gpu": BS_PER_GPU, "params_per_sample": PARAMS_PER_SAMPLE,
            "parallelism": f"dp{world} (sample shards, no data-path collective)",
            "l2": "working set per step (236 MB ss image + adjoint) exceeds the 126 MB L2; no explicit flush"}


STAGES = ["k_unconstrain", "k_prep", "k_raytrace_fwd", "k_conv_fwd", "k_conv_bwd", "k_raytrace_bwd", "k_sample_bwd"]


def _load_json(*path):
    try:
        return json.load(open(os.path.join(ROOT, *path)))
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": float(self.samples[0][1]) if self.samples else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def _dist_setup():
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend=backend, **kw)
    return rank, world, local


# ----------------------------------------------------------------------------------------------------------------
# the reference's CPU path (oracle port: torch CPU fp32 + autograd), on a bounded sample of the workload
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_rate(bs, reps):
    """evals/s of the oracle port on ALL host cores (torchrun exports OMP_NUM_THREADS=1: overridden explicitly)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch

    import oracle_bridge
    from gigalens_b200 import workloads
    from gigalens_b200.model import ProbabilisticModel

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = workloads.c2_workload()
    z = ProbabilisticModel(wl["prior"]).bij_inverse(wl["prior"].sample(bs, seed=0))
    sim, pm = oracle_bridge.build_oracle(wl, bs, torch.float32)

    def step():
        zt = torch.as_tensor(z).clone().requires_grad_(True)
        logp, _ = pm.log_prob(sim, zt)
        logp.sum().backward()
        return float(logp[0])

    step()
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    med = sorted(times)[len(times) // 2]
    return bs / med, med, torch.get_num_threads()


def run_reference(args):
    rank, world, _ = _dist_setup()
    if rank != 0:
        return
    bs = 128   # one "step" = a bs=128 slice of the 4096-sample batch
    t0 = time.perf_counter()
    rate, med, cores = cpu_reference_rate(bs, max(1, args.steps))
    sample = (f"each step = a bs={bs} slice of the bs={BS_PER_GPU} batch (evals_per_step = {bs}), median of {max(1, args.steps)} steps after "
              f"1 warm-up; torch-CPU fp32 oracle port with autograd on {cores} host threads (the reference's TF/JAX stack is not installable offline)")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(args.gpus), "evals_per_step": bs,
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------
# CUDA arm
# ----------------------------------------------------------------------------------------------------------------
class Ctx:
    pass


def _sync_all(ctx):
    import torch

    torch.cuda.synchronize()
    if ctx.world > 1:
        ctx.dist.barrier()
        torch.cuda.synchronize()


def _max_over_ranks(ctx, x):
    import torch

    t = torch.tensor([float(x)], device="cuda")
    if ctx.world > 1:
        ctx.dist.all_reduce(t, op=ctx.dist.ReduceOp.MAX)
    return float(t.item())


def _sum_over_ranks(ctx, x):
    import torch

    t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
    if ctx.world > 1:
        ctx.dist.all_reduce(t)
    return float(t.item())


def _timed_steps(ctx, fn, steps, warmup):
    """W warm-up calls, then K calls between CUDA events, barrier + synchronize on both sides; ms (max over ranks)."""
    import torch

    for _ in range(warmup):
        fn()
    _sync_all(ctx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    _sync_all(ctx)
    return _max_over_ranks(ctx, ms), out


def _stage_times(ctx, sim, fn, steps):
    """Per-kernel device time (ms per step) from CUDA events the library records on the launch stream around every kernel
    of `steps` further timed steps (gl_plan_set_option "timing")."""
    n = min(steps, 256)
    sim.set_option("timing", n)
    for _ in range(n):
        fn()
    stage = (C.c_float * 7)()
    ncalls = C.c_int32(0)
    from gigalens_b200 import _cabi

    _cabi.check(ctx.lib.gl_plan_get_timings(sim._plan, stage, C.byref(ncalls)), ctx.lib)
    sim.set_option("timing", 0)
    return [stage[k] / max(1, ncalls.value) for k in range(7)]


def measure_fp32_peak(ctx):
    a, b = C.c_float(0), C.c_float(0)
    from gigalens_b200 import _cabi

    _cabi.check(ctx.lib.gl_fp32_peak(ctx.local, C.byref(a), C.byref(b)), ctx.lib)
    return {"ffma_TFLOPs": a.value, "ffma2_TFLOPs": b.value, "peak_TFLOPs": max(a.value, b.value),
            "how": "gl_fp32_peak: 16 independent FMA chains per thread, 8 CTAs x 256 threads per SM, no memory; best of 5 launches, CUDA events",
            "derived_TFLOPs_at_max_clock": 148 * 128 * 2 * (ctx.peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12}


def c2_nominal_flops(npix, mean_trips):
    """SURVEY.md 8d's hand count (FMA = 2) per eval: per ss pixel EPL 60 + 14 I, shear 6, SersicEllipse 45 x 2; adjoint = 2 x forward;
    conv 4.87 M each way.  I = mean series length of THIS batch (the kernels run per-sample counts)."""
    fwd = npix * (60.0 + 14.0 * mean_trips + 6.0 + 90.0)
    return {"k_raytrace_fwd": fwd, "k_conv_fwd": 4.87e6, "k_conv_bwd": 4.87e6, "k_raytrace_bwd": 2.0 * fwd}


def build_roofline(ctx, sim, stage_ms, step_ms, flops_tab, alg_bytes, ncu_key, fp32):
    """`roofline` of the dominant kernel on its BINDING ceiling (FP32 FMA, measured peak) + the per-kernel table against both
    ceilings (north_star: FP32 utilisation for the profile kernels, HBM GB/s for the conv / likelihood)."""
    bs = sim.bs
    hbm_peak = ctx.peaks.get("hbm_gbs", 6650.0)
    which = "MEASURED_PEAKS.json" if ctx.peaks else "fallback 6650 GB/s"
    ncu = (_load_json("profiles", "r02_traffic.json") or _load_json("profiles", "r01_traffic.json") or {}).get(ncu_key, {})
    per_kernel = {}
    for k, nm in enumerate(STAGES):
        if nm not in alg_bytes or stage_ms[k] <= 0:
            continue
        rec = next((ncu[key] for key in (nm + "_p", nm + "_tma", nm) if key in ncu), {})
        gbs = alg_bytes[nm] * bs / (stage_ms[k] * 1e-3) / 1e9
        tf = flops_tab[nm] * bs / (stage_ms[k] * 1e-3) / 1e12
        ex = rec.get("executed_flops")   # ncu source page: FFMA(2) x 2(4) + FMUL(2) / FADD(2) x 1(2) predicated-on thread instructions, per launch
        if ex and rec.get("ncu_bs"):
            ex = ex * bs / rec["ncu_bs"]  # the capture ran a smaller batch: work per launch scales with the batch
        per_kernel[nm] = {
            "ms": stage_ms[k], "share_of_step": stage_ms[k] / step_ms,
            "fp32_TFLOPs_useful": tf, "fp32_frac_useful": tf / fp32["peak_TFLOPs"],
            "fp32_TFLOPs_executed": (ex / (stage_ms[k] * 1e-3) / 1e12) if ex else None,
            "fp32_frac_executed": (ex / (stage_ms[k] * 1e-3) / 1e12 / fp32["peak_TFLOPs"]) if ex else None,
            "hbm_GBs": gbs, "hbm_frac": gbs / hbm_peak, "algorithmic_bytes": alg_bytes[nm] * bs,
            "ncu": {k2: rec.get(k2) for k2 in ("traffic_bytes", "pipe_fma_pct", "pipe_xu_pct", "issue_active_pct", "registers", "ncu_bs")} if rec else None,
        }
    dom = max(per_kernel, key=lambda nm: per_kernel[nm]["ms"])
    d = per_kernel[dom]
    traffic = (d["ncu"] or {}).get("traffic_bytes")
    if traffic and (d["ncu"] or {}).get("ncu_bs") and d["ncu"]["ncu_bs"] != bs:
        traffic = traffic * bs / d["ncu"]["ncu_bs"]
    roof = {"bound": "fp32_fma", "kernel": dom, "achieved": d["fp32_TFLOPs_useful"], "peak": fp32["peak_TFLOPs"], "unit": "TFLOP/s",
            "frac": d["fp32_frac_useful"], "traffic": traffic, "kernel_ms": d["ms"], "kernel_share_of_step": d["share_of_step"],
            "achieved_executed": d["fp32_TFLOPs_executed"], "frac_executed": d["fp32_frac_executed"],
            "note": "achieved = SURVEY 8d nominal (useful) flops per eval x evals per launch / CUDA-event time of the kernel inside bench.py; "
                    "peak = FP32 FMA peak MEASURED on this GPU by gl_fp32_peak in this run; *_executed = flops the kernel issues (ncu sass op counts, "
                    "profiles/) / the same time; traffic = ncu dram bytes per launch.  HBM view of the same kernel: "
                    f"{d['hbm_GBs']:.0f} GB/s = {d['hbm_frac']:.3f} of {hbm_peak:.0f} GB/s ({which}) -- not the binding ceiling"}
    whole = sum(flops_tab.values())
    return roof, per_kernel, {"flops_per_eval_nominal": whole, "TFLOPs": whole * bs / (step_ms * 1e-3) / 1e12,
                              "frac": whole * bs / (step_ms * 1e-3) / 1e12 / fp32["peak_TFLOPs"]}


def bench_c2(ctx, args):
    import numpy as np
    import torch

    from gigalens_b200 import _cabi, workloads
    from gigalens_b200.model import ForwardProbModel
    from gigalens_b200.simulator import LensSimulator

    lib, rank, world = ctx.lib, ctx.rank, ctx.world
    wl = workloads.c2_workload()
    bs = BS_PER_GPU
    sim = LensSimulator(wl["phys_model"], wl["sim_config"], bs=bs)
    pmod = ForwardProbModel(wl["prior"], wl["observed"], background_rms=wl["background_rms"], exp_time=wl["exp_time"])
    draw = wl["prior"].sample(bs, seed=rank)
    z_host = torch.as_tensor(pmod.bij_inverse(draw)).pin_memory()
    # mean EPL series length of this batch (per-sample count at the plan's default tolerance 1e-9, cap 50): enters the nominal flop count
    _e = np.hypot(np.asarray(draw["lens_mass"][0]["e1"], dtype=np.float64), np.asarray(draw["lens_mass"][0]["e2"], dtype=np.float64))
    _f = np.clip(_e, 1e-30, 1.0 - 1e-9)
    mean_trips = float(np.mean(np.clip(np.ceil(np.log(1e-9) / np.log(_f) + 2.0) - 1.0, 0, 50)))
    z = z_host.cuda()
    d = z.shape[1]
    assert d == PARAMS_PER_SAMPLE
    warm = max(args.warmup, 3)

    # ---- device-resident throughput (no per-kernel instrumentation inside this region)
    sampler = ClockSampler(ctx.local) if rank == 0 else None
    if sampler:
        sampler.start()                      # clocks are sampled under load: warm-up + timed regions
    step = lambda: pmod.log_prob_and_grad(sim, z)
    launches0 = lib.gl_launch_count()
    ms_max, out = _timed_steps(ctx, step, args.steps, warm)
    launches = (lib.gl_launch_count() - launches0) * args.steps // (args.steps + warm)
    assert bool(torch.isfinite(out[0]).all()), "non-finite log-prob in the benchmark batch"
    # ---- the same steps again with CUDA events around every kernel (kernel durations for the roofline)
    stage_ms = _stage_times(ctx, sim, step, args.steps)
    torch.cuda.synchronize()

    # ---- end to end through the C ABI with host buffers
    logp_h = torch.empty(bs, dtype=torch.float32).pin_memory()
    chi_h = torch.empty(bs, dtype=torch.float32).pin_memory()
    dz_h = torch.empty((bs, d), dtype=torch.float32).pin_memory()
    pmod._bind(sim)

    def e2e_step():
        _cabi.check(lib.gl_logprob_grad_host(sim._plan, z_host.data_ptr(), logp_h.data_ptr(), chi_h.data_ptr(), dz_h.data_ptr()), lib)

    for _ in range(3):
        e2e_step()
    _sync_all(ctx)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    e2e_s = _max_over_ranks(ctx, time.perf_counter() - t0)
    _sync_all(ctx)
    clocks = sampler.stop() if sampler else None
    assert np.allclose(logp_h.numpy(), out[0].cpu().numpy(), rtol=1e-6), "host-buffer path disagrees with device path"

    npix = (sim.numPix * sim.supersample) ** 2
    P = sim.numPix ** 2
    alg_bytes = {"k_raytrace_fwd": 4 * npix, "k_conv_fwd": 4 * npix + 8 * P, "k_conv_bwd": 4 * P + 4 * npix, "k_raytrace_bwd": 4 * npix + 4 * 144}
    roof, per_kernel, whole = build_roofline(ctx, sim, stage_ms, sum(stage_ms), c2_nominal_flops(npix, mean_trips), alg_bytes, "kernels", ctx.fp32)
    whole["mean_epl_trips"] = mean_trips
    value = bs * world * args.steps / (ms_max * 1e-3)
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": make_config(world),
        "e2e": {"value": bs * world * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": bs * d * 4,
                "d2h_bytes_per_step": bs * (d + 2) * 4, "api": "gl_logprob_grad_host (pinned host z -> logp, red_chi2, dz)"},
        "gpu_launches": int(launches),
        "clocks": clocks, "roofline": roof, "fp32_peak_measured": ctx.fp32, "roofline_kernels": per_kernel, "whole_step": whole,
        "kernel_ms": {STAGES[k]: stage_ms[k] for k in range(len(STAGES))}, "ms_per_step_sum_of_kernels": sum(stage_ms),
    }


def bench_c4(ctx, args):
    """BASELINE.json configs[3]: cluster model, log-prob + gradient, bs 1024 per GPU."""
    import torch

    from gigalens_b200 import workloads
    from gigalens_b200.model import ForwardProbModel
    from gigalens_b200.simulator import LensSimulator

    bs, steps = 1024, max(3, min(args.steps, 10))
    ctx.c4_obs = workloads.c4_observation()
    wl = workloads.c4_workload(observed=ctx.c4_obs)
    sim = LensSimulator(wl["phys_model"], wl["sim_config"], bs=bs)
    pmod = ForwardProbModel(wl["prior"], wl["observed"], background_rms=wl["background_rms"], exp_time=wl["exp_time"])
    z = torch.as_tensor(pmod.bij_inverse(wl["prior"].sample(bs, seed=ctx.rank)), device="cuda")
    step = lambda: pmod.log_prob_and_grad(sim, z)
    ms, out = _timed_steps(ctx, step, steps, 3)
    finite = float(torch.isfinite(out[0]).float().mean())
    stage_ms = _stage_times(ctx, sim, step, steps)
    npix, P = (sim.numPix * sim.supersample) ** 2, sim.numPix ** 2
    # SURVEY 8d nominal: per ss pixel NFW 50 + 30 dPIE members x 90 + shear 6 + SersicEllipse 45 forward, adjoint 2x; conv 2 P E^2 each way
    fwd = npix * (50.0 + 30 * 90.0 + 6.0 + 45.0)
    conv = 2.0 * P * 26 * 26
    flops = {"k_raytrace_fwd": fwd, "k_conv_fwd": conv, "k_conv_bwd": conv, "k_raytrace_bwd": 2.0 * fwd}
    alg_bytes = {"k_raytrace_fwd": 4 * npix, "k_conv_fwd": 4 * npix + 8 * P, "k_conv_bwd": 4 * P + 4 * npix, "k_raytrace_bwd": 4 * npix}
    roof, per_kernel, whole = build_roofline(ctx, sim, stage_ms, sum(stage_ms), flops, alg_bytes, "c4_kernels", ctx.fp32)
    # The taped forward kernel carries the member loop of BOTH directions (forward-mode Jacobian), so SURVEY's per-kernel split
    # (forward N x 2801, adjoint twice that) does not apply to it: its roofline is the EXECUTED flop rate (ncu op counts of the same
    # kernel, profiles/r02_c4_ncu_summary.md); the nominal count is applied to the whole step (`whole_step`).
    d = per_kernel.get(roof["kernel"], {})
    for nm in ("k_raytrace_fwd", "k_raytrace_bwd"):
        if nm in per_kernel:
            per_kernel[nm]["fp32_TFLOPs_useful"] = per_kernel[nm]["fp32_frac_useful"] = None
    if d.get("fp32_TFLOPs_executed"):
        roof.update(achieved=d["fp32_TFLOPs_executed"], frac=d["fp32_frac_executed"],
                    note="achieved = FP32 flops the kernel executes (ncu source-page op counts of the same kernel, scaled to this batch) / "
                         "CUDA-event time inside bench.py; peak = FP32 FMA peak measured in this run.  Half of the kernel's FP32 instructions "
                         "are FMUL2 / FADD2 (1 flop per lane-cycle) and the XU (MUFU) pipe is co-limiting (57 % busy, ncu): FMA pipe 76 % busy")
    else:
        roof.update(achieved=None, frac=None, note="no ncu op counts available (profiles/r02_traffic.json): see whole_step for the nominal rate")
    del sim
    torch.cuda.empty_cache()
    return {"workload": wl["name"] + " (BASELINE.json configs[3])", "batch_per_gpu": bs, "global_batch": bs * ctx.world, "steps": steps,
            "ms_per_step": ms / steps, "value": bs * ctx.world * steps / (ms * 1e-3), "unit": UNIT, "scaling": "weak",
            "finite_logp_fraction": finite, "roofline": roof, "roofline_kernels": per_kernel, "whole_step": whole,
            "nominal_ceiling_evals_per_s_per_gpu": ctx.fp32["peak_TFLOPs"] * 1e12 / whole["flops_per_eval_nominal"]}


def bench_c3(ctx, args):
    """BASELINE.json configs[2]: Shapelets(n_max=10) source through lstsq_simulate (BackwardProbModel), bs 2048 per GPU."""
    import torch

    from gigalens_b200 import workloads
    from gigalens_b200.model import BackwardProbModel
    from gigalens_b200.simulator import LensSimulator

    bs, steps = 2048, max(3, min(args.steps, 10))
    wl = workloads.c3_workload(n_max=10, observed=workloads.c3_observation(n_max=10))
    sim = LensSimulator(wl["phys_model"], wl["sim_config"], bs=bs)
    pmod = BackwardProbModel(wl["prior"], wl["observed"], wl["background_rms"], wl["exp_time"])
    z = torch.as_tensor(pmod.bij_inverse(wl["prior"].sample(bs, seed=ctx.rank)), device="cuda")
    step = lambda: pmod.log_prob_and_grad(sim, z)
    ms, out = _timed_steps(ctx, step, steps, 3)
    finite = float(torch.isfinite(out[0]).float().mean())
    flops_per_eval = 0.4e9   # depthwise conv of the 66 channels 321 M + components 9.8 M + Gram 31 M + adjoint
    rate_gpu = bs * steps / (ms * 1e-3)
    del sim
    torch.cuda.empty_cache()
    return {"workload": wl["name"] + " (BASELINE.json configs[2])", "batch_per_gpu": bs, "global_batch": bs * ctx.world, "steps": steps,
            "ms_per_step": ms / steps, "value": rate_gpu * ctx.world, "unit": UNIT, "scaling": "weak", "finite_logp_fraction": finite,
            "linear_components": 66,
            "whole_step": {"flops_per_eval_nominal": flops_per_eval, "TFLOPs": flops_per_eval * rate_gpu / 1e12,
                           "frac": flops_per_eval * rate_gpu / 1e12 / ctx.fp32["peak_TFLOPs"]},
            "nominal_ceiling_evals_per_s_per_gpu": ctx.fp32["peak_TFLOPs"] * 1e12 / flops_per_eval}


def bench_c5(ctx, args):
    """BASELINE.json configs[4]: SVI and HMC (ModellingSequence) on the cluster model, GLOBAL batch 16384 over the N ranks.
    Mirrors jax/inference.py:113-128 (per-step pmean of the ELBO gradient) and :175-202 (HMC)."""
    import numpy as np
    import torch

    from gigalens_b200 import workloads
    from gigalens_b200.inference import Adam, ModellingSequence
    from gigalens_b200.model import ForwardProbModel

    obs = getattr(ctx, "c4_obs", None)
    if obs is None:
        obs = workloads.c4_observation()
    wl = workloads.c4_workload(observed=obs)
    pmod = ForwardProbModel(wl["prior"], wl["observed"], background_rms=wl["background_rms"], exp_time=wl["exp_time"])
    seq = ModellingSequence(wl["phys_model"], pmod, wl["sim_config"])
    n = C5_GLOBAL_BATCH
    truth = wl["prior"].sample(1, seed=11)   # the draw c4_observation simulates: start the surrogate next to it
    z0 = pmod.bij_inverse(truth)[0] + 0.01
    svi_steps = args.c5_svi_steps
    hmc_burn = max(2, args.c5_hmc_steps // 3)           # adaptive steps: 80 % of them all-reduce the ChEES / step-size statistics
    hmc_res = max(1, args.c5_hmc_steps - hmc_burn - 1)
    out = {}
    # ---- SVI
    seq.time_collectives(True)
    seq.SVI(optimizer=Adam(1e-3), start_mean=z0, n_vi=n, num_steps=2, seed=2)   # warm-up: plan creation (the drivers reuse it), NCCL channels
    seq.time_collectives(True)
    _sync_all(ctx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    q_z, losses = seq.SVI(optimizer=Adam(1e-3), start_mean=z0, n_vi=n, num_steps=svi_steps, seed=2)
    e1.record()
    torch.cuda.synchronize()
    ms = _max_over_ranks(ctx, e0.elapsed_time(e1))
    coll_ms, n_coll = seq.collective_ms()
    coll_ms = _max_over_ranks(ctx, coll_ms)
    out["c5_svi"] = {"workload": "SVI (full-rank MVN surrogate) on the C4 cluster model, ModellingSequence.SVI (BASELINE.json configs[4])",
                     "global_batch": n, "batch_per_gpu": n // ctx.world, "steps": svi_steps, "ms_per_step": ms / svi_steps,
                     "value": n * svi_steps / (ms * 1e-3), "unit": UNIT, "scaling": "strong",
                     "includes": "per-step surrogate sampling, autograd of the ELBO surrogate and Adam in torch around the fused log-prob + gradient call",
                     "allreduce": {"count": n_coll, "floats_per_call": 1 + 16 + 16 * 17 // 2, "ms_total": coll_ms,
                                   "share_of_step": coll_ms / ms, "backend": "nccl" if ctx.world > 1 else "none (single rank)"},
                     "elbo_first_last": [losses[0], losses[-1]], "elbo_finite": bool(np.isfinite(losses).all())}
    # ---- HMC, two timed calls: (a) adaptive burn-in (dual averaging + ChEES: every step all-reduces its statistics; the
    # adapted step size decides the leapfrog count), (b) sampling at a fixed step size with 5 leapfrogs per step.
    phases = {}
    tot_ms = tot_evals = tot_coll = 0.0
    # warm-up, like the SVI call above: one adaptive and one sampling step (3 evaluations of the batch) take the first-use costs of the
    # torch side of the driver (cuSOLVER / cuBLAS handles of cholesky_ex and solve_triangular, the CUDA generator) out of the timed phases
    seq.HMC(q_z, n_hmc=n, max_leapfrog_steps=1, seed=3, init_eps=0.02, init_l=1000, num_burnin_steps=1, num_results=1)
    for tag, kw in (("adaptive_burnin", dict(init_eps=0.02, init_l=1000, num_burnin_steps=hmc_burn, num_results=1)),
                    ("sampling_5_leapfrogs", dict(init_eps=0.02, init_l=1000, num_burnin_steps=0, num_results=hmc_res))):
        seq.time_collectives(True)
        _sync_all(ctx)
        e0.record()
        samples, stats = seq.HMC(q_z, n_hmc=n, max_leapfrog_steps=5, seed=3, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = _max_over_ranks(ctx, e0.elapsed_time(e1))
        coll_ms, n_coll = seq.collective_ms()
        coll_ms = _max_over_ranks(ctx, coll_ms)
        evals = _sum_over_ranks(ctx, stats["n_evals"]) + n     # + the initial log-prob of the chains
        phases[tag] = {"steps": kw["num_burnin_steps"] + kw["num_results"], "leapfrogs_per_step": stats["num_leapfrog"], "evals": evals,
                       "ms": ms, "evals_per_s": evals / (ms * 1e-3),
                       "allreduce": {"count": n_coll, "ms_total": coll_ms, "share_of_time": coll_ms / ms},
                       "accept_prob_mean": float(np.mean(stats["accept_prob"])), "samples_finite": bool(torch.isfinite(samples).all())}
        tot_ms, tot_evals, tot_coll = tot_ms + ms, tot_evals + evals, tot_coll + coll_ms
    n_steps = sum(p["steps"] for p in phases.values())
    out["c5_hmc"] = {"workload": "HMC (SVI-preconditioned, dual averaging + ChEES) on the C4 cluster model, ModellingSequence.HMC (configs[4])",
                     "global_batch": n, "batch_per_gpu": n // ctx.world, "steps": n_steps, "evals": tot_evals, "ms_per_step": tot_ms / n_steps,
                     "value": tot_evals / (tot_ms * 1e-3), "unit": UNIT, "scaling": "strong",
                     "includes": "initial log-prob of the chains + momentum draws / leapfrog updates in torch around the fused log-prob + gradient call",
                     "warmup": "one untimed HMC call of 2 steps (3 batch evaluations) before the two timed phases",
                     "allreduce": {"ms_total": tot_coll, "share_of_step": tot_coll / tot_ms, "backend": "nccl" if ctx.world > 1 else "none (single rank)"},
                     "phases": phases}
    return out


def run_cuda(args):
    import torch

    rank, world, local = _dist_setup()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this repo has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    import torch.distributed as dist

    from gigalens_b200 import _cabi

    ctx = Ctx()
    ctx.rank, ctx.world, ctx.local, ctx.dist = rank, world, local, dist
    ctx.lib = _cabi.load()
    ctx.peaks = _load_json("MEASURED_PEAKS.json") or {}
    ctx.fp32 = measure_fp32_peak(ctx)
    t_start = time.perf_counter()
    line = bench_c2(ctx, args)
    extras = {}
    if not args.headline_only:
        for name, fn in (("c4", bench_c4), ("c3", bench_c3)):
            t0 = time.perf_counter()
            extras[name] = fn(ctx, args)
            extras[name]["wall_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        extras.update(bench_c5(ctx, args))
        extras["c5_svi"]["wall_s_svi_plus_hmc"] = time.perf_counter() - t0
    launches_total = int(ctx.lib.gl_launch_count())
    if rank != 0:
        return
    cpu = None
    if world == 1:  # the CPU baseline is reported at N=1 only (rank 0)
        rate, med, cores = cpu_reference_rate(128, 5)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"bs=128 slice of the bs={BS_PER_GPU} batch, median of 5 after 1 warm-up (torch-CPU fp32 oracle port with autograd, {cores} host threads)"}
    line["cpu_baseline"] = cpu
    line.update(extras)
    line["gpu_launches_whole_run"] = launches_total
    line["wall_s"] = time.perf_counter() - t_start
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--headline-only", action="store_true", help="skip the c3 / c4 / c5 extra keys")
    ap.add_argument("--c5-svi-steps", type=int, default=50)
    ap.add_argument("--c5-hmc-steps", type=int, default=16)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    try:
        main()
    finally:
        try:
            import torch.distributed as _dist
            if _dist.is_available() and _dist.is_initialized():
                _dist.destroy_process_group()
        except Exception:
            pass
