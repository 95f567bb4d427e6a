#!/usr/bin/env python
"""bench.py -- log-prob + gradient evaluations per second on the C2 workload (BASELINE.json
configs[1]: EPL+shear lens, SersicEllipse lens light + source, 60x60, supersample 2, 13x13 PSF,
batch 4096 per GPU), one process per GPU.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A "step" is one pass of the hot path over one batch: ``ForwardProbModel.log_prob`` and its
gradient w.r.t. ``z`` for ``bs`` samples (bijector + prior, ray-shooting + light, PSF conv + pool,
chi^2 likelihood, and the hand adjoint of all of it).  ``value`` times it with ``z`` resident in
HBM; ``e2e`` times the same call through the C ABI's host-buffer entry point
(``gl_logprob_grad_host``: pinned host ``z`` in, ``logp`` / ``red_chi2`` / ``dz`` out, copies inside
the timed region).  Samples are independent, so N GPUs run N shards with no data-path collective
("weak" scaling); timing is CUDA events, max over ranks.
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


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
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


def _dist_setup(n_gpus):
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


def cpu_reference_rate(bs, reps, threads=None):
    """The reference's CPU path (oracle port: torch CPU fp32 + autograd) on a bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import torch

    import oracle_bridge
    from gigalens_b200 import workloads
    from gigalens_b200.model import ProbabilisticModel

    if threads:
        torch.set_num_threads(threads)
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
    rank, world, _ = _dist_setup(args.gpus)
    if rank != 0:
        return
    bs = 64
    # each "step" = one bounded sample of the workload: bs=64 of the 4096-sample batch
    t0 = time.perf_counter()
    rate, med, cores = cpu_reference_rate(bs, max(1, args.steps), None)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: EPL+shear / SersicEllipse x2, 60x60, ss=2, PSF 13x13, MAP logprob+grad",
                   "global_batch": BS_PER_GPU * args.gpus, "sample_batch": bs},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"bs={bs} slice of the bs=4096 batch per step, torch-CPU fp32 oracle port with autograd "
                                   f"(the reference's TF/JAX stack is not installable offline)"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


def run_cuda(args):
    import numpy as np
    import torch

    rank, world, local = _dist_setup(args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this repo has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    import torch.distributed as dist

    from gigalens_b200 import _cabi, workloads
    from gigalens_b200.model import ForwardProbModel
    from gigalens_b200.simulator import LensSimulator

    lib = _cabi.load()
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

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput -----------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()                      # clocks are sampled under load: warm-up + timed region
    for _ in range(max(args.warmup, 3)):
        out = pmod.log_prob_and_grad(sim, z)
    sync_all()
    pmod._bind(sim)
    sim.set_option("timing", min(args.steps, 256))     # CUDA events around every kernel of the timed steps
    launches0 = lib.gl_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = pmod.log_prob_and_grad(sim, z)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = lib.gl_launch_count() - launches0
    stage = (C.c_float * 7)()
    ncalls = C.c_int32(0)
    _cabi.check(lib.gl_plan_get_timings(sim._plan, stage, C.byref(ncalls)), lib)
    sim.set_option("timing", 0)
    stage_ms = [stage[k] / max(1, ncalls.value) for k in range(7)]
    sync_all()
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    assert bool(torch.isfinite(out[0]).all()), "non-finite log-prob in the benchmark batch"

    # ---- end to end through the C ABI with host buffers ---------------------------------------
    logp_h = torch.empty(bs, dtype=torch.float32).pin_memory()
    chi_h = torch.empty(bs, dtype=torch.float32).pin_memory()
    dz_h = torch.empty((bs, d), dtype=torch.float32).pin_memory()
    pmod._bind(sim)

    def e2e_step():
        _cabi.check(lib.gl_logprob_grad_host(sim._plan, z_host.data_ptr(), logp_h.data_ptr(), chi_h.data_ptr(),
                                             dz_h.data_ptr()), lib)

    for _ in range(3):
        e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    assert np.allclose(logp_h.numpy(), out[0].cpu().numpy(), rtol=1e-6), "host-buffer path disagrees with device path"

    # ---- roofline of the dominant kernel (ray-shooting adjoint), timed live --------------------
    roof = None
    if rank == 0:
        roof = kernel_roofline(sim, stage_ms, ms / args.steps, d, mean_trips)

    if rank != 0:
        return
    cpu = None
    if world == 1:  # the CPU baseline is reported at N=1 only (rank 0)
        rate, med, cores = cpu_reference_rate(32, 3)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "bs=32 slice of the batch, median of 3 (torch-CPU fp32 oracle port with autograd)"}
    value = bs * world * args.steps / (ms_max * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: EPL+shear / SersicEllipse x2, 60x60, ss=2, PSF 13x13, MAP logprob+grad (BASELINE.json configs[1])",
                   "global_batch": bs * world, "batch_per_gpu": bs, "params_per_sample": d, "parallelism": f"dp{world} (sample shards, no collective)",
                   "l2": "working set per step (236 MB ss image + adjoint) exceeds the 126 MB L2; no explicit flush"},
        "e2e": {"value": bs * world * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": bs * d * 4,
                "d2h_bytes_per_step": bs * (d + 2) * 4, "api": "gl_logprob_grad_host (pinned host z -> logp, red_chi2, dz)"},
        "gpu_launches": int(launches),
        "clocks": clocks, "roofline": roof["roofline"], "roofline_fp32": roof["fp32"], "kernel_ms": roof["kernel_ms"], "roofline_kernels": roof["per_kernel"],
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))


STAGES = ["k_unconstrain", "k_prep", "k_raytrace_fwd", "k_conv_fwd", "k_conv_bwd", "k_raytrace_bwd", "k_sample_bwd"]


def kernel_roofline(sim, stage_ms, step_ms, d, mean_trips):
    """Roofline of the dominant kernel (the ray-tracing adjoint) from its CUDA-event time measured on
    the launch stream inside the timed region (gl_plan_get_timings), with DESIGN.md's algorithmic
    bytes / flops per eval.  `traffic` is the ncu DRAM byte count of the same kernel (profiles/)."""
    peaks = _peaks() or {}
    bs = sim.bs
    npix = (sim.numPix * sim.supersample) ** 2
    dom = max(range(len(STAGES)), key=lambda k: stage_ms[k])
    name = STAGES[dom]
    # algorithmic HBM bytes per eval of each kernel: N = ss pixels, P = pixels
    P = sim.numPix ** 2
    alg_bytes = {"k_raytrace_fwd": 4 * npix, "k_conv_fwd": 4 * npix + 8 * P, "k_conv_bwd": 4 * P + 4 * npix,
                 "k_raytrace_bwd": 4 * npix + 4 * 144}.get(name, 4 * npix)
    # nominal flops per eval (SURVEY.md section 8d's hand count, FMA = 2): per ss pixel EPL 60 + 14 I, shear 6, SersicEllipse 45 x 2,
    # backward = 2 x forward, conv 4.87 M each way.  I = mean series length of THIS batch (SURVEY quotes I = 29, the batch-global
    # count of the reference; the kernels run per-sample counts, so I = 29 would overstate the work they do)
    fwd_flops = npix * (60.0 + 14.0 * mean_trips + 6.0 + 90.0)
    flops_tab = {"k_raytrace_fwd": fwd_flops, "k_conv_fwd": 4.87e6, "k_conv_bwd": 4.87e6, "k_raytrace_bwd": 2.0 * fwd_flops}
    step_flops = 3.0 * fwd_flops + 2 * 4.87e6
    alg_flops = flops_tab.get(name, 0.0)
    ms = stage_ms[dom]
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    sm_mhz = peaks.get("sm_max_mhz", 1965.0)
    fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    achieved_gbs = alg_bytes * bs / (ms * 1e-3) / 1e9
    traffic, ncu_pipe = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))["kernels"]
        rec = {}
        for key in (name + "_p", name + "_tma", name):   # kernel variants the C2 plan launches
            if key in tr:
                rec = tr[key]
                break
        traffic = rec.get("traffic_bytes")
        if "pipe_fma_pct" in rec:   # from the committed ncu --set full capture of the same kernel (profiles/)
            ncu_pipe = {"sm__pipe_fma_cycles_active_pct": rec["pipe_fma_pct"], "smsp__issue_active_pct": rec["issue_active_pct"]}
    except Exception:
        pass
    which = "of measured" if _peaks() else "of fallback"
    # every hot kernel against both ceilings (north_star: FP32 utilisation for the profile kernels, HBM GB/s for the conv / likelihood)
    all_bytes = {"k_raytrace_fwd": 4 * npix, "k_conv_fwd": 4 * npix + 8 * P, "k_conv_bwd": 4 * P + 4 * npix, "k_raytrace_bwd": 4 * npix + 4 * 144}
    all_flops = flops_tab
    per_kernel = {}
    try:
        tr_all = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))["kernels"]
    except Exception:
        tr_all = {}
    for k, nm in enumerate(STAGES):
        if nm not in all_bytes or stage_ms[k] <= 0:
            continue
        rec = next((tr_all[key] for key in (nm + "_p", nm + "_tma", nm) if key in tr_all), {})
        gbs = all_bytes[nm] * bs / (stage_ms[k] * 1e-3) / 1e9
        tf = all_flops[nm] * bs / (stage_ms[k] * 1e-3) / 1e12
        per_kernel[nm] = {"ms": stage_ms[k], "hbm_GBs": gbs, "hbm_frac": gbs / hbm_peak, "fp32_TFLOPs_nominal": tf,
                          "fp32_frac_nominal": tf / fp32_peak, "ncu_pipe_fma_pct": rec.get("pipe_fma_pct"),
                          "ncu_dram_bytes": rec.get("traffic_bytes"), "algorithmic_bytes": all_bytes[nm] * bs}
    return {
        "kernel_ms": {STAGES[k]: stage_ms[k] for k in range(len(STAGES))},
        "per_kernel": per_kernel,
        "roofline": {"bound": "hbm", "kernel": name, "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved_gbs / hbm_peak, "traffic": traffic, "kernel_ms": ms,
                     "kernel_share_of_step": ms / step_ms,
                     "note": f"algorithmic {alg_bytes} B/eval x {bs} evals per launch / CUDA-event time of the kernel; peak {which}. "
                             "The kernel is FP32-issue bound, not HBM bound (DESIGN.md 3.1): see roofline_fp32"},
        "fp32": {"bound": "fp32_fma", "kernel": name, "achieved": alg_flops * bs / (ms * 1e-3) / 1e12, "peak": fp32_peak,
                 "unit": "TFLOP/s", "frac": alg_flops * bs / (ms * 1e-3) / 1e12 / fp32_peak,
                 "note": "SURVEY 8d's nominal flops/eval formula (FMA=2) at the batch's mean series length; for the adjoint the kernel itself executes "
                         "fewer (closed-form EPL f-derivative, series state reused from the forward sweep), so this is a rate of useful work -- "
                         "the pipe occupancy actually measured is `ncu`.  peak = 148 SM x 128 lanes x 2 x max SM clock (derived, not measured)",
                 "ncu": ncu_pipe,
                 "mean_epl_trips": mean_trips,
                 "whole_step": {"achieved": step_flops * bs / (step_ms * 1e-3) / 1e12, "frac": step_flops * bs / (step_ms * 1e-3) / 1e12 / fp32_peak,
                                "flops_per_eval": step_flops}},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
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
