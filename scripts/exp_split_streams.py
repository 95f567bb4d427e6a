"""Experiment: does running the C2 log-prob + gradient step as K sub-batches on K streams (one plan each) beat one plan over the
whole batch?  The per-sample kernels (unconstrain / prep / sample_bwd, ~80 us on 32 SMs) and every kernel's last wave would overlap
the other sub-batch's heavy kernels.  usage: python scripts/exp_split_streams.py [bs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gigalens_b200 import workloads
from gigalens_b200.model import ForwardProbModel
from gigalens_b200.simulator import LensSimulator

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
wl = workloads.c2_workload()
z_all = None
for K in (1, 2, 4, 1, 2, 4):
    nb = bs // K
    sims = [LensSimulator(wl["phys_model"], wl["sim_config"], bs=nb) for _ in range(K)]
    pms = [ForwardProbModel(wl["prior"], wl["observed"], background_rms=0.2, exp_time=100.0) for _ in range(K)]
    if z_all is None:
        z_all = torch.as_tensor(pms[0].bij_inverse(wl["prior"].sample(bs, seed=0)), device="cuda")
    zs = [z_all[k * nb:(k + 1) * nb].contiguous() for k in range(K)]
    streams = [torch.cuda.Stream() for _ in range(K)]

    def step():
        cur = torch.cuda.current_stream()
        outs = []
        for k in range(K):
            streams[k].wait_stream(cur)
            with torch.cuda.stream(streams[k]):
                outs.append(pms[k].log_prob_and_grad(sims[k], zs[k]))
        for k in range(K):
            cur.wait_stream(streams[k])
        return outs

    for _ in range(5): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 40
    e0.record()
    for _ in range(n): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"K={K}: {ms:.4f} ms/step  {bs / ms * 1e3 / 1e6:.4f} M evals/s", flush=True)
    del sims, pms
