"""BASELINE.json configs[4]: VI + HMC pipeline on the cluster config, sample batch sharded over the
ranks (torchrun, one process per GPU), NCCL for the SVI gradient all-reduce and the HMC adaptation
statistics.  Prints one JSON line with log-prob+gradient evaluations per second of each stage.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      scripts/run_c5.py --batch 16384 --svi-steps 20 --hmc-steps 5
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from gigalens_b200 import workloads
from gigalens_b200.inference import Adam, ModellingSequence
from gigalens_b200.model import ForwardProbModel

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16384)
ap.add_argument("--num-pix", type=int, default=200)
ap.add_argument("--members", type=int, default=30)
ap.add_argument("--svi-steps", type=int, default=20)
ap.add_argument("--hmc-steps", type=int, default=5)
ap.add_argument("--leapfrogs", type=int, default=5)
args = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl")

obs = workloads.c4_observation(args.num_pix, args.members)
wl = workloads.c4_workload(args.num_pix, args.members, observed=obs)
prob = ForwardProbModel(wl["prior"], wl["observed"], background_rms=wl["background_rms"], exp_time=wl["exp_time"])
seq = ModellingSequence(wl["phys_model"], prob, wl["sim_config"])
start = prob.bij_inverse(wl["prior"].sample(1, seed=11))[0]     # the truth draw of c4_observation


def timed(fn):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1: dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return out, float(dt)

(q_z, losses), t_svi = timed(lambda: seq.SVI(Adam(2e-3), start_mean=start, n_vi=args.batch, init_scales=1e-3, num_steps=args.svi_steps))
(samples, stats), t_hmc = timed(lambda: seq.HMC(q_z, init_eps=0.1, init_l=args.leapfrogs, n_hmc=args.batch, num_burnin_steps=0,
                                                num_results=args.hmc_steps, max_leapfrog_steps=args.leapfrogs))
ne = torch.tensor([float(stats["n_evals"])], device="cuda")
if world > 1: dist.all_reduce(ne)
if rank == 0:
    print(json.dumps({
        "workload": wl["name"] + f", VI+HMC, batch {args.batch} over {world} GPU(s)",
        "n_gpus": world, "batch": args.batch,
        "svi": {"steps": args.svi_steps, "seconds": t_svi, "evals_per_s": args.batch * args.svi_steps / t_svi,
                "loss_first": losses[0], "loss_last": losses[-1], "collective": "1 all-reduce(sum) of 1 + d + d(d+1)/2 floats per step"},
        "hmc": {"steps": args.hmc_steps, "leapfrogs": stats["num_leapfrog"], "seconds": t_hmc, "evals_per_s": float(ne) / t_hmc,
                "accept_prob_mean": float(np.mean(stats["accept_prob"]))},
    }), flush=True)
if world > 1:
    dist.destroy_process_group()
