"""BASELINE configs C4 / C5 sharded over the GPUs of one box (torchrun --nproc-per-node N tools/run_config_sharded.py C5).

Rank r solves the interleaved shard i = r (mod N) of the batch on its GPU (dispatches_b200.sweep.solve_sharded), one
all_gather assembles objective / status / iterations on every rank; rank 0 prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from dispatches_b200 import scenarios as SC, solver as S, sweep, templates as TP

cfg = sys.argv[1] if len(sys.argv) > 1 else "C5"
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
if cfg == "C5":
    t = TP.wind_battery(24); lmp, cf, w, b = SC.c5(); rp_all = TP.wind_battery_rparams(24, cf, w, b)
elif cfg == "C4":
    t = TP.fossil_surrogate(168); lmp = SC.c4(2000); rp_all = None
else:
    raise SystemExit("config must be C4 or C5")
sol = S.BatchLPSolver(t)
N = lmp.shape[0]
idx = sweep.shard_indices(N, rank, world)
cp = torch.tensor(lmp[idx], device=dev)
rp = torch.tensor(rp_all[idx], device=dev) if rp_all is not None else None
out = sol.solve(cp, rp)                      # warm-up
torch.cuda.synchronize()


def solve_fn(ix):
    assert np.array_equal(ix, idx)
    o = sol.solve(cp, rp, out=out)
    return dict(obj=o.obj, status=o.status, iters=o.iters)


if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
res = sweep.solve_sharded(solve_fn, N)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    st = res["status"].cpu().numpy()
    print(json.dumps({"config": cfg, "n_gpus": world, "N": int(N), "ms": float(ms), "lps": N / float(ms) * 1e3,
                      "non_optimal": int((st != 0).sum()), "iters_mean": float(res["iters"].float().mean()),
                      "obj_checksum": float(res["obj"].sum())}))
if world > 1:
    dist.destroy_process_group()
