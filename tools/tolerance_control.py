#!/usr/bin/env python
"""fp32-vs-fp64 tolerance control for the update: per update, from an identical state, the element-wise
parameter deviation |dp| / max(|p|, 1e-3) of (a) the CUDA path and (b) the fp32 CPU oracle, both against the SAME update done in
double precision by the oracle, plus the CUDA-vs-fp32-oracle deviation in element-wise and max-norm form. Also the free-running
numbers (no re-synchronisation of the parameters between updates). Writes a markdown table and a JSON next to it.
Usage (GPU box): python tools/tolerance_control.py --updates 3 --out gpurun_out/r2_tolerance_control"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_util import run_baseline_updates  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--updates", type=int, default=3)
ap.add_argument("--envs", default="coordsum,lbf,rware")
ap.add_argument("--out", default="gpurun_out/r2_tolerance_control")
args = ap.parse_args()
torch.set_num_threads(os.cpu_count() or 1)
dev = torch.device("cuda:0")
res = {}
lines = ["Columns: max-norm = max |dp| / max |p| over the tensors with max |p| > 0.02; `lr` columns = max |dp| in units of the learning rate "
         "(2.5e-4; one update = 8 Adam steps of at most ~lr each); `viol` = max |dp| / (1e-4 |p| + 0.1 lr), the asserted bound (<= 1); "
         "element-wise = max |dp| / max(|p|, 1e-3). o32 / o64 = the CPU oracle's update in float32 / float64 from the identical state.", "",
         "| config (E=16, U=2, T=128, P=4, M=2) | mode | update | actions / rewards / obs identical | loss dev | cuda vs o32 max-norm | "
         "cuda vs o32 [lr] | cuda vs o64 [lr] | o32 vs o64 [lr] | cuda vs o64 viol | o32 vs o64 viol | cuda vs o64 element-wise | o32 vs o64 element-wise | "
         "worst tensor (cuda vs o64) |", "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
for env in args.envs.split(","):
    for resync in (True, False):
        rows = run_baseline_updates(env, dev, updates=args.updates, with_fp64=True, resync=resync)
        res[f"{env}/{'resync' if resync else 'free'}"] = rows
        for r in rows:
            ok = r["actions_exact"] and r["rewards_exact"] and r["obs_exact"]
            lines.append(f"| {env} | {'per-update (re-synced)' if resync else 'free-running'} | {r['update']} | {ok} | {r['loss_dev']:.1e} | "
                         f"{r['cuda_vs_o32_maxnorm']:.1e} | {r['cuda_vs_o32_lr']:.4f} | {r['cuda_vs_o64_lr']:.4f} | {r['o32_vs_o64_lr']:.4f} | "
                         f"{r['cuda_vs_o64_viol']:.3f} | {r['o32_vs_o64_viol']:.3f} | {r['cuda_vs_o64']:.1e} | {r['o32_vs_o64']:.1e} | "
                         f"`{r['cuda_vs_o64_tensor']}` |")
            print(lines[-1], flush=True)
os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
with open(args.out + ".md", "w") as f:
    f.write("\n".join(lines) + "\n")
with open(args.out + ".json", "w") as f:
    json.dump(res, f, indent=1)
