#!/usr/bin/env python
"""Few-query InfoNCE backward (a rank's 16 queries against the gathered regions of W ranks: streaming kernel
`infonce_bwd_kernel`): CUDA-event time of the C-ABI call, median over --iters.  python benchmarks/nce_small_bench.py"""
import argparse, json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cor_b200 import ops, synth

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=30)
args = ap.parse_args()
dev = torch.device("cuda:0")
for W in (1, 2, 4, 8):
    Nq, Nr, D = 16, 1024 * W, 256
    g = synth.make_gallery(3, Nr, Nq, D=D)
    t = torch.from_numpy((np.arange(Nq) * 7) % Nr).to(dev)
    r = torch.from_numpy(g["regions"]).to(dev).requires_grad_(True)
    q = torch.from_numpy(g["queries"]).to(dev).requires_grad_(True)
    for _ in range(3):
        ops.infonce_loss(r, q, t, tau=0.07, engine="stream").backward()
    ops.TIMING["events"] = {}
    for _ in range(args.iters):
        ops.infonce_loss(r, q, t, tau=0.07, engine="stream").backward()
    torch.cuda.synchronize()
    ev = ops.TIMING["events"]
    ops.TIMING["events"] = None
    out = {"case": f"infonce stream Nq{Nq} Nr{Nr} (W={W})", "lib": os.environ.get("COR_B200_LIB", "default")}
    for name, pairs in ev.items():
        if "infonce" in name or "sim" in name:
            out[name + "_us"] = round(1000 * statistics.median(a.elapsed_time(b) for a, b in pairs), 2)
    print(json.dumps(out), flush=True)
