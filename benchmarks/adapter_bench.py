#!/usr/bin/env python
"""MaskAdapterPooling (lib/support_model/mask_adapter.py:28-80) forward and forward+backward: our path
(cor_b200.mask_adapter: tcgen05 GEMMs, LN / depth-wise kernels, feature half of `fuse` shared across masks) against the
REFERENCE module (oracle/_ref copy) run by PyTorch eager on the same B200, fp32 and under bf16 autocast (how the reference
trains, utils/trainer_v3_g.py:51).  CUDA events, median of --iters after warm-up.

    python benchmarks/adapter_bench.py > profiles/rNN_adapter_bench.jsonl
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cor_b200.mask_adapter import MaskAdapterPooling  # noqa: E402
from oracle import ref_step  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=7)
ap.add_argument("--cases", default="10x1,16x16,16x64")
args = ap.parse_args()
dev = torch.device("cuda:0")
kw = dict(x_in_channel=768, mask_adatpet_network_in_channel=512, mask_downscaling_mid_channel=16, mask_adatpet_network_mid_channel=256,
          num_output_maps=8)       # lib/support_branch.py:30-36
Ref = ref_step.module("lib/support_model/mask_adapter.py").MaskAdapterPooling if ref_step.available() else None


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(args.iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


for case in args.cases.split(","):
    B, Q = (int(v) for v in case.split("x"))
    torch.manual_seed(0)
    ours = MaskAdapterPooling(**kw).to(dev)
    feat = torch.randn(B, 768, 24, 24, device=dev)
    mask = (torch.rand(B, Q, 24, 24, device=dev) > 0.6).float()
    gy = torch.randn(B, Q, 768, device=dev)

    def fwd(m, autocast=False):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            return m(feat, mask)

    def fwd_bwd(m, autocast=False):
        m.zero_grad(set_to_none=True)
        f = feat.detach().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out = m(f, mask)
        out.float().backward(gy)

    rec = {"case": f"MaskAdapterPooling B={B} Q={Q} C=768 24x24", "ours_fwd_ms": round(timed(lambda: fwd(ours)), 3),
           "ours_fwd_bwd_ms": round(timed(lambda: fwd_bwd(ours)), 3)}
    if Ref is not None:
        ref = Ref(**kw).to(dev)
        ref.load_state_dict(ours.state_dict())
        try:
            rec["ref_eager_fp32_fwd_ms"] = round(timed(lambda: fwd(ref)), 3)
            rec["ref_eager_fp32_fwd_bwd_ms"] = round(timed(lambda: fwd_bwd(ref)), 3)
            rec["ref_eager_bf16_autocast_fwd_ms"] = round(timed(lambda: fwd(ref, True)), 3)
            rec["ref_eager_bf16_autocast_fwd_bwd_ms"] = round(timed(lambda: fwd_bwd(ref, True)), 3)
        except Exception as e:  # noqa: BLE001  (e.g. out of memory at the largest case)
            rec["ref_error"] = f"{type(e).__name__}: {e}"[:160]
        del ref
    print(json.dumps(rec), flush=True)
    del ours
    torch.cuda.empty_cache()
