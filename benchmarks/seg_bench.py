#!/usr/bin/env python
"""Segmentation-loss kernels alone (utils/loss_func.py:5-32 is what train_stage calls every step): the op is captured
into a CUDA graph (so the Python wrapper's allocations and launch overhead stay out of the number: at B=16 they are
several times the kernel) and ONE replay is timed with CUDA events after an L2 flush, median of --iters: forward
(tile/strip kernel + finalize + the two 8-float copies of the wrapper) and forward+backward, for the
strip kernel and for the 64x64 tile kernel (COR_SEG_STRIP=0), against the HBM peak on ALGORITHMIC bytes
(18 B/pixel with bf16 logits and 4 fp32 taps) and on the DRAM-sector floor (rows 4y+1, 4y+2 whole: 34 B/pixel).

    python benchmarks/seg_bench.py > profiles/rNN_seg_bench.jsonl
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cor_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=15)
ap.add_argument("--batches", default="16,128")
args = ap.parse_args()
HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")


def timed(fn):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fn()
    fn = graph.replay
    fn()
    ts = []
    for _ in range(args.iters):
        flush.add_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


g = torch.Generator(device="cuda").manual_seed(0)
for B in [int(b) for b in args.batches.split(",")]:
    pred = torch.randn(B, 1, 256, 256, device="cuda", generator=g).bfloat16()
    full = (torch.rand(B, 1, 1024, 1024, device="cuda", generator=g) > 0.5).float()
    cases = {"f32_4x": full, "u8_4x": (full * 255).to(torch.uint8), "f32_same": torch.nn.functional.avg_pool2d(full, 4)}
    for name, mask in cases.items():
        esz, taps = (4 if mask.dtype == torch.float32 else 1), (4 if mask.shape[-1] == 1024 else 1)
        alg = B * 256 * 256 * (2 + taps * esz)
        sect = B * 256 * 256 * 2 + (mask.numel() * esz // (2 if taps == 4 else 1))
        for strip in (1, 0):
            os.environ["COR_SEG_STRIP"] = "2" if strip else "0"
            t_f = timed(lambda: ops.seg_loss(pred, mask))
            pg = pred.detach().requires_grad_(True)
            t_fb = timed(lambda: ops.seg_loss(pg, mask).backward())
            print(json.dumps({"case": f"seg_loss B={B} {name}", "kernel": "strip" if strip else "tile", "fwd_us": round(t_f * 1e6, 2),
                              "fwd_bwd_us": round(t_fb * 1e6, 2), "algorithmic_bytes": alg, "sector_floor_bytes": sect,
                              "hbm_frac_algorithmic": round(alg / t_f / 1e9 / HBM, 3), "hbm_frac_sector_floor": round(sect / t_f / 1e9 / HBM, 3)}),
                  flush=True)
os.environ.pop("COR_SEG_STRIP", None)
