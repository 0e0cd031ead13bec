#!/usr/bin/env python
"""Many-query InfoNCE at BASELINE configs[3]'s global shape (1024 queries x 102 400 regions x 256): the forward kernel
(similarity + log-sum-exp), the tensor-core backward (dQ and dR, csrc/nce_bwd_umma.cu) as bare C-ABI calls, and the
public op forward+backward; CUDA events after an L2 flush, median of --iters.

    python benchmarks/nce_bench.py > profiles/rNN_nce_bench.jsonl
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cor_b200 import _lib as L  # noqa: E402
from cor_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--shapes", default="1024x102400,256x4096,4096x16384")
args = ap.parse_args()
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
TCB = pk.get("bf16_tflops", 1590.0)
flush = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")
lib = L.load()


def timed(fn):
    for _ in range(2):
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
D, tau = 256, 0.07
for shp in args.shapes.split(","):
    Nq, Nr = (int(v) for v in shp.split("x"))
    R32 = torch.nn.functional.normalize(torch.randn(Nr, D, device="cuda", generator=g), dim=-1)
    Q32 = torch.nn.functional.normalize(torch.randn(Nq, D, device="cuda", generator=g), dim=-1)
    R, Q = R32.bfloat16(), Q32.bfloat16()
    tg = (torch.arange(Nq, device="cuda") * 97) % Nr
    _, lse = ops._sim_forward(R, Q, 1 / tau, False, True, "umma")
    gl = torch.ones(1, device="cuda")
    gq = torch.empty(Nq, D, device="cuda")
    gr = torch.empty(Nr, D, device="cuda")
    work = ops._work(lib.cor_infonce_bwd_umma_work_bytes(Nq, Nr, D), R.device)
    flops = 2.0 * Nq * Nr * D
    t_f = timed(lambda: ops._sim_lse_parts(R, Q, 1 / tau, "umma"))
    t_q = timed(lambda: ops._call("cor_infonce_bwd_umma", R.device, ops.ptr(R), ops.ptr(Q), Nr, Nq, D, ops._f(1 / tau), ops.ptr(lse), ops.ptr(tg),
                                  ops.ptr(gl), ops._f(1.0), None, ops.ptr(gq), ops.ptr(work)))
    t_r = timed(lambda: ops._call("cor_infonce_bwd_umma", R.device, ops.ptr(R), ops.ptr(Q), Nr, Nq, D, ops._f(1 / tau), ops.ptr(lse), ops.ptr(tg),
                                  ops.ptr(gl), ops._f(1.0), ops.ptr(gr), None, ops.ptr(work)))

    def fb():
        r = R32.detach().requires_grad_(True)
        q = Q32.detach().requires_grad_(True)
        ops.infonce_loss(r, q, tg, tau=tau, regions_bf16=R).backward()

    t_fb = timed(fb)
    print(json.dumps({"case": f"infonce {Nq}x{Nr}x{D}", "fwd_lse_us": round(t_f * 1e6, 1), "bwd_dQ_us": round(t_q * 1e6, 1),
                      "bwd_dR_us": round(t_r * 1e6, 1), "op_fwd_bwd_us": round(t_fb * 1e6, 1),
                      "fwd_tflops": round(flops / t_f / 1e12, 1), "dQ_tflops": round(2 * flops / t_q / 1e12, 1),
                      "dR_tflops": round(2 * flops / t_r / 1e12, 1), "fwd_frac_burst": round(flops / t_f / 1e12 / TCB, 3),
                      "bwd_frac_burst": round(4 * flops / (t_q + t_r) / 1e12 / TCB, 3)}), flush=True)
