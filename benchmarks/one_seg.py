"""A few seg_loss forward launches (for ncu captures): python benchmarks/one_seg.py B"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cor_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
g = torch.Generator(device="cuda").manual_seed(0)
pred = torch.randn(B, 1, 256, 256, device="cuda", generator=g).bfloat16()
mask = (torch.rand(B, 1, 1024, 1024, device="cuda", generator=g) > 0.5).float()
for _ in range(3):
    ops.seg_loss(pred, mask)
torch.cuda.synchronize()
print("ok")
