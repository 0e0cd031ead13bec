"""Single sim_umma launch at the tensor-bound sweep point (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cor_b200 import ops
Nq, Nr = int(sys.argv[1]), int(sys.argv[2])
what = sys.argv[3] if len(sys.argv) > 3 else "lse"
g = torch.Generator(device="cuda").manual_seed(0)
R = torch.nn.functional.normalize(torch.randn(Nr, 256, device="cuda", generator=g), dim=-1).bfloat16()
Q = torch.nn.functional.normalize(torch.randn(Nq, 256, device="cuda", generator=g), dim=-1).bfloat16()
for _ in range(3):
    ops._sim_forward(R, Q, 1 / 0.07, what == "S", what == "lse", "umma")
torch.cuda.synchronize()
print("ok")
