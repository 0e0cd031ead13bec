"""One forward + backward of MaskAdapterPooling on our path between cudaProfilerStart/Stop (for ncu launch lists):
python benchmarks/one_adapter.py B Q"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cor_b200.mask_adapter import MaskAdapterPooling
B, Q = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16, 16)
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = MaskAdapterPooling(x_in_channel=768, mask_adatpet_network_in_channel=512, mask_downscaling_mid_channel=16,
                       mask_adatpet_network_mid_channel=256, num_output_maps=8).to(dev)
feat = torch.randn(B, 768, 24, 24, device=dev)
mask = (torch.rand(B, Q, 24, 24, device=dev) > 0.6).float()
gy = torch.randn(B, Q, 768, device=dev)
def step():
    m.zero_grad(set_to_none=True)
    f = feat.detach().requires_grad_(True)
    m(f, mask).backward(gy)
for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
