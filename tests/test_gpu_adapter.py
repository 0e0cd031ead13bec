"""The learned half of MaskAdapterPooling on our kernels -- ``cor_b200.mask_adapter.adapter_maps``:
channels-last rows, tcgen05 GEMMs with fused epilogues, ``cor_ln_rows`` -- against the REFERENCE module itself
(``oracle/_ref/lib/support_model/mask_adapter.py``, the unmodified copy) run on the same GPU in fp32 with the same
parameters: activation maps, the pooled output of the whole ``MaskAdapterPooling.forward``, and the gradients of every
parameter and of the feature map.  Run with ``-m gpu`` on a B200."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-20))


def _pair(C=768, cin=512, mid=256, nmaps=8):
    from cor_b200.mask_adapter import MaskAdapterPooling
    from oracle import ref_step
    if not ref_step.available():
        pytest.skip("oracle/_ref not built")
    Ref = ref_step.module("lib/support_model/mask_adapter.py").MaskAdapterPooling
    torch.manual_seed(3)
    kw = dict(x_in_channel=C, mask_adatpet_network_in_channel=cin, mask_downscaling_mid_channel=16,
              mask_adatpet_network_mid_channel=mid, num_output_maps=nmaps)
    ref = Ref(**kw).to(dev())
    with torch.no_grad():                      # the shipped init (gamma = 1e-6, unit norms) would hide most of the network
        for n, p in ref.named_parameters():
            if n.endswith("gamma"):
                p.copy_(0.3 + 0.4 * torch.rand_like(p))
            elif "norm" in n and n.endswith("weight"):
                p.copy_(0.8 + 0.4 * torch.rand_like(p))
            elif n.endswith("bias"):
                p.copy_(0.1 * torch.randn_like(p))
    ours = MaskAdapterPooling(**kw).to(dev())
    ours.load_state_dict(ref.state_dict(), strict=True)
    return ref, ours


@pytest.mark.parametrize("shape", [(2, 3, 768, 24, 24), (1, 5, 768, 16, 24)])
def test_adapter_maps_and_pooled_output_vs_reference_module(shape):
    from cor_b200.mask_adapter import adapter_maps
    B, Q, C, h, w = shape
    ref, ours = _pair(C)
    g = torch.Generator(device=dev()).manual_seed(5)
    feat = torch.randn(B, C, h, w, device=dev(), generator=g)
    mask = (torch.rand(B, Q, h, w, device=dev(), generator=g) > 0.6).float()
    with torch.no_grad():
        want_maps = ref.get_mask_map(ref.channel_clip_to_maskadapter(feat), mask)
        got_maps = adapter_maps(ours, feat, mask)
        want = ref(feat, mask)
        got = ours(feat, mask)
    assert got_maps.shape == want_maps.shape == (B, Q * 8, h, w)
    assert rel(got_maps, want_maps) < 2e-2, rel(got_maps, want_maps)          # bf16 GEMM operands through 3 ConvNeXt blocks
    assert got.shape == want.shape == (B, Q, C)
    assert rel(got, want) < 5e-3, rel(got, want)


def test_adapter_backward_vs_reference_module():
    B, Q, C, h, w = 2, 2, 768, 24, 24
    ref, ours = _pair(C)
    g = torch.Generator(device=dev()).manual_seed(7)
    feat = torch.randn(B, C, h, w, device=dev(), generator=g)
    mask = (torch.rand(B, Q, h, w, device=dev(), generator=g) > 0.5).float()
    gy = torch.randn(B, Q, C, device=dev(), generator=g)
    f1 = feat.clone().requires_grad_(True)
    ours(f1, mask).backward(gy)
    f2 = feat.clone().requires_grad_(True)
    ref(f2, mask).backward(gy)
    assert rel(f1.grad, f2.grad) < 3e-2, rel(f1.grad, f2.grad)
    pr = dict(ref.named_parameters())
    worst = {}
    for n, p in ours.named_parameters():
        assert p.grad is not None, n
        worst[n] = rel(p.grad, pr[n].grad)
    flat = lambda ps: torch.cat([p.grad.reshape(-1) for p in ps])
    total = rel(flat([p for _, p in ours.named_parameters()]), flat([pr[n] for n, _ in ours.named_parameters()]))
    assert total < 3e-2, (total, sorted(worst.items(), key=lambda kv: -kv[1])[:5])
    assert max(worst.values()) < 2e-1, sorted(worst.items(), key=lambda kv: -kv[1])[:5]
