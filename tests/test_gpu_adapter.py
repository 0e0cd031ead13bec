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


@pytest.mark.parametrize("shape", [(3, 24, 24, 256), (2, 9, 30, 64), (1, 40, 17, 32), (20, 24, 24, 256), (37, 12, 12, 32)])
def test_depthwise_7x7_channels_last_vs_torch(shape):
    """cor_dwconv7_cl forward / input gradient / weight gradient against F.conv2d(groups=C) in fp32.  The last two shapes
    give every persistent CTA several (image, band) items, so the two-stage cp.async ring wraps."""
    import torch.nn.functional as F
    from cor_b200.linear import dwconv7_rows
    n, h, w, C = shape
    g = torch.Generator(device=dev()).manual_seed(h * w)
    x = torch.randn(n * h * w, C, device=dev(), generator=g, requires_grad=True)
    wt = (0.2 * torch.randn(C, 1, 7, 7, device=dev(), generator=g)).requires_grad_(True)
    b = torch.randn(C, device=dev(), generator=g, requires_grad=True)
    y = dwconv7_rows(x, wt, b, n, h, w)
    gy = torch.randn_like(y)
    y.backward(gy)
    # float64 reference: the sums of the weight gradient run over n*h*w terms, an fp32 reference is itself off by 1e-4 relative
    x2, w2, b2 = (t.detach().double().requires_grad_(True) for t in (x, wt, b))
    ref = F.conv2d(x2.view(n, h, w, C).permute(0, 3, 1, 2), w2, b2, padding=3, groups=C).permute(0, 2, 3, 1).reshape(n * h * w, C)
    ref.backward(gy.double())
    torch.testing.assert_close(y.double(), ref, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(x.grad.double(), x2.grad, rtol=1e-4, atol=1e-4)
    scale = float(w2.grad.abs().max())
    torch.testing.assert_close(wt.grad.double(), w2.grad, rtol=1e-4, atol=1e-5 * scale)
    torch.testing.assert_close(b.grad.double(), b2.grad, rtol=1e-4, atol=1e-5 * float(b2.grad.abs().max()))


@pytest.mark.parametrize("shape", [(1000, 512), (77, 256), (33, 1000), (50, 250)])
@pytest.mark.parametrize("act", [None, "gelu"])
def test_ln_rows_forward_backward_vs_torch(shape, act):
    import torch.nn.functional as F
    from cor_b200.linear import ln_rows
    rows, C = shape
    g = torch.Generator(device=dev()).manual_seed(rows)
    x = (2 * torch.randn(rows, C, device=dev(), generator=g) + 0.5).requires_grad_(True)
    wt = (1 + 0.2 * torch.randn(C, device=dev(), generator=g)).requires_grad_(True)
    b = (0.3 * torch.randn(C, device=dev(), generator=g)).requires_grad_(True)
    y = ln_rows(x, wt, b, 1e-6, act)
    gy = torch.randn_like(y)
    y.backward(gy)
    x2, w2, b2 = (t.detach().clone().requires_grad_(True) for t in (x, wt, b))
    ref = F.layer_norm(x2, (C,), w2, b2, 1e-6)
    if act == "gelu":
        ref = F.gelu(ref)
    ref.backward(gy)
    torch.testing.assert_close(y, ref, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(x.grad, x2.grad, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(wt.grad, w2.grad, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(b.grad, b2.grad, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("shape", [(2, 3, 8, 768, 24, 24), (3, 16, 8, 128, 16, 24), (1, 70, 2, 64, 8, 8)])
def test_pool_tail_on_tensor_cores_vs_streaming_kernel(shape):
    """softmax_map_pool_tail: the batched-GEMM path (csrc/adapter_tail.cu + gemm_umma.cu: one bf16 weight row per mask)
    against the exact fp32 streaming engine, forward and both gradients (mask_adapter.py:62-79)."""
    from cor_b200 import mask_adapter as ma, ops
    B, Q, G, C, h, w = shape
    g = torch.Generator(device=dev()).manual_seed(B * Q + C)
    maps = (2.0 * torch.randn(B, Q * G, h, w, device=dev(), generator=g)).requires_grad_(True)
    feat = torch.randn(B, C, h, w, device=dev(), generator=g, requires_grad=True)
    gy = torch.randn(B, Q, C, device=dev(), generator=g)
    assert ma._tail_gemm_ok(maps, feat, G)
    got = ma.softmax_map_pool_tail(maps, feat, G)
    got.backward(gy)
    m2, f2 = maps.detach().clone().requires_grad_(True), feat.detach().clone().requires_grad_(True)
    want = ops.region_pool(f2, m2, transform=ops.W_SIGMOID, normalize=False, group=G, eps=0.0, engine="stream").fg
    want.backward(gy)
    assert got.shape == want.shape == (B, Q, C)
    assert rel(got, want) < 4e-3, rel(got, want)                # bf16 weights and features, fp32 accumulation
    assert rel(feat.grad, f2.grad) < 6e-3, rel(feat.grad, f2.grad)
    assert rel(maps.grad, m2.grad) < 1e-2, rel(maps.grad, m2.grad)


@pytest.mark.parametrize("shape", [(5, 4, 48, 48), (3, 16, 24, 24), (2, 7, 5, 9), (1, 32, 3, 3)])
@pytest.mark.parametrize("act", [None, "gelu"])
def test_ln_channels_first_forward_backward_vs_reference_layernorm(shape, act):
    """cor_ln_cf_fwd / bwd against the module's own channels-first LayerNorm (mean / pow / sqrt over dim 1,
    mask_adapter.py:240-251) + nn.GELU in fp32 eager."""
    import torch.nn.functional as F
    from cor_b200.linear import ln_channels_first
    from cor_b200.mask_adapter import LayerNorm
    N, C, H, W = shape
    g = torch.Generator(device=dev()).manual_seed(N * C + H)
    x = (1.5 * torch.randn(N, C, H, W, device=dev(), generator=g) + 0.3).requires_grad_(True)
    ln = LayerNorm(C, eps=1e-6, data_format="channels_first").to(dev())
    with torch.no_grad():
        ln.weight.copy_(0.8 + 0.4 * torch.rand(C, device=dev(), generator=g))
        ln.bias.copy_(0.2 * torch.randn(C, device=dev(), generator=g))
    y = ln_channels_first(x, ln.weight, ln.bias, ln.eps, act)
    gy = torch.randn(N, C, H, W, device=dev(), generator=g)
    y.backward(gy)
    got = (y.detach(), x.grad.clone(), ln.weight.grad.clone(), ln.bias.grad.clone())
    x2 = x.detach().clone().requires_grad_(True)
    ln.weight.grad = ln.bias.grad = None
    ref = ln(x2)
    if act == "gelu":
        ref = F.gelu(ref)
    ref.backward(gy)
    torch.testing.assert_close(got[0], ref.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(got[1], x2.grad, rtol=1e-3, atol=1e-5)
    torch.testing.assert_close(got[2], ln.weight.grad, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(got[3], ln.bias.grad, rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("case", [(6, 24, 24, 96, 96), (3, 16, 24, 64, 96), (2, 27, 27, 108, 108)])
def test_mask_resample_kernel_upsamples_like_interpolate(case):
    """The x4 bilinear up-sampling at the head of mask_downscaling (mask_adapter.py:157) through cor_mask_prep."""
    import torch.nn.functional as F
    from cor_b200 import ops
    n, hm, wm, h, w = case
    g = torch.Generator(device=dev()).manual_seed(n + hm)
    m = torch.rand(n, hm, wm, device=dev(), generator=g)
    got = ops.mask_prep(m, (h, w))[0].view(n, h, w)
    want = F.interpolate(m[:, None], size=(h, w), mode="bilinear", align_corners=False)[:, 0]
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-6)
