"""tcgen05 GEMM (csrc/gemm_umma.cu) in every operand orientation, the fused linear layer's forward / backward, and the
composed-query head (cor_b200/support_head.py) against the reference's own CirFuseModule + dim_proj chain
(tests/golden/support_head.npz, generated from lib/support_model/cir_feature_fuse.py and lib/support_branch.py:47-54,60-86).
Run with ``-m gpu`` on a B200."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("case", [(300, 200, 136, False, False, 1), (304, 200, 136, True, False, 1), (300, 200, 136, False, True, 1),
                                  (136, 264, 72, True, True, 1), (16, 768, 1536, False, False, 1), (768, 1536, 16, True, True, 1),
                                  (576, 512, 768, True, False, 3), (1000, 256, 1024, False, False, 2), (4096, 1024, 256, False, False, 1),
                                  (1024, 256, 36864, True, True, 1)])
def test_gemm_orientations_vs_torch(case):
    """C = A B^T with A / B stored K-major or MN-major, batched, ragged tiles, automatic split-K; bf16 operands -> compare
    with an fp32 matmul of the same bf16-rounded values.  (An MN-major operand's contiguous extent must be a multiple of 8:
    TMA needs a 16-byte row pitch.)"""
    from cor_b200 import linear as lin
    M, N, K, a_mn, b_mn, batch = case
    g = torch.Generator(device=dev()).manual_seed(M + N + K)
    A = torch.randn(batch, M, K, device=dev(), generator=g).bfloat16()
    B = torch.randn(batch, N, K, device=dev(), generator=g).bfloat16()
    a2d = (A.transpose(1, 2).contiguous().view(batch * K, M) if a_mn else A.view(batch * M, K))
    b2d = (B.transpose(1, 2).contiguous().view(batch * K, N) if b_mn else B.view(batch * N, K))
    C = lin.gemm(a2d, b2d, M, N, K, a_mn=a_mn, b_mn=b_mn, batch=batch, a_batch_rows=(K if a_mn else M), b_batch_rows=(K if b_mn else N))
    ref = torch.bmm(A.float(), B.float().transpose(1, 2)).view(batch * M, N)
    torch.testing.assert_close(C, ref, rtol=2e-3, atol=2e-3 * float(K) ** 0.5)


def test_gemm_epilogue_bias_act_scale_residual_bf16_out():
    from cor_b200 import linear as lin
    g = torch.Generator(device=dev()).manual_seed(3)
    M, N, K = 200, 256, 192
    A = torch.randn(M, K, device=dev(), generator=g).bfloat16()
    B = (0.1 * torch.randn(N, K, device=dev(), generator=g)).bfloat16()
    bias = torch.randn(N, device=dev(), generator=g)
    scale = torch.rand(N, device=dev(), generator=g)
    res = torch.randn(M, N, device=dev(), generator=g)
    emul = (torch.rand(M, N, device=dev(), generator=g) > 0.5).float() * 2
    C, pre = lin.gemm(A, B, M, N, K, bias=bias, act=lin.ACT_GELU, emul=emul, colscale=scale, residual=res, out_dtype=torch.bfloat16, want_pre=True)
    z = A.float() @ B.float().t() + bias
    ref = F.gelu(z) * emul * scale + res
    torch.testing.assert_close(pre.float(), z, rtol=1e-2, atol=2e-2)
    torch.testing.assert_close(C.float(), ref, rtol=1e-2, atol=2e-2)


@pytest.mark.parametrize("act", [None, "relu", "gelu", "sigmoid"])
def test_linear_forward_backward_vs_torch(act):
    from cor_b200.linear import linear
    g = torch.Generator(device=dev()).manual_seed(11)
    rows, cin, cin2, cout = 24, 320, 192, 264
    x = torch.randn(rows, cin, device=dev(), generator=g, requires_grad=True)
    x2 = torch.randn(rows, cin2, device=dev(), generator=g, requires_grad=True)
    w = (0.05 * torch.randn(cout, cin + cin2, device=dev(), generator=g)).requires_grad_(True)
    b = torch.randn(cout, device=dev(), generator=g, requires_grad=True)
    mask = (torch.rand(rows, cout, device=dev(), generator=g) > 0.3).float() / 0.7 if act in ("relu", "gelu") else None
    y = linear(x, w, b, act, mask, x2)
    gy = torch.randn_like(y)
    y.backward(gy)
    xr, x2r, wr, br = (t.detach().clone().requires_grad_(True) for t in (x, x2, w, b))
    z = F.linear(torch.cat((xr, x2r), -1).bfloat16().float(), wr.bfloat16().float(), br)
    yr = {None: lambda t: t, "relu": F.relu, "gelu": F.gelu, "sigmoid": torch.sigmoid}[act](z)
    if mask is not None:
        yr = yr * mask
    yr.backward(gy)
    rel = lambda a, c: float((a - c).norm() / c.norm().clamp_min(1e-12))
    assert rel(y, yr) < 2e-3
    assert rel(x.grad, xr.grad) < 1e-2 and rel(x2.grad, x2r.grad) < 1e-2
    assert rel(w.grad, wr.grad) < 1e-2 and rel(b.grad, br.grad) < 1e-2


class _Branch(nn.Module):
    """The sub-modules composed_query uses, built exactly as lib/support_branch.py:42-54 builds them."""

    def __init__(self, dim, cir_cls=None):
        super().__init__()
        from cor_b200.mask_adapter import LayerNorm
        self.ln_channel_last = LayerNorm(normalized_shape=dim, eps=1e-6, data_format="channels_last")
        self.cir_fuse = cir_cls(image_embed_dim=dim, text_embed_dim=dim)
        self.dim_proj = nn.Sequential(nn.Linear(dim, 512), nn.GELU(), nn.Dropout(0.8), nn.Linear(512, 256), nn.GELU(), nn.Dropout(0.8))


def _ref_head(branch, support_feat, text_feat):
    """lib/support_branch.py:60-86 verbatim in behaviour, through the reference's CirFuseModule."""
    x = branch.ln_channel_last(support_feat).squeeze(1)
    rep = branch.cir_fuse.compose_img_text(x, text_feat.squeeze(1))["repres"]
    return F.normalize(branch.dim_proj(rep), p=2, dim=-1).unsqueeze(1)


@pytest.mark.parametrize("training", [False, True])
def test_composed_query_head_vs_reference_module(training):
    """Same parameters, same inputs, same seed: the head on our kernels against the reference's CirFuseModule
    (oracle/_ref copy) + dim_proj in fp32 -- eval mode, and train mode where the dropout masks must coincide."""
    from cor_b200.support_head import composed_query
    from oracle import ref_step
    if not ref_step.available():
        pytest.skip("oracle/_ref not built")
    Cir = ref_step.module("lib/support_model/cir_feature_fuse.py").CirFuseModule
    torch.manual_seed(5)
    dim, n = 768, 16
    branch = _Branch(dim, Cir).to(dev())
    branch.train(training)
    sf = torch.randn(n, 1, dim, device=dev(), requires_grad=True)
    tf_ = torch.randn(n, 1, dim, device=dev())
    torch.manual_seed(99)
    out = composed_query(branch, sf, tf_)
    gy = torch.randn_like(out)
    out.backward(gy)
    grads = {k: p.grad.clone() for k, p in branch.named_parameters()}
    g_sf = sf.grad.clone()
    branch.zero_grad()
    sf2 = sf.detach().clone().requires_grad_(True)
    torch.manual_seed(99)
    ref = _ref_head(branch, sf2, tf_)
    ref.backward(gy)
    rel = lambda a, c: float((a - c).norm() / c.norm().clamp_min(1e-12))
    assert out.shape == ref.shape == (n, 1, 256)
    assert rel(out, ref) < 1e-2, rel(out, ref)
    assert rel(g_sf, sf2.grad) < 3e-2, rel(g_sf, sf2.grad)
    # bf16 operands: every gradient within a few per cent of its own norm (the tiny ones -- the dynamic-scalar gate -- are
    # the noisiest), and the whole gradient vector within 2 %
    for k, p in branch.named_parameters():
        assert rel(grads[k], p.grad) < 1e-1, (k, rel(grads[k], p.grad))
    flat = lambda d: torch.cat([v.reshape(-1) for v in d])
    assert rel(flat([grads[k] for k, _ in branch.named_parameters()]), flat([p.grad for _, p in branch.named_parameters()])) < 2e-2
