"""Hypernetwork product of the SAM decoder (csrc/hyper_logits.cu) against torch's own bmm formulation of
lib/sam_model/mask_decoder.py:135-137, forward and backward.  Run with ``-m gpu`` on a B200."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", [(3, 4, 32, 64, 64, torch.float32, (0, 4)), (2, 4, 32, 256, 256, torch.float32, (0, 1)),
                                  (2, 4, 32, 128, 64, torch.bfloat16, (1, 3)), (1, 2, 48, 20, 12, torch.float32, (0, 2))])
def test_hyper_mask_logits_forward_backward(case):
    from cor_b200.mask_decoder import hyper_mask_logits
    B, T, C, H, W, dt, tok = case
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(5)
    h = torch.randn(B, T, C, device=dev, generator=g, requires_grad=True)
    up = torch.randn(B, C, H, W, device=dev, generator=g).to(dt).requires_grad_(True)
    out = hyper_mask_logits(h, up, tok, torch.float32)
    go = torch.randn_like(out)
    out.backward(go)
    h2 = h.detach().clone().requires_grad_(True)
    up2 = up.detach().float().clone().requires_grad_(True)
    ref = (h2 @ up2.view(B, C, H * W)).view(B, -1, H, W)[:, tok[0]:tok[1]]            # mask_decoder.py:137 + :97-102
    ref.backward(go)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(h.grad, h2.grad, rtol=2e-4, atol=2e-3)
    tol = dict(rtol=1e-5, atol=1e-5) if dt == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(up.grad.float(), up2.grad, **tol)


def test_logits_and_seg_loss_matches_two_step_path():
    """bf16 logits of token 0 fed to the seg-loss kernel == the loss of the reference's fp32 product rounded to bf16."""
    from cor_b200 import ops
    from cor_b200.mask_decoder import logits_and_seg_loss
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(9)
    h = (0.3 * torch.randn(2, 4, 32, device=dev, generator=g)).requires_grad_(True)
    up = torch.randn(2, 32, 256, 256, device=dev, generator=g, requires_grad=True)
    mask = (torch.rand(2, 1, 1024, 1024, device=dev, generator=g) > 0.6).float()
    loss, logits = logits_and_seg_loss(h, up, mask)
    loss.backward()
    h2, up2 = h.detach().clone().requires_grad_(True), up.detach().clone().requires_grad_(True)
    ref_logits = (h2 @ up2.view(2, 32, -1)).view(2, 4, 256, 256)[:, 0:1]
    ref = ops.seg_loss(ref_logits.bfloat16(), mask)
    ref.backward()
    torch.testing.assert_close(logits.float(), ref_logits.bfloat16().float(), rtol=0, atol=1e-2)
    assert abs(float(loss) - float(ref)) < 2e-3 * abs(float(ref))
    assert float((h.grad - h2.grad).norm() / h2.grad.norm()) < 2e-2
    assert float((up.grad - up2.grad).norm() / up2.grad.norm()) < 2e-2
