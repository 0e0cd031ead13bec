"""Segmentation loss on the GPU: the TMA row-strip kernel (csrc/seg_strip.cu) against the oracle and against the
64x64 tile kernel it replaces on the reference's shapes, plus the Class-N variants (dice / focal / plain) forward and
backward against their restatement in oracle/aten_port.py.  Run with ``-m gpu`` on a B200."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as tf

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def _inputs(seed, N, H, W, scale, mask_dtype, pred_dtype=torch.bfloat16, soft=True):
    from cor_b200 import synth
    rng = np.random.default_rng(seed)
    masks = synth.make_masks(rng, N, 1, H * scale, W * scale, soft=soft, degenerate=False)
    if N > 2:
        masks[1] = 0.0                      # an empty target
        masks[2] = 1.0                      # a full one
    pred = synth.make_logits(rng, N, H, W)
    m = torch.from_numpy(masks)
    if mask_dtype == torch.uint8:
        m = torch.round(m * 255).to(torch.uint8)
    else:
        m = m.to(mask_dtype)
    return torch.from_numpy(pred).to(pred_dtype), m


def _run(pred, mask, strip, **kw):
    from cor_b200 import ops
    os.environ["COR_SEG_STRIP"] = "2" if strip else "0"
    try:
        p = pred.to(dev()).requires_grad_(True)
        loss, extra = ops.seg_loss(p, mask.to(dev()), return_extras=True, **kw)
        loss.backward()
        torch.cuda.synchronize()
        return loss.detach().cpu(), extra.detach().cpu(), p.grad.float().cpu()
    finally:
        os.environ.pop("COR_SEG_STRIP", None)


@pytest.mark.parametrize("case", [
    (3, 256, 256, 4, torch.float32, torch.bfloat16),       # the shipped shape: 1024^2 fp32 mask, 256^2 bf16 logits
    (3, 256, 256, 4, torch.uint8, torch.bfloat16),         # uint8 transport
    (2, 256, 256, 4, torch.bfloat16, torch.float32),
    (5, 256, 256, 1, torch.float32, torch.float32),        # wbce_with_wiou_loss(pred, target) as the reference calls it
    (2, 80, 64, 4, torch.float32, torch.bfloat16),         # ragged strips (80 rows), narrow image
    (3, 100, 48, 1, torch.float32, torch.float32),
    (2, 40, 16, 4, torch.uint8, torch.float32),
    (17, 64, 128, 1, torch.float32, torch.bfloat16),       # many samples: larger strips
])
def test_strip_kernel_equals_tile_kernel(case):
    N, H, W, scale, mdt, pdt = case
    pred, mask = _inputs(7 + N + H, N, H, W, scale, mdt, pdt)
    a = _run(pred, mask, True)
    b = _run(pred, mask, False)
    torch.testing.assert_close(a[1], b[1], rtol=2e-6, atol=1e-7)           # all eight outputs incl. dice / focal / bce / iou / wdice
    torch.testing.assert_close(a[2], b[2], rtol=2e-5, atol=1e-9 if pdt == torch.float32 else 1e-7)


@pytest.mark.parametrize("mdt", [torch.float32, torch.uint8])
def test_strip_kernel_full_resolution_vs_oracle(mdt):
    """1024^2 mask -> 256^2 logits (trainer_v3_g.py:67-68) against the numpy oracle and the port's gradient."""
    from oracle import aten_port as ap
    from oracle import np_oracle as no
    pred, mask = _inputs(3, 3, 256, 256, 4, mdt, torch.float32)
    loss, _, g = _run(pred, mask, True)
    mf = mask.float() / 255.0 if mdt == torch.uint8 else mask.float()
    np.testing.assert_allclose(float(loss), float(no.segmentation_loss(pred.numpy(), mf.numpy())), rtol=3e-5)
    pc = pred.clone().requires_grad_(True)
    ap.seg_loss_fullres(pc, mf).backward()
    torch.testing.assert_close(g, pc.grad, rtol=2e-4, atol=1e-9)


@pytest.mark.parametrize("strip", [True, False])
@pytest.mark.parametrize("name", ["bce_with_iou_loss", "bce_with_dice_loss", "wbce_with_wdice_loss", "focal_loss_with_iou_loss"])
def test_class_n_variants_forward_backward(name, strip):
    """Drop-in names of the stale .pyc (Class N, parity unpinned by the reference): value and d/d pred vs the port."""
    from cor_b200 import loss_func as lf
    from oracle import aten_port as ap
    pred, mask = _inputs(29, 4, 96, 64, 1, torch.float32, torch.float32)
    os.environ["COR_SEG_STRIP"] = "2" if strip else "0"
    try:
        p = pred.to(dev()).requires_grad_(True)
        v = getattr(lf, name)(p, mask.to(dev()))
        (2.0 * v).backward()
    finally:
        os.environ.pop("COR_SEG_STRIP", None)
    pc = pred.clone().requires_grad_(True)
    ref = getattr(ap, name)(pc, mask)
    (2.0 * ref).backward()
    np.testing.assert_allclose(float(v), float(ref), rtol=3e-5)
    torch.testing.assert_close(p.grad.cpu(), pc.grad, rtol=5e-4, atol=1e-9)


def test_class_n_variant_with_full_resolution_mask():
    from cor_b200 import loss_func as lf
    from oracle import aten_port as ap
    pred, mask = _inputs(31, 2, 256, 256, 4, torch.float32, torch.float32)
    p = pred.to(dev()).requires_grad_(True)
    v = lf.wbce_with_wdice_loss(p, mask.to(dev()), smooth=2.0)
    v.backward()
    pc = pred.clone().requires_grad_(True)
    t = tf.interpolate(mask, size=(256, 256), mode="bilinear", align_corners=False)
    ref = ap.wbce_with_wdice_loss(pc, t, smooth=2.0)
    ref.backward()
    np.testing.assert_allclose(float(v), float(ref), rtol=3e-5)
    torch.testing.assert_close(p.grad.cpu(), pc.grad, rtol=5e-4, atol=1e-9)
