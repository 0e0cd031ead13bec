"""Randomised cross-check of the two CPU oracles: the numpy restatement (oracle/np_oracle.py, written from the ATen
formulas) against the torch port (oracle/aten_port.py, which CALLS ATen).  The golden fixtures pin both to the
reference at a handful of shapes; this sweeps odd sizes and scale factors (non-integer, up-sampling, 1-pixel maps)
so that the resample index arithmetic, the 31x31 zero-padded box mean and the eps conventions agree everywhere."""
import numpy as np
import pytest
import torch

from cor_b200 import synth
from oracle import aten_port as ap
from oracle import np_oracle as no

SHAPES = [  # (B, C, h, w, H, W)
    (2, 8, 5, 7, 33, 41),      # non-integer down-scale, odd everything
    (1, 4, 9, 9, 6, 6),        # up-sampling (feature map finer than the mask)
    (3, 16, 1, 1, 17, 5),      # a single feature pixel
    (2, 8, 12, 10, 12, 10),    # identity resample
    (1, 8, 6, 4, 96, 64),      # exact 16x
]


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


@pytest.mark.parametrize("shape", SHAPES)
def test_pooling_and_fgbg_agree(shape):
    B, C, h, w, H, W = shape
    rng = np.random.default_rng(hash(shape) % (2 ** 31))
    emb = rng.standard_normal((B, C, h, w)).astype(np.float32)
    mask = (rng.random((B, 1, H, W)) * 1.4 - 0.2).astype(np.float32)      # outside [0,1]: the clamp matters
    comb = synth.unit_rows(rng, B, 1, C)
    np.testing.assert_allclose(no.bilinear_resize(mask, (h, w)), ap._resize(t(mask), (h, w)).numpy(), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(no.masked_pooling(emb, mask), ap.plain_masked_mean(t(emb), t(mask)).numpy(), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(no.mask_pooling(emb, mask), ap.unit_region_feature(t(emb), t(mask)).numpy(), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(no.fg_feat_similarity_loss(emb, comb, mask), float(ap.fg_loss(t(emb), t(comb), t(mask))), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(no.bg_feat_similarity_loss(emb, comb, mask), float(ap.bg_loss(t(emb), t(comb), t(mask))), rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("hw,HW", [((20, 20), (20, 20)), ((33, 47), (70, 90)), ((16, 64), (64, 256)), ((40, 8), (13, 5))])
def test_seg_loss_agrees(hw, HW):
    rng = np.random.default_rng(hw[0] * 131 + HW[1])
    pred = (2 * rng.standard_normal((2, 1, *hw))).astype(np.float32)
    mask = rng.random((2, 1, *HW)).astype(np.float32)
    target = no.bilinear_resize(mask, hw)
    np.testing.assert_allclose(no.wbce_with_wiou_loss(pred, target), float(ap.edge_weighted_seg_loss(t(pred), t(target))), rtol=2e-5)
    np.testing.assert_allclose(no.segmentation_loss(pred, mask), float(ap.seg_loss_fullres(t(pred), t(mask))), rtol=2e-5)


@pytest.mark.parametrize("hw,out", [((8, 8), (32, 32)), ((7, 9), (20, 31)), ((16, 16), None)])
def test_val_post_agrees(hw, out):
    rng = np.random.default_rng(hw[1])
    pred = (3 * rng.standard_normal((2, 1, *hw))).astype(np.float32)
    np.testing.assert_allclose(no.val_postprocess(pred, out), ap.val_post(t(pred), out).numpy(), rtol=1e-5, atol=1e-6)
