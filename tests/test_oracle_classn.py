"""Class-N oracles (no reference implementation) reduce to the reference-pinned Class-R ones."""
import numpy as np
import torch

from cor_b200 import synth
from oracle import aten_port as ap
from oracle import np_oracle as no


def test_multi_mask_pool_row_equals_mask_pooling():
    d = synth.make_triplets(3, B=2, M=3, C=16, h=8, w=8, H=64, W=64, hp=16, wp=16, degenerate=False)
    rows = no.multi_mask_pool(d["emb"], d["masks"])
    for b in range(2):
        for m in range(3):
            ref = no.mask_pooling(d["emb"][b:b + 1], d["masks"][b:b + 1, m:m + 1])[0, 0]
            np.testing.assert_allclose(rows[b, m], ref, rtol=1e-6, atol=1e-7)
    t = ap.multi_mask_regions(torch.from_numpy(d["emb"]), torch.from_numpy(d["masks"])).numpy()
    np.testing.assert_allclose(t, rows, rtol=1e-5, atol=1e-6)


def test_similarity_target_column_is_the_reference_cosine():
    d = synth.make_triplets(4, B=3, M=2, C=16, h=8, w=8, H=64, W=64, hp=16, wp=16, degenerate=False)
    rows = no.multi_mask_pool(d["emb"], d["masks"]).reshape(6, 16)
    q = d["comb"][:, 0, :]
    S = no.region_query_similarity(rows, q, bf16_operands=False)
    for b in range(3):
        cos = no.cosine_similarity(rows[b * 2], q[b], axis=-1)
        np.testing.assert_allclose(S[b, b * 2], cos, rtol=1e-5, atol=1e-6)
    # fg loss == 1 - mean of that column (all masks valid here)
    fg = no.fg_feat_similarity_loss(d["emb"], d["comb"], d["masks"][:, 0:1])
    np.testing.assert_allclose(fg, 1 - np.mean([S[b, b * 2] for b in range(3)]), rtol=1e-5)


def test_infonce_limits_and_port_agreement():
    g = synth.make_gallery(5, 64, 8, D=32)
    t = np.arange(8) * 3
    # tau -> inf: uniform softmax, loss -> log(Nr)
    assert abs(float(no.infonce_loss(g["regions"], g["queries"], t, tau=1e6)) - np.log(64)) < 1e-3
    a = no.infonce_loss(g["regions"], g["queries"], t, tau=0.07)
    b = ap.infonce(torch.from_numpy(g["regions"]), torch.from_numpy(g["queries"]), torch.from_numpy(t), 0.07)
    np.testing.assert_allclose(a, b.item(), rtol=1e-5)


def test_topk_total_order_and_ties():
    g = synth.make_gallery(6, 128, 4, D=32, duplicate=True)
    idx, sc = no.topk_retrieve(g["regions"], g["queries"], 128)
    assert (np.diff(sc.astype(np.float64), axis=1) <= 0).all()
    # the duplicated rows (3 and 64) score identically and appear index-ascending
    for qi in range(4):
        p3, p64 = list(idx[qi]).index(3), list(idx[qi]).index(64)
        assert sc[qi, p3] == sc[qi, p64] and p3 + 1 == p64
    s = torch.from_numpy(no.region_query_similarity(g["regions"], g["queries"]))
    order = torch.sort(s.double(), dim=1, descending=True, stable=True).indices
    assert (order.numpy() == idx).all()


def test_bf16_rounding_matches_torch():
    x = np.random.default_rng(0).standard_normal(4096).astype(np.float32) * 3
    np.testing.assert_array_equal(no.to_bf16(x), torch.from_numpy(x).bfloat16().float().numpy())


def test_dice_focal_textbook():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, 1, 8, 8)).astype(np.float32)
    t = (rng.random((2, 1, 8, 8)) > 0.5).astype(np.float32)
    p = torch.sigmoid(torch.from_numpy(x))
    tt = torch.from_numpy(t)
    dice = (1 - (2 * (p * tt).sum((2, 3)) + 1) / (p.sum((2, 3)) + tt.sum((2, 3)) + 1)).mean()
    np.testing.assert_allclose(no.dice_loss(x, t), dice.item(), rtol=1e-5)
    bce = torch.nn.functional.binary_cross_entropy_with_logits(torch.from_numpy(x), tt, reduction="none")
    pt = p * tt + (1 - p) * (1 - tt)
    focal = ((0.25 * tt + 0.75 * (1 - tt)) * (1 - pt) ** 2 * bce).mean()
    np.testing.assert_allclose(no.focal_loss(x, t), focal.item(), rtol=1e-5)


def test_port_seg_variants_agree_with_numpy_restatement():
    """The torch restatement of the Class-N segmentation variants (oracle/aten_port.py) against the numpy one."""
    import torch
    from cor_b200 import synth
    from oracle import aten_port as ap
    rng = np.random.default_rng(5)
    t = synth.make_masks(rng, 3, 1, 48, 40, soft=True, degenerate=False)
    x = synth.make_logits(rng, 3, 48, 40)
    terms = ap._soft_terms(torch.from_numpy(x), torch.from_numpy(t))
    np.testing.assert_allclose(float(terms["dice"].mean()), float(no.dice_loss(x, t)), rtol=1e-5)
    np.testing.assert_allclose(float(terms["focal"].mean()), float(no.focal_loss(x, t)), rtol=1e-5)
    np.testing.assert_allclose(float((terms["wbce"] + terms["wiou"]).mean()), float(no.wbce_with_wiou_loss(x, t)), rtol=1e-5)
    np.testing.assert_allclose(float(ap.bce_with_dice_loss(torch.from_numpy(x), torch.from_numpy(t))),
                               float(terms["bce"].mean() + terms["dice"].mean()), rtol=1e-6)
