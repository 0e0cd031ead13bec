"""The oracles (numpy restatement + ATen port) against golden vectors from the unmodified
reference (oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import aten_port as ap
from oracle import np_oracle as no

from conftest import load_golden

RT, AT = 2e-5, 2e-6


def close(a, b, rtol=RT, atol=AT):
    np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=rtol, atol=atol)


def tt(x, grad=False):
    return torch.from_numpy(np.ascontiguousarray(x)).clone().requires_grad_(grad)


@pytest.mark.parametrize("tag", ["int16", "so400m", "same"])
def test_masked_pooling(tag):
    g = load_golden(f"masked_pooling_{tag}")
    close(no.masked_pooling(g["feat"], g["mask"]), g["out"])
    f = tt(g["feat"], True)
    out = ap.plain_masked_mean(f, tt(g["mask"]))
    close(out.detach().numpy(), g["out"])
    out.backward(tt(g["gout"]))
    close(f.grad.numpy(), g["gfeat"])


def test_mask_adapter_tail():
    g = load_golden("mask_adapter_tail")
    G = int(g["num_output_maps"])
    close(no.mask_adapter_pool_tail(g["maps"], g["feat"], G), g["out"], rtol=1e-4, atol=1e-6)
    f, m = tt(g["feat"], True), tt(g["maps"], True)
    out = ap.softmax_map_pool(m, f, G)
    close(out.detach().numpy(), g["out"], rtol=1e-4, atol=1e-6)
    out.backward(tt(g["gout"]))
    close(f.grad.numpy(), g["gfeat"], rtol=1e-4, atol=1e-7)
    close(m.grad.numpy(), g["gmaps"], rtol=1e-4, atol=1e-7)


def test_mask_pooling():
    g = load_golden("mask_pooling")
    close(no.mask_pooling(g["emb"], g["mask"]), g["out"])
    e = tt(g["emb"], True)
    out = ap.unit_region_feature(e, tt(g["mask"]))
    close(out.detach().numpy(), g["out"])
    out.backward(tt(g["gout"]))
    close(e.grad.numpy(), g["gemb"], atol=1e-7)


@pytest.mark.parametrize("tag", ["b5", "b1"])
def test_fg_bg_losses(tag):
    g = load_golden(f"fgbg_{tag}")
    close(no.fg_feat_similarity_loss(g["emb"], g["comb"], g["mask"]), g["fg"])
    close(no.bg_feat_similarity_loss(g["emb"], g["comb"], g["mask"]), g["bg"])
    for name, fn in (("fg", ap.fg_loss), ("bg", ap.bg_loss)):
        e, c = tt(g["emb"], True), tt(g["comb"], True)
        v = fn(e, c, tt(g["mask"]))
        close(v.item(), g[name])
        v.backward()
        close(e.grad.numpy(), g["gemb_" + name], atol=1e-7)
        close(c.grad.numpy(), g["gcomb_" + name], atol=1e-7)


def test_bg_loss_is_not_the_paired_form():
    """Documents the reference's broadcasting at loss_func.py:120-123: the value differs from the
    per-sample cosine the docstring describes, and d(bg)/d(emb) is rounding noise."""
    g = load_golden("fgbg_b5")
    paired = no.bg_feat_similarity_loss_paired(g["emb"], g["comb"], g["mask"])
    assert abs(float(paired) - float(g["bg"])) > 1e-3
    assert np.abs(g["gemb_bg"]).max() < 1e-7


def test_all_invalid_returns_gradless_zero():
    g = load_golden("fgbg_allinvalid")
    z = np.zeros((2, 1, 64, 64), np.float32)
    assert float(g["fg"]) == 0.0 and float(g["bg"]) == 0.0
    assert not bool(g["fg_requires_grad"]) and not bool(g["bg_requires_grad"])
    assert float(no.fg_feat_similarity_loss(g["emb"], g["comb"], z)) == 0.0
    assert float(no.bg_feat_similarity_loss(g["emb"], g["comb"], z + 1)) == 0.0
    assert float(ap.fg_loss(tt(g["emb"]), tt(g["comb"]), tt(z))) == 0.0


@pytest.mark.parametrize("tag", ["sq64", "rect", "tiny"])
def test_wbce_wiou(tag):
    g = load_golden(f"wbce_wiou_{tag}")
    close(no.wbce_with_wiou_loss(g["pred"], g["mask"]), g["loss"])
    p = tt(g["pred"], True)
    v = ap.edge_weighted_seg_loss(p, tt(g["mask"]))
    close(v.item(), g["loss"])
    v.backward()
    close(p.grad.numpy(), g["gpred"], atol=1e-8)


def test_wbce_wiou_weights():
    g = load_golden("wbce_wiou_weights")
    close(no.wbce_with_wiou_loss(g["pred"], g["mask"], float(g["w1"]), float(g["w2"])), g["loss"])


def test_trainer_step():
    g = load_golden("trainer_step")
    close(no.bilinear_resize(g["masks"], g["pred"].shape[2:]), g["target"], atol=1e-6)
    close(no.segmentation_loss(g["pred"], g["masks"]), g["seg"])
    close(no.region_path_loss(g["pred"], g["emb"], g["comb"], g["masks"]), g["total"])
    p, e, c = tt(g["pred"], True), tt(g["emb"], True), tt(g["comb"], True)
    v = ap.trainer_loss(p, e, c, tt(g["masks"]))
    close(v.item(), g["total"])
    v.backward()
    close(p.grad.numpy(), g["gpred"], atol=1e-8)
    close(e.grad.numpy(), g["gemb"], atol=1e-7)
    close(c.grad.numpy(), g["gcomb"], atol=1e-7)


def test_val_post():
    g = load_golden("val_post")
    up = no.val_postprocess(g["pred"], (128, 128))
    close(up, g["post_up"], atol=1e-6)
    close(no.val_postprocess(g["pred"]), g["post_same"], atol=1e-6)
    agree = (no.binarize(up) == g["hard_up"]).mean()
    assert agree >= 0.999
    close(ap.val_post(tt(g["pred"]), (128, 128)).numpy(), g["post_up"], atol=1e-6)
    m = no.soft_metrics(g["post_up"], g["gt"])
    for k in ("dice", "mae", "iou", "mdice", "miou"):
        if k in g:
            close(m[k], g[k], rtol=1e-5)


def test_vailder_offline_binarisation():
    """utils/vailder.py:427-430,466,473 (sigmoid + min-max at the logit size, cv2.resize to the GT size, > 0.5)
    generated with the real cv2; the oracle's bilinear restatement must agree on >= 99.9 % of the pixels."""
    g = load_golden("vailder_hard")
    out = no.vailder_postprocess(g["pred"], tuple(int(v) for v in g["gt_hw"]))
    close(out, g["resized"], rtol=1e-5, atol=2e-6)
    assert (no.binarize(out) == g["hard"]).mean() >= 0.999


def test_reference_copy_step_equals_port():
    """oracle/_ref (the reference's own modules, copied by oracle/build_ref.py) composed into the bench step equals the
    ATen port: the two CPU baselines of bench.py are interchangeable, value and gradients."""
    import pytest
    import torch
    from cor_b200 import synth
    from oracle import aten_port as ap
    from oracle import ref_step
    if not ref_step.available():
        pytest.skip("oracle/_ref not built (no reference checkout on this machine)")
    assert ref_step.verified(), "oracle/_ref differs from its MANIFEST"
    d = synth.make_triplets(5, B=3, M=4, C=32, h=16, w=16, H=64, W=64, hp=32, wp=32)
    grads = []
    for fn in (ref_step.region_step_loss, ap.region_step_loss):
        t = {k: torch.from_numpy(v).clone().requires_grad_(k != "masks") for k, v in d.items()}
        loss, rows = fn(t["pred"], t["emb"], t["comb"], t["masks"], tau=0.07)
        loss.backward()
        grads.append((float(loss), rows.detach(), t["pred"].grad, t["emb"].grad, t["comb"].grad))
    assert abs(grads[0][0] - grads[1][0]) < 1e-6 * abs(grads[1][0])
    for a, b in zip(grads[0][1:], grads[1][1:]):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-7)
