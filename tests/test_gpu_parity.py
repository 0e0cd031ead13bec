"""GPU parity: the CUDA path (through the C ABI) against the golden vectors of the unmodified
reference and against the CPU oracle on seeded inputs.  Run with ``-m gpu`` on a B200.

Tolerances (BASELINE.json north_star): similarity / loss values within 1e-3 relative (fp32
accumulate); top-k indices identical; binarised masks >= 99.9 % pixel agreement.  Most fp32 paths
are held to a much tighter 2e-5 here so that regressions show.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

RT, AT = 1e-3, 1e-3          # north_star tolerance
TIGHT = dict(rtol=3e-5, atol=3e-6)


def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def cu(x, grad=False, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x)).to(dev())
    if dtype is not None:
        t = t.to(dtype)
    return t.requires_grad_(grad)


def close(a, b, rtol=RT, atol=AT):
    a = a.detach().float().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    np.testing.assert_allclose(a.astype(np.float64), np.asarray(b, dtype=np.float64), rtol=rtol, atol=atol)


# ---------------------------------------------------------------------------------------- Class R
@pytest.mark.parametrize("tag", ["int16", "so400m", "same"])
def test_masked_pooling_golden(tag):
    from cor_b200.mask_adapter import MaskedPooling
    g = load_golden(f"masked_pooling_{tag}")
    f = cu(g["feat"], True)
    out = MaskedPooling()(f, cu(g["mask"]))
    assert out.shape == g["out"].shape
    close(out, g["out"], **TIGHT)
    out.backward(cu(g["gout"]))
    close(f.grad, g["gfeat"], **TIGHT)


def test_mask_adapter_tail_golden():
    from cor_b200.mask_adapter import softmax_map_pool_tail
    g = load_golden("mask_adapter_tail")
    f, m = cu(g["feat"], True), cu(g["maps"], True)
    out = softmax_map_pool_tail(m, f, int(g["num_output_maps"]))
    close(out, g["out"], rtol=1e-4, atol=1e-6)
    out.backward(cu(g["gout"]))
    close(f.grad, g["gfeat"], rtol=1e-4, atol=1e-6)
    close(m.grad, g["gmaps"], rtol=1e-3, atol=1e-7)


def test_mask_adapter_module_matches_oracle_with_shared_weights():
    """Whole MaskAdapterPooling module vs the ATen port of the tail: exactly, on the maps our learned half produces
    (cor_b200.mask_adapter.adapter_maps: bf16-operand GEMMs), and within the bf16 operand rounding on the maps of the module's
    eager fp32 definition.  tests/test_gpu_adapter.py checks the learned half against the reference module itself."""
    from cor_b200.mask_adapter import MaskAdapterPooling, adapter_maps
    from oracle import aten_port as ap
    torch.manual_seed(0)
    mod = MaskAdapterPooling(x_in_channel=48, mask_adatpet_network_in_channel=32, mask_downscaling_mid_channel=16,
                             mask_adatpet_network_mid_channel=32, num_output_maps=8).to(dev())
    feat = torch.randn(2, 48, 24, 24, device=dev())
    mask = (torch.rand(2, 3, 96, 96, device=dev()) > 0.5).float()
    out = mod(feat, mask)
    assert out.shape == (2, 3, 48)
    with torch.no_grad():
        m = torch.nn.functional.interpolate(mask, size=(24, 24), mode="bilinear", align_corners=False)
        maps = mod.get_mask_map(mod.channel_clip_to_maskadapter(feat), m)
        maps_ours = adapter_maps(mod, feat, m)
    close(out, ap.softmax_map_pool(maps_ours.cpu(), feat.cpu(), 8).numpy(), rtol=1e-4, atol=1e-5)
    close(out, ap.softmax_map_pool(maps.cpu(), feat.cpu(), 8).numpy(), rtol=2e-2, atol=2e-3)
    out.sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in mod.parameters())


def test_mask_pooling_golden():
    from cor_b200.loss_func import mask_pooling
    g = load_golden("mask_pooling")
    e = cu(g["emb"], True)
    out = mask_pooling(e, cu(g["mask"]))
    assert out.shape == g["out"].shape
    close(out, g["out"], **TIGHT)
    out.backward(cu(g["gout"]))
    close(e.grad, g["gemb"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("tag", ["b5", "b1"])
def test_fg_bg_losses_golden(tag):
    from cor_b200 import loss_func as lf
    g = load_golden(f"fgbg_{tag}")
    for name, fn in (("fg", lf.fg_feat_similarity_loss), ("bg", lf.bg_feat_similarity_loss)):
        e, c = cu(g["emb"], True), cu(g["comb"], True)
        v = fn(e, c, cu(g["mask"]))
        close(v, g[name], **TIGHT)
        v.backward()
        close(c.grad, g["gcomb_" + name], rtol=1e-4, atol=1e-6)
        ge = e.grad if e.grad is not None else torch.zeros_like(e)
        close(ge, g["gemb_" + name], rtol=1e-4, atol=1e-6)


def test_fg_bg_shared_pass_and_paired_mode():
    from cor_b200 import loss_func as lf
    from cor_b200 import ops
    from oracle import np_oracle as no
    g = load_golden("fgbg_b5")
    e, c, m = cu(g["emb"]), cu(g["comb"]), cu(g["mask"])
    n0 = ops.LAUNCHES["count"]
    fg = lf.fg_feat_similarity_loss(e, c, m)
    n1 = ops.LAUNCHES["count"]
    bg = lf.bg_feat_similarity_loss(e, c, m)
    assert ops.LAUNCHES["count"] == n1 and n1 > n0, "bg must reuse the pass the fg call made"
    close(fg, g["fg"], **TIGHT)
    close(bg, g["bg"], **TIGHT)
    _, bgp = lf.fg_bg_feat_similarity_loss(e, c, m, bg_mode=lf.BG_PAIRED)
    close(bgp, no.bg_feat_similarity_loss_paired(g["emb"], g["comb"], g["mask"]), **TIGHT)


def test_all_invalid_masks_give_zero():
    from cor_b200 import loss_func as lf
    g = load_golden("fgbg_allinvalid")
    z = torch.zeros(2, 1, 64, 64, device=dev())
    e, c = cu(g["emb"], True), cu(g["comb"], True)
    fg, _ = lf.fg_bg_feat_similarity_loss(e, c, z)
    _, bg = lf.fg_bg_feat_similarity_loss(e, c, z + 1)
    assert float(fg) == 0.0 and float(bg) == 0.0
    (fg + bg).backward()
    assert float(c.grad.abs().max()) == 0.0


@pytest.mark.parametrize("tag", ["sq64", "rect", "tiny"])
def test_wbce_wiou_golden(tag):
    from cor_b200.loss_func import wbce_with_wiou_loss
    g = load_golden(f"wbce_wiou_{tag}")
    p = cu(g["pred"], True)
    v = wbce_with_wiou_loss(p, cu(g["mask"]))
    close(v, g["loss"], **TIGHT)
    v.backward()
    close(p.grad, g["gpred"], rtol=1e-4, atol=1e-8)


def test_wbce_wiou_weights_and_extras():
    from cor_b200 import ops
    from cor_b200.loss_func import wbce_with_wiou_loss
    from oracle import np_oracle as no
    g = load_golden("wbce_wiou_weights")
    close(wbce_with_wiou_loss(cu(g["pred"]), cu(g["mask"]), float(g["w1"]), float(g["w2"])), g["loss"], **TIGHT)
    _, extra = ops.seg_loss(cu(g["pred"]), cu(g["mask"]), return_extras=True)
    close(extra[1], no.dice_loss(g["pred"], g["mask"]), rtol=1e-4, atol=1e-6)      # Class N
    close(extra[2], no.focal_loss(g["pred"], g["mask"]), rtol=1e-3, atol=1e-6)     # Class N


def test_trainer_step_golden():
    """utils/trainer_v3_g.py:67-73 through the drop-in functions, values and all three gradients."""
    from cor_b200 import loss_func as lf
    g = load_golden("trainer_step")
    p, e, c, qm = cu(g["pred"], True), cu(g["emb"], True), cu(g["comb"], True), cu(g["masks"])
    seg = lf.wbce_with_wiou_loss(p, qm)          # full-res mask: resample fused in the kernel
    close(seg, g["seg"], **TIGHT)
    total = seg + 5 * lf.fg_feat_similarity_loss(e, c, qm) + 5 * lf.bg_feat_similarity_loss(e, c, qm)
    close(total, g["total"], **TIGHT)
    total.backward()
    close(p.grad, g["gpred"], rtol=1e-4, atol=1e-8)
    close(e.grad, g["gemb"], rtol=1e-4, atol=1e-6)
    close(c.grad, g["gcomb"], rtol=1e-4, atol=1e-6)
    close(lf.region_path_loss(cu(g["pred"]), cu(g["emb"]), cu(g["comb"]), qm), g["total"], **TIGHT)


def test_val_post_golden():
    from cor_b200 import ops
    g = load_golden("val_post")
    r = ops.val_postprocess(cu(g["pred"]), size=(128, 128), gt=cu(g["gt"]), want_hard=True)
    close(r["post"], g["post_up"], rtol=1e-5, atol=2e-6)
    agree = (r["hard"].cpu().numpy() == g["hard_up"]).mean()
    assert agree >= 0.999, agree
    for i, k in enumerate(("dice", "mae", "iou", "mdice", "miou")):
        close(r["metrics"][:, i], g[k], rtol=1e-4, atol=1e-6)
    close(ops.val_postprocess(cu(g["pred"]))["post"], g["post_same"], rtol=1e-5, atol=2e-6)


# ---------------------------------------------------------------------- dtypes / shapes vs oracle
@pytest.mark.parametrize("mask_dtype", ["u8", "bf16", "f32"])
@pytest.mark.parametrize("feat_dtype", [torch.float32, torch.bfloat16])
def test_mask_pooling_dtypes(mask_dtype, feat_dtype):
    from cor_b200 import synth
    from cor_b200.loss_func import mask_pooling
    from oracle import np_oracle as no
    d = synth.make_triplets(11, B=3, M=1, C=64, h=32, w=32, H=512, W=512, hp=32, wp=32, soft=(mask_dtype == "u8"), degenerate=False)
    emb = torch.from_numpy(d["emb"]).to(feat_dtype)
    masks = d["masks"]
    if mask_dtype == "u8":
        m_dev = cu(np.round(masks * 255).astype(np.uint8))
    elif mask_dtype == "bf16":
        m_dev = cu(masks, dtype=torch.bfloat16)
    else:
        m_dev = cu(masks)
    out = mask_pooling(emb.to(dev()), m_dev)
    ref = no.mask_pooling(emb.float().numpy(), masks)
    close(out, ref, rtol=1e-4, atol=1e-5)


def test_mask_prep_stats_exact_for_hard_masks():
    from cor_b200 import ops, synth
    rng = np.random.default_rng(5)
    m = synth.make_masks(rng, 2, 6, 256, 320, degenerate=True)
    m[0, 1] = 0.0
    m[1, 2] = 1.0
    m[1, 3] = 1.0
    m[1, 3, 17, 5] = 254.0 / 255.0              # one not-quite-foreground pixel keeps bg valid
    flat = m.reshape(12, 256, 320)
    for t in (cu(flat), cu(np.round(flat * 255).astype(np.uint8))):
        w32, stats = ops.mask_prep(t, (16, 20), ops.W_CLAMP)
        s = stats.cpu().numpy()
        np.testing.assert_allclose(s[:, 0], flat.astype(np.float64).sum((1, 2)), rtol=1e-6)
        assert (s[:, 0] > 0).tolist() == (flat.sum((1, 2)) > 0).tolist()
        assert (s[:, 1] > 0).tolist() == ((1 - flat.astype(np.float64)).sum((1, 2)) > 0).tolist()
        from oracle import np_oracle as no
        close(w32.view(12, 16, 20), no.bilinear_resize(flat, (16, 20)), rtol=1e-5, atol=1e-6)


def test_seg_loss_bf16_logits_fullres_mask():
    from cor_b200 import synth
    from cor_b200.loss_func import segmentation_loss
    from oracle import np_oracle as no
    d = synth.make_triplets(13, B=3, M=1, C=8, h=8, w=8, H=512, W=512, hp=128, wp=128, soft=True, degenerate=False)
    pred16 = torch.from_numpy(d["pred"]).bfloat16()
    ref = no.segmentation_loss(pred16.float().numpy(), d["masks"])
    close(segmentation_loss(pred16.to(dev()), cu(d["masks"])), ref, rtol=1e-4, atol=1e-6)
    close(segmentation_loss(pred16.to(dev()), cu(np.round(d["masks"] * 255).astype(np.uint8))), ref, rtol=1e-4, atol=1e-6)


# ---------------------------------------------------------------------------------------- Class N
def test_pool_regions_multi_mask_vs_oracle():
    from cor_b200 import region, synth
    from oracle import np_oracle as no
    d = synth.make_triplets(17, B=2, M=5, C=32, h=16, w=16, H=256, W=256, hp=16, wp=16, degenerate=False)
    p = region.pool_regions(cu(d["emb"]), cu(d["masks"]), background=True, engine="stream")
    close(p.fg, no.multi_mask_pool(d["emb"], d["masks"]), rtol=1e-4, atol=1e-5)
    close(p.bg, no.multi_mask_pool(d["emb"], d["masks"], background=True), rtol=1e-4, atol=1e-5)


def test_similarity_and_reduction_to_reference_cosine():
    from cor_b200 import region, synth
    from oracle import np_oracle as no
    g = synth.make_gallery(19, 300, 20, D=256)
    S = region.region_query_similarity(cu(g["regions"]), cu(g["queries"]), engine="stream")
    ref = no.region_query_similarity(g["regions"], g["queries"])
    close(S, ref, rtol=1e-3, atol=2e-6)
    # the target column equals the cosine inside fg_feat_similarity_loss (bf16-rounded operands)
    d = synth.make_triplets(23, B=4, M=3, C=64, h=16, w=16, H=128, W=128, hp=16, wp=16, degenerate=False)
    p = region.pool_regions(cu(d["emb"]), cu(d["masks"]), engine="stream")
    S2 = region.region_query_similarity(p.fg.reshape(12, 64), cu(d["comb"][:, 0, :]), engine="stream").cpu().numpy()
    fg = float(no.fg_feat_similarity_loss(d["emb"], d["comb"], d["masks"][:, 0:1]))
    diag = np.array([S2[b, b * 3] for b in range(4)], dtype=np.float64)
    # (a) exactly the cosine of the bf16-rounded operands the contraction consumes
    rows16 = p.fg.reshape(12, 64).bfloat16().float().cpu().numpy().astype(np.float64)[::3]
    q16 = torch.from_numpy(d["comb"][:, 0, :]).bfloat16().float().numpy().astype(np.float64)
    np.testing.assert_allclose(diag, (rows16 * q16).sum(-1), rtol=0, atol=2e-6)
    # (b) the reference's fp32 cosine up to that operand rounding: |d cos| <= 2 * 2^-9 per unit-row pair worst case,
    #     ~1e-3 / sqrt(C) typical
    assert abs((1 - diag.mean()) - fg) < 1e-3


@pytest.mark.parametrize("shape", [(300, 16, 256), (1000, 40, 128), (64, 3, 64)])
def test_infonce_value_and_grads(shape):
    from cor_b200 import ops, synth
    from oracle import aten_port as ap
    from oracle import np_oracle as no
    Nr, Nq, D = shape
    g = synth.make_gallery(29, Nr, Nq, D=D)
    t = (np.arange(Nq) * 7) % Nr
    r, q = cu(g["regions"], True), cu(g["queries"], True)
    loss = ops.infonce_loss(r, q, cu(t), tau=0.07, engine="stream")
    close(loss, no.infonce_loss(g["regions"], g["queries"], t, 0.07), rtol=1e-3, atol=1e-5)
    loss.backward()
    rc, qc = torch.from_numpy(g["regions"]).requires_grad_(True), torch.from_numpy(g["queries"]).requires_grad_(True)
    ap.infonce(rc, qc, torch.from_numpy(t), 0.07).backward()
    close(r.grad, rc.grad.numpy(), rtol=2e-3, atol=2e-5)
    close(q.grad, qc.grad.numpy(), rtol=2e-3, atol=2e-5)


@pytest.mark.parametrize("k", [1, 5, 10, 50])
def test_topk_identical_indices(k):
    from cor_b200 import region, synth
    from oracle import np_oracle as no
    g = synth.make_gallery(31, 2048, 24, D=256, duplicate=True)
    idx, sc = region.topk_regions(cu(g["regions"]), cu(g["queries"]), k, engine="stream")
    ridx, rsc = no.topk_retrieve(g["regions"], g["queries"], k)
    assert (idx.cpu().numpy() == ridx).all(), "top-k indices must be identical to the oracle"
    np.testing.assert_array_equal(sc.cpu().numpy(), rsc)


def test_l2_normalize_and_grad():
    from cor_b200 import ops
    x = torch.randn(33, 256, device=dev(), requires_grad=True)
    y = ops.l2_normalize(x)
    ref = torch.nn.functional.normalize(x.detach().cpu().requires_grad_(True), dim=-1)
    close(y, ref.detach().numpy(), **TIGHT)
    g = torch.randn(33, 256)
    y.backward(g.to(dev()))
    xr = x.detach().cpu().requires_grad_(True)
    torch.nn.functional.normalize(xr, dim=-1).backward(g)
    close(x.grad, xr.grad.numpy(), rtol=1e-4, atol=1e-6)


def test_region_step_vs_cpu_port():
    """The benchmark's step (fwd + bwd) on a small batch against the ATen port, incl. gradients."""
    from cor_b200 import region, synth
    from oracle import aten_port as ap
    d = synth.make_triplets(37, B=4, M=6, C=64, h=16, w=16, H=256, W=256, hp=64, wp=64, degenerate=False)
    d["masks"][1, 3] = 0.0
    p, e, c = cu(d["pred"], True), cu(d["emb"], True), cu(d["comb"], True)
    out = region.region_step(p, e, c, cu(d["masks"]), tau=0.07, gather=False, pool_engine="stream", sim_engine="stream")
    pc, ec, cc = (torch.from_numpy(d[k]).requires_grad_(True) for k in ("pred", "emb", "comb"))
    ref, _ = ap.region_step_loss(pc, ec, cc, torch.from_numpy(d["masks"]), tau=0.07)
    close(out.loss, ref.item(), rtol=1e-3, atol=1e-4)
    out.loss.backward()
    ref.backward()
    close(p.grad, pc.grad.numpy(), rtol=2e-3, atol=1e-7)
    close(c.grad, cc.grad.numpy(), rtol=5e-3, atol=5e-5)
    close(e.grad, ec.grad.numpy(), rtol=5e-3, atol=5e-5)


def test_hooks_install_rebinds_reference_names():
    import types
    from cor_b200 import hooks, loss_func
    fake = types.ModuleType("fake_loss_func")
    for n in ("wbce_with_wiou_loss", "mask_pooling", "fg_feat_similarity_loss", "bg_feat_similarity_loss"):
        setattr(fake, n, lambda *a, **k: None)
    hooks.install(loss_module=fake, trainer_module=None, adapter_module=None)
    assert fake.wbce_with_wiou_loss is loss_func.wbce_with_wiou_loss
    hooks.uninstall()


# ------------------------------------------------------------------ full-size property checks
def test_full_size_properties_config2():
    """BASELINE config 2 shape (16 triplets x 64 masks, 1024^2, bf16): size-independent properties."""
    from cor_b200 import region
    B, M = 16, 64
    g = torch.Generator(device="cuda").manual_seed(0)
    emb = torch.randn(B, 256, 64, 64, device=dev(), generator=g).bfloat16()
    masks = torch.zeros(B, M, 1024, 1024, device=dev())
    for m in range(M):
        y0, x0 = (37 * m) % 700, (91 * m) % 700
        masks[:, m, y0:y0 + 64 + 4 * m, x0:x0 + 300] = 1.0
    masks[:, 5] = masks[:, 4]                 # duplicate mask -> identical rows
    masks[3, 7] = 0.0                         # empty
    masks[4, 8] = 1.0                         # full
    p = region.pool_regions(emb, masks, background=True, engine="stream")
    fg, bg, st = p.fg, p.bg, p.stats.view(B, M, 4)
    n = fg.norm(dim=-1)
    valid = st[..., 2] > 0
    assert torch.allclose(n[valid], torch.ones_like(n[valid]), atol=1e-4)          # unit rows
    assert torch.equal(fg[:, 5], fg[:, 4])                                          # duplicates
    assert float(st[3, 7, 0]) == 0.0 and float(fg[3, 7].abs().max()) == 0.0         # empty mask -> zero row, invalid
    assert float(st[4, 8, 1]) == 0.0                                                # full mask -> bg invalid
    assert torch.allclose(st[..., 0], masks.sum((2, 3)), rtol=1e-6)                  # exact full-res sums
    # linearity: den-weighted fg + bg means reassemble the global mean of the feature map
    P = 64 * 64
    raw = region.ops.region_pool(emb, masks, transform=region.ops.W_CLAMP, normalize=False, pair=True, engine="stream")
    den = st[..., 2:3]
    recon = (raw.fg * den + raw.bg * (P - den)) / P
    mean = emb.float().mean((2, 3))[:, None, :].expand_as(recon)
    assert torch.allclose(recon, mean, atol=2e-4)
    # similarity symmetry and top-k sortedness on the pooled rows
    R = fg.reshape(B * M, 256)
    q = torch.nn.functional.normalize(torch.randn(B, 256, device=dev(), generator=g), dim=-1)
    S = region.region_query_similarity(R, q, engine="stream")
    St = region.region_query_similarity(q, R[:64], engine="stream")
    assert torch.allclose(S[:, :64].t(), St, atol=1e-6)
    idx, sc = region.topk_regions(R, q, 10, engine="stream")
    assert (sc[:, 1:] <= sc[:, :-1]).all() and int(idx.max()) < B * M


def test_config1_full_size_fp32_vs_cpu_port():
    """BASELINE config 1 at full size: 1 triplet, 64 candidate 1024x1024 masks, fp32 everything.
    The CUDA path (exact fp32 streaming engines) against the ATen port of the reference, loss parts
    and the gradient of the composed query."""
    from cor_b200 import region, synth
    from oracle import aten_port as ap
    d = synth.make_triplets(101, B=1, M=64, C=256, h=64, w=64, H=1024, W=1024, hp=256, wp=256, degenerate=True)
    p, e, c = cu(d["pred"], True), cu(d["emb"], True), cu(d["comb"], True)
    out = region.region_step(p, e, c, cu(d["masks"]), tau=0.07, gather=False, pool_engine="stream", sim_engine="stream")
    pc, ec, cc = (torch.from_numpy(d[k]).requires_grad_(True) for k in ("pred", "emb", "comb"))
    ref, regions = ap.region_step_loss(pc, ec, cc, torch.from_numpy(d["masks"]), tau=0.07)
    close(out.loss, ref.item(), rtol=1e-3, atol=1e-4)
    close(out.regions.reshape(64, 256), regions.detach().numpy(), rtol=1e-3, atol=1e-4)
    out.loss.backward()
    ref.backward()
    close(c.grad, cc.grad.numpy(), rtol=5e-3, atol=5e-5)
    close(p.grad, pc.grad.numpy(), rtol=2e-3, atol=1e-8)


def test_ragged_and_odd_shapes():
    """Shapes that hit every scalar / tail path: odd widths, non-multiple-of-16 resample ratios, masks
    narrower than a vector, P not a multiple of the vector width, C not a multiple of the CTA tile."""
    from cor_b200 import ops, synth
    from oracle import np_oracle as no
    rng = np.random.default_rng(7)
    for (B, M, C, h, w, H, W) in [(2, 3, 17, 9, 13, 37, 51), (1, 5, 40, 27, 27, 384, 384), (3, 1, 8, 5, 7, 5, 7), (1, 2, 130, 6, 6, 100, 90),
                                  (1, 2, 8, 9, 9, 6, 6), (2, 2, 16, 1, 1, 17, 5)]:          # up-sampled masks; a single feature pixel
        emb = rng.standard_normal((B, C, h, w)).astype(np.float32)
        masks = synth.make_masks(rng, B, M, H, W, soft=True, degenerate=False)
        p = ops.region_pool(cu(emb), cu(masks), transform=ops.W_CLAMP, normalize=True, pair=True, engine="stream")
        close(p.fg, no.multi_mask_pool(emb, masks), rtol=2e-4, atol=2e-5)
        close(p.bg, no.multi_mask_pool(emb, masks, background=True), rtol=2e-4, atol=2e-5)
        st = p.stats.cpu().numpy()
        np.testing.assert_allclose(st[:, 0], masks.reshape(B * M, -1).astype(np.float64).sum(1), rtol=1e-5)
    for (N, H, W, Hm, Wm) in [(2, 33, 47, 70, 131), (1, 65, 64, 65, 64), (3, 7, 200, 28, 800), (2, 40, 8, 13, 5)]:   # last: mask coarser than the logits
        pred = rng.standard_normal((N, 1, H, W)).astype(np.float32)
        mask = synth.make_masks(rng, N, 1, Hm, Wm, soft=True, degenerate=False)
        close(ops.seg_loss(cu(pred), cu(mask)), no.segmentation_loss(pred, mask), rtol=1e-4, atol=1e-6)


def test_val_post_generic_scale_and_u8_gt():
    from cor_b200 import ops, synth
    from oracle import np_oracle as no
    rng = np.random.default_rng(9)
    pred = synth.make_logits(rng, 2, 24, 40)
    gt = synth.make_masks(rng, 2, 1, 60, 100, degenerate=False)
    r = ops.val_postprocess(cu(pred), size=(60, 100), gt=cu((gt * 255).astype(np.uint8)), want_hard=True)
    ref = no.val_postprocess(pred, (60, 100))
    close(r["post"], ref, rtol=1e-5, atol=3e-6)
    assert (r["hard"].cpu().numpy() == no.binarize(ref)).mean() >= 0.999
    m = no.soft_metrics(ref, gt)
    for i, k in enumerate(("dice", "mae", "iou", "mdice", "miou")):
        close(r["metrics"][:, i], m[k], rtol=1e-4, atol=1e-6)


def test_collective_free_infonce_backward_emulating_two_ranks():
    """The multi-rank step computes d(sum_j loss_j)/d(my regions) locally from (my regions, ALL queries) and
    d loss_me / d(my queries) from (ALL regions, my queries): emulate world_size 2 on one GPU and compare with
    autograd over the concatenated problem."""
    from cor_b200 import _lib as L
    from cor_b200 import ops, synth
    ws, B, M, D = 2, 5, 12, 128
    n_local = B * M
    g = synth.make_gallery(131, ws * n_local, ws * B, D=D)
    R = cu(g["regions"], True)
    Q = cu(g["queries"], True)
    tau = 0.07
    tg = [torch.arange(B, device=dev()) * M + j * n_local for j in range(ws)]
    total = sum(ops.infonce_loss(R, Q[j * B:(j + 1) * B], tg[j], tau, engine="stream") for j in range(ws))
    total.backward()
    r16 = R.detach().to(torch.bfloat16).contiguous()
    q16 = Q.detach().to(torch.bfloat16).contiguous()
    _, lse_all = ops._sim_forward(r16, q16, 1 / tau, False, True, "stream")
    one = torch.ones(1, device=dev())
    lib = L.load()
    for rank in range(ws):
        g_regions = torch.empty(n_local, D, device=dev())
        g_q = torch.empty(B, D, device=dev())
        work = ops._work(lib.cor_sim_work_bytes(ws * B, ws * n_local, D), dev())
        offset = rank * n_local
        tgt_all = torch.cat(tg) - offset
        my_q = q16[rank * B:(rank + 1) * B].contiguous()
        my_r = r16[offset:offset + n_local].contiguous()
        ops._call("cor_infonce_bwd", dev(), ops.ptr(r16), ops.ptr(my_q), ops.ptr(tg[rank].contiguous()), ops.ptr(lse_all[rank * B:(rank + 1) * B].contiguous()),
                  ws * n_local, B, D, ops._f(1 / tau), ops.ptr(one), ops._f(1.0), None, ops.ptr(g_q), ops.ptr(work))
        ops._call("cor_infonce_bwd", dev(), ops.ptr(my_r), ops.ptr(q16), ops.ptr(tgt_all.contiguous()), ops.ptr(lse_all), n_local, ws * B, D,
                  ops._f(1 / tau), ops.ptr(one), ops._f(float(ws)), ops.ptr(g_regions), None, None)
        close(g_q, Q.grad[rank * B:(rank + 1) * B].cpu().numpy(), rtol=1e-4, atol=1e-6)
        close(g_regions, R.grad[offset:offset + n_local].cpu().numpy(), rtol=1e-4, atol=1e-6)


def test_vailder_offline_order_golden():
    """Offline evaluator order (vailder.py:427-430,466,473) on the device vs the cv2-generated golden."""
    from cor_b200 import ops
    g = load_golden("vailder_hard")
    hw = tuple(int(v) for v in g["gt_hw"])
    r = ops.val_postprocess(cu(g["pred"]), size=hw, want_hard=True, post_first=True)
    close(r["post"], g["resized"], rtol=1e-5, atol=3e-6)
    assert (r["hard"].cpu().numpy() == g["hard"]).mean() >= 0.999


def test_trainer_usage_under_autocast_trains_parameters():
    """The drop-in functions used exactly like utils/trainer_v3_g.py:51-76 (bf16 autocast, fp32 masks,
    loss.backward()) on a toy model: gradients reach the parameters and match the ATen port."""
    from cor_b200 import loss_func as lf
    from oracle import aten_port as ap
    torch.manual_seed(3)
    B = 4
    enc = torch.nn.Conv2d(3, 64, 3, padding=1).to(dev())
    head = torch.nn.Conv2d(64, 1, 1).to(dev())
    proj = torch.nn.Linear(64, 64).to(dev())
    img = torch.randn(B, 3, 32, 32, device=dev())
    qmask = (torch.rand(B, 1, 128, 128, device=dev()) > 0.6).float()
    qmask[1] = 0.0

    def forward(fns):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            emb = enc(img)                                             # bf16 [B,64,32,32]
            pred = head(emb)                                           # bf16 [B,1,32,32]
            comb = torch.nn.functional.normalize(proj(emb.mean((2, 3))).float(), dim=-1).unsqueeze(1)
            target = torch.nn.functional.interpolate(qmask, size=pred.shape[2:], mode="bilinear", align_corners=False)
            return fns[0](pred, target) + 5 * fns[1](emb, comb, qmask) + 5 * fns[2](emb, comb, qmask)

    params = list(enc.parameters()) + list(head.parameters()) + list(proj.parameters())
    loss = forward((lf.wbce_with_wiou_loss, lf.fg_feat_similarity_loss, lf.bg_feat_similarity_loss))
    loss.backward()
    ours = [p.grad.clone() for p in params]
    for p in params:
        p.grad = None
    ref = forward((lambda a, b: ap.edge_weighted_seg_loss(a.float(), b), lambda e, c, m: ap.fg_loss(e.float(), c, m),
                   lambda e, c, m: ap.bg_loss(e.float(), c, m)))
    ref.backward()
    close(loss, ref.item(), rtol=2e-3, atol=2e-3)
    for a, p in zip(ours, params):
        assert torch.isfinite(a).all()
        close(a, p.grad.cpu().numpy(), rtol=5e-2, atol=5e-3 * float(p.grad.abs().max()) + 1e-6)
    with torch.no_grad():
        v = forward((lf.wbce_with_wiou_loss, lf.fg_feat_similarity_loss, lf.bg_feat_similarity_loss))
    assert not v.requires_grad and abs(float(v) - float(loss)) < 1e-4


def test_drop_in_metric_functions_golden():
    """compute_dice / mae / iou / mdice / miou (trainer_v3_g.py:381-443) through cor_b200.metrics: one pass serves
    all five calls, values equal the reference's."""
    from cor_b200 import metrics as mt
    from cor_b200 import ops
    g = load_golden("val_post")
    pred, gt = cu(g["post_up"]), cu(g["gt"])
    n0 = ops.LAUNCHES["count"]
    res = {"dice": mt.compute_dice(pred, gt), "mae": mt.compute_mae(pred, gt), "iou": mt.compute_iou(pred, gt),
           "mdice": mt.compute_mdice(pred, gt), "miou": mt.compute_miou(pred, gt)}
    assert ops.LAUNCHES["count"] - n0 == 2, "five metric calls on the same tensors must share one pass"
    for k, v in res.items():
        assert v.shape == (2,)
        close(v, g[k], rtol=1e-5, atol=1e-7)
    close(mt.compute_dice(pred, gt, smooth=1.0), (2 * (g["post_up"] * g["gt"]).reshape(2, -1).sum(1) + 1.0) /
          (g["post_up"].reshape(2, -1).sum(1) + g["gt"].reshape(2, -1).sum(1) + 1.0), rtol=1e-5, atol=1e-7)
