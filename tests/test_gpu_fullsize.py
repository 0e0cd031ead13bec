"""Full-size parity: the tensor-core paths at BASELINE.json's own sizes, DIRECTLY against the
oracle (oracle/aten_port.py = the reference's ATen op sequence), not against another kernel of ours.

  * config 2 (16 triplets x 64 masks of 1024^2, bf16 features): the fused tcgen05 step -- loss parts, region rows and
    all three gradients -- against the port looped per image (bounded host memory);
  * config 5 corners (C = 1152, 128^2 feature map, 256 masks from 1024^2): pooled unit rows against the port;
  * config 4 "global" (1024 queries x 102 400 regions): InfoNCE value and both gradients against the port.

Run with ``-m gpu`` on a B200; the CPU side of each test takes seconds on the box's host cores."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def relnorm(a, b):
    a, b = a.detach().float().cpu().double(), b.detach().float().cpu().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _port_step_per_image(pred, emb, comb, masks, tau):
    """aten_port.region_step_loss on the whole batch, evaluated image by image: regions first (no graph), then the
    batch losses with the region rows as a leaf, then the pooling backward per image (268 MB of graph at a time)."""
    from oracle import aten_port as ap
    B, M = masks.shape[:2]
    with torch.no_grad():
        rows = torch.cat([ap.multi_mask_regions(emb[b:b + 1], masks[b:b + 1]).reshape(M, -1) for b in range(B)])
    rows = rows.requires_grad_(True)
    p, e, c = (t.detach().clone().requires_grad_(True) for t in (pred, emb, comb))
    gt = masks[:, 0:1]
    seg = ap.seg_loss_fullres(p, gt)
    fg = ap.fg_loss(e, c, gt)
    bg = ap.bg_loss(e, c, gt)
    targets = torch.arange(B) * M
    nce = ap.infonce(rows, c[:, 0, :].float(), targets, tau)
    loss = seg + 5 * fg + 5 * bg + nce
    loss.backward()
    g_emb = e.grad.clone()
    for b in range(B):
        eb = emb[b:b + 1].detach().clone().requires_grad_(True)
        rb = ap.multi_mask_regions(eb, masks[b:b + 1]).reshape(M, -1)
        rb.backward(rows.grad[b * M:(b + 1) * M])
        g_emb[b] += eb.grad[0]
    parts = {k: float(v) for k, v in (("loss", loss), ("seg", seg), ("fg", fg), ("bg", bg), ("nce", nce))}
    return parts, rows.detach(), p.grad, c.grad, g_emb


def test_config2_full_size_fused_step_vs_port():
    """BASELINE configs[1] exactly as bench.py builds it (same generator, seed 1234)."""
    import bench
    from cor_b200 import region
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    cfg = dict(bench.CFG)
    inp = bench.device_inputs(dev(), 1234, cfg, "f32")
    B, M = cfg["B"], cfg["M"]
    p = inp["pred"].clone().requires_grad_(True)
    e = inp["emb"].clone().requires_grad_(True)
    c = inp["comb"].clone().requires_grad_(True)
    assert region._fused_ok(e, inp["masks"], "auto")
    out = region.region_step(p, e, c, inp["masks"], tau=cfg["tau"], gather=False)          # auto engines = the benched path
    out.loss.backward()
    host = {k: v.detach().float().cpu() for k, v in inp.items()}
    parts, rows, g_pred, g_comb, g_emb = _port_step_per_image(host["pred"], host["emb"], host["comb"], host["masks"], cfg["tau"])
    for k in ("loss", "seg", "fg", "bg", "nce"):
        got = float(getattr(out, k))
        assert abs(got - parts[k]) <= 1e-3 * abs(parts[k]) + 1e-6, (k, got, parts[k])
    got_rows = out.regions.reshape(B * M, -1).float().cpu()
    np.testing.assert_allclose(got_rows.numpy(), rows.numpy(), rtol=1e-3, atol=1e-3)
    assert relnorm(got_rows, rows) < 2e-3
    assert relnorm(p.grad, g_pred) < 5e-3
    assert relnorm(c.grad, g_comb) < 5e-3
    assert relnorm(e.grad, g_emb) < 5e-3
    np.testing.assert_allclose(p.grad.float().cpu().numpy(), g_pred.numpy(), rtol=2e-2, atol=2e-3 * float(g_pred.abs().max()))


def _port_rows_chunked(emb, masks, chunk=16):
    """Unit rows [M, C] of ONE image through aten_port.multi_mask_regions, `chunk` masks at a time."""
    from oracle import aten_port as ap
    M = masks.shape[1]
    with torch.no_grad():
        return torch.cat([ap.multi_mask_regions(emb, masks[:, m:m + chunk]).reshape(-1, emb.shape[1]) for m in range(0, M, chunk)])


@pytest.mark.parametrize("cfg", [(1152, 128, 256, 1024), (1152, 64, 64, 1024), (256, 128, 16, 1024), (768, 96, 100, 384)])
def test_config5_corners_pooling_vs_port(cfg):
    """BASELINE configs[4] corners: C x (h = w) x M masks resampled from Hm^2; bf16 features; hard masks with soft
    (8-bit quantised) edges on every third mask so that the resampled weights are NOT all exact in bf16."""
    from cor_b200 import ops
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    C, h, M, Hm = cfg
    g = torch.Generator(device="cuda").manual_seed(7 + M)
    emb = torch.randn(1, C, h, h, device=dev(), generator=g).bfloat16()
    masks = torch.zeros(1, M, Hm, Hm, device=dev())
    yy = torch.arange(Hm, device=dev()).view(Hm, 1).float()
    xx = torch.arange(Hm, device=dev()).view(1, Hm).float()
    for m in range(M):
        y0, x0 = (37 * m) % (Hm // 2), (91 * m) % (Hm // 2)
        hh, ww = Hm // 16 + (5 * m) % (Hm // 3), Hm // 16 + (11 * m) % (Hm // 3)
        if m % 3 == 2:      # soft ellipse, 8-bit quantised like a PNG through ToTensor
            d2 = ((yy - y0 - hh / 2) / (hh / 2)) ** 2 + ((xx - x0 - ww / 2) / (ww / 2)) ** 2
            masks[0, m] = torch.round(torch.clamp(1.5 - d2, 0, 1) * 255) / 255
        else:
            masks[0, m, y0:y0 + hh, x0:x0 + ww] = 1.0
    pair = M < 256
    assert ops.umma_pool_eligible(emb, M, h * h, ops.W_CLAMP, 1, pair)
    got = ops.region_pool(emb, masks, transform=ops.W_CLAMP, normalize=True, pair=pair, engine="umma").fg[0].float().cpu()
    want = _port_rows_chunked(emb.float().cpu(), masks.cpu(), chunk=8 if C * h * h > (1 << 23) else 32)
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-3, atol=1e-3 / np.sqrt(C / 256.0))
    assert relnorm(got, want) < 1e-3


def test_config4_global_infonce_fwd_bwd_vs_port():
    """1024 queries x 102 400 regions x 256 (BASELINE configs[3], negatives of all 8 ranks): value and both gradients."""
    from cor_b200 import ops, synth
    from oracle import aten_port as ap
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    Nq, Nr, D = 1024, 102400, 256
    g = synth.make_gallery(77, Nr, Nq, D=D)
    t = (np.arange(Nq, dtype=np.int64) * 97) % Nr
    r = torch.from_numpy(g["regions"]).to(dev()).requires_grad_(True)
    q = torch.from_numpy(g["queries"]).to(dev()).requires_grad_(True)
    loss = ops.infonce_loss(r, q, torch.from_numpy(t).to(dev()), tau=0.07)
    loss.backward()
    rc, qc = torch.from_numpy(g["regions"]).requires_grad_(True), torch.from_numpy(g["queries"]).requires_grad_(True)
    ref = ap.infonce(rc, qc, torch.from_numpy(t), 0.07)
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-3 * abs(float(ref))
    assert relnorm(q.grad, qc.grad) < 5e-3
    assert relnorm(r.grad, rc.grad) < 5e-3
    assert float((r.grad.cpu() - rc.grad).abs().max() / rc.grad.abs().max()) < 2e-2


def test_config4_per_rank_few_queries_long_gallery_vs_port():
    """16 local queries x 102 400 gathered regions: the shape a rank scores at 8 GPUs (auto engine = tensor cores)."""
    from cor_b200 import ops, synth
    from oracle import aten_port as ap
    Nq, Nr, D = 16, 102400, 256
    assert ops._sim_engine(Nq, Nr, D, "auto") == "umma"
    g = synth.make_gallery(79, Nr, Nq, D=D)
    t = (np.arange(Nq, dtype=np.int64) * 6400) % Nr
    r = torch.from_numpy(g["regions"]).to(dev()).requires_grad_(True)
    q = torch.from_numpy(g["queries"]).to(dev()).requires_grad_(True)
    loss = ops.infonce_loss(r, q, torch.from_numpy(t).to(dev()), tau=0.07)
    loss.backward()
    rc, qc = torch.from_numpy(g["regions"]).requires_grad_(True), torch.from_numpy(g["queries"]).requires_grad_(True)
    ref = ap.infonce(rc, qc, torch.from_numpy(t), 0.07)
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-4 * abs(float(ref))
    assert relnorm(q.grad, qc.grad) < 1e-3
    assert relnorm(r.grad, rc.grad) < 1e-3
