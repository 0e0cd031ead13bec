"""GPU parity of the tcgen05 / TMEM / TMA kernels (pool_umma.cu, sim_umma.cu) against the CPU oracle and
against the exact-fp32 streaming kernels.  Run with ``-m gpu`` on a B200."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def cu(x, dtype=None, grad=False):
    t = torch.from_numpy(np.ascontiguousarray(x)).to(dev())
    if dtype is not None:
        t = t.to(dtype)
    return t.requires_grad_(grad)


def close(a, b, rtol=1e-3, atol=1e-3):
    a = a.detach().float().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().float().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    np.testing.assert_allclose(a.astype(np.float64), b.astype(np.float64), rtol=rtol, atol=atol)


# (B, M, C, h, w, Hm, Wm, soft): integer 4x hard masks (weights exact in bf16), soft 8-bit masks (they are not), and
# non-integer resample ratios (100 -> 24 = 4.17x, 384 -> 24 x 40 = 16x / 9.6x, SURVEY 7 "Bilinear parity")
@pytest.mark.parametrize("shape", [(2, 16, 128, 16, 16, 64, 64, False), (3, 40, 256, 32, 32, 128, 128, False),
                                   (1, 100, 384, 64, 64, 256, 256, False), (2, 255, 128, 8, 8, 32, 32, False),
                                   (2, 48, 128, 32, 32, 128, 128, True), (2, 33, 256, 24, 24, 100, 100, True),
                                   (1, 64, 128, 24, 40, 384, 384, True), (2, 20, 128, 16, 16, 16, 16, True)])
def test_pool_umma_vs_oracle_and_stream(shape):
    """bf16 features, M masks: tensor-core pooling == oracle within the north-star tolerance and
    == the fp32 streaming kernel fed the same bf16 features."""
    from cor_b200 import ops, synth
    from oracle import np_oracle as no
    B, M, C, h, w, Hm, Wm, soft = shape
    d = synth.make_triplets(41 + M, B=B, M=M, C=C, h=h, w=w, H=Hm, W=Wm, hp=8, wp=8, soft=soft, degenerate=True)
    emb16 = torch.from_numpy(d["emb"]).bfloat16()
    masks = cu(d["masks"])
    pu = ops.region_pool(emb16.to(dev()), masks, transform=ops.W_CLAMP, normalize=True, pair=True, engine="umma", want_bf16=True)
    ps = ops.region_pool(emb16.to(dev()), masks, transform=ops.W_CLAMP, normalize=True, pair=True, engine="stream")
    ref_fg = no.multi_mask_pool(emb16.float().numpy(), d["masks"])
    ref_bg = no.multi_mask_pool(emb16.float().numpy(), d["masks"], background=True)
    close(pu.fg, ref_fg, rtol=1e-3, atol=1e-3)
    close(pu.bg, ref_bg, rtol=1e-3, atol=1e-3)
    close(pu.fg, ps.fg, rtol=1e-3, atol=5e-4)
    close(pu.bg, ps.bg, rtol=1e-3, atol=5e-4)
    close(pu.fg_bf16.view_as(pu.fg), pu.fg, rtol=1e-2, atol=4e-3)
    assert torch.equal(pu.stats[:, :2], ps.stats[:, :2])
    # determinism: split-K partials are summed in fixed order
    pu2 = ops.region_pool(emb16.to(dev()), masks, transform=ops.W_CLAMP, normalize=True, pair=True, engine="umma")
    assert torch.equal(pu.fg, pu2.fg) and torch.equal(pu.bg, pu2.bg)


def test_pool_umma_256_masks_foreground_only():
    """The sweep's upper end (256 masks per image): all 256 UMMA N columns are masks when no background
    rows are asked for."""
    from cor_b200 import ops, synth
    from oracle import np_oracle as no
    d = synth.make_triplets(49, B=1, M=256, C=128, h=16, w=16, H=32, W=32, hp=8, wp=8, degenerate=False)
    emb16 = torch.from_numpy(d["emb"]).bfloat16()
    p = ops.region_pool(emb16.to(dev()), cu(d["masks"]), transform=ops.W_CLAMP, normalize=True, pair=False, engine="umma")
    close(p.fg, no.multi_mask_pool(emb16.float().numpy(), d["masks"]), rtol=1e-3, atol=1e-3)
    with pytest.raises(Exception):
        ops.region_pool(emb16.to(dev()), cu(d["masks"]), transform=ops.W_CLAMP, normalize=True, pair=True, engine="umma")


def test_pool_umma_unnormalised_sums_hard_masks_exact_weights():
    """Hard masks at 4x scale resample to {0,.25,.5,.75,1}: exact in bf16, so with bf16 features the
    tensor-core sums equal the fp32 streaming sums up to accumulation order."""
    from cor_b200 import ops, synth
    d = synth.make_triplets(43, B=2, M=24, C=128, h=32, w=32, H=128, W=128, hp=8, wp=8, degenerate=False)
    emb16 = torch.from_numpy(d["emb"]).bfloat16().to(dev())
    a = ops.region_pool(emb16, cu(d["masks"]), transform=ops.W_CLAMP, normalize=False, engine="umma")
    b = ops.region_pool(emb16, cu(d["masks"]), transform=ops.W_CLAMP, normalize=False, engine="stream")
    close(a.fg, b.fg, rtol=2e-5, atol=2e-5)


def test_pool_umma_backward_matches_stream():
    from cor_b200 import ops, synth
    d = synth.make_triplets(47, B=2, M=20, C=128, h=16, w=16, H=64, W=64, hp=8, wp=8, degenerate=False)
    g = torch.randn(2, 20, 128, device=dev())
    grads = []
    for eng in ("umma", "stream"):
        e = torch.from_numpy(d["emb"]).bfloat16().to(dev()).requires_grad_(True)
        p = ops.region_pool(e, cu(d["masks"]), transform=ops.W_CLAMP, normalize=True, pair=True, engine=eng)
        (p.fg * g).sum().backward(retain_graph=True)
        (p.bg * g.flip(0)).sum().backward()
        grads.append(e.grad.float())
    close(grads[0], grads[1], rtol=2e-2, atol=2e-3)


@pytest.mark.parametrize("shape", [(300, 20, 256), (4096, 256, 256), (1000, 130, 128), (257, 1, 64), (5000, 16, 256), (770, 300, 192)])
def test_sim_umma_vs_oracle(shape):
    from cor_b200 import ops, synth
    from oracle import np_oracle as no
    Nr, Nq, D = shape
    g = synth.make_gallery(53 + Nq, Nr, Nq, D=D)
    S = ops.similarity(cu(g["regions"]), cu(g["queries"]), engine="umma")
    ref = no.region_query_similarity(g["regions"], g["queries"])
    close(S, ref, rtol=1e-3, atol=2e-6)
    Ss = ops.similarity(cu(g["regions"]), cu(g["queries"]), engine="stream")
    close(S, Ss, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("shape", [(300, 20, 256), (4096, 256, 256), (1000, 130, 128), (20000, 16, 256)])
def test_infonce_umma(shape):
    from cor_b200 import ops, synth
    from oracle import np_oracle as no
    Nr, Nq, D = shape
    g = synth.make_gallery(59 + Nq, Nr, Nq, D=D)
    t = (np.arange(Nq) * 13) % Nr
    r, q = cu(g["regions"], grad=True), cu(g["queries"], grad=True)
    loss = ops.infonce_loss(r, q, cu(t), tau=0.07, engine="umma")
    close(loss, no.infonce_loss(g["regions"], g["queries"], t, 0.07), rtol=1e-3, atol=1e-5)
    ls = ops.infonce_loss(cu(g["regions"]), cu(g["queries"]), cu(t), tau=0.07, engine="stream")
    close(loss, ls, rtol=1e-5, atol=1e-6)
    loss.backward()
    from oracle import aten_port as ap
    rc, qc = torch.from_numpy(g["regions"]).requires_grad_(True), torch.from_numpy(g["queries"]).requires_grad_(True)
    ap.infonce(rc, qc, torch.from_numpy(t), 0.07).backward()
    for got, want in ((r.grad, rc.grad), (q.grad, qc.grad)):
        assert float((got.cpu() - want).norm() / want.norm()) < 5e-3


@pytest.mark.parametrize("k", [1, 10, 50])
def test_topk_umma_identical_indices(k):
    """BASELINE config 3: 256 queries x 4096 regions; indices identical to the oracle's total order."""
    from cor_b200 import region, synth
    from oracle import np_oracle as no
    g = synth.make_gallery(61, 4096, 256, D=256, duplicate=True)
    idx, sc = region.topk_regions(cu(g["regions"]), cu(g["queries"]), k, engine="umma")
    ridx, rsc = no.topk_retrieve(g["regions"], g["queries"], k)
    assert (idx.cpu().numpy() == ridx).all()
    np.testing.assert_array_equal(sc.cpu().numpy(), rsc)


def test_region_step_auto_engines_vs_cpu_port():
    """The fused step with the tensor-core pooling engine (bf16 features) against the ATen port."""
    from cor_b200 import region, synth
    from oracle import aten_port as ap
    d = synth.make_triplets(67, B=2, M=16, C=128, h=16, w=16, H=128, W=128, hp=32, wp=32, degenerate=False)
    emb16 = torch.from_numpy(d["emb"]).bfloat16()
    p, c = cu(d["pred"], grad=True), cu(d["comb"], grad=True)
    e = emb16.to(dev()).requires_grad_(True)
    out = region.region_step(p, e, c, cu(d["masks"]), tau=0.07, gather=False, pool_engine="umma", sim_engine="stream")
    pc, cc = (torch.from_numpy(d[k]).requires_grad_(True) for k in ("pred", "comb"))
    ec = emb16.float().requires_grad_(True)
    ref, _ = ap.region_step_loss(pc, ec, cc, torch.from_numpy(d["masks"]), tau=0.07)
    close(out.loss, ref.item(), rtol=2e-3, atol=2e-3)
    out.loss.backward()
    ref.backward()
    close(c.grad, cc.grad.numpy(), rtol=2e-2, atol=2e-3)
    close(p.grad, pc.grad.numpy(), rtol=2e-3, atol=1e-7)


@pytest.mark.parametrize("cfg", [(2, 64, 256, 64, 64, torch.bfloat16, True), (1, 255, 128, 16, 16, torch.float32, True),
                                 (2, 7, 128, 32, 32, torch.float32, False), (1, 100, 384, 32, 16, torch.bfloat16, True)])
def test_pool_bwd_umma_vs_cuda_core(cfg):
    """Tensor-core backward of the pooling contraction vs the fp32 CUDA-core kernel (same saved weights)."""
    from cor_b200 import ops, synth
    B, M, C, h, w, dt, pair = cfg
    d = synth.make_triplets(71 + M, B=B, M=M, C=C, h=h, w=w, H=2 * h, W=2 * w, hp=8, wp=8, soft=True, degenerate=False)
    g1 = torch.randn(B, M, C, device=dev())
    g2 = torch.randn(B, M, C, device=dev())
    grads = []
    for eng in ("auto", "stream"):
        e = torch.from_numpy(d["emb"]).to(dt).to(dev()).requires_grad_(True)
        p = ops.region_pool(e, cu(d["masks"]), transform=ops.W_CLAMP, normalize=False, pair=pair, engine=eng)
        loss = (p.fg * g1).sum() + ((p.bg * g2).sum() if pair else 0.0)
        loss.backward()
        grads.append(e.grad.float())
    scale = float(grads[1].abs().max())
    close(grads[0], grads[1], rtol=2e-2, atol=1e-2 * scale)
    # relative Frobenius error: bf16 operand rounding only
    err = float((grads[0] - grads[1]).norm() / grads[1].norm())
    assert err < 6e-3, err


@pytest.mark.parametrize("sim_engine", ["stream", "umma"])
@pytest.mark.parametrize("bg_mode", [0, 1])
def test_fused_step_equals_modular_ops(bg_mode, sim_engine):
    """region_step's single-node fused path (hand-written backward, log-sum-exp partials of either similarity
    producer merged by the one-launch tail kernel) against the composition of the modular autograd ops on the same
    inputs: loss parts and all three gradients."""
    from cor_b200 import region, synth
    d = synth.make_triplets(73, B=3, M=20, C=128, h=16, w=16, H=64, W=64, hp=32, wp=32, degenerate=False)
    d["masks"][1, 0] = 0.0                      # an invalid GT mask
    outs = []
    for fused in (True, False):
        p = cu(d["pred"], grad=True)
        c = cu(d["comb"], grad=True)
        e = torch.from_numpy(d["emb"]).bfloat16().to(dev()).requires_grad_(True)
        o = region.region_step(p, e, c, cu(d["masks"]), tau=0.07, gather=False, bg_mode=bg_mode, fused=fused, sim_engine=sim_engine)
        o.loss.backward()
        outs.append((o, p.grad.float(), c.grad.float(), e.grad.float()))
    (a, pa, ca, ea), (b, pb, cb, eb) = outs
    for k in ("loss", "seg", "fg", "bg", "nce"):
        close(getattr(a, k), getattr(b, k), rtol=1e-5, atol=1e-6)
    close(a.regions, b.regions, rtol=1e-5, atol=1e-6)
    close(pa, pb, rtol=1e-5, atol=1e-9)
    close(ca, cb, rtol=1e-4, atol=1e-6)
    close(ea, eb, rtol=2e-2, atol=2e-3 * float(eb.abs().max()))


def test_full_size_config2_tensor_core_step_vs_exact_engines():
    """BASELINE config 2 at full size (16 triplets x 64 masks of 1024^2, bf16 features): the fused tensor-core
    step against the composition of the exact fp32 streaming engines -- loss parts, region rows, gradients --
    plus size-independent properties (unit rows, zero rows for empty masks, top-1 of each query among its own
    image's regions being reproducible)."""
    from cor_b200 import region
    B, M = 16, 64
    g = torch.Generator(device="cuda").manual_seed(5)
    emb = torch.randn(B, 256, 64, 64, device=dev(), generator=g).bfloat16()
    comb = torch.nn.functional.normalize(torch.randn(B, 1, 256, device=dev(), generator=g), dim=-1)
    pred = (2 * torch.randn(B, 1, 256, 256, device=dev(), generator=g)).bfloat16()
    masks = torch.zeros(B, M, 1024, 1024, device=dev())
    for m in range(M):
        y0, x0 = (53 * m) % 640, (97 * m) % 640
        masks[:, m, y0:y0 + 96 + 4 * m, x0:x0 + 128 + 3 * m] = 1.0
    masks[2, 0] = 0.0                       # empty GT mask -> invalid fg sample
    masks[5, 9] = 0.0
    res = []
    for fused, pe, se in ((True, "umma", "auto"), (False, "stream", "stream")):
        p = pred.clone().requires_grad_(True)
        c = comb.clone().requires_grad_(True)
        e = emb.clone().requires_grad_(True)
        o = region.region_step(p, e, c, masks, tau=0.07, gather=False, pool_engine=pe, sim_engine=se, fused=fused)
        o.loss.backward()
        res.append((o, p.grad.float(), c.grad.float(), e.grad.float()))
    (a, pa, ca, ea), (b, pb, cb, eb) = res
    for k in ("loss", "seg", "fg", "bg", "nce"):
        close(getattr(a, k), getattr(b, k), rtol=1e-3, atol=1e-3)
    close(a.regions, b.regions, rtol=1e-3, atol=1e-3)
    close(pa, pb, rtol=1e-4, atol=1e-9)
    close(ca, cb, rtol=2e-2, atol=2e-3 * float(cb.abs().max()))
    assert float((ea - eb).norm() / eb.norm()) < 2e-2
    n = a.regions.norm(dim=-1)
    assert float(a.regions[5, 9].abs().max()) == 0.0 and float(a.regions[2, 0].abs().max()) == 0.0
    ok = n > 0
    assert torch.allclose(n[ok], torch.ones_like(n[ok]), atol=2e-3)
    ia, _ = region.topk_regions(a.regions.reshape(B * M, 256), comb.reshape(B, 256), 5)
    ib, _ = region.topk_regions(a.regions.reshape(B * M, 256), comb.reshape(B, 256), 5, engine="stream")
    assert torch.equal(ia, ib)


@pytest.mark.parametrize("shape", [(4096, 256, 256), (5001, 130, 128), (3000, 300, 192), (4100, 128, 64), (9000, 129, 256)])
def test_infonce_dense_backward_vs_streaming_and_oracle(shape):
    """Many-query InfoNCE backward on the tcgen05 kernel of csrc/nce_bwd_umma.cu (S tile recomputed, bf16 coefficients
    formed in registers and fed back through shared memory as the A operand, the region / query tile reused as an
    MN-major B operand; no S, no P, no library GEMM) against the exact streaming backward and the ATen port.
    bf16 coefficients: gradients agree to a few 1e-3 of their norm.  Ragged tiles in both dimensions, all four D."""
    from cor_b200 import ops, synth
    from oracle import aten_port as ap
    Nr, Nq, D = shape
    assert ops._infonce_bwd_dense(Nq, Nr, D, "auto")
    g = synth.make_gallery(31, Nr, Nq, D=D)
    t = (np.arange(Nq) * 11) % Nr
    grads = {}
    for eng in ("auto", "stream"):
        r, q = cu(g["regions"], grad=True), cu(g["queries"], grad=True)
        loss = ops.infonce_loss(r, q, cu(t), tau=0.07, engine=eng)
        (3.0 * loss).backward()
        grads[eng] = (float(loss), r.grad.float(), q.grad.float())
    rc, qc = torch.from_numpy(g["regions"]).requires_grad_(True), torch.from_numpy(g["queries"]).requires_grad_(True)
    lc = ap.infonce(rc, qc, torch.from_numpy(t), 0.07)
    (3.0 * lc).backward()
    assert abs(grads["auto"][0] - float(lc)) < 1e-3 * abs(float(lc))
    for i, ref in ((1, rc.grad), (2, qc.grad)):
        a, s_ = grads["auto"][i].cpu(), grads["stream"][i].cpu()
        assert float((a - s_).norm() / s_.norm()) < 4e-3
        assert float((a - ref).norm() / ref.norm()) < 4e-3
        assert float((a - ref).abs().max() / ref.abs().max()) < 1e-2


def test_tensor_core_paths_from_a_thread_without_a_cuda_context():
    """TMA descriptors are encoded through the driver API, which needs a context current on the CALLING thread; a fresh
    Python thread (or autograd's backward thread when our node is the first to run) has none.  The library binds the
    context of the device that owns the operand."""
    import threading
    from cor_b200 import ops, synth
    g = synth.make_gallery(5, 512, 160, D=128)
    r, q = cu(g["regions"]), cu(g["queries"])
    want = ops.similarity(r, q, engine="stream").cpu()
    torch.cuda.synchronize()
    box = {}

    def work():
        try:
            box["S"] = ops.similarity(r, q, engine="umma").cpu()
        except Exception as e:  # noqa: BLE001
            box["err"] = e

    th = threading.Thread(target=work)
    th.start()
    th.join()
    assert "err" not in box, box.get("err")
    torch.testing.assert_close(box["S"], want, rtol=1e-3, atol=1e-3)
    # autograd thread, our backward as the first node, dense (tensor-core) path
    g2 = synth.make_gallery(6, 4096, 256, D=256)
    r2, q2 = cu(g2["regions"], grad=True), cu(g2["queries"], grad=True)
    ops.infonce_loss(r2, q2, cu((np.arange(256) * 3) % 4096), tau=0.07).backward()
    assert torch.isfinite(r2.grad).all() and torch.isfinite(q2.grad).all()


def test_captured_step_with_tensor_core_similarity_replays_bit_equal():
    """At 8 GPUs the gathered gallery (>= 8192 rows) sends the fused step's similarity stage to the tcgen05 kernel with the
    queries in tensor memory (sim_umma_ts.cu); here the same path is forced on one GPU (sim_engine="umma") and captured
    into a CUDA graph: replays must reproduce the eager step bit for bit, loss and all three gradients."""
    from cor_b200 import region, synth
    d = synth.make_triplets(91, B=4, M=32, C=256, h=16, w=16, H=64, W=64, hp=32, wp=32, degenerate=False)
    t = {k: torch.from_numpy(v).to(dev()) for k, v in d.items()}
    t["emb"] = t["emb"].bfloat16()
    bufs = region.StepBuffers(4, 32, C=256, h=16, w=16, H=64, W=64, hp=32, wp=32, device=dev(), emb_dtype=torch.bfloat16)
    bufs.load(t)
    kw = dict(gather=False, sim_engine="umma")
    loss_e, grads_e = bufs._step(True, True, kw)
    loss_e = loss_e.clone()
    grads_e = {k: v.clone() for k, v in grads_e.items()}
    bufs.capture(**kw)
    for _ in range(3):
        bufs.replay()
    torch.cuda.synchronize()
    assert torch.equal(bufs.loss.reshape(()), loss_e.reshape(()))
    for k in ("pred", "emb", "comb"):
        assert torch.equal(bufs.grads[k], grads_e[k]), k
