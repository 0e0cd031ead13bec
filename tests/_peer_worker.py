"""Worker of tests/test_gpu_peer.py: launched by torchrun with >= 2 GPUs (NCCL).  Checks the NVLink peer-memory
exchange against the NCCL collectives it replaces, eagerly and under CUDA-graph replay."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, ws, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    from cor_b200 import peer, region, synth

    n, Cc = 96, 64
    px = peer.get_exchange(n, Cc, dev)
    assert px is not None and px.ok, "peer exchange unavailable"
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    # ---- gather / reduce vs NCCL, several epochs, buffers rewritten between calls ----
    for it in range(5):
        rows = torch.randn(n, Cc, device=dev, generator=g).to(torch.bfloat16)
        px.before_produce(0)
        px.pub.copy_(rows)
        px.signal(0)
        got = px.gather()
        want = torch.empty_like(got)
        dist.all_gather_into_tensor(want, rows)
        assert torch.equal(got, want), f"gather mismatch at epoch {it}"
        grad = torch.randn(ws * n, Cc, device=dev, generator=g)
        px.before_produce(1)
        px.gall.copy_(grad)
        px.signal(1)
        red = px.reduce()
        parts = [torch.empty_like(grad) for _ in range(ws)]
        dist.all_gather(parts, grad)
        ref = torch.zeros(n, Cc, device=dev)
        for p in range(ws):                      # the kernel's fixed order: rank ascending
            ref = ref + parts[p][rank * n:(rank + 1) * n]
        assert torch.equal(red, ref), f"reduce mismatch at epoch {it}: {(red - ref).abs().max().item()}"
    for it in range(3):                          # same channel back to back: the wait_exit path
        rows = torch.randn(n, Cc, device=dev, generator=g).to(torch.bfloat16)
        px.before_produce(0)
        px.pub.copy_(rows)
        px.signal(0)
        got = px.gather()
        want = torch.empty_like(got)
        dist.all_gather_into_tensor(want, rows)
        assert torch.equal(got, want), f"forward-only gather mismatch at {it}"
    # ---- the fused step: peer exchange vs NCCL collectives, identical inputs.  Shapes the tensor-core fused step takes
    #      (bf16 features, >= 16 masks, C % 128 == 0, P % 64 == 0), otherwise region_step would use the modular NCCL path in both runs
    d = synth.make_triplets(11 + rank, B=4, M=16, C=128, h=16, w=16, H=128, W=128, hp=32, wp=32, degenerate=False)
    t = {k: torch.from_numpy(v).to(dev) for k, v in d.items()}
    t["emb"] = t["emb"].bfloat16()
    assert region._fused_ok(t["emb"], t["masks"], "auto"), "test shapes must be eligible for the fused step"

    def run(use_peer):
        os.environ["COR_PEER"] = "1" if use_peer else "0"
        pred = t["pred"].clone().requires_grad_(True)
        emb = t["emb"].clone().requires_grad_(True)
        comb = t["comb"].clone().requires_grad_(True)
        out = region.region_step(pred, emb, comb, t["masks"], fused=True)
        out.loss.backward()
        return out.loss.detach().clone(), pred.grad.clone(), emb.grad.float().clone(), comb.grad.clone()

    os.environ["COR_PEER_FUSED"] = "0"           # separate gather and similarity kernels: the arithmetic of the NCCL path
    a = run(True)
    assert any(k[0] == 4 * 16 and k[1] == 128 and px_.ok for k, px_ in peer._CACHE.items()), "the fused step did not use the peer exchange"
    b = run(False)
    names = ("loss", "g_pred", "g_emb", "g_comb")
    for x, y, nm in zip(a, b, names):
        assert torch.isfinite(x).all(), nm
        if ws == 2:
            assert torch.equal(x, y), f"{nm}: peer vs nccl differ by {(x - y).abs().max().item()}"
        else:                                    # NCCL's reduction order is its own for ws > 2
            torch.testing.assert_close(x, y, rtol=2e-2 if nm == "g_emb" else 1e-4, atol=1e-5, msg=nm)
    # the fused gather + similarity kernel (default): same rows, log-sum-exp partials merged in another order
    os.environ["COR_PEER_FUSED"] = "2"           # forced: at 8 GPUs the default would pick the separate kernels
    assert region._peer_fused_ok(4, 128, ws * 64)
    f = run(True)
    for x, y, nm in zip(f, a, names):
        assert torch.isfinite(x).all(), nm
        torch.testing.assert_close(x, y, rtol=2e-2 if nm == "g_emb" else 1e-4, atol=1e-5, msg="fused gather+sim: " + nm)
    os.environ["COR_PEER_FUSED"] = "1"
    # ---- graph replay: epochs advance on the device; replays reproduce the eager step bit for bit ----
    os.environ["COR_PEER"] = "1"
    bufs = region.StepBuffers(4, 16, C=128, h=16, w=16, H=128, W=128, hp=32, wp=32, device=dev, emb_dtype=torch.bfloat16)
    bufs.load(t)
    loss_e, grads_e = bufs._step(True, True, dict(fused=True))
    loss_e, gemb_e = loss_e.clone(), grads_e["emb"].clone()
    bufs.capture(fused=True)
    for _ in range(4):
        bufs.replay()
    torch.cuda.synchronize()
    assert torch.equal(bufs.loss.reshape(()), loss_e.reshape(())), "graph replay differs from the eager step"
    assert torch.equal(bufs.grads["emb"], gemb_e), "graph replay gradient differs from the eager step"
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print("PEER_OK", ws, flush=True)
    peer.release_all()
    os._exit(0)


if __name__ == "__main__":
    main()
