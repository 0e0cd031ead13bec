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
    # ---- the fused step: peer exchange vs NCCL collectives, identical inputs ----
    d = synth.make_triplets(11 + rank, B=4, M=8, C=64, h=16, w=16, H=128, W=128, hp=32, wp=32, degenerate=False)
    t = {k: torch.from_numpy(v).to(dev) for k, v in d.items()}

    def run(use_peer):
        os.environ["COR_PEER"] = "1" if use_peer else "0"
        pred = t["pred"].clone().requires_grad_(True)
        emb = t["emb"].clone().requires_grad_(True)
        comb = t["comb"].clone().requires_grad_(True)
        out = region.region_step(pred, emb, comb, t["masks"], fused=True)
        out.loss.backward()
        return out.loss.detach().clone(), pred.grad.clone(), emb.grad.clone(), comb.grad.clone()

    a = run(True)
    b = run(False)
    names = ("loss", "g_pred", "g_emb", "g_comb")
    for x, y, nm in zip(a, b, names):
        if ws == 2:
            assert torch.equal(x, y), f"{nm}: peer vs nccl differ by {(x - y).abs().max().item()}"
        else:                                    # NCCL's reduction order is its own for ws > 2
            torch.testing.assert_close(x, y, rtol=1e-5, atol=1e-6, msg=nm)
    # ---- graph replay: epochs advance on the device ----
    os.environ["COR_PEER"] = "1"
    bufs = region.StepBuffers(4, 8, C=64, h=16, w=16, H=128, W=128, hp=32, wp=32, device=dev, emb_dtype=torch.float32)
    bufs.load(t)
    bufs.capture(fused=True)
    for _ in range(4):
        bufs.replay()
    torch.cuda.synchronize()
    assert torch.equal(bufs.loss.reshape(()), a[0].reshape(())), "graph replay differs from the eager step"
    assert torch.equal(bufs.grads["emb"], a[2]), "graph replay gradient differs from the eager step"
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print("PEER_OK", ws, flush=True)
    peer.release_all()
    os._exit(0)


if __name__ == "__main__":
    main()
