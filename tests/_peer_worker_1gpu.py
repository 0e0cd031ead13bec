"""Worker of tests/test_gpu_peer.py::test_peer_exchange_two_ranks_on_one_gpu: TWO processes that share cuda:0 (the
driver's GPU-test box has one GPU).  The processes exchange CUDA IPC handles exactly like ranks on different GPUs do;
torch.distributed runs over gloo (NCCL refuses two ranks on one device), which only carries the handles, the
barriers and the CPU copies the results are checked against.  Kernels of the two processes time-slice on the GPU, so a
flag wait costs a scheduling quantum instead of a microsecond: slow, but the same code path
(peer_signal / peer_gather / peer_reduce / peer_wait_exit and the fused step around them)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def cpu_all_gather(x):
    parts = [torch.empty_like(x) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, x.contiguous())
    return parts


def main():
    rank, ws = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    os.environ["COR_PEER_ANY_BACKEND"] = "1"
    os.environ.setdefault("COR_PEER_TIMEOUT_S", "30")
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    dist.init_process_group("gloo")
    torch.set_num_threads(4)
    from cor_b200 import peer, region, synth
    from oracle import aten_port as ap

    n, Cc = 96, 64
    px = peer.get_exchange(n, Cc, dev)
    assert px is not None and px.ok, "peer exchange unavailable (CUDA IPC between two processes on one device)"
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    for it in range(3):
        rows = torch.randn(n, Cc, device=dev, generator=g).to(torch.bfloat16)
        px.before_produce(0)
        px.pub.copy_(rows)
        px.signal(0)
        got = px.gather()
        want = torch.cat(cpu_all_gather(rows.float().cpu()))
        assert torch.equal(got.float().cpu(), want), f"gather mismatch at epoch {it}"
        grad = torch.randn(ws * n, Cc, device=dev, generator=g)
        px.before_produce(1)
        px.gall.copy_(grad)
        px.signal(1)
        red = px.reduce()
        parts = cpu_all_gather(grad.cpu())
        ref = torch.zeros(n, Cc)
        for p in range(ws):
            ref = ref + parts[p][rank * n:(rank + 1) * n]
        assert torch.equal(red.cpu(), ref), f"reduce mismatch at epoch {it}"
    for it in range(2):                          # same channel back to back: the wait_exit path
        rows = torch.randn(n, Cc, device=dev, generator=g).to(torch.bfloat16)
        px.before_produce(0)
        px.pub.copy_(rows)
        px.signal(0)
        got = px.gather()
        assert torch.equal(got.float().cpu(), torch.cat(cpu_all_gather(rows.float().cpu()))), f"forward-only gather mismatch at {it}"
    px.check()

    # ---- the fused tensor-core step with gathered negatives, against the ATen port evaluated on BOTH ranks' inputs ----
    B, M, C = 3, 16, 128
    ds = [synth.make_triplets(11 + r, B=B, M=M, C=C, h=16, w=16, H=64, W=64, hp=32, wp=32, degenerate=False) for r in range(ws)]
    for d in ds:
        d["emb"] = torch.from_numpy(d["emb"]).bfloat16().float().numpy()
    t = {k: torch.from_numpy(v).to(dev) for k, v in ds[rank].items()}
    pred = t["pred"].clone().requires_grad_(True)
    emb = t["emb"].bfloat16().requires_grad_(True)
    comb = t["comb"].clone().requires_grad_(True)
    assert region._fused_ok(emb, t["masks"], "auto")
    out = region.region_step(pred, emb, comb, t["masks"], tau=0.07, gather=True)
    out.loss.backward()
    torch.cuda.synchronize()
    peer.check_all()
    assert any(k[0] == B * M and k[1] == C and p_.ok for k, p_ in peer._CACHE.items()), "the fused step did not use the peer exchange"
    # port: every rank's loss is its own trainer loss + InfoNCE of ITS queries against ALL regions; this rank's
    # gradients are those of the SUM over ranks (what the reduce of d loss / d gathered rows delivers)
    leaves, losses, rows = [], [], []
    for r in range(ws):
        lv = {k: torch.from_numpy(ds[r][k]).requires_grad_(k != "masks") for k in ("pred", "emb", "comb", "masks")}
        leaves.append(lv)
        rows.append(ap.multi_mask_regions(lv["emb"], lv["masks"]).reshape(B * M, -1))
    all_rows = torch.cat(rows)
    for r in range(ws):
        lv = leaves[r]
        l_r, _ = ap.region_step_loss(lv["pred"], lv["emb"], lv["comb"], lv["masks"], tau=0.07, regions_all=all_rows, target_offset=r * B * M)
        losses.append(l_r)
    torch.stack(losses).sum().backward()
    want = leaves[rank]

    def rel(a, b):
        a, b = a.detach().float().cpu().double(), b.detach().double()
        return float((a - b).norm() / b.norm())

    assert abs(float(out.loss) - float(losses[rank])) < 1e-3 * abs(float(losses[rank])), (float(out.loss), float(losses[rank]))
    assert rel(pred.grad, want["pred"].grad) < 5e-3
    assert rel(comb.grad, want["comb"].grad) < 5e-3
    assert rel(emb.grad, want["emb"].grad) < 1e-2, rel(emb.grad, want["emb"].grad)

    # ---- CUDA-graph replay reproduces the eager step, with an eager forward-only call between replays ----
    bufs = region.StepBuffers(B, M, C=C, h=16, w=16, H=64, W=64, hp=32, wp=32, device=dev, emb_dtype=torch.bfloat16)
    bufs.load(t)
    loss_e, grads_e = bufs._step(True, True, dict(gather=True))
    loss_e, gemb_e = loss_e.clone(), grads_e["emb"].clone()
    bufs.capture(gather=True)
    bufs.replay()
    with torch.no_grad():
        region.region_step(t["pred"], t["emb"].bfloat16(), t["comb"], t["masks"], tau=0.07, gather=True)   # eval between replays
    bufs.replay()
    torch.cuda.synchronize()
    peer.check_all()
    assert torch.equal(bufs.loss.reshape(()), loss_e.reshape(())), "graph replay differs from the eager step"
    assert torch.equal(bufs.grads["emb"], gemb_e), "graph replay gradient differs from the eager step"
    dist.barrier()
    if rank == 0:
        print("PEER1GPU_OK", ws, flush=True)
    peer.release_all()
    os._exit(0)


if __name__ == "__main__":
    main()
