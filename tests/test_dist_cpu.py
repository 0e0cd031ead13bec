"""Host-side multi-rank logic on CPU (gloo, world_size 2): the autograd-aware all-gather of region
embeddings, the target offsetting, and the sharded top-k merge.  The arithmetic here is the oracle's
(tests may use it); what is under test is cor_b200/dist.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cor_b200 import dist as cdist
from cor_b200 import synth
from oracle import aten_port as ap
from oracle import np_oracle as no


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, ws, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        ret[rank] = fn(rank, ws)
    finally:
        dist.destroy_process_group()


def run_ranks(fn, ws=2):
    mgr = mp.Manager()
    for attempt in range(2):                 # the probed port can be taken before gloo binds it: one retry on a fresh port
        ret = mgr.dict()
        try:
            mp.spawn(_worker, args=(ws, _free_port(), fn, ret), nprocs=ws, join=True)
            return [ret[r] for r in range(ws)]
        except Exception:                     # noqa: BLE001
            if attempt:
                raise


def _nce_rank(rank, ws):
    n_local, nq_local, D = 12, 3, 16
    g = synth.make_gallery(100, n_local * ws, nq_local * ws, D=D)
    R = torch.from_numpy(g["regions"][rank * n_local:(rank + 1) * n_local]).requires_grad_(True)
    Q = torch.from_numpy(g["queries"][rank * nq_local:(rank + 1) * nq_local]).requires_grad_(True)
    tgt_local = torch.arange(nq_local) * 4
    allR = cdist.all_gather_rows(R)
    assert allR.shape == (n_local * ws, D)
    loss = ap.infonce(allR, Q, tgt_local + rank * n_local, 0.07)
    loss.backward()
    return loss.item(), R.grad.numpy(), Q.grad.numpy()


def test_all_gather_rows_forward_backward_matches_full_batch():
    ws, n_local, nq_local, D = 2, 12, 3, 16
    out = run_ranks(_nce_rank, ws)
    g = synth.make_gallery(100, n_local * ws, nq_local * ws, D=D)
    R = torch.from_numpy(g["regions"]).requires_grad_(True)
    Q = torch.from_numpy(g["queries"]).requires_grad_(True)
    tgt = torch.cat([torch.arange(nq_local) * 4 + r * n_local for r in range(ws)])
    full = ap.infonce(R, Q, tgt, 0.07)
    full.backward()
    np.testing.assert_allclose(np.mean([o[0] for o in out]), full.item(), rtol=1e-6)
    for r in range(ws):
        # per-rank mean losses: summed region grads == ws * d(full-batch mean)/dR  (DDP then divides by ws)
        np.testing.assert_allclose(out[r][1], ws * R.grad.numpy()[r * n_local:(r + 1) * n_local], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(out[r][2], ws * Q.grad.numpy()[r * nq_local:(r + 1) * nq_local], rtol=1e-4, atol=1e-7)


def _topk_rank(rank, ws):
    g = synth.make_gallery(101, 64, 5, D=16, duplicate=True)
    lo, hi = cdist.shard_range(64, rank, ws)
    idx, sc = no.topk_retrieve(g["regions"][lo:hi], g["queries"], 6)
    gi, gs = cdist.merge_topk(torch.from_numpy(idx), torch.from_numpy(sc), lo, 6)
    return gi.numpy(), gs.numpy()


def test_sharded_topk_merge_equals_global_topk():
    out = run_ranks(_topk_rank, 2)
    g = synth.make_gallery(101, 64, 5, D=16, duplicate=True)
    ridx, rsc = no.topk_retrieve(g["regions"], g["queries"], 6)
    for gi, gs in out:
        np.testing.assert_array_equal(gi, ridx)
        np.testing.assert_array_equal(gs, rsc)


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 102400):
        for ws in (1, 2, 3, 8):
            spans = [cdist.shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_is_identity():
    x = torch.randn(4, 3, requires_grad=True)
    assert cdist.all_gather_rows(x) is x
    i, s = cdist.merge_topk(torch.arange(6).view(2, 3), torch.ones(2, 3), 10, 2)
    assert i.tolist() == [[10, 11], [13, 14]]


def _peer_rank(rank, ws):
    from cor_b200 import peer
    # gloo group on CPU: the NVLink peer exchange must step aside (-> the torch.distributed collectives), never raise
    return peer.get_exchange(64, 32, torch.device("cpu")) is None and peer.enabled()


def test_peer_exchange_steps_aside_without_nccl(monkeypatch):
    from cor_b200 import peer
    assert peer.get_exchange(64, 32, torch.device("cpu")) is None          # not distributed at all
    assert all(run_ranks(_peer_rank))
    monkeypatch.setenv("COR_PEER", "0")
    assert not peer.enabled()


def test_peer_exchange_wait_policy(monkeypatch):
    """Host-side half of the peer protocol (cor_b200/peer.py): a wait_exit kernel is issued before a producer only when
    no exchange on the OTHER channel ran since the last one on this channel -- or ALWAYS while capturing a CUDA graph,
    because the host-side history would be frozen into the graph and an eager forward-only call between replays would
    break it.  Every exchange needs exactly one signal before it."""
    from cor_b200 import peer
    px = object.__new__(peer.PeerExchange)
    px._last = None
    calls = []
    px._ctl = lambda name, ch: calls.append((name, ch))
    capturing = {"v": False}
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: capturing["v"])

    def exchange(ch):                      # what gather()/reduce() record
        px._last = ch

    # training loop: gather, reduce, gather, reduce -> never a wait
    for _ in range(3):
        px.before_produce(0); px.signal(0); exchange(0)
        px.before_produce(1); px.signal(1); exchange(1)
    assert [c for c in calls if c[0] == "cor_peer_wait_exit"] == []
    assert [c for c in calls if c[0] == "cor_peer_signal"] == [("cor_peer_signal", 0), ("cor_peer_signal", 1)] * 3
    # forward-only loop: the same channel twice in a row -> wait before the second producer
    calls.clear()
    px._last = None
    px.before_produce(0); exchange(0)
    px.before_produce(0); exchange(0)
    assert calls == [("cor_peer_wait_exit", 0)]
    # capture: the wait is always part of the graph, whatever ran before the capture
    for last in (None, 0, 1):
        calls.clear()
        px._last = last
        capturing["v"] = True
        px.before_produce(0)
        px.before_produce(1)
        assert calls == [("cor_peer_wait_exit", 0), ("cor_peer_wait_exit", 1)]
