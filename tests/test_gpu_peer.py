"""NVLink peer-memory exchange (cor_b200/peer.py, csrc/peer.cu) against the NCCL collectives it replaces.  The NCCL
comparison needs two GPUs; on a single-GPU box it is skipped and the lock-step emulation below still drives every peer
kernel with all ranks' regions on one device.  The host-side logic of the multi-rank step is covered by
tests/test_dist_cpu.py under gloo."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_peer_exchange_matches_nccl():
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "_peer_worker.py")]
    env = dict(os.environ, NCCL_DEBUG="WARN")
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "PEER_OK" in r.stdout, r.stdout[-3000:] + "\n" + r.stderr[-3000:]


class _VirtualRank:
    """One rank of a W-rank exchange whose W "peer" regions all live on cuda:0 of THIS process (plain torch buffers,
    no IPC): what cor_b200.peer.PeerExchange holds per rank, built by hand so that the test can drive the protocol in
    lock step."""

    def __init__(self, rank, world, n, Cc, flags, pubs, galls):
        from cor_b200 import _lib as L
        lib = L.load()
        dev = flags[0].device
        i64 = dict(dtype=torch.int64, device=dev)
        self.rank, self.world, self.n, self.C = rank, world, n, Cc
        self.flag_ptrs = torch.tensor([f.data_ptr() for f in flags], **i64)
        self.pub_ptrs = torch.tensor([p.data_ptr() for p in pubs], **i64)
        self.gall_ptrs = torch.tensor([g.data_ptr() for g in galls], **i64)
        self.state = torch.zeros(max(lib.cor_peer_state_bytes() // 4, 12), dtype=torch.int32, device=dev)
        self.pub, self.gall = pubs[rank], galls[rank]
        self.err_word = lib.cor_peer_error_word()

    def ctl(self, name, channel):
        from cor_b200 import ops
        ops._call(name, self.pub.device, ops.ptr(self.flag_ptrs), ops.ptr(self.state), self.rank, self.world, channel)

    def gather(self):
        from cor_b200 import ops
        out = torch.empty((self.world * self.n, self.C), dtype=torch.bfloat16, device=self.pub.device)
        ops._call("cor_peer_gather_rows", out.device, ops.ptr(self.pub_ptrs), ops.ptr(out), ops._ll(self.n * self.C * 2),
                  ops.ptr(self.flag_ptrs), ops.ptr(self.state), self.rank, self.world, 0)
        return out

    def gather_sim(self, q16, inv_tau):
        """Fused all-gather + similarity: (gathered rows, lse of this rank's queries merged from the kernel's partials)."""
        import ctypes as C_
        from cor_b200 import _lib as L, ops
        lib = L.load()
        out = torch.empty((self.world * self.n, self.C), dtype=torch.bfloat16, device=self.pub.device)
        work = torch.zeros(lib.cor_peer_gather_sim_work_bytes() // 4, dtype=torch.float32, device=out.device)
        nparts = C_.c_int(0)
        ops._call("cor_peer_gather_sim", out.device, ops.ptr(self.pub_ptrs), ops.ptr(out), ops._ll(self.n), self.C, ops.ptr(q16),
                  int(q16.shape[0]), ops._f(inv_tau), ops.ptr(work), C_.byref(nparts), ops.ptr(self.flag_ptrs), ops.ptr(self.state),
                  self.rank, self.world, 0)
        parts = work[:nparts.value * 16 * 2].view(nparts.value, 16, 2).double()
        m, sm = parts[..., 0], parts[..., 1]
        M = m.max(dim=0).values
        lse = M + torch.log((sm * torch.exp(m - M)).nan_to_num(0.0).sum(dim=0))
        return out, lse[:q16.shape[0]]

    def reduce(self):
        from cor_b200 import ops
        out = torch.empty((self.n, self.C), dtype=torch.float32, device=self.pub.device)
        ops._call("cor_peer_reduce_rows", out.device, ops.ptr(self.gall_ptrs), ops.ptr(out), ops._ll(self.n * self.C),
                  ops.ptr(self.flag_ptrs), ops.ptr(self.state), self.rank, self.world, 1)
        return out


@pytest.mark.parametrize("world", [2, 4])
def test_peer_kernels_lock_step_on_one_gpu(world):
    """The driver's GPU-test box has ONE GPU.  All W ranks' regions are placed on it and the protocol is driven in lock
    step on ONE stream: every rank's producer + signal first, then every rank's exchange kernel -- so each wait is
    already satisfied when its kernel starts and no kernel ever spins on a later launch (two processes or streams
    spinning on each other on one GPU is what B200_PROFILING.md warns against).  Same kernels, same flag / epoch /
    exit-flag arithmetic as across GPUs: peer_signal, peer_gather, peer_reduce and peer_wait_exit over 4 epochs,
    results bit-equal to the concatenation / the fixed-order sum."""
    from cor_b200 import _lib as L
    lib = L.load()
    os.environ.setdefault("COR_PEER_TIMEOUT_S", "5")
    dev = torch.device("cuda:0")
    n, Cc = 96, 64
    flags = [torch.zeros(lib.cor_peer_flag_bytes() // 4, dtype=torch.int32, device=dev) for _ in range(world)]
    pubs = [torch.zeros(n, Cc, dtype=torch.bfloat16, device=dev) for _ in range(world)]
    galls = [torch.zeros(world * n, Cc, dtype=torch.float32, device=dev) for _ in range(world)]
    ranks = [_VirtualRank(r, world, n, Cc, flags, pubs, galls) for r in range(world)]
    g = torch.Generator(device=dev).manual_seed(3)
    for it in range(4):
        rows = [torch.randn(n, Cc, device=dev, generator=g).bfloat16() for _ in range(world)]
        for r in ranks:
            if it >= 2:
                r.ctl("cor_peer_wait_exit", 0)           # forward-only pattern: same channel twice in a row
            r.pub.copy_(rows[r.rank])
            r.ctl("cor_peer_signal", 0)
        want = torch.cat(rows)
        for r in ranks:
            if it % 2 == 0:
                assert torch.equal(r.gather(), want), f"gather mismatch, rank {r.rank}, epoch {it}"
            else:
                # the fused gather + similarity kernel: same protocol, same gathered rows, plus the log-sum-exp of the rank's
                # own queries against every row
                nq = 16 if r.rank % 2 == 0 else 5
                q16 = torch.nn.functional.normalize(torch.randn(nq, Cc, device=dev, generator=g), dim=-1).bfloat16()
                got, lse = r.gather_sim(q16, 1.0 / 0.07)
                assert torch.equal(got, want), f"fused gather mismatch, rank {r.rank}, epoch {it}"
                ref = torch.logsumexp((q16.double() @ want.double().t()) / 0.07, dim=-1)
                torch.testing.assert_close(lse, ref, rtol=1e-5, atol=2e-3)
        if it >= 2:
            continue
        grads = [torch.randn(world * n, Cc, device=dev, generator=g) for _ in range(world)]
        for r in ranks:
            r.gall.copy_(grads[r.rank])
            r.ctl("cor_peer_signal", 1)
        for r in ranks:
            ref = torch.zeros(n, Cc, device=dev)
            for p in range(world):                       # the kernel's fixed order: rank ascending
                ref = ref + grads[p][r.rank * n:(r.rank + 1) * n]
            assert torch.equal(r.reduce(), ref), f"reduce mismatch, rank {r.rank}, epoch {it}"
    torch.cuda.synchronize()
    for r in ranks:
        assert int(r.state[r.err_word].item()) == 0, "a wait expired"
    # an unsatisfied wait does not trap: it records the epoch in the error word and the kernel returns
    os.environ["COR_PEER_TIMEOUT_S"] = "0.5"
    lonely = ranks[0]
    lonely.pub.copy_(rows[0])
    lonely.ctl("cor_peer_signal", 0)
    lonely.gather()                                      # the other ranks never signalled this epoch
    torch.cuda.synchronize()
    os.environ["COR_PEER_TIMEOUT_S"] = "5"
    assert int(lonely.state[lonely.err_word].item()) != 0
