"""pytest configuration: the ``gpu`` marker and shared fixtures."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture
def golden():
    return load_golden


@pytest.fixture(autouse=True)
def _work_buffer_canary(request, monkeypatch):
    """GPU tests: every scratch buffer handed to the C ABI (``ops._work``, sized by the library's ``*_work_bytes``) gets a
    4 KB canary behind it, checked when the test ends -- a launcher that writes more partials than its ``*_work_bytes``
    promised corrupts whatever the allocator placed next and fails a DIFFERENT test, depending on the order."""
    if request.node.get_closest_marker("gpu") is None:
        yield
        return
    import torch
    from cor_b200 import ops
    pad, held = 4096, []

    def verify():
        torch.cuda.synchronize()
        for buf, n in held:
            assert bool((buf[n:] == 0xA5).all()), f"a kernel wrote past its {n}-byte work buffer"
        held.clear()

    def guarded(nbytes, dev):
        n = (max(int(nbytes), 16) + 15) // 16 * 16
        if not torch.cuda.is_current_stream_capturing() and (len(held) >= 64 or sum(b.numel() for b, _ in held) > (1 << 30)):
            verify()
        buf = torch.empty(n + pad, dtype=torch.uint8, device=dev)
        buf[n:] = 0xA5
        held.append((buf, n))
        return buf[:n]

    monkeypatch.setattr(ops, "_work", guarded)
    yield
    verify()
