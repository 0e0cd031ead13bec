"""Deterministic synthetic inputs for the region path.

Used by bench.py, the tests and oracle/gen_golden.py, so that every consumer sees the same
tensors for the same (seed, shape).  numpy PCG64 only -- identical on every host.

Shapes follow the reference's model outputs (lib/sam_with_sup_branch.py:57-104):
  emb   [B,C,h,w]     SAM neck output (post-LayerNorm2d, ~N(0,1))
  masks [B,M,H,W]     candidate masks in [0,1]; mask 0 of each image is the GT (query_mask)
  comb  [B,1,D]       composed query, unit rows (lib/support_branch.py:85)
  pred  [B,1,hp,wp]   mask logits from the decoder
"""
from __future__ import annotations

import numpy as np


def make_masks(rng: np.random.Generator, B: int, M: int, H: int, W: int, soft: bool = False,
               degenerate: bool = True) -> np.ndarray:
    """Union of 1-3 axis-aligned rectangles + one ellipse per mask, area fraction roughly
    0.5%-30%.  With ``degenerate`` one mask in 32 is all-zero and one in 256 all-one (exercise the
    ``valid`` paths of loss_func.py:73-77,103-110).  ``soft`` quantises a blurred edge to k/255
    like an 8-bit PNG through ``ToTensor`` (utils/dataloader.py:190)."""
    out = np.zeros((B, M, H, W), dtype=np.float32)
    yy = np.arange(H, dtype=np.float32)[:, None]
    xx = np.arange(W, dtype=np.float32)[None, :]
    for b in range(B):
        for m in range(M):
            idx = b * M + m
            if degenerate and idx % 32 == 31:
                continue
            if degenerate and idx % 256 == 129:
                out[b, m] = 1.0
                continue
            frac = rng.uniform(0.005, 0.30)
            side = np.sqrt(frac)
            a = out[b, m]
            for _ in range(int(rng.integers(1, 4))):
                rh = max(2, int(H * side * rng.uniform(0.4, 1.0)))
                rw = max(2, int(W * side * rng.uniform(0.4, 1.0)))
                y0 = int(rng.integers(0, max(1, H - rh)))
                x0 = int(rng.integers(0, max(1, W - rw)))
                a[y0:y0 + rh, x0:x0 + rw] = 1.0
            cy, cx = rng.uniform(0.2, 0.8) * H, rng.uniform(0.2, 0.8) * W
            ry, rx = max(2.0, H * side * 0.4), max(2.0, W * side * 0.4)
            a[((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0] = 1.0
            if soft:
                # 3x3 box blur of the edge, quantised to 8 bits
                p = np.pad(a, 1, mode="edge")
                s = sum(p[i:i + H, j:j + W] for i in range(3) for j in range(3)) / 9.0
                out[b, m] = np.round(s * 255.0).astype(np.float32) / np.float32(255.0)
    return out


def make_logits(rng: np.random.Generator, B: int, H: int, W: int) -> np.ndarray:
    """2*randn low-pass filtered with a 5x5 box so sigmoid has spatial structure."""
    x = rng.standard_normal((B, 1, H + 4, W + 4)).astype(np.float32) * 2.0
    c = np.pad(x, [(0, 0), (0, 0), (1, 0), (1, 0)]).cumsum(-2).cumsum(-1)
    s = c[..., 5:, 5:] - c[..., :-5, 5:] - c[..., 5:, :-5] + c[..., :-5, :-5]
    return (s / 5.0).astype(np.float32)


def unit_rows(rng: np.random.Generator, *shape) -> np.ndarray:
    x = rng.standard_normal(shape).astype(np.float32)
    return (x / np.maximum(np.linalg.norm(x, axis=-1, keepdims=True), 1e-12)).astype(np.float32)


def make_triplets(seed: int, B: int, M: int, C: int = 256, h: int = 64, w: int = 64, H: int = 1024,
                  W: int = 1024, D: int | None = None, hp: int = 256, wp: int = 256, soft: bool = False,
                  degenerate: bool = True) -> dict:
    """One batch of synthetic triplets as float32 numpy arrays."""
    rng = np.random.default_rng(seed)
    D = C if D is None else D
    return {
        "emb": rng.standard_normal((B, C, h, w)).astype(np.float32),
        "masks": make_masks(rng, B, M, H, W, soft=soft, degenerate=degenerate),
        "comb": unit_rows(rng, B, 1, D),
        "pred": make_logits(rng, B, hp, wp),
    }


def make_gallery(seed: int, n_regions: int, n_queries: int, D: int = 256, duplicate: bool = False) -> dict:
    """Validation retrieval gallery (BASELINE config 3): unit rows, tie-free unless ``duplicate``
    plants one exact duplicate region row (tests the (score desc, index asc) tie rule)."""
    rng = np.random.default_rng(seed)
    r = unit_rows(rng, n_regions, D)
    q = unit_rows(rng, n_queries, D)
    if duplicate and n_regions > 8:
        r[n_regions // 2] = r[3]
    return {"regions": r, "queries": q}
