"""Drop-in for the metric functions of the reference's ``utils/trainer_v3_g.py:381-443``:
``compute_dice / compute_mae / compute_iou / compute_mdice / compute_miou(pred, gt[, smooth])`` -> [B].

``val_stage`` calls all five on the same ``(pred, gt)`` (trainer_v3_g.py:233-237); the first call runs ONE pass
over the two 1024x1024 maps on the device and parks the other four results for the calls that follow.
"""
from __future__ import annotations

import weakref

import torch

from . import ops

__all__ = ["compute_dice", "compute_mae", "compute_iou", "compute_mdice", "compute_miou"]

_COLS = {"dice": 0, "mae": 1, "iou": 2, "mdice": 3, "miou": 4}
_memo = {}


def _all(pred: torch.Tensor, gt: torch.Tensor, smooth: float) -> torch.Tensor:
    assert pred.shape == gt.shape, f"Shape mismatch: pred {pred.shape} vs gt {gt.shape}"
    hit = _memo.get("entry")
    if hit is not None:
        (rp, rg, vp, vg, sm), out = hit
        if rp() is pred and rg() is gt and vp == pred._version and vg == gt._version and sm == smooth:
            return out
    out = ops.soft_metrics(pred, gt, smooth)
    _memo["entry"] = ((weakref.ref(pred), weakref.ref(gt), pred._version, gt._version, smooth), out)
    return out


def compute_dice(pred: torch.Tensor, gt: torch.Tensor, smooth=1e-5):
    return _all(pred, gt, smooth)[:, 0]


def compute_mae(pred: torch.Tensor, gt: torch.Tensor):
    return _all(pred, gt, 1e-5)[:, 1]


def compute_iou(pred: torch.Tensor, gt: torch.Tensor, smooth=1e-5):
    return _all(pred, gt, smooth)[:, 2]


def compute_mdice(pred: torch.Tensor, gt: torch.Tensor, smooth=1e-5):
    return _all(pred, gt, smooth)[:, 3]


def compute_miou(pred: torch.Tensor, gt: torch.Tensor, smooth=1e-5):
    return _all(pred, gt, smooth)[:, 4]
