"""The bench 'step' through the REFERENCE'S OWN functions (``oracle/_ref``, made by oracle/build_ref.py).

TEST / BASELINE INFRASTRUCTURE ONLY.  Used by ``bench.py --impl reference`` and the ``cpu_baseline`` leg when
``oracle/_ref`` exists (``cpu_baseline.kind == "reference"``); otherwise those legs time oracle/aten_port.py
(``"port"``).  Class R parts call the unmodified reference modules; multi-mask inputs use SURVEY.md 8c's recipe
(``repeat_interleave`` of the feature map, masks flattened to [B*M,1,H,W]) because ``utils/loss_func.mask_pooling``
takes one mask per sample (loss_func.py:35-56).  The Class-N InfoNCE has no reference code: it is the
same ``F.cross_entropy`` restatement as oracle/aten_port.infonce on top of the reference-pooled rows.
"""
from __future__ import annotations

import importlib.util
import json
import hashlib
import os

import torch
import torch.nn.functional as tf

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
_mods = {}


def available() -> bool:
    return os.path.exists(os.path.join(REF, "MANIFEST.json")) and os.path.exists(os.path.join(REF, "utils", "loss_func.py"))


def verified() -> bool:
    """Every file under oracle/_ref still has the sha256 its MANIFEST recorded at copy time."""
    try:
        man = json.load(open(os.path.join(REF, "MANIFEST.json")))
        return all(hashlib.sha256(open(os.path.join(REF, rel), "rb").read()).hexdigest() == h for rel, h in man["files"].items())
    except Exception:
        return False


def module(rel: str):
    """Import one reference file by path under a private name (no sys.path games, no clash with the reference checkout)."""
    if rel not in _mods:
        spec = importlib.util.spec_from_file_location("cor_ref_" + rel.replace("/", "_")[:-3], os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        _mods[rel] = m
    return _mods[rel]


def region_step_loss(pred, emb, comb, masks, tau: float = 0.07, nce_weight: float = 1.0):
    """trainer_v3_g.py:67-73 on the GT masks through the reference's functions + Class-N InfoNCE over all B*M regions."""
    lf = module("utils/loss_func.py")
    B, M = masks.shape[:2]
    gt = masks[:, 0:1]
    target = tf.interpolate(gt, size=pred.shape[2:], mode="bilinear", align_corners=False)      # trainer_v3_g.py:67
    loss = lf.wbce_with_wiou_loss(pred, target)                                                   # :68
    loss = loss + 5 * lf.fg_feat_similarity_loss(emb, comb, gt)                                   # :69
    loss = loss + 5 * lf.bg_feat_similarity_loss(emb, comb, gt)                                   # :71
    rows = lf.mask_pooling(emb.repeat_interleave(M, 0), masks.reshape(B * M, 1, *masks.shape[2:])).reshape(B * M, -1)
    q = comb[:, 0, :].float()
    r16 = rows + (rows.bfloat16().to(rows.dtype) - rows).detach()
    q16 = q + (q.bfloat16().to(q.dtype) - q).detach()
    nce = tf.cross_entropy((q16 @ r16.t()) / tau, torch.arange(B, device=masks.device) * M)
    return loss + nce_weight * nce, rows
