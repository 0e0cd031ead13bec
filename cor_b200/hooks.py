"""Installing the B200 path into an unmodified checkout of the reference.

The reference has no plugin registry; its hot path is reached through Python names
(SURVEY.md 8b).  ``install()`` rebinds exactly those names:

  utils.loss_func.{wbce_with_wiou_loss, mask_pooling, fg_feat_similarity_loss, bg_feat_similarity_loss}
  utils.trainer_v3_g.{wbce_with_wiou_loss, fg_feat_similarity_loss, bg_feat_similarity_loss}   (imported by name, :5-9)
  utils.trainer_v3_g.{compute_dice, compute_mae, compute_iou, compute_mdice, compute_miou}     (val_stage, :233-237)
  lib.support_model.mask_adapter.{MaskedPooling.forward, MaskAdapterPooling.forward}
      -> so SupportBranch(mask_pooling="MaskedPooling"|"MaskAdapterPooling") (lib/support_branch.py:29-40)
         built through build_model_with_query_support_feat(..., mask_pooling=) (lib/build_model.py:14-20,72)
         runs the CUDA tails with its own parameters and state_dict untouched.
  lib.support_branch.SupportBranch.forward                  (composed-query head on the tcgen05 GEMM, support_head.py)
  lib.sam_model.mask_decoder.MaskDecoder.predict_masks      (hypernetwork product on csrc/hyper_logits.cu)
      -> both use the module's own parameters in place; opt out with install(head=False, decoder=False).

Nothing else in the reference changes: my_train_a.py / trainer loops / checkpoints keep working.
"""
from __future__ import annotations

import importlib
import sys

import torch.nn.functional as F

from . import loss_func as _lf
from . import mask_adapter as _ma
from . import mask_decoder as _md
from . import metrics as _mt
from . import support_head as _sh

_LOSS_NAMES = ("wbce_with_wiou_loss", "mask_pooling", "fg_feat_similarity_loss", "bg_feat_similarity_loss")
_METRIC_NAMES = ("compute_dice", "compute_mae", "compute_iou", "compute_mdice", "compute_miou")
_saved = {}


def _masked_forward(self, clip_feature, mask):
    return _ma.masked_pool_tail(clip_feature, mask)


def _adapter_forward(self, clip_feature, mask):
    if mask.shape[-2:] != clip_feature.shape[-2:]:
        mask = F.interpolate(mask, size=clip_feature.shape[-2:], mode="bilinear", align_corners=False)
    if _ma._adapter_maps_ok(clip_feature):
        maps = _ma.adapter_maps(self, clip_feature, mask)        # the reference module's own parameters, our kernels
    else:
        maps = self.get_mask_map(self.channel_clip_to_maskadapter(clip_feature), mask)
    return _ma.softmax_map_pool_tail(maps, clip_feature, self.num_output_maps)


def _maybe(name):
    if name in sys.modules:
        return sys.modules[name]
    try:
        return importlib.import_module(name)
    except Exception:
        return None


def install(loss_module="utils.loss_func", trainer_module="utils.trainer_v3_g",
            adapter_module="lib.support_model.mask_adapter", branch_module="lib.support_branch",
            decoder_module="lib.sam_model.mask_decoder", head=True, decoder=True):
    """Rebind the reference's hot-path names to the CUDA implementations.  Modules that cannot be
    imported (e.g. the trainer without ``accelerate``) are skipped.  Returns the list patched."""
    done = []
    lm = _maybe(loss_module) if isinstance(loss_module, str) else loss_module
    if lm is not None:
        for n in _LOSS_NAMES:
            _saved.setdefault((lm.__name__, n), getattr(lm, n))
            setattr(lm, n, getattr(_lf, n))
        done.append(lm.__name__)
    tm = _maybe(trainer_module) if isinstance(trainer_module, str) else trainer_module
    if tm is not None:
        for n in _LOSS_NAMES:
            if hasattr(tm, n):
                _saved.setdefault((tm.__name__, n), getattr(tm, n))
                setattr(tm, n, getattr(_lf, n))
        for n in _METRIC_NAMES:          # val_stage's five soft metrics (trainer_v3_g.py:233-237, :381-443)
            if hasattr(tm, n):
                _saved.setdefault((tm.__name__, n), getattr(tm, n))
                setattr(tm, n, getattr(_mt, n))
        done.append(tm.__name__)
    am = _maybe(adapter_module) if isinstance(adapter_module, str) else adapter_module
    if am is not None:
        _saved.setdefault((am.__name__, "MaskedPooling.forward"), am.MaskedPooling.forward)
        _saved.setdefault((am.__name__, "MaskAdapterPooling.forward"), am.MaskAdapterPooling.forward)
        am.MaskedPooling.forward = _masked_forward
        am.MaskAdapterPooling.forward = _adapter_forward
        done.append(am.__name__)
    bm = (_maybe(branch_module) if isinstance(branch_module, str) else branch_module) if head else None
    if bm is not None and hasattr(bm, "SupportBranch"):
        _saved.setdefault((bm.__name__, "SupportBranch.forward"), bm.SupportBranch.forward)
        bm.SupportBranch.forward = _sh.support_branch_forward
        done.append(bm.__name__)
    dm = (_maybe(decoder_module) if isinstance(decoder_module, str) else decoder_module) if decoder else None
    if dm is not None and hasattr(dm, "MaskDecoder"):
        _saved.setdefault((dm.__name__, "MaskDecoder.predict_masks"), dm.MaskDecoder.predict_masks)
        dm.MaskDecoder.predict_masks = _md.predict_masks
        done.append(dm.__name__)
    return done


def uninstall():
    """Restore every name ``install`` replaced."""
    for (mod, name), val in list(_saved.items()):
        m = sys.modules.get(mod)
        if m is None:
            continue
        if "." in name:
            cls, attr = name.split(".")
            setattr(getattr(m, cls), attr, val)
        else:
            setattr(m, name, val)
    _saved.clear()
