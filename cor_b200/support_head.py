"""Composed-query head of the support branch on the B200 path: everything between the pooled
support feature and the composed query ``comb_support_feat`` the region losses consume --

    ln_channel_last -> CirFuseModule.compose_img_text -> dim_proj -> F.normalize        lib/support_branch.py:60-86
                       (lib/support_model/cir_feature_fuse.py:44-64)

Every ``nn.Linear`` of the head (three gate MLPs of ``CirFuseModule``, the two projection layers) runs forward and
backward on the tcgen05 GEMM through :func:`cor_b200.linear.linear`, with bias, ReLU / Sigmoid / GELU and the dropout
mask in the GEMM epilogue and ``torch.cat`` folded into the operand cast; both L2-normalisations run on
``cor_l2_normalize``.  What is left to torch is [N, C]-sized glue on N <= a few dozen rows: the LayerNorm of the pooled
feature, the two sigmoid gates' element-wise products and the convex mix (five tiny element-wise ops), and the dropout
MASKS, which are drawn with torch's own RNG in the reference's call order so that a seeded run reproduces the
reference's masks exactly (``nn.Dropout`` at cir_feature_fuse.py:25,32,39 and support_branch.py:50,53).

:func:`composed_query` takes the reference's own sub-modules (``SupportBranch.ln_channel_last``, ``.cir_fuse``,
``.dim_proj``) and uses their parameters in place: state_dict keys, checkpoints and optimizers are untouched.
``hooks.install()`` rebinds ``SupportBranch.forward`` to :func:`support_branch_forward`.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops
from .linear import linear

__all__ = ["composed_query", "compose_img_text", "support_branch_forward"]


def _mask(like: torch.Tensor, p: float, training: bool):
    """What ``nn.Dropout(p)`` multiplies by (0 or 1/(1-p)), drawn from torch's generator exactly like the reference's own
    dropout call on a tensor of this shape; None in eval mode."""
    if not training or p <= 0.0:
        return None
    return F.dropout(torch.ones_like(like), p=p, training=True)


def _mlp2(seq, x, x2, final_act, training):
    """Linear -> ReLU -> Dropout -> Linear -> Sigmoid (cir_feature_fuse.py:22-43): two fused GEMMs."""
    lin1, drop, lin2 = seq[0], seq[2], seq[3]
    rows = x.shape[0]
    m = _mask(torch.empty((rows, lin1.out_features), device=x.device), drop.p, training)
    h = linear(x, lin1.weight, lin1.bias, "relu", m, x2)
    if lin2.out_features % 8:          # the C -> 1 layer of dynamic_scalar: a 16-byte TMA row pitch needs 8 bf16 columns;
        return torch.sigmoid(F.linear(h, lin2.weight.float(), lin2.bias.float()))       # N x C MACs, left to torch
    return linear(h, lin2.weight, lin2.bias, final_act)


def compose_img_text(cir_fuse, image_features: torch.Tensor, text_features: torch.Tensor, training: bool = None) -> dict:
    """``CirFuseModule.compose_img_text`` (cir_feature_fuse.py:44-64) with the module's own parameters."""
    training = cir_fuse.training if training is None else training
    img, txt = image_features.float(), text_features.float()
    # dropout masks are drawn in the order the reference draws them: atten_Image, atten_Text, dynamic_scalar
    atten_i = _mlp2(cir_fuse.atten_Image, img, txt, "sigmoid", training)          # raw_combined = cat(img, txt): fused into the cast
    atten_t = _mlp2(cir_fuse.atten_Text, img, txt, "sigmoid", training)
    img2, txt2 = atten_i * img, atten_t * txt
    dyn = _mlp2(cir_fuse.dynamic_scalar, img2, txt2, "sigmoid", training)         # [N, 1]
    com = dyn * img2 + (1 - dyn) * txt2
    return {"repres": ops.l2_normalize(com), "fuseimg": img2, "fusetxt": txt2, "dynamic_scalar": dyn}


def composed_query(branch, support_feat: torch.Tensor, text_feat: torch.Tensor) -> torch.Tensor:
    """lib/support_branch.py:60-86: pooled support feature [N,(1,)C] + text feature [N,(1,)C] -> comb_support_feat [N,1,256]."""
    training = branch.training
    x = branch.ln_channel_last(support_feat)
    x = x.squeeze(1) if x.dim() == 3 else x
    t = text_feat.squeeze(1) if text_feat.dim() == 3 else text_feat
    rep = compose_img_text(branch.cir_fuse, x, t, training)["repres"]
    lin1, drop1, lin2, drop2 = branch.dim_proj[0], branch.dim_proj[2], branch.dim_proj[3], branch.dim_proj[5]
    rows = rep.shape[0]
    m1 = _mask(torch.empty((rows, lin1.out_features), device=rep.device), drop1.p, training)
    h = linear(rep, lin1.weight, lin1.bias, "gelu", m1)
    m2 = _mask(torch.empty((rows, lin2.out_features), device=rep.device), drop2.p, training)
    h = linear(h, lin2.weight, lin2.bias, "gelu", m2)
    return ops.l2_normalize(h).unsqueeze(1)


def support_branch_forward(self, support_input, change_text, mask_input):
    """Drop-in for ``SupportBranch.forward`` (lib/support_branch.py:56-87): SigLIP and the two LayerNorms are the module's
    own; pooling runs on the CUDA tails (hooks.install patches the pooling classes) and the head on :func:`composed_query`."""
    _, text_feat, _, dense = self.siglip(support_input, change_text)
    dense = self.ln_channel_first(dense)
    support_feat = self.mask_pooling(dense, mask_input)
    return composed_query(self, support_feat, text_feat)
