// Pieces shared by the two forward kernels of the segmentation loss (seg_loss.cu: generic 64x64 tiles; seg_strip.cu:
// TMA-streamed row strips for same-size and exact-4x masks), its finalize kernel and its backward.
#pragma once
#include "common.cuh"

namespace cor {

// Per-sample partial sums both forward kernels produce (fp64 across threads, fp32 in `per_sample`):
//   0 sum w          1 sum w*bce      2 sum p*t*w     3 sum (p+t)*w    4 sum p*t
//   5 sum p          6 sum t          7 sum focal     8 sum bce        9 (unused)
// with t the (resampled) target, w = 1 + 5|boxmean31(t) - t| (utils/loss_func.py:18-19), p = sigmoid(logit),
// bce = (1-t)x - logsigmoid(x) (ATen binary_cross_entropy_with_logits), focal = a_t (1-p_t)^gamma bce.
constexpr int kNP = 10;

// Loss terms per sample, combined as loss = mean_n sum_k coef[k] * term_k[n]  (COR_SEG_* indices, include/cor_b200.h)
//   0 wbce  = s1/s0                                   (loss_func.py:21-22)
//   1 wiou  = 1 - (s2 + 1e-6)/(s3 - s2 + 1e-6)        (loss_func.py:25-29)
//   2 dice  = 1 - (2 s4 + sm)/(s5 + s6 + sm)          [Class N]
//   3 bce   = s8 / HW                                 [Class N]
//   4 iou   = 1 - (s4 + 1e-6)/(s5 + s6 - s4 + 1e-6)   [Class N]
//   5 wdice = 1 - (2 s2 + sm)/(s3 + sm)               [Class N]
//   6 focal = s7 / HW                                 [Class N]
constexpr int kNC = 7;
struct SegCoef {
  float c[kNC];
};

// One pixel's contribution to the partial sums (identical in both forward kernels).
__device__ __forceinline__ void seg_pixel_terms(float zz, float t, float boxsum, float focal_alpha, float focal_gamma, float (&f)[kNP],
                                                float& wgt_out) {
  const float wgt = 1.f + 5.f * fabsf(boxsum * (1.f / 961.f) - t);
  const float e = __expf(-fabsf(zz));
  const float bce = (1.f - t) * zz - (fminf(zz, 0.f) - __logf(1.f + e));   // (1-t)x - logsigmoid(x)
  const float inv = __fdividef(1.f, 1.f + e);
  const float p = zz >= 0.f ? inv : e * inv;
  f[0] += wgt;
  f[1] = fmaf(wgt, bce, f[1]);
  f[2] = fmaf(p * t, wgt, f[2]);
  f[3] = fmaf(p + t, wgt, f[3]);
  f[4] = fmaf(p, t, f[4]);
  f[5] += p;
  f[6] += t;
  f[8] += bce;
  if (focal_gamma >= 0.f) {
    const float pt = p * t + (1.f - p) * (1.f - t);
    const float at = focal_alpha * t + (1.f - focal_alpha) * (1.f - t);
    f[7] = fmaf(at * __powf(fmaxf(1.f - pt, 0.f), focal_gamma), bce, f[7]);
  }
  wgt_out = wgt;
}

// seg_strip.cu: returns COR_OK after launching, or COR_EINVAL *without setting an error* when the shape is not one the
// strip kernel serves (the caller then takes the tile kernel).  part: [N * (*strips)][kNP] doubles.
int seg_strip_try_launch(const void* pred, int pred_dtype, const void* mask, int mask_dtype, float mscale, int N, int H, int W, int Hm,
                         int Wm, long long mask_nstride, float focal_alpha, float focal_gamma, float* t_save, float* w_save,
                         double* part, int* strips, cudaStream_t st);
int seg_strip_max_strips(int H);   // upper bound of *strips for work-buffer sizing

}  // namespace cor
