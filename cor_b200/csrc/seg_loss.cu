// Segmentation loss: weighted BCE + weighted IoU with the 31x31 box-filter edge weight
// (utils/loss_func.py:5-32), with the target resample of utils/trainer_v3_g.py:67 fused in.
//
// One CTA per 64x64 logit tile.  The (resampled) target incl. a 15-pixel zero-padded halo is built
// once in shared memory (4 taps per value straight from the full-resolution mask), the box filter is
// separable sliding-window sums in shared memory, and the per-pixel terms are reduced in registers ->
// warp shuffles -> one partial record per tile.  Compulsory HBM traffic only: logits once, the mask
// sectors the taps touch once; ~15 full passes over [B,1,256,256] in the eager reference become one.
//
// Two forward kernels share the per-pixel arithmetic (seg_common.cuh): the row-strip streaming kernel of seg_strip.cu
// serves the shapes the reference produces (target at the logit size, or a mask at exactly 4x it); this file's 64x64
// tile kernel serves every other resample ratio, wide images and oddly aligned buffers.
#include <stdlib.h>

#include "seg_common.cuh"

namespace cor {

constexpr int kT = 64;            // tile edge
constexpr int kHalo = 15;         // (31-1)/2
constexpr int kTH = kT + 2 * kHalo;   // 94
constexpr int kTS = kTH + 1;      // padded row stride (odd -> conflict-free column walks)
constexpr int kHS = kT + 1;
constexpr int kSegThreads = 512;  // 64 columns x 8 row segments of 8 rows
constexpr int kRowsPT = kT / (kSegThreads / 64);   // rows per thread in step 3

struct SegSmem {
  float t[kTH * kTS];
  float hs[kTH * kHS];
};

// One bilinear sample of the mask at logit pixel (gy, gx): issue the loads (tap values into v[4]) ...
template <typename TM>
__device__ __forceinline__ void tap_load(const TM* __restrict__ mbase, int Hm, int Wm, float sh, float sw, int gy, int gx, bool same,
                                         float (&v)[4], float (&l)[4]) {
  if (same) {
    v[0] = to_f<TM>(__ldg(mbase + (long long)gy * Wm + gx));
    v[1] = v[2] = v[3] = 0.f;
    l[0] = 1.f; l[1] = 0.f; l[2] = 1.f; l[3] = 0.f;
    return;
  }
  int y0, y1, x0, x1;
  src_index(sh, gy, Hm, y0, y1, l[0], l[1]);
  src_index(sw, gx, Wm, x0, x1, l[2], l[3]);
  const TM* r0 = mbase + (long long)y0 * Wm;
  const TM* r1 = mbase + (long long)y1 * Wm;
  v[0] = to_f<TM>(__ldg(r0 + x0)); v[1] = to_f<TM>(__ldg(r0 + x1));
  v[2] = to_f<TM>(__ldg(r1 + x0)); v[3] = to_f<TM>(__ldg(r1 + x1));
}
// ... and combine them exactly like ATen's upsample_bilinear2d: ly0*(lx0*v00 + lx1*v01) + ly1*(lx0*v10 + lx1*v11)
__device__ __forceinline__ float tap_combine(const float (&v)[4], const float (&l)[4]) {
  return l[0] * (l[2] * v[0] + l[3] * v[1]) + l[1] * (l[2] * v[2] + l[3] * v[3]);
}

constexpr int kFill = 4;   // halo pixels whose taps are in flight per thread (memory-level parallelism)

// FAST4: fp32 mask at exactly 4x the logit resolution (the shipped 1024^2 -> 256^2 case): the 2x2 taps of
// pixel (gy,gx) are elements .y/.z of the aligned float4 at column 4*gx in rows 4*gy+1 and 4*gy+2.
template <typename TP, typename TM, bool FAST4>
__global__ void __launch_bounds__(kSegThreads) seg_loss_tile_kernel(const TP* __restrict__ pred, const TM* __restrict__ mask, float mscale, int H,
                                                            int W, int Hm, int Wm, long long mask_nstride, float focal_alpha, float focal_gamma,
                                                            float* __restrict__ t_save, float* __restrict__ w_save,
                                                            double* __restrict__ part) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SegSmem& sm = *reinterpret_cast<SegSmem*>(smem_raw);
  __shared__ double scratch[kNP * 32];
  const int tiles_x = (W + kT - 1) / kT, tiles_y = (H + kT - 1) / kT;
  const int n = blockIdx.x / (tiles_x * tiles_y);
  const int tile = blockIdx.x % (tiles_x * tiles_y);
  const int ty0 = (tile / tiles_x) * kT, tx0 = (tile % tiles_x) * kT;
  const bool same = (Hm == H && Wm == W);
  const float sh = (float)Hm / (float)H, sw = (float)Wm / (float)W;
  const TM* mbase = mask + (long long)n * mask_nstride;

  // 0. prefetch this thread's logits (consumed in step 3) so their latency hides behind steps 1-2
  const int px = threadIdx.x & 63, py0 = (threadIdx.x >> 6) * kRowsPT;
  float z[kRowsPT];
#pragma unroll
  for (int y = 0; y < kRowsPT; ++y) {
    const int gy = ty0 + py0 + y, gx = tx0 + px;
    z[y] = (gy < H && gx < W) ? to_f<TP>(__ldg(pred + ((long long)n * H + gy) * W + gx)) : 0.f;
  }

  // 1. target tile with halo (zero outside the image: avg_pool2d zero padding, count_include_pad);
  //    kFill pixels per thread are gathered before any is consumed
  for (int i0 = threadIdx.x; i0 < kTH * kTH; i0 += kFill * blockDim.x) {
    float v[kFill][4], l[kFill][4];
    bool ok[kFill];
#pragma unroll
    for (int u = 0; u < kFill; ++u) {
      const int i = i0 + u * blockDim.x;
      const int hy = i / kTH, hx = i % kTH;
      const int gy = ty0 + hy - kHalo, gx = tx0 + hx - kHalo;
      ok[u] = i < kTH * kTH && gy >= 0 && gy < H && gx >= 0 && gx < W;
      if (ok[u]) {
        if (FAST4) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(mbase + (long long)(4 * gy + 1) * Wm) + gx);
          const float4 b = __ldg(reinterpret_cast<const float4*>(mbase + (long long)(4 * gy + 2) * Wm) + gx);
          v[u][0] = a.y; v[u][1] = a.z; v[u][2] = b.y; v[u][3] = b.z;
          l[u][0] = l[u][1] = l[u][2] = l[u][3] = 0.5f;
        } else {
          tap_load<TM>(mbase, Hm, Wm, sh, sw, gy, gx, same, v[u], l[u]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kFill; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < kTH * kTH) sm.t[(i / kTH) * kTS + (i % kTH)] = ok[u] ? tap_combine(v[u], l[u]) * mscale : 0.f;
    }
  }
  __syncthreads();

  // 2. horizontal 31-sums: item = (segment of 16 columns, halo row)
  for (int it = threadIdx.x; it < 4 * kTH; it += blockDim.x) {
    const int hy = it % kTH, x0 = (it / kTH) * 16;
    const float* row = sm.t + hy * kTS;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < 31; ++d) s += row[x0 + d];
    sm.hs[hy * kHS + x0] = s;
#pragma unroll
    for (int x = 1; x < 16; ++x) {
      s += row[x0 + x + 30] - row[x0 + x - 1];
      sm.hs[hy * kHS + x0 + x] = s;
    }
  }
  __syncthreads();

  // 3. vertical 31-sums + per-pixel terms: thread = (column, segment of kRowsPT rows)
  double acc[kNP];
  {
    const int x = px, y0 = py0;
    const int gx = tx0 + x;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < 31; ++d) s += sm.hs[(y0 + d) * kHS + x];
    float f[kNP];
#pragma unroll
    for (int k = 0; k < kNP; ++k) f[k] = 0.f;
#pragma unroll
    for (int y = 0; y < kRowsPT; ++y) {
      if (y > 0) s += sm.hs[(y0 + y + 30) * kHS + x] - sm.hs[(y0 + y - 1) * kHS + x];
      const int gy = ty0 + y0 + y;
      if (gy < H && gx < W) {
        const float t = sm.t[(y0 + y + kHalo) * kTS + x + kHalo];
        float wgt;
        seg_pixel_terms(z[y], t, s, focal_alpha, focal_gamma, f, wgt);
        if (t_save) {
          const long long o = ((long long)n * H + gy) * W + gx;
          t_save[o] = t;
          w_save[o] = wgt;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kNP; ++k) acc[k] = (double)f[k];
  }
  block_sum<kNP>(acc, scratch);
  if (threadIdx.x == 0) {
    double* o = part + (long long)blockIdx.x * kNP;
#pragma unroll
    for (int k = 0; k < kNP; ++k) o[k] = acc[k];
  }
}

// The seven per-sample loss terms (seg_common.cuh) from the per-sample sums, in double.
__device__ __forceinline__ void seg_terms(const double* s, double HW, double sm, double (&T)[kNC]) {
  T[0] = s[1] / s[0];
  T[1] = 1.0 - (s[2] + 1e-6) / (s[3] - s[2] + 1e-6);
  T[2] = 1.0 - (2.0 * s[4] + sm) / (s[5] + s[6] + sm);
  T[3] = s[8] / HW;
  T[4] = 1.0 - (s[4] + 1e-6) / (s[5] + s[6] - s[4] + 1e-6);
  T[5] = 1.0 - (2.0 * s[2] + sm) / (s[3] + sm);
  T[6] = s[7] / HW;
}

// per_sample[n][kNP] = tile / strip sums in fixed order; out8 = {loss, dice, focal, wbce, wiou, bce, iou, wdice}, each the
// mean over samples, loss = mean_n sum_k coef[k] term_k[n]
__global__ void __launch_bounds__(256) seg_loss_finalize_kernel(const double* __restrict__ part, int N, int tiles, int HW, SegCoef coef,
                                                                float dice_smooth, float* __restrict__ per_sample,
                                                                float* __restrict__ out8) {
  __shared__ double scratch[8 * 32];
  __shared__ double ssum[16][kNP + 1];      // per-sample sums for a batch of 16 samples
  double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int n0 = 0; n0 < N; n0 += 16) {
    // thread (s, k) = (sample n0 + tid / 16, partial k = tid % 16) folds the tiles in fixed order
    const int sidx = threadIdx.x >> 4, k = threadIdx.x & 15, n = n0 + sidx;
    if (k < kNP) {
      double acc = 0.0;
      if (n < N)
        for (int t = 0; t < tiles; ++t) acc += part[((long long)n * tiles + t) * kNP + k];
      ssum[sidx][k] = acc;
      if (n < N) per_sample[(long long)n * kNP + k] = (float)acc;
    }
    __syncthreads();
    if (threadIdx.x < 16 && n0 + threadIdx.x < N) {
      double T[kNC];
      seg_terms(ssum[threadIdx.x], (double)HW, (double)dice_smooth, T);
      double l = 0.0;
#pragma unroll
      for (int c = 0; c < kNC; ++c) l += (double)coef.c[c] * T[c];
      v[0] += l;
      v[1] += T[2]; v[2] += T[6]; v[3] += T[0]; v[4] += T[1]; v[5] += T[3]; v[6] += T[4]; v[7] += T[5];
    }
    __syncthreads();
  }
  block_sum<8>(v, scratch);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) out8[k] = (float)(v[k] / N);
  }
}

// d loss / d pred, elementwise.  With p = sigmoid(z), dp = p(1-p) and the per-sample sums s[]:
//   wbce: w (p - t) / s0                       bce: (p - t) / HW
//   ratio terms 1 - A/B:  -(dA B - A dB)/B^2 dp   with (dA, dB) per pixel = wiou (t w, w - t w... see below)
//   focal: a_t [ (1-p_t)^g (p - t) - g (1-p_t)^(g-1) (2t-1) dp bce ] / HW,  p_t = p t + (1-p)(1-t)
template <typename TP, typename TG>
__global__ void __launch_bounds__(256) seg_loss_bwd_kernel(const TP* __restrict__ pred, const float* __restrict__ t_save,
                                                           const float* __restrict__ w_save, const float* __restrict__ per_sample,
                                                           int N, long long HW, SegCoef coef, float dice_smooth, float focal_alpha,
                                                           float focal_gamma, const float* __restrict__ g_loss, TG* __restrict__ g_pred) {
  const long long total = (long long)N * HW;
  const float g = g_loss[0] / (float)N;
  const float c0 = coef.c[0], c1 = coef.c[1], c2 = coef.c[2], c3 = coef.c[3], c4 = coef.c[4], c5 = coef.c[5], c6 = coef.c[6];
  const bool extras = c2 != 0.f || c3 != 0.f || c4 != 0.f || c5 != 0.f || c6 != 0.f;
  const float inv_hw = 1.f / (float)HW;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(o / HW);
    const float* s = per_sample + (long long)n * kNP;
    const float sw = s[0], I = s[2] + 1e-6f, U = s[3] - s[2] + 1e-6f;
    const float z = to_f<TP>(pred[o]), t = t_save[o], w = w_save[o];
    const float e = expf(-fabsf(z));
    const float p = z >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
    const float dp = p * (1.f - p);
    const float dwbce = w * (p - t) / sw;
    const float dratio = (t * w * U - I * w * (1.f - t)) / (U * U);   // d (I/U) / d p
    float d = c0 * dwbce - c1 * dratio * dp;
    if (extras) {
      const float sm = dice_smooth;
      if (c2 != 0.f) {           // dice: A = 2 s4 + sm, B = s5 + s6 + sm; dA = 2t, dB = 1
        const float A = 2.f * s[4] + sm, B = s[5] + s[6] + sm;
        d -= c2 * (2.f * t * B - A) / (B * B) * dp;
      }
      if (c3 != 0.f) d += c3 * (p - t) * inv_hw;
      if (c4 != 0.f) {           // iou: A = s4 + eps, B = s5 + s6 - s4 + eps; dA = t, dB = 1 - t
        const float A = s[4] + 1e-6f, B = s[5] + s[6] - s[4] + 1e-6f;
        d -= c4 * (t * B - A * (1.f - t)) / (B * B) * dp;
      }
      if (c5 != 0.f) {           // wdice: A = 2 s2 + sm, B = s3 + sm; dA = 2 t w, dB = w
        const float A = 2.f * s[2] + sm, B = s[3] + sm;
        d -= c5 * (2.f * t * w * B - A * w) / (B * B) * dp;
      }
      if (c6 != 0.f) {
        const float bce = (1.f - t) * z - (fminf(z, 0.f) - log1pf(e));
        const float pt = p * t + (1.f - p) * (1.f - t);
        const float at = focal_alpha * t + (1.f - focal_alpha) * (1.f - t);
        const float om = fmaxf(1.f - pt, 0.f);
        const float pw = powf(om, focal_gamma);
        const float dpw = om > 0.f ? focal_gamma * powf(om, focal_gamma - 1.f) : 0.f;
        d += c6 * at * (pw * (p - t) - dpw * (2.f * t - 1.f) * dp * bce) * inv_hw;
      }
    }
    g_pred[o] = from_f<TG>(g * d);
  }
}

template <typename TP, typename TM, bool FAST4>
static int launch_tiles_impl(const void* pred, const void* mask, float mscale, int N, int H, int W, int Hm, int Wm, long long ns, float fa, float fg_,
                             float* t_save, float* w_save, double* part, cudaStream_t st) {
  const int tiles = ceil_div(H, kT) * ceil_div(W, kT);
  auto k = seg_loss_tile_kernel<TP, TM, FAST4>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SegSmem));
  if (e != cudaSuccess) {
    set_error("seg_loss: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return COR_ECUDA;
  }
  k<<<N * tiles, kSegThreads, sizeof(SegSmem), st>>>((const TP*)pred, (const TM*)mask, mscale, H, W, Hm, Wm, ns, fa, fg_, t_save, w_save, part);
  return check_launch("seg_loss_tile_kernel");
}

template <typename TP, typename TM>
static int launch_tiles(const void* pred, const void* mask, float mscale, int N, int H, int W, int Hm, int Wm, long long ns, float fa, float fg_,
                        float* t_save, float* w_save, double* part, cudaStream_t st) {
  if (sizeof(TM) == 4 && Hm == 4 * H && Wm == 4 * W && (((uintptr_t)mask) & 15) == 0 && ns % 4 == 0)
    return launch_tiles_impl<TP, float, true>(pred, mask, mscale, N, H, W, Hm, Wm, ns, fa, fg_, t_save, w_save, part, st);
  return launch_tiles_impl<TP, TM, false>(pred, mask, mscale, N, H, W, Hm, Wm, ns, fa, fg_, t_save, w_save, part, st);
}

}  // namespace cor

using namespace cor;

extern "C" int cor_seg_loss_npartials(void) { return kNP; }

extern "C" size_t cor_seg_loss_work_bytes(int N, int H, int W) {
  const int tiles = ceil_div(H, kT) * ceil_div(W, kT), strips = seg_strip_max_strips(H);
  return (size_t)N * (tiles > strips ? tiles : strips) * kNP * sizeof(double);
}

static SegCoef make_coef(const float* coef7) {
  SegCoef c;
  for (int k = 0; k < kNC; ++k) c.c[k] = coef7 ? coef7[k] : (k < 2 ? 1.f : 0.f);    // default: wbce + wiou (loss_func.py:31)
  return c;
}

extern "C" int cor_seg_loss_fwd(const void* pred, int pred_dtype, const void* mask, int mask_dtype, float mask_scale, int N,
                                int H, int W, int Hm, int Wm, long long mask_nstride, const float* coef7, float focal_alpha,
                                float focal_gamma, float dice_smooth, float* out8, float* per_sample, float* t_save, float* w_save,
                                void* work, cor_stream_t stream) {
  COR_REQUIRE(pred && mask && out8 && per_sample && work, "cor_seg_loss_fwd: null pointer");
  COR_REQUIRE(N > 0 && H > 0 && W > 0 && Hm > 0 && Wm > 0, "cor_seg_loss_fwd: bad shape");
  COR_REQUIRE((t_save == nullptr) == (w_save == nullptr), "cor_seg_loss_fwd: t_save and w_save go together");
  cudaStream_t st = as_stream(stream);
  double* part = reinterpret_cast<double*>(work);
  int rc = COR_EINVAL;
  if (mask_nstride <= 0) mask_nstride = (long long)Hm * Wm;
  COR_REQUIRE(mask_nstride >= (long long)Hm * Wm, "cor_seg_loss_fwd: mask_nstride too small");
  const SegCoef coef = make_coef(coef7);
  int tiles = 0;
  const char* knob = getenv("COR_SEG_STRIP");                                  // A/B knob, read per call
  const bool no_strip = knob && atoi(knob) == 0;
  rc = no_strip ? COR_EINVAL
                : seg_strip_try_launch(pred, pred_dtype, mask, mask_dtype, mask_scale, N, H, W, Hm, Wm, mask_nstride, focal_alpha,
                                       focal_gamma, t_save, w_save, part, &tiles, st);
  if (rc == COR_ECUDA) return rc;
  if (rc != COR_OK) {
    tiles = ceil_div(H, kT) * ceil_div(W, kT);
#define COR_SEG(TP, TM) rc = launch_tiles<TP, TM>(pred, mask, mask_scale, N, H, W, Hm, Wm, mask_nstride, focal_alpha, focal_gamma, t_save, w_save, part, st)
    if (pred_dtype == COR_F32 && mask_dtype == COR_F32) COR_SEG(float, float);
    else if (pred_dtype == COR_BF16 && mask_dtype == COR_F32) COR_SEG(bf16, float);
    else if (pred_dtype == COR_F32 && mask_dtype == COR_U8) COR_SEG(float, uint8_t);
    else if (pred_dtype == COR_BF16 && mask_dtype == COR_U8) COR_SEG(bf16, uint8_t);
    else if (pred_dtype == COR_F32 && mask_dtype == COR_BF16) COR_SEG(float, bf16);
    else if (pred_dtype == COR_BF16 && mask_dtype == COR_BF16) COR_SEG(bf16, bf16);
    else COR_REQUIRE(false, "cor_seg_loss_fwd: unsupported dtypes pred=%d mask=%d", pred_dtype, mask_dtype);
#undef COR_SEG
    if (rc) return rc;
  }
  seg_loss_finalize_kernel<<<1, 256, 0, st>>>(part, N, tiles, H * W, coef, dice_smooth, per_sample, out8);
  return check_launch("seg_loss_finalize_kernel");
}

extern "C" int cor_seg_loss_bwd(const void* pred, int pred_dtype, const float* t_save, const float* w_save,
                                const float* per_sample, int N, int H, int W, const float* coef7, float dice_smooth, float focal_alpha,
                                float focal_gamma, const float* g_loss, void* g_pred, int g_dtype, cor_stream_t stream) {
  COR_REQUIRE(pred && t_save && w_save && per_sample && g_loss && g_pred, "cor_seg_loss_bwd: null pointer");
  const long long HW = (long long)H * W;
  const int blocks = (int)min((long long)sm_count() * 8, (N * HW + 255) / 256);
  cudaStream_t st = as_stream(stream);
  const SegCoef coef = make_coef(coef7);
#define COR_SEGB(TP, TG) seg_loss_bwd_kernel<TP, TG><<<blocks, 256, 0, st>>>((const TP*)pred, t_save, w_save, per_sample, N, HW, coef, dice_smooth, focal_alpha, focal_gamma, g_loss, (TG*)g_pred)
  if (pred_dtype == COR_F32 && g_dtype == COR_F32) COR_SEGB(float, float);
  else if (pred_dtype == COR_BF16 && g_dtype == COR_BF16) COR_SEGB(bf16, bf16);
  else if (pred_dtype == COR_BF16 && g_dtype == COR_F32) COR_SEGB(bf16, float);
  else COR_REQUIRE(false, "cor_seg_loss_bwd: unsupported dtypes pred=%d grad=%d", pred_dtype, g_dtype);
#undef COR_SEGB
  return check_launch("seg_loss_bwd_kernel");
}
