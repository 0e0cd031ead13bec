// Row-wise LayerNorm (+ optional GELU) over channels-last rows, forward and backward: the normalisations of the
// MaskAdapter map generator -- ChannelReduction's LayerNorm(channels_first) + GELU (lib/support_model/mask_adapter.py:83-94;
// over channels per pixel it is the same arithmetic as a channels-last LayerNorm on [pixels][C] rows, biased variance,
// :240-251), the ConvNeXt blocks' LayerNorm (:200-209) and the final norm (:171-173).
//   y = act( (x - mean) * rstd * w + b ),  rstd = 1 / sqrt(var + eps)
// One warp per row (C <= 1024: up to 32 values per lane in registers, one pass over global memory), fp32 in, fp32 or
// bf16 out (bf16 = the next GEMM's A operand, no separate cast).  Backward recomputes xhat from the saved (mean, rstd);
// the affine gradients are per-CTA partial column sums folded in fixed order (deterministic).
#include "common.cuh"

namespace cor {

constexpr int kLnMaxPerLane = 32;      // C <= 1024
constexpr int kLnWarps = 8;

__device__ __forceinline__ float gelu_f(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float x) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
  return cdf + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

template <typename TO, int PL>      // PL = values per lane: C <= 32 * PL
__global__ void __launch_bounds__(kLnWarps * 32) ln_rows_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                    const float* __restrict__ b, long long rows, int C, float eps, int act,
                                                                    TO* __restrict__ y, float* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * C;
  float v[PL];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PL; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < C ? xr[c] : 0.f;
    s += v[i];
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PL; ++i) {
    const int c = lane + 32 * i;
    const float d = c < C ? v[i] - mean : 0.f;
    q = fmaf(d, d, q);
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
  for (int i = 0; i < PL; ++i) {
    const int c = lane + 32 * i;
    if (c < C) {
      float o = (v[i] - mean) * rstd * w[c] + b[c];
      if (act == COR_ACT_GELU) o = gelu_f(o);
      y[row * C + c] = from_f<TO>(o);
    }
  }
  if (lane == 0 && stats) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
}

// dx = rstd * (g - mean_c(g) - xhat * mean_c(g * xhat)),  g = dy * act'(.) * w;  dw_part / db_part [gridDim.x][C]
template <typename TD, int PL>
__global__ void __launch_bounds__(kLnWarps * 32) ln_rows_bwd_kernel(const TD* __restrict__ dy, const float* __restrict__ x,
                                                                    const float* __restrict__ w, const float* __restrict__ b,
                                                                    const float* __restrict__ stats, long long rows, int C, int act,
                                                                    long long rows_per_cta, float* __restrict__ dx,
                                                                    float* __restrict__ dw_part, float* __restrict__ db_part) {
  __shared__ float red[kLnWarps][2];
  extern __shared__ float acc_s[];             // [2][kLnWarps][C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float aw[PL], ab[PL];
#pragma unroll
  for (int i = 0; i < PL; ++i) aw[i] = ab[i] = 0.f;
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  for (long long row = r0 + warp; row < r1; row += kLnWarps) {
    const float mean = stats[2 * row], rstd = stats[2 * row + 1];
    float g[PL], xh[PL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < PL; ++i) {
      const int c = lane + 32 * i;
      g[i] = xh[i] = 0.f;
      if (c < C) {
        xh[i] = (x[row * C + c] - mean) * rstd;
        float d = to_f<TD>(dy[row * C + c]);
        if (act == COR_ACT_GELU) d *= gelu_grad(xh[i] * w[c] + b[c]);
        aw[i] = fmaf(d, xh[i], aw[i]);
        ab[i] += d;
        g[i] = d * w[c];
        s1 += g[i];
        s2 = fmaf(g[i], xh[i], s2);
      }
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
#pragma unroll
    for (int i = 0; i < PL; ++i) {
      const int c = lane + 32 * i;
      if (c < C) dx[row * C + c] = rstd * (g[i] - s1 - xh[i] * s2);
    }
  }
  (void)red;
  // fold the 8 warps' column sums in fixed order, publish this CTA's partial
#pragma unroll
  for (int i = 0; i < PL; ++i) {
    const int c = lane + 32 * i;
    if (c < C) { acc_s[warp * C + c] = aw[i]; acc_s[(kLnWarps + warp) * C + c] = ab[i]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float sw = 0.f, sb = 0.f;
    for (int k = 0; k < kLnWarps; ++k) { sw += acc_s[k * C + c]; sb += acc_s[(kLnWarps + k) * C + c]; }
    dw_part[(long long)blockIdx.x * C + c] = sw;
    db_part[(long long)blockIdx.x * C + c] = sb;
  }
}

// grid = (ceil(C / 32), 2): y = 0 folds d weight, y = 1 d bias
__global__ void ln_fold_kernel(const float* __restrict__ dw_part, const float* __restrict__ db_part, int nparts, int C,
                               float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float sm[kFoldTy][32];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const float t = fold_parts(blockIdx.y ? db_part : dw_part, nparts, C, c, c < C, sm);
  if (threadIdx.y == 0 && c < C) (blockIdx.y ? db : dw)[c] = t;
}

static int ln_parts(long long rows) {
  long long p = (rows + kLnWarps * 4 - 1) / (kLnWarps * 4);        // >= 4 rows per warp
  const long long cap = (long long)sm_count() * 4;
  if (p > cap) p = cap;
  return (int)(p < 1 ? 1 : p);
}

}  // namespace cor

using namespace cor;

extern "C" size_t cor_ln_rows_work_bytes(long long rows, int C) { return (size_t)ln_parts(rows) * C * 2 * sizeof(float) + 16; }

extern "C" int cor_ln_rows_fwd(const float* x, const float* weight, const float* bias, long long rows, int C, float eps, int act,
                               void* y, int y_dtype, float* stats, cor_stream_t stream) {
  COR_REQUIRE(x && weight && bias && y, "cor_ln_rows_fwd: null pointer");
  COR_REQUIRE(rows > 0 && C > 0 && C <= 32 * kLnMaxPerLane, "cor_ln_rows_fwd: need 0 < C <= %d (C=%d)", 32 * kLnMaxPerLane, C);
  COR_REQUIRE(act == COR_ACT_NONE || act == COR_ACT_GELU, "cor_ln_rows_fwd: act %d", act);
  const unsigned blocks = (unsigned)((rows + kLnWarps - 1) / kLnWarps);
  cudaStream_t st = as_stream(stream);
  COR_REQUIRE(y_dtype == COR_F32 || y_dtype == COR_BF16, "cor_ln_rows_fwd: output dtype %d", y_dtype);
#define COR_LN_F(TO, PL) ln_rows_fwd_kernel<TO, PL><<<blocks, kLnWarps * 32, 0, st>>>(x, weight, bias, rows, C, eps, act, (TO*)y, stats)
  if (y_dtype == COR_F32) { if (C <= 256) COR_LN_F(float, 8); else if (C <= 512) COR_LN_F(float, 16); else COR_LN_F(float, 32); }
  else { if (C <= 256) COR_LN_F(bf16, 8); else if (C <= 512) COR_LN_F(bf16, 16); else COR_LN_F(bf16, 32); }
#undef COR_LN_F
  return check_launch("ln_rows_fwd_kernel");
}

extern "C" int cor_ln_rows_bwd(const void* dy, int dy_dtype, const float* x, const float* weight, const float* bias, const float* stats,
                               long long rows, int C, int act, float* dx, float* dweight, float* dbias, void* work, cor_stream_t stream) {
  COR_REQUIRE(dy && x && weight && bias && stats && dx && dweight && dbias && work, "cor_ln_rows_bwd: null pointer");
  COR_REQUIRE(rows > 0 && C > 0 && C <= 32 * kLnMaxPerLane, "cor_ln_rows_bwd: need 0 < C <= %d (C=%d)", 32 * kLnMaxPerLane, C);
  const int parts = ln_parts(rows);
  const long long per = (rows + parts - 1) / parts;
  float* dwp = reinterpret_cast<float*>(work);
  float* dbp = dwp + (size_t)parts * C;
  cudaStream_t st = as_stream(stream);
  const size_t smem = (size_t)2 * kLnWarps * C * sizeof(float);
  COR_REQUIRE(dy_dtype == COR_F32 || dy_dtype == COR_BF16, "cor_ln_rows_bwd: gradient dtype %d", dy_dtype);
#define COR_LN_B(TD, PL)                                                                                                          \
  do {                                                                                                                            \
    COR_CUDA(cudaFuncSetAttribute(ln_rows_bwd_kernel<TD, PL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
    ln_rows_bwd_kernel<TD, PL><<<parts, kLnWarps * 32, smem, st>>>(reinterpret_cast<const TD*>(dy), x, weight, bias, stats, rows, \
                                                                   C, act, per, dx, dwp, dbp);                                    \
  } while (0)
  if (dy_dtype == COR_F32) { if (C <= 256) COR_LN_B(float, 8); else if (C <= 512) COR_LN_B(float, 16); else COR_LN_B(float, 32); }
  else { if (C <= 256) COR_LN_B(bf16, 8); else if (C <= 512) COR_LN_B(bf16, 16); else COR_LN_B(bf16, 32); }
#undef COR_LN_B
  int rc = check_launch("ln_rows_bwd_kernel");
  if (rc) return rc;
  ln_fold_kernel<<<dim3((C + 31) / 32, 2), dim3(32, kFoldTy), 0, st>>>(dwp, dbp, parts, C, dweight, dbias);
  return check_launch("ln_fold_kernel");
}
