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

// VEC adjacent columns per lane and load (4 when C % 4 == 0: 128-bit fp32 / 64-bit bf16 accesses; else 1)
template <typename T, int VEC>
__device__ __forceinline__ void ldv(const T* p, float* v);
template <>
__device__ __forceinline__ void ldv<float, 1>(const float* p, float* v) { v[0] = *p; }
template <>
__device__ __forceinline__ void ldv<bf16, 1>(const bf16* p, float* v) { v[0] = __bfloat162float(*p); }
template <>
__device__ __forceinline__ void ldv<float, 4>(const float* p, float* v) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void ldv<bf16, 4>(const bf16* p, float* v) {
  const uint2 t = *reinterpret_cast<const uint2*>(p);
  v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
  v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
}
template <typename T, int VEC>
__device__ __forceinline__ void stv(T* p, const float* v);
template <>
__device__ __forceinline__ void stv<float, 1>(float* p, const float* v) { *p = v[0]; }
template <>
__device__ __forceinline__ void stv<bf16, 1>(bf16* p, const float* v) { *p = __float2bfloat16_rn(v[0]); }
template <>
__device__ __forceinline__ void stv<float, 4>(float* p, const float* v) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
template <>
__device__ __forceinline__ void stv<bf16, 4>(bf16* p, const float* v) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
}

template <typename TO, int PL, int VEC>      // PL = values per lane: C <= 32 * PL; column of value (i, k): (lane + 32 i) VEC + k
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
  for (int i = 0; i < PL / VEC; ++i) {
    const int c = (lane + 32 * i) * VEC;
    if (c < C) ldv<float, VEC>(xr + c, v + i * VEC);
    else
#pragma unroll
      for (int k = 0; k < VEC; ++k) v[i * VEC + k] = 0.f;
#pragma unroll
    for (int k = 0; k < VEC; ++k) s += v[i * VEC + k];
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PL / VEC; ++i) {
    const int c = (lane + 32 * i) * VEC;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float d = c < C ? v[i * VEC + k] - mean : 0.f;
      q = fmaf(d, d, q);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
  for (int i = 0; i < PL / VEC; ++i) {
    const int c = (lane + 32 * i) * VEC;
    if (c < C) {
      float wv[VEC], bv[VEC], o[VEC];
      ldv<float, VEC>(w + c, wv);
      ldv<float, VEC>(b + c, bv);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        o[k] = (v[i * VEC + k] - mean) * rstd * wv[k] + bv[k];
        if (act == COR_ACT_GELU) o[k] = gelu_f(o[k]);
      }
      stv<TO, VEC>(y + row * C + c, o);
    }
  }
  if (lane == 0 && stats) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
}

// dx = rstd * (g - mean_c(g) - xhat * mean_c(g * xhat)),  g = dy * act'(.) * w;  dw_part / db_part [gridDim.x][C]
template <typename TD, int PL, int VEC>
__global__ void __launch_bounds__(kLnWarps * 32) ln_rows_bwd_kernel(const TD* __restrict__ dy, const float* __restrict__ x,
                                                                    const float* __restrict__ w, const float* __restrict__ b,
                                                                    const float* __restrict__ stats, long long rows, int C, int act,
                                                                    long long rows_per_cta, float* __restrict__ dx,
                                                                    float* __restrict__ dw_part, float* __restrict__ db_part) {
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
    for (int i = 0; i < PL / VEC; ++i) {
      const int c = (lane + 32 * i) * VEC;
#pragma unroll
      for (int k = 0; k < VEC; ++k) g[i * VEC + k] = xh[i * VEC + k] = 0.f;
      if (c < C) {
        float xv[VEC], dv[VEC], wv[VEC], bv[VEC];
        ldv<float, VEC>(x + row * C + c, xv);
        ldv<TD, VEC>(dy + row * C + c, dv);
        ldv<float, VEC>(w + c, wv);
        if (act == COR_ACT_GELU) ldv<float, VEC>(b + c, bv);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const int j = i * VEC + k;
          xh[j] = (xv[k] - mean) * rstd;
          float d = dv[k];
          if (act == COR_ACT_GELU) d *= gelu_grad(xh[j] * wv[k] + bv[k]);
          aw[j] = fmaf(d, xh[j], aw[j]);
          ab[j] += d;
          g[j] = d * wv[k];
          s1 += g[j];
          s2 = fmaf(g[j], xh[j], s2);
        }
      }
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
#pragma unroll
    for (int i = 0; i < PL / VEC; ++i) {
      const int c = (lane + 32 * i) * VEC;
      if (c < C) {
        float o[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) o[k] = rstd * (g[i * VEC + k] - s1 - xh[i * VEC + k] * s2);
        stv<float, VEC>(dx + row * C + c, o);
      }
    }
  }
  // fold the 8 warps' column sums in fixed order, publish this CTA's partial
#pragma unroll
  for (int i = 0; i < PL / VEC; ++i) {
    const int c = (lane + 32 * i) * VEC;
#pragma unroll
    for (int k = 0; k < VEC; ++k)
      if (c + k < C) { acc_s[warp * C + c + k] = aw[i * VEC + k]; acc_s[(kLnWarps + warp) * C + c + k] = ab[i * VEC + k]; }
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
  const bool vec = C % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(weight) |
                                   reinterpret_cast<uintptr_t>(bias)) & 15) == 0;
#define COR_LN_F(TO, PL, V) ln_rows_fwd_kernel<TO, PL, V><<<blocks, kLnWarps * 32, 0, st>>>(x, weight, bias, rows, C, eps, act, (TO*)y, stats)
#define COR_LN_FP(TO, V) do { if (C <= 256) COR_LN_F(TO, 8, V); else if (C <= 512) COR_LN_F(TO, 16, V); else COR_LN_F(TO, 32, V); } while (0)
  if (y_dtype == COR_F32) { if (vec) COR_LN_FP(float, 4); else COR_LN_FP(float, 1); }
  else { if (vec) COR_LN_FP(bf16, 4); else COR_LN_FP(bf16, 1); }
#undef COR_LN_FP
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
  const bool vec = C % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(weight) |
                                   reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0;
#define COR_LN_B(TD, PL, V)                                                                                                          \
  do {                                                                                                                               \
    COR_CUDA(cudaFuncSetAttribute(ln_rows_bwd_kernel<TD, PL, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
    ln_rows_bwd_kernel<TD, PL, V><<<parts, kLnWarps * 32, smem, st>>>(reinterpret_cast<const TD*>(dy), x, weight, bias, stats, rows, \
                                                                      C, act, per, dx, dwp, dbp);                                    \
  } while (0)
#define COR_LN_BP(TD, V) do { if (C <= 256) COR_LN_B(TD, 8, V); else if (C <= 512) COR_LN_B(TD, 16, V); else COR_LN_B(TD, 32, V); } while (0)
  if (dy_dtype == COR_F32) { if (vec) COR_LN_BP(float, 4); else COR_LN_BP(float, 1); }
  else { if (vec) COR_LN_BP(bf16, 4); else COR_LN_BP(bf16, 1); }
#undef COR_LN_BP
#undef COR_LN_B
  int rc = check_launch("ln_rows_bwd_kernel");
  if (rc) return rc;
  ln_fold_kernel<<<dim3((C + 31) / 32, 2), dim3(32, kFoldTy), 0, st>>>(dwp, dbp, parts, C, dweight, dbias);
  return check_launch("ln_fold_kernel");
}

// ---- channels-first LayerNorm (+ GELU) over a FEW channels: the two normalisations of mask_downscaling ----------------
// (lib/support_model/mask_adapter.py:128-142: Conv2d(1,4,3,2) -> LayerNorm2d -> GELU -> Conv2d(4,16,3,2) -> LayerNorm2d ->
// GELU; LayerNorm(channels_first) :240-251 is mean / pow / sqrt over dim 1, i.e. ~8 element-wise launches forward and ~20
// backward per norm in eager PyTorch.)  x [N][C][P] fp32, C <= 32: one thread per pixel walks the C channel planes
// (coalesced across the warp), everything in registers, one launch each way; the affine gradients are per-CTA partial sums
// folded in fixed order.
namespace cor {

template <int CMAX>
__global__ void __launch_bounds__(256) ln_cf_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                       long long N, int C, long long P, float eps, int act, float* __restrict__ y) {
  const long long total = N * P;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / P, p = i - n * P;
    const float* xp = x + n * C * P + p;
    float v[CMAX];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      v[c] = c < C ? xp[(long long)c * P] : 0.f;
      s += v[c];
    }
    const float mean = s / (float)C;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      const float d = c < C ? v[c] - mean : 0.f;
      q = fmaf(d, d, q);
    }
    const float rstd = rsqrtf(q / (float)C + eps);
    float* yp = y + n * C * P + p;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      if (c < C) {
        float o = (v[c] - mean) * rstd * w[c] + b[c];
        if (act == COR_ACT_GELU) o = gelu_f(o);
        yp[(long long)c * P] = o;
      }
    }
  }
}

template <int CMAX>
__global__ void __launch_bounds__(256) ln_cf_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ b, long long N, int C, long long P, float eps, int act,
                                                       float* __restrict__ dx, float* __restrict__ dw_part, float* __restrict__ db_part) {
  __shared__ float red[8][2 * CMAX];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float aw[CMAX], ab[CMAX];
#pragma unroll
  for (int c = 0; c < CMAX; ++c) aw[c] = ab[c] = 0.f;
  const long long total = N * P;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / P, p = i - n * P;
    const float* xp = x + n * C * P + p;
    const float* gp = dy + n * C * P + p;
    float v[CMAX], g[CMAX];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      v[c] = c < C ? xp[(long long)c * P] : 0.f;
      g[c] = c < C ? gp[(long long)c * P] : 0.f;
      s += v[c];
    }
    const float mean = s / (float)C;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      const float d = c < C ? v[c] - mean : 0.f;
      q = fmaf(d, d, q);
    }
    const float rstd = rsqrtf(q / (float)C + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      if (c < C) {
        const float xh = (v[c] - mean) * rstd;
        float d = g[c];
        if (act == COR_ACT_GELU) d *= gelu_grad(xh * w[c] + b[c]);
        aw[c] = fmaf(d, xh, aw[c]);
        ab[c] += d;
        v[c] = xh;
        g[c] = d * w[c];
        s1 += g[c];
        s2 = fmaf(g[c], xh, s2);
      }
    }
    s1 /= (float)C;
    s2 /= (float)C;
    float* op = dx + n * C * P + p;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) op[(long long)c * P] = rstd * (g[c] - s1 - v[c] * s2);
  }
#pragma unroll
  for (int c = 0; c < CMAX; ++c) {
    const float a = warp_sum(aw[c]), bb = warp_sum(ab[c]);
    if (lane == 0) { red[warp][c] = a; red[warp][CMAX + c] = bb; }
  }
  __syncthreads();
  if (threadIdx.x < 2 * CMAX) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    const int c = threadIdx.x < CMAX ? threadIdx.x : threadIdx.x - CMAX;
    if (c < C) (threadIdx.x < CMAX ? dw_part : db_part)[(long long)blockIdx.x * C + c] = t;
  }
}

static int ln_cf_blocks(long long N, long long P) {
  long long bl = (N * P + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (bl > cap) bl = cap;
  return (int)(bl < 1 ? 1 : bl);
}

}  // namespace cor

extern "C" size_t cor_ln_cf_work_bytes(long long N, int C, long long P) { return (size_t)ln_cf_blocks(N, P) * C * 2 * sizeof(float) + 16; }

extern "C" int cor_ln_cf_fwd(const float* x, const float* weight, const float* bias, long long N, int C, long long P, float eps, int act,
                             float* y, cor_stream_t stream) {
  COR_REQUIRE(x && weight && bias && y, "cor_ln_cf_fwd: null pointer");
  COR_REQUIRE(N > 0 && P > 0 && C > 0 && C <= 32, "cor_ln_cf_fwd: need 0 < C <= 32 (C=%d)", C);
  COR_REQUIRE(act == COR_ACT_NONE || act == COR_ACT_GELU, "cor_ln_cf_fwd: act %d", act);
  const int blocks = ln_cf_blocks(N, P);
  cudaStream_t st = as_stream(stream);
  if (C <= 4) ln_cf_fwd_kernel<4><<<blocks, 256, 0, st>>>(x, weight, bias, N, C, P, eps, act, y);
  else if (C <= 16) ln_cf_fwd_kernel<16><<<blocks, 256, 0, st>>>(x, weight, bias, N, C, P, eps, act, y);
  else ln_cf_fwd_kernel<32><<<blocks, 256, 0, st>>>(x, weight, bias, N, C, P, eps, act, y);
  return check_launch("ln_cf_fwd_kernel");
}

extern "C" int cor_ln_cf_bwd(const float* dy, const float* x, const float* weight, const float* bias, long long N, int C, long long P, float eps,
                             int act, float* dx, float* dweight, float* dbias, void* work, cor_stream_t stream) {
  COR_REQUIRE(dy && x && weight && bias && dx && dweight && dbias && work, "cor_ln_cf_bwd: null pointer");
  COR_REQUIRE(N > 0 && P > 0 && C > 0 && C <= 32, "cor_ln_cf_bwd: need 0 < C <= 32 (C=%d)", C);
  COR_REQUIRE(act == COR_ACT_NONE || act == COR_ACT_GELU, "cor_ln_cf_bwd: act %d", act);
  const int blocks = ln_cf_blocks(N, P);
  float* dwp = reinterpret_cast<float*>(work);
  float* dbp = dwp + (size_t)blocks * C;
  cudaStream_t st = as_stream(stream);
  if (C <= 4) ln_cf_bwd_kernel<4><<<blocks, 256, 0, st>>>(dy, x, weight, bias, N, C, P, eps, act, dx, dwp, dbp);
  else if (C <= 16) ln_cf_bwd_kernel<16><<<blocks, 256, 0, st>>>(dy, x, weight, bias, N, C, P, eps, act, dx, dwp, dbp);
  else ln_cf_bwd_kernel<32><<<blocks, 256, 0, st>>>(dy, x, weight, bias, N, C, P, eps, act, dx, dwp, dbp);
  int rc = check_launch("ln_cf_bwd_kernel");
  if (rc) return rc;
  ln_fold_kernel<<<dim3((C + 31) / 32, 2), dim3(32, kFoldTy), 0, st>>>(dwp, dbp, blocks, C, dweight, dbias);
  return check_launch("ln_fold_kernel");
}
