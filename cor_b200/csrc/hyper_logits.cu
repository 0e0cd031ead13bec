// Mask-logit producer of the SAM decoder: the hypernetwork product
//   masks[b, t, p] = sum_c hyper_in[b, t, c] * upscaled[b, c, p]          lib/sam_model/mask_decoder.py:135-137
// ([B,T,32] @ [B,32,65536]) whose output the segmentation loss reads straight back.  It is a 32-deep contraction over a
// 65 536-pixel map: HBM-bound on one pass over `upscaled`, so it runs on CUDA cores with 128-bit streaming loads -- only
// the token rows that are consumed are produced (the reference computes all four mask tokens and slices one,
// mask_decoder.py:97-102), in the dtype the loss kernel wants.
//
// Backward, ONE pass over `upscaled` and the logit gradient:
//   d upscaled[b, c, p] = sum_t hyper_in[b, t, c] * g[b, t, p]
//   d hyper_in[b, t, c] = sum_p g[b, t, p] * upscaled[b, c, p]            (per-CTA partials, fixed-order fold)
#include "common.cuh"

namespace cor {

constexpr int kHyMaxT = 4;       // mask tokens (SAM: num_mask_tokens = 4, mask_decoder.py:49)
constexpr int kHyMaxC = 64;      // channels of the upscaled embedding (SAM: transformer_dim / 8 = 32)
constexpr int kHyThreads = 256;
constexpr int kHyVec = 4;        // pixels per thread

template <typename T>
__device__ __forceinline__ void ld4(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void ld4<float>(const float* p, float (&v)[4]) {
  const uint4 u = ld_stream16(p);
  v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y); v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
}
template <>
__device__ __forceinline__ void ld4<bf16>(const bf16* p, float (&v)[4]) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  v[0] = bf16lo(u.x); v[1] = bf16hi(u.x); v[2] = bf16lo(u.y); v[3] = bf16hi(u.y);
}
template <typename T>
__device__ __forceinline__ void st4(T* p, const float (&v)[4]);
template <>
__device__ __forceinline__ void st4<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <>
__device__ __forceinline__ void st4<bf16>(bf16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}

// grid = (pixel blocks, B); hyper [B, T_all, C] f32 (tokens t0 .. t0+T-1 are used), up [B, C, P], out [B, T, P]
template <typename TU, typename TO>
__global__ void __launch_bounds__(kHyThreads) hyper_logits_fwd_kernel(const float* __restrict__ hyper, const TU* __restrict__ up,
                                                                      TO* __restrict__ out, int T_all, int t0, int T, int Cc, long long P) {
  __shared__ float h[kHyMaxT][kHyMaxC];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < T * Cc; i += blockDim.x) h[i / Cc][i % Cc] = hyper[((long long)b * T_all + t0 + i / Cc) * Cc + i % Cc];
  __syncthreads();
  const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * kHyVec;
  if (p >= P) return;
  const TU* u = up + (long long)b * Cc * P + p;
  float acc[kHyMaxT][4];
#pragma unroll
  for (int t = 0; t < kHyMaxT; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[t][j] = 0.f;
#pragma unroll 8
  for (int c = 0; c < Cc; ++c) {
    float v[4];
    ld4<TU>(u + (long long)c * P, v);
#pragma unroll
    for (int t = 0; t < kHyMaxT; ++t)
      if (t < T) {
        const float w = h[t][c];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[t][j] = fmaf(w, v[j], acc[t][j]);
      }
  }
#pragma unroll
  for (int t = 0; t < kHyMaxT; ++t)
    if (t < T) st4<TO>(out + ((long long)b * T + t) * P + p, acc[t]);
}

// grid = (pixel blocks, B): g [B, T, P] (TG), up [B, C, P] -> d_up [B, C, P] (TU), partials [B][gridDim.x][T][C] f32
template <typename TU, typename TG>
__global__ void __launch_bounds__(kHyThreads) hyper_logits_bwd_kernel(const float* __restrict__ hyper, const TU* __restrict__ up,
                                                                      const TG* __restrict__ g, TU* __restrict__ d_up,
                                                                      float* __restrict__ part, int T_all, int t0, int T, int Cc, long long P) {
  __shared__ float h[kHyMaxT][kHyMaxC];
  __shared__ float red[kHyThreads / 32][kHyMaxT];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < T * Cc; i += blockDim.x) h[i / Cc][i % Cc] = hyper[((long long)b * T_all + t0 + i / Cc) * Cc + i % Cc];
  __syncthreads();
  const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * kHyVec;
  const bool ok = p < P;
  float gv[kHyMaxT][4];
#pragma unroll
  for (int t = 0; t < kHyMaxT; ++t) {
#pragma unroll
    for (int j = 0; j < 4; ++j) gv[t][j] = 0.f;
    if (t < T && ok) ld4<TG>(g + ((long long)b * T + t) * P + p, gv[t]);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* my_part = part + (((long long)b * gridDim.x + blockIdx.x) * T) * Cc;
  for (int c = 0; c < Cc; ++c) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (ok) ld4<TU>(up + ((long long)b * Cc + c) * P + p, v);
    float du[4] = {0.f, 0.f, 0.f, 0.f};
    float dh[kHyMaxT];
#pragma unroll
    for (int t = 0; t < kHyMaxT; ++t) {
      dh[t] = 0.f;
      if (t < T) {
        const float w = h[t][c];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          du[j] = fmaf(w, gv[t][j], du[j]);
          dh[t] = fmaf(gv[t][j], v[j], dh[t]);
        }
      }
    }
    if (d_up && ok) st4<TU>(d_up + ((long long)b * Cc + c) * P + p, du);
    // block-wide sum of dh[t] for this channel (fixed tree: deterministic)
#pragma unroll
    for (int t = 0; t < kHyMaxT; ++t) dh[t] = warp_sum(dh[t]);
    if (lane == 0) {
#pragma unroll
      for (int t = 0; t < kHyMaxT; ++t) red[warp][t] = dh[t];
    }
    __syncthreads();
    if (threadIdx.x < T) {
      float s = 0.f;
      for (int w = 0; w < kHyThreads / 32; ++w) s += red[w][threadIdx.x];
      my_part[threadIdx.x * Cc + c] = s;
    }
    __syncthreads();
  }
}

// d_hyper[b, t0 + t, c] = sum over pixel blocks (ascending) of the partials; other tokens get zero
__global__ void hyper_fold_kernel(const float* __restrict__ part, float* __restrict__ d_hyper, int B, int nblk, int T_all, int t0, int T, int Cc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * T_all * Cc) return;
  const int c = i % Cc, t = (i / Cc) % T_all, b = i / (Cc * T_all);
  float s = 0.f;
  if (t >= t0 && t < t0 + T)
    for (int k = 0; k < nblk; ++k) s += part[(((long long)b * nblk + k) * T + (t - t0)) * Cc + c];
  d_hyper[i] = s;
}

static int hy_blocks(long long P) { return (int)((P + (long long)kHyThreads * kHyVec - 1) / ((long long)kHyThreads * kHyVec)); }

}  // namespace cor

using namespace cor;

extern "C" size_t cor_hyper_logits_work_bytes(int B, int T, int C, long long P) {
  return (size_t)B * hy_blocks(P) * T * C * sizeof(float) + 16;
}

extern "C" int cor_hyper_logits_fwd(const float* hyper, const void* up, int up_dtype, void* out, int out_dtype, int B, int T_all, int t0,
                                    int T, int C, long long P, cor_stream_t stream) {
  COR_REQUIRE(hyper && up && out, "cor_hyper_logits_fwd: null pointer");
  COR_REQUIRE(B > 0 && T >= 1 && T <= kHyMaxT && t0 >= 0 && t0 + T <= T_all && C >= 1 && C <= kHyMaxC && P > 0 && P % kHyVec == 0,
              "cor_hyper_logits_fwd: need 1 <= T <= %d tokens, C <= %d, P %% 4 == 0 (T=%d C=%d P=%lld)", kHyMaxT, kHyMaxC, T, C, P);
  COR_REQUIRE(((uintptr_t)up & 15) == 0 && ((uintptr_t)out & 15) == 0, "cor_hyper_logits_fwd: 16-byte alignment");
  const dim3 grid(hy_blocks(P), B);
  cudaStream_t st = as_stream(stream);
#define COR_HY(TU, TO) hyper_logits_fwd_kernel<TU, TO><<<grid, kHyThreads, 0, st>>>(hyper, (const TU*)up, (TO*)out, T_all, t0, T, C, P)
  if (up_dtype == COR_F32 && out_dtype == COR_F32) COR_HY(float, float);
  else if (up_dtype == COR_F32 && out_dtype == COR_BF16) COR_HY(float, bf16);
  else if (up_dtype == COR_BF16 && out_dtype == COR_BF16) COR_HY(bf16, bf16);
  else if (up_dtype == COR_BF16 && out_dtype == COR_F32) COR_HY(bf16, float);
  else COR_REQUIRE(false, "cor_hyper_logits_fwd: unsupported dtypes up=%d out=%d", up_dtype, out_dtype);
#undef COR_HY
  return check_launch("hyper_logits_fwd_kernel");
}

extern "C" int cor_hyper_logits_bwd(const float* hyper, const void* up, int up_dtype, const void* g, int g_dtype, void* d_up,
                                    float* d_hyper, int B, int T_all, int t0, int T, int C, long long P, void* work, cor_stream_t stream) {
  COR_REQUIRE(hyper && up && g && d_hyper && work, "cor_hyper_logits_bwd: null pointer");
  COR_REQUIRE(B > 0 && T >= 1 && T <= kHyMaxT && t0 >= 0 && t0 + T <= T_all && C >= 1 && C <= kHyMaxC && P > 0 && P % kHyVec == 0,
              "cor_hyper_logits_bwd: bad shape (T=%d C=%d P=%lld)", T, C, P);
  const dim3 grid(hy_blocks(P), B);
  cudaStream_t st = as_stream(stream);
  float* part = reinterpret_cast<float*>(work);
#define COR_HYB(TU, TG) hyper_logits_bwd_kernel<TU, TG><<<grid, kHyThreads, 0, st>>>(hyper, (const TU*)up, (const TG*)g, (TU*)d_up, part, T_all, t0, T, C, P)
  if (up_dtype == COR_F32 && g_dtype == COR_F32) COR_HYB(float, float);
  else if (up_dtype == COR_F32 && g_dtype == COR_BF16) COR_HYB(float, bf16);
  else if (up_dtype == COR_BF16 && g_dtype == COR_BF16) COR_HYB(bf16, bf16);
  else if (up_dtype == COR_BF16 && g_dtype == COR_F32) COR_HYB(bf16, float);
  else COR_REQUIRE(false, "cor_hyper_logits_bwd: unsupported dtypes up=%d g=%d", up_dtype, g_dtype);
#undef COR_HYB
  int rc = check_launch("hyper_logits_bwd_kernel");
  if (rc) return rc;
  const int total = B * T_all * C;
  hyper_fold_kernel<<<(total + 255) / 256, 256, 0, st>>>(part, d_hyper, B, (int)grid.x, T_all, t0, T, C);
  return check_launch("hyper_fold_kernel");
}
