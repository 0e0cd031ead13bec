// Region pooling, CUDA-core streaming variant + row epilogues + backward.
//
//   fg_sum[b,r,c] = sum_p w[b,r,p] * F[b,c,p]          (mask_adapter.py:22-23, loss_func.py:50-52,
//   bg_sum[b,r,c] = sum_p (1-w[b,r,p]) * F[b,c,p]       and the bmm at mask_adapter.py:72-75)
//
// HBM-bound for small R: each feature row is read once with 128-bit loads while the (transformed)
// weight rows sit in shared memory; the reduction over mask pixels is per-lane FMA chains closed by
// warp shuffles.  Exact fp32 (no tensor cores): this is the path the reference-shaped M=1 calls and the
// 8-map MaskAdapter tail take; many-mask bf16 pooling goes to pool_umma.cu.
#include "common.cuh"

namespace cor {

constexpr int kPTBudget = 8192;   // floats of weights in shared memory: the tile covers 8192 / RT mask pixels
constexpr int kWarps = 8;
constexpr int kCPW = 2;         // channels per warp per pass (weight registers reused across them)
constexpr int kCPB = kWarps * kCPW;

__device__ __forceinline__ float transform_w(float r, int transform) {
  if (transform == COR_W_CLAMP) return fminf(fmaxf(r, 0.f), 1.f);
  if (transform == COR_W_SIGMOID) return sigmoid_acc(r);
  return r;
}

template <typename TF>
struct FeatVec;
template <>
struct FeatVec<float> {
  static constexpr int N = 4;
  __device__ static __forceinline__ void unpack(uint4 v, float (&f)[4]) {
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
  }
};
template <>
struct FeatVec<bf16> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void unpack(uint4 v, float (&f)[8]) {
    f[0] = bf16lo(v.x); f[1] = bf16hi(v.x); f[2] = bf16lo(v.y); f[3] = bf16hi(v.y);
    f[4] = bf16lo(v.z); f[5] = bf16hi(v.z); f[6] = bf16lo(v.w); f[7] = bf16hi(v.w);
  }
};

// grid = (ceil(C/kCPB), B, ceil(R/RT)); block = 256.
template <typename TF, int RT, bool PAIR>
__global__ void __launch_bounds__(kWarps * 32) pool_stream_kernel(const TF* __restrict__ feat, const float* __restrict__ wts,
                                                                  long long ldw, int C, int P, int R, int transform,
                                                                  int vec_ok, float* __restrict__ fg_sum,
                                                                  float* __restrict__ bg_sum) {
  constexpr int kPT = kPTBudget / RT;   // RT=1 (the reference's single mask): the whole 64x64 map in one tile
  __shared__ __align__(16) float ws[RT][kPT];
  const int b = blockIdx.y, rbase = blockIdx.z * RT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.x * kCPB + warp * kCPW;
  constexpr int VE = FeatVec<TF>::N;

  float afg[kCPW][RT], abg[kCPW][RT];
#pragma unroll
  for (int k = 0; k < kCPW; ++k)
#pragma unroll
    for (int r = 0; r < RT; ++r) afg[k][r] = abg[k][r] = 0.f;

  const TF* frow[kCPW];
#pragma unroll
  for (int k = 0; k < kCPW; ++k) frow[k] = feat + ((long long)b * C + min(c0 + k, C - 1)) * P;

  for (int p0 = 0; p0 < P; p0 += kPT) {
    const int pt = min(kPT, P - p0);
    __syncthreads();
    for (int i = threadIdx.x; i < RT * kPT; i += blockDim.x) {
      const int r = i / kPT, p = i % kPT;
      float v = 0.f;
      if (rbase + r < R && p < pt) v = transform_w(__ldg(wts + ((long long)b * R + rbase + r) * ldw + p0 + p), transform);
      ws[r][p] = v;
    }
    __syncthreads();
    if (c0 >= C) continue;
    if (vec_ok) {
      const int nvec = pt / VE;
      constexpr int kU = (RT <= 2) ? 4 : 1;      // vectors in flight per lane and channel (small RT: latency-bound)
      for (int v0 = lane; v0 < nvec; v0 += 32 * kU) {
        float f[kU][kCPW][VE];
#pragma unroll
        for (int u = 0; u < kU; ++u)
#pragma unroll
          for (int k = 0; k < kCPW; ++k) {
            const int v = v0 + 32 * u;
            if (v < nvec) {
              uint4 raw = ld_stream16(reinterpret_cast<const uint4*>(frow[k] + p0) + v);
              FeatVec<TF>::unpack(raw, f[u][k]);
            }
          }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int v = v0 + 32 * u;
          if (v >= nvec) break;
#pragma unroll
          for (int r = 0; r < RT; ++r) {
            float wv[VE];
#pragma unroll
            for (int q = 0; q < VE / 4; ++q) {
              float4 t = *reinterpret_cast<const float4*>(&ws[r][v * VE + q * 4]);
              wv[q * 4 + 0] = t.x; wv[q * 4 + 1] = t.y; wv[q * 4 + 2] = t.z; wv[q * 4 + 3] = t.w;
            }
#pragma unroll
            for (int k = 0; k < kCPW; ++k)
#pragma unroll
              for (int e = 0; e < VE; ++e) {
                afg[k][r] = fmaf(f[u][k][e], wv[e], afg[k][r]);
                if (PAIR) abg[k][r] = fmaf(f[u][k][e], 1.f - wv[e], abg[k][r]);
              }
          }
        }
      }
      // pt is a multiple of VE whenever vec_ok (P % VE == 0 and kPT % VE == 0)
    } else {
      for (int p = lane; p < pt; p += 32) {
        float f[kCPW];
#pragma unroll
        for (int k = 0; k < kCPW; ++k) f[k] = to_f<TF>(__ldg(frow[k] + p0 + p));
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          const float wv = ws[r][p];
#pragma unroll
          for (int k = 0; k < kCPW; ++k) {
            afg[k][r] = fmaf(f[k], wv, afg[k][r]);
            if (PAIR) abg[k][r] = fmaf(f[k], 1.f - wv, abg[k][r]);
          }
        }
      }
    }
  }
  if (c0 >= C) return;
#pragma unroll
  for (int k = 0; k < kCPW; ++k)
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const float a = warp_sum(afg[k][r]);
      const float g = PAIR ? warp_sum(abg[k][r]) : 0.f;
      if (lane == 0 && c0 + k < C && rbase + r < R) {
        const long long o = ((long long)b * R + rbase + r) * C + c0 + k;
        fg_sum[o] = a;
        if (PAIR) bg_sum[o] = g;
      }
    }
}

template <typename TF, bool PAIR>
static int launch_pool_stream(const TF* feat, const float* wts, long long ldw, int B, int C, int P, int R, int transform,
                              float* fg, float* bg, cudaStream_t st) {
  const int vec_ok = ((P * sizeof(TF)) % 16 == 0) && (((uintptr_t)feat & 15) == 0);
  const int rt = R == 1 ? 1 : R == 2 ? 2 : R <= 4 ? 4 : 8;
  dim3 grid(ceil_div(C, kCPB), B, ceil_div(R, rt));
#define COR_LAUNCH(RT_) pool_stream_kernel<TF, RT_, PAIR><<<grid, kWarps * 32, 0, st>>>(feat, wts, ldw, C, P, R, transform, vec_ok, fg, bg)
  if (rt == 1) COR_LAUNCH(1);
  else if (rt == 2) COR_LAUNCH(2);
  else if (rt == 4) COR_LAUNCH(4);
  else COR_LAUNCH(8);
#undef COR_LAUNCH
  return check_launch("pool_stream_kernel");
}

// ---- row epilogue: divide, group mean, L2-normalise ---------------------------------------------
// grid = rows_out, block = 256, dynamic smem = C floats.
__global__ void __launch_bounds__(256) rows_finalize_kernel(const float* __restrict__ sums, int rows_per_image, long long img_stride,
                                                            int nsplit, long long split_stride, const float* __restrict__ den, int den_stride, float eps, int C, int G,
                                                            int normalize, const float* __restrict__ all, float p_total,
                                                            float* __restrict__ out_f32, bf16* __restrict__ out_bf16,
                                                            float* __restrict__ inv_norm) {
  extern __shared__ float vals[];
  __shared__ float scratch[32];
  const int j = blockIdx.x;
  float ss[1] = {0.f};
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int g = 0; g < G; ++g) {
      const long long i = (long long)j * G + g;
      const long long img = i / rows_per_image;
      float d = den[i * den_stride];
      const long long so = img * img_stride + (i % rows_per_image) * C + c;
      float s = sums[so];
      for (int k = 1; k < nsplit; ++k) s += sums[so + k * split_stride];
      if (all) {
        float a = all[img * img_stride + c];
        for (int k = 1; k < nsplit; ++k) a += all[img * img_stride + c + k * split_stride];
        s = a - s;
        d = p_total - d;
      }
      acc += s / (d + eps);
    }
    acc = acc / (float)G;
    vals[c] = acc;
    ss[0] += acc * acc;
  }
  float inv = 1.f;
  if (normalize) {
    block_sum<1>(ss, scratch);
    inv = 1.f / fmaxf(sqrtf(ss[0]), 1e-12f);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float v = vals[c] * inv;
    out_f32[(long long)j * C + c] = v;
    if (out_bf16) out_bf16[(long long)j * C + c] = __float2bfloat16_rn(v);
  }
  if (threadIdx.x == 0 && inv_norm) inv_norm[j] = inv;
}

// g_out [rows_out,C] -> g_sums [rows_in,C]   (d out / d sums; for `all`-derived rows the caller treats
// g_sums as the gradient of the background sum itself).
__global__ void __launch_bounds__(256) rows_finalize_bwd_kernel(const float* __restrict__ g_out, const float* __restrict__ out_f32,
                                                                const float* __restrict__ inv_norm, const float* __restrict__ den,
                                                                int den_stride, float eps, int C, int G, int normalize,
                                                                int bg_from_all, float p_total, float* __restrict__ g_sums) {
  __shared__ float scratch[32];
  const int j = blockIdx.x;
  float dot[1] = {0.f};
  float inv = 1.f;
  if (normalize) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) dot[0] += g_out[(long long)j * C + c] * out_f32[(long long)j * C + c];
    block_sum<1>(dot, scratch);
    inv = inv_norm[j];
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float gv = g_out[(long long)j * C + c];
    if (normalize) gv = inv * (gv - out_f32[(long long)j * C + c] * dot[0]);
    gv /= (float)G;
    for (int g = 0; g < G; ++g) {
      const long long i = (long long)j * G + g;
      float d = den[i * den_stride];
      if (bg_from_all) d = p_total - d;
      g_sums[i * C + c] = gv / (d + eps);
    }
  }
}

// ---- backward w.r.t. features ---------------------------------------------------------------------
// g_feat[b,c,p] = sum_r g_fg[b,r,c] w[b,r,p] + g_bg[b,r,c] (1 - w[b,r,p])
// grid = (ceil(P/256), ceil(C/32), B), block = 256 (64 p-quads x 4 channel lanes), r in chunks of 16.
constexpr int kBR = 16, kBP = 256, kBC = 32;
template <typename TG, bool PAIR>
__global__ void __launch_bounds__(256) pool_bwd_feat_kernel(const float* __restrict__ g_fg, const float* __restrict__ g_bg,
                                                            const float* __restrict__ wts, long long ldw, int C, int P, int R,
                                                            int transform, TG* __restrict__ g_feat) {
  __shared__ __align__(16) float ws[kBR][kBP];
  __shared__ float gf[kBR][kBC], gb[kBR][kBC];
  const int b = blockIdx.z, c0 = blockIdx.y * kBC, p0 = blockIdx.x * kBP;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  float acc[kBC / 4][4];
#pragma unroll
  for (int k = 0; k < kBC / 4; ++k) acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f;
  for (int r0 = 0; r0 < R; r0 += kBR) {
    __syncthreads();
    for (int i = threadIdx.x; i < kBR * kBP; i += blockDim.x) {
      const int r = i / kBP, p = i % kBP;
      float v = 0.f;
      if (r0 + r < R && p0 + p < P) v = transform_w(__ldg(wts + ((long long)b * R + r0 + r) * ldw + p0 + p), transform);
      ws[r][p] = v;
    }
    for (int i = threadIdx.x; i < kBR * kBC; i += blockDim.x) {
      const int r = i / kBC, c = i % kBC;
      const bool ok = r0 + r < R && c0 + c < C;
      const long long o = ((long long)b * R + r0 + r) * C + c0 + c;
      gf[r][c] = ok ? g_fg[o] : 0.f;
      gb[r][c] = (PAIR && ok) ? g_bg[o] : 0.f;
    }
    __syncthreads();
    const int rn = min(kBR, R - r0);
    for (int r = 0; r < rn; ++r) {
      const float4 w4 = *reinterpret_cast<const float4*>(&ws[r][tx * 4]);
#pragma unroll
      for (int k = 0; k < kBC / 4; ++k) {
        const float a = gf[r][ty + 4 * k];
        acc[k][0] = fmaf(a, w4.x, acc[k][0]); acc[k][1] = fmaf(a, w4.y, acc[k][1]);
        acc[k][2] = fmaf(a, w4.z, acc[k][2]); acc[k][3] = fmaf(a, w4.w, acc[k][3]);
        if (PAIR) {
          const float bb = gb[r][ty + 4 * k];
          acc[k][0] = fmaf(bb, 1.f - w4.x, acc[k][0]); acc[k][1] = fmaf(bb, 1.f - w4.y, acc[k][1]);
          acc[k][2] = fmaf(bb, 1.f - w4.z, acc[k][2]); acc[k][3] = fmaf(bb, 1.f - w4.w, acc[k][3]);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kBC / 4; ++k) {
    const int c = c0 + ty + 4 * k;
    if (c >= C) continue;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int p = p0 + tx * 4 + e;
      if (p < P) g_feat[((long long)b * C + c) * P + p] = from_f<TG>(acc[k][e]);
    }
  }
}

// ---- backward w.r.t. sigmoid maps (MaskAdapter tail, mask_adapter.py:71-79) -------------------------
//   pooled[r,c] = sum_p s[r,p] F[c,p] / D_r,  s = sigmoid(x), D_r = sum_p s[r,p]
//   g_x[r,p] = s(1-s)/D_r * ( sum_c g[r,c] F[c,p] - sum_c g[r,c] pooled[r,c] ),  g = dL/d pooled[r]
// grid = (ceil(P/256), B, ceil(R/8)); block 256: one thread per mask pixel, loop over channels.
constexpr int kMR = 8;
template <typename TF>
__global__ void __launch_bounds__(256) pool_bwd_maps_kernel(const TF* __restrict__ feat, const float* __restrict__ maps, long long ldw,
                                                            const float* __restrict__ g_pooled, const float* __restrict__ pooled,
                                                            const float* __restrict__ den, int C, int P, int R,
                                                            float* __restrict__ g_maps) {
  extern __shared__ float gs[];   // [kMR][C]
  __shared__ float kr[kMR];
  __shared__ float scratch[kMR * 32];
  const int b = blockIdx.y, r0 = blockIdx.z * kMR, p = blockIdx.x * 256 + threadIdx.x;
  float dots[kMR];
#pragma unroll
  for (int r = 0; r < kMR; ++r) dots[r] = 0.f;
  for (int i = threadIdx.x; i < kMR * C; i += blockDim.x) {
    const int r = i / C, c = i % C;
    float g = 0.f;
    if (r0 + r < R) {
      const long long o = ((long long)b * R + r0 + r) * C + c;
      g = g_pooled[o];
#pragma unroll
      for (int q = 0; q < kMR; ++q)
        if (q == r) dots[q] += g * pooled[o];
    }
    gs[r * C + c] = g;
  }
  block_sum<kMR>(dots, scratch);
  if (threadIdx.x < kMR) kr[threadIdx.x] = 0.f;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int r = 0; r < kMR; ++r) kr[r] = dots[r];
  }
  __syncthreads();
  if (p >= P) return;
  float acc[kMR];
#pragma unroll
  for (int r = 0; r < kMR; ++r) acc[r] = 0.f;
  const TF* f = feat + (long long)b * C * P + p;
  for (int c = 0; c < C; ++c) {
    const float fv = to_f<TF>(__ldg(f + (long long)c * P));
#pragma unroll
    for (int r = 0; r < kMR; ++r) acc[r] = fmaf(gs[r * C + c], fv, acc[r]);
  }
#pragma unroll
  for (int r = 0; r < kMR; ++r) {
    if (r0 + r >= R) break;
    const long long row = (long long)b * R + r0 + r;
    const float s = sigmoid_acc(maps[row * ldw + p]);
    g_maps[row * ldw + p] = s * (1.f - s) / den[row] * (acc[r] - kr[r]);
  }
}

}  // namespace cor

using namespace cor;

extern "C" int cor_pool_stream_fwd(const void* feat, int feat_dtype, const float* wts, long long ldw, int B, int C, int P,
                                   int R, int transform, float* fg_sum, float* bg_sum, cor_stream_t stream) {
  COR_REQUIRE(feat && wts && fg_sum, "cor_pool_stream_fwd: null pointer");
  COR_REQUIRE(B > 0 && C > 0 && P > 0 && R > 0 && ldw >= P, "cor_pool_stream_fwd: bad shape B=%d C=%d P=%d R=%d ldw=%lld", B, C,
              P, R, ldw);
  COR_REQUIRE(B <= 65535 && ceil_div(R, 8) <= 65535, "cor_pool_stream_fwd: B or R too large for the grid");
  cudaStream_t st = as_stream(stream);
  if (feat_dtype == COR_F32)
    return bg_sum ? launch_pool_stream<float, true>((const float*)feat, wts, ldw, B, C, P, R, transform, fg_sum, bg_sum, st)
                  : launch_pool_stream<float, false>((const float*)feat, wts, ldw, B, C, P, R, transform, fg_sum, bg_sum, st);
  if (feat_dtype == COR_BF16)
    return bg_sum ? launch_pool_stream<bf16, true>((const bf16*)feat, wts, ldw, B, C, P, R, transform, fg_sum, bg_sum, st)
                  : launch_pool_stream<bf16, false>((const bf16*)feat, wts, ldw, B, C, P, R, transform, fg_sum, bg_sum, st);
  COR_REQUIRE(false, "cor_pool_stream_fwd: unsupported feature dtype %d", feat_dtype);
}

extern "C" int cor_rows_finalize(const float* sums, int rows_per_image, long long img_stride, int nsplit, long long split_stride,
                                 const float* den, int den_stride, float eps, int rows_in, int C, int G, int normalize,
                                 const float* all_sum, float p_total, float* out_f32, void* out_bf16, float* inv_norm,
                                 cor_stream_t stream) {
  COR_REQUIRE(sums && den && out_f32, "cor_rows_finalize: null pointer");
  COR_REQUIRE(rows_in > 0 && C > 0 && G > 0 && rows_in % G == 0 && den_stride > 0, "cor_rows_finalize: bad shape rows=%d C=%d G=%d",
              rows_in, C, G);
  COR_REQUIRE(C * sizeof(float) <= 48 * 1024, "cor_rows_finalize: C=%d too large", C);
  if (rows_per_image <= 0) { rows_per_image = rows_in; img_stride = (long long)rows_in * C; }
  COR_REQUIRE(img_stride >= (long long)rows_per_image * C, "cor_rows_finalize: img_stride too small");
  rows_finalize_kernel<<<rows_in / G, 256, C * sizeof(float), as_stream(stream)>>>(
      sums, rows_per_image, img_stride, nsplit, split_stride, den, den_stride, eps, C, G, normalize, all_sum, p_total, out_f32,
      (bf16*)out_bf16, inv_norm);
  return check_launch("rows_finalize_kernel");
}

extern "C" int cor_rows_finalize_bwd(const float* g_out, const float* out_f32, const float* inv_norm, const float* den,
                                     int den_stride, float eps, int rows_in, int C, int G, int normalize, int bg_from_all,
                                     float p_total, float* g_sums, cor_stream_t stream) {
  COR_REQUIRE(g_out && den && g_sums, "cor_rows_finalize_bwd: null pointer");
  COR_REQUIRE(!normalize || (out_f32 && inv_norm), "cor_rows_finalize_bwd: normalize needs out_f32 and inv_norm");
  COR_REQUIRE(rows_in > 0 && C > 0 && G > 0 && rows_in % G == 0, "cor_rows_finalize_bwd: bad shape");
  rows_finalize_bwd_kernel<<<rows_in / G, 256, 0, as_stream(stream)>>>(g_out, out_f32, inv_norm, den, den_stride, eps, C, G,
                                                                       normalize, bg_from_all, p_total, g_sums);
  return check_launch("rows_finalize_bwd_kernel");
}

extern "C" int cor_pool_bwd_feat(const float* g_fg, const float* g_bg, const float* wts, long long ldw, int B, int C, int P,
                                 int R, int transform, void* g_feat, int feat_dtype, cor_stream_t stream) {
  COR_REQUIRE(g_fg && wts && g_feat, "cor_pool_bwd_feat: null pointer");
  COR_REQUIRE(B > 0 && C > 0 && P > 0 && R > 0 && B <= 65535, "cor_pool_bwd_feat: bad shape");
  dim3 grid(ceil_div(P, kBP), ceil_div(C, kBC), B);
  cudaStream_t st = as_stream(stream);
  if (feat_dtype == COR_F32) {
    if (g_bg) pool_bwd_feat_kernel<float, true><<<grid, 256, 0, st>>>(g_fg, g_bg, wts, ldw, C, P, R, transform, (float*)g_feat);
    else pool_bwd_feat_kernel<float, false><<<grid, 256, 0, st>>>(g_fg, g_bg, wts, ldw, C, P, R, transform, (float*)g_feat);
  } else if (feat_dtype == COR_BF16) {
    if (g_bg) pool_bwd_feat_kernel<bf16, true><<<grid, 256, 0, st>>>(g_fg, g_bg, wts, ldw, C, P, R, transform, (bf16*)g_feat);
    else pool_bwd_feat_kernel<bf16, false><<<grid, 256, 0, st>>>(g_fg, g_bg, wts, ldw, C, P, R, transform, (bf16*)g_feat);
  } else {
    COR_REQUIRE(false, "cor_pool_bwd_feat: unsupported dtype %d", feat_dtype);
  }
  return check_launch("pool_bwd_feat_kernel");
}

extern "C" int cor_pool_bwd_maps(const void* feat, int feat_dtype, const float* maps, long long ldw, const float* g_pooled,
                                 const float* pooled, const float* den, int B, int C, int P, int R, float* g_maps,
                                 cor_stream_t stream) {
  COR_REQUIRE(feat && maps && g_pooled && pooled && den && g_maps, "cor_pool_bwd_maps: null pointer");
  COR_REQUIRE(B > 0 && C > 0 && P > 0 && R > 0 && B <= 65535, "cor_pool_bwd_maps: bad shape");
  const size_t smem = (size_t)kMR * C * sizeof(float);
  COR_REQUIRE(smem <= 48 * 1024, "cor_pool_bwd_maps: C=%d too large", C);
  dim3 grid(ceil_div(P, 256), B, ceil_div(R, kMR));
  cudaStream_t st = as_stream(stream);
  if (feat_dtype == COR_F32)
    pool_bwd_maps_kernel<float><<<grid, 256, smem, st>>>((const float*)feat, maps, ldw, g_pooled, pooled, den, C, P, R, g_maps);
  else if (feat_dtype == COR_BF16)
    pool_bwd_maps_kernel<bf16><<<grid, 256, smem, st>>>((const bf16*)feat, maps, ldw, g_pooled, pooled, den, C, P, R, g_maps);
  else
    COR_REQUIRE(false, "cor_pool_bwd_maps: unsupported dtype %d", feat_dtype);
  return check_launch("pool_bwd_maps_kernel");
}
