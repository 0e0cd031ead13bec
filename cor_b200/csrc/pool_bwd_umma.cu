// Backward of region pooling w.r.t. the feature map, on tcgen05 tensor cores.
//
//   g_feat[b,c,p] = sum_r g_fg[b,r,c] * w[b,r,p] + g_bg[b,r,c] * (1 - w[b,r,p])
//                 = sum_{r<R} (g_fg - g_bg)[b,r,c] * w[b,r,p]  +  (sum_r g_bg[b,r,c]) * 1
//
// i.e. per image a [C x K] x [K x P] GEMM with K = R (+1 "ones" row for the background term), tiny
// in flops and bound by the write of g_feat (B*C*P elements).  One CTA per (image, 128 channels,
// 256 pixels): the operands are built in shared memory straight from the fp32 tensors the forward
// saved (coalesced global reads, bf16 pack, 16-byte swizzled stores in the canonical K-major 128B-swizzle
// layout the UMMA descriptors expect), one thread issues K/16 tcgen05.mma (M128 x N256 x K16) into TMEM,
// and all 8 warps drain TMEM with tcgen05.ld and write full 32-byte sectors of g_feat.
#include "umma.cuh"

namespace cor {

using namespace umma;

constexpr int kGM = 128;     // channels per CTA (UMMA M)
constexpr int kGN = 256;     // pixels per CTA   (UMMA N)

__device__ __forceinline__ float transform_wb(float r, int transform) {
  if (transform == COR_W_CLAMP) return fminf(fmaxf(r, 0.f), 1.f);
  if (transform == COR_W_SIGMOID) return sigmoid_acc(r);
  return r;
}

// byte offset of the 16-byte chunk holding k-elements [8*c16, 8*c16+8) of `row` inside one 64-wide K chunk
__device__ __forceinline__ uint32_t sw128_off(int row, int c16) { return (uint32_t)row * 128u + (uint32_t)((c16 ^ (row & 7)) << 4); }

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <typename TG>
__global__ void __launch_bounds__(256, 2) pool_bwd_umma_kernel(const float* __restrict__ g_fg, const float* __restrict__ g_bg,
                                                               const float* __restrict__ wts, long long ldw, int C, int P, int R,
                                                               int transform, int Kp, TG* __restrict__ g_feat) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  const int nchunk = (Kp + 63) / 64;
  uint8_t* a_smem = base;                                   // nchunk x [128 rows x 128 B]
  uint8_t* b_smem = base + (size_t)nchunk * kGM * 128;      // nchunk x [256 rows x 128 B]
  uint64_t* done = reinterpret_cast<uint64_t*>(b_smem + (size_t)nchunk * kGN * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ptiles = P / kGN, cblocks = C / kGM;
  const int pt = blockIdx.x % ptiles, cb = (blockIdx.x / ptiles) % cblocks, b = blockIdx.x / (ptiles * cblocks);
  const int p0 = pt * kGN, c0 = cb * kGM;
  const bool has_bg = g_bg != nullptr;

  if (threadIdx.x == 0) { mbar_init(done, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc(tmem_slot, kGN);

  // ---- B operand: w[p][r] (row = pixel, K = mask) from the fp32 weights, transform on the fly -------
  // The build is pure load latency: keep 32 independent loads in flight per thread before packing.
  {
    const int p = threadIdx.x;                               // 256 threads <-> 256 pixels of the tile
    const float* wp = wts + (long long)b * R * ldw + p0 + p;
    for (int rb = 0; rb < Kp; rb += 32) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int r = rb + j;
        v[j] = r < R ? __ldg(wp + (long long)r * ldw) : 0.f;
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int r8 = (rb >> 3) + g;
        if (r8 * 8 < Kp) {
          float t[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int r = r8 * 8 + j;
            t[j] = r < R ? transform_wb(v[g * 8 + j], transform) : ((has_bg && r == R) ? 1.f : 0.f);
          }
          uint4 q = make_uint4(pack_bf16x2(t[0], t[1]), pack_bf16x2(t[2], t[3]), pack_bf16x2(t[4], t[5]), pack_bf16x2(t[6], t[7]));
          *reinterpret_cast<uint4*>(b_smem + (size_t)(r8 >> 3) * kGN * 128 + sw128_off(p, r8 & 7)) = q;
        }
      }
    }
  }
  // ---- A operand: (g_fg - g_bg)[c][r] (row = channel, K = mask); ones-row column = sum_r g_bg[r][c] ---
  {
    const int c = threadIdx.x & (kGM - 1), half = threadIdx.x >> 7;
    const float* gf = g_fg + (long long)b * R * C + c0 + c;
    const float* gb = has_bg ? g_bg + (long long)b * R * C + c0 + c : nullptr;
    float bgsum = 0.f;
    if (has_bg && (R / 8) % 2 == half) {                     // the thread half that will write column R
      float part8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int r = 0; r < R; r += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (r + j < R) part8[j] += __ldg(gb + (long long)(r + j) * C);
      }
      bgsum = ((part8[0] + part8[1]) + (part8[2] + part8[3])) + ((part8[4] + part8[5]) + (part8[6] + part8[7]));
    }
    for (int r8 = half; r8 < Kp / 8; r8 += 4) {              // two 8-groups (r8, r8+2) per round: 32 loads in flight
      float vf[2][8], vb[2][8];
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int r = (r8 + 2 * u) * 8 + j;
          vf[u][j] = r < R ? __ldg(gf + (long long)r * C) : 0.f;
          vb[u][j] = (has_bg && r < R) ? __ldg(gb + (long long)r * C) : 0.f;
        }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int g8 = r8 + 2 * u;
        if (g8 < Kp / 8) {
          float t[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int r = g8 * 8 + j;
            t[j] = r < R ? vf[u][j] - vb[u][j] : ((has_bg && r == R) ? bgsum : 0.f);
          }
          uint4 q = make_uint4(pack_bf16x2(t[0], t[1]), pack_bf16x2(t[2], t[3]), pack_bf16x2(t[4], t[5]), pack_bf16x2(t[6], t[7]));
          *reinterpret_cast<uint4*>(a_smem + (size_t)(g8 >> 3) * kGM * 128 + sw128_off(c, g8 & 7)) = q;
        }
      }
    }
  }
  fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0 && lane == 0) {
    const uint32_t idesc = make_idesc_bf16(kGM, kGN);
    for (int t = 0; t < Kp / 16; ++t) {
      const int ch = t >> 2, k = t & 3;
      const uint64_t da = make_desc_sw128(smem_u32(a_smem + (size_t)ch * kGM * 128)) + 2 * k;
      const uint64_t db = make_desc_sw128(smem_u32(b_smem + (size_t)ch * kGN * 128)) + 2 * k;
      mma_bf16_ss(tmem, da, db, idesc, t != 0);
    }
    mma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after();

  // ---- epilogue: warp w reads lane quarter w%4, column half w/4; thread = one channel row --------------
  {
    const int q = warp & 3, hf = warp >> 2;
    const int c = c0 + q * 32 + lane;
    TG* out = g_feat + ((long long)b * C + c) * P + p0 + hf * 128;
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * 128);
#pragma unroll 2
    for (int col = 0; col < 128; col += 16) {
      uint32_t v[16];
      tmem_ld_16(taddr + (uint32_t)col, v);
      tmem_ld_wait();
      if (sizeof(TG) == 2) {
        uint4 lo = make_uint4(pack_bf16x2(__uint_as_float(v[0]), __uint_as_float(v[1])), pack_bf16x2(__uint_as_float(v[2]), __uint_as_float(v[3])),
                              pack_bf16x2(__uint_as_float(v[4]), __uint_as_float(v[5])), pack_bf16x2(__uint_as_float(v[6]), __uint_as_float(v[7])));
        uint4 hi = make_uint4(pack_bf16x2(__uint_as_float(v[8]), __uint_as_float(v[9])), pack_bf16x2(__uint_as_float(v[10]), __uint_as_float(v[11])),
                              pack_bf16x2(__uint_as_float(v[12]), __uint_as_float(v[13])), pack_bf16x2(__uint_as_float(v[14]), __uint_as_float(v[15])));
        uint4* o = reinterpret_cast<uint4*>(out + col);
        o[0] = lo;
        o[1] = hi;
      } else {
        uint4* o = reinterpret_cast<uint4*>(out + col);
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, kGN);
}

}  // namespace cor

using namespace cor;

// 1 if the tensor-core backward can serve this shape (else callers use cor_pool_bwd_feat).
extern "C" int cor_pool_bwd_umma_ok(int B, int C, int P, int R, int has_bg) {
  const int K = R + (has_bg ? 1 : 0);
  return B > 0 && C % kGM == 0 && P % kGN == 0 && K >= 1 && K <= 256;
}

extern "C" int cor_pool_bwd_umma(const float* g_fg, const float* g_bg, const float* wts, long long ldw, int B, int C, int P, int R,
                                 int transform, void* g_feat, int feat_dtype, cor_stream_t stream) {
  COR_REQUIRE(g_fg && wts && g_feat, "cor_pool_bwd_umma: null pointer");
  COR_REQUIRE(cor_pool_bwd_umma_ok(B, C, P, R, g_bg != nullptr), "cor_pool_bwd_umma: need C %% 128 == 0, P %% 256 == 0, R (+1) <= 256 (C=%d P=%d R=%d)", C, P, R);
  COR_REQUIRE((((uintptr_t)g_feat) & 15) == 0, "cor_pool_bwd_umma: g_feat must be 16-byte aligned");
  const int K = R + (g_bg ? 1 : 0);
  const int Kp = (K + 15) / 16 * 16;
  const int nchunk = (Kp + 63) / 64;
  const size_t smem = (size_t)nchunk * (kGM + kGN) * 128 + 64 + 1024;
  const long long grid = (long long)B * (C / kGM) * (P / kGN);
  COR_REQUIRE(grid < 2147483647LL, "cor_pool_bwd_umma: grid too large");
  cudaStream_t st = as_stream(stream);
  if (feat_dtype == COR_BF16) {
    COR_CUDA(cudaFuncSetAttribute(pool_bwd_umma_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pool_bwd_umma_kernel<bf16><<<(unsigned)grid, 256, smem, st>>>(g_fg, g_bg, wts, ldw, C, P, R, transform, Kp, (bf16*)g_feat);
  } else if (feat_dtype == COR_F32) {
    COR_CUDA(cudaFuncSetAttribute(pool_bwd_umma_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pool_bwd_umma_kernel<float><<<(unsigned)grid, 256, smem, st>>>(g_fg, g_bg, wts, ldw, C, P, R, transform, Kp, (float*)g_feat);
  } else {
    COR_REQUIRE(false, "cor_pool_bwd_umma: unsupported dtype %d", feat_dtype);
  }
  return check_launch("pool_bwd_umma_kernel");
}
