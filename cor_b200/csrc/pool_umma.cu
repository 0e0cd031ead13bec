// Region pooling on the 5th-gen tensor cores (tcgen05 / TMEM), operands staged by TMA.
//
//   D[b][c][r] = sum_p F[b,c,p] * W[b,r,p]        F: bf16 feature map (A operand, 128 channels x 64 pixels
//                                                  per stage, K-major), W: bf16 mask weights (B operand,
//                                                  Rp masks x 64 pixels), D: fp32 in TMEM (128 lanes x Rp cols)
//
// With M >= 16 masks the contraction has 15-200 flop/B: it only stays HBM-bound if it
// runs on tensor cores.  The grid is (image, 128-channel block, K-split): split-K over the mask pixels
// spreads one 2 MB feature map over ~18 SMs; every CTA streams its slice ONCE through a 4-stage
// TMA -> mbarrier -> tcgen05.mma pipeline and writes one fp32 partial tile; the partials are summed in
// fixed order by the row epilogue (deterministic, no atomics).  Row Rp-1.. of W may hold an all-ones
// row so that sum_p F (background by subtraction) falls out of the same MMA.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (one elected lane),
// warps 2..5 = epilogue (TMEM -> registers -> coalesced global stores, one TMEM lane quarter each).
#include "umma.cuh"

namespace cor {

using namespace umma;

constexpr int kPoolStages = 4;
constexpr int kBK = 64;                 // pixels per stage = one 128-byte swizzle row of bf16
constexpr int kBM = 128;                // channels per CTA = UMMA M
constexpr int kABytes = kBM * kBK * 2;  // 16 KB

struct PoolSmemTail {
  uint64_t full[kPoolStages], empty[kPoolStages], accum;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(192, 2) pool_umma_kernel(const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmW,
                                                           int B, int C, int P, int Rp, int ksplit, uint32_t tmem_cols,
                                                           float* __restrict__ part) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t b_bytes = (uint32_t)Rp * kBK * 2;
  const uint32_t stage_bytes = kABytes + b_bytes;
  // dynamic smem is only guaranteed 16-byte aligned: round up to the 1024 B the 128B swizzle needs
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  PoolSmemTail* tail = reinterpret_cast<PoolSmemTail*>(base + kPoolStages * stage_bytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cblocks = C / kBM;
  const int s = blockIdx.x % ksplit;
  const int cb = (blockIdx.x / ksplit) % cblocks;
  const int b = blockIdx.x / (ksplit * cblocks);
  const int nkb = P / kBK;
  const int kb0 = (int)((long long)s * nkb / ksplit), kb1 = (int)((long long)(s + 1) * nkb / ksplit);
  const int iters = kb1 - kb0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmF);
    prefetch_tmap(&tmW);
    for (int i = 0; i < kPoolStages; ++i) { mbar_init(&tail->full[i], 1); mbar_init(&tail->empty[i], 1); }
    mbar_init(&tail->accum, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tail->tmem_base, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tail->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < iters; ++it) {
        const int st = it % kPoolStages;
        const uint32_t ph = (it / kPoolStages) & 1;
        mbar_wait(&tail->empty[st], ph ^ 1);
        uint8_t* a = base + st * stage_bytes;
        mbar_expect_tx(&tail->full[st], stage_bytes);
        tma_load_2d(a, &tmF, &tail->full[st], (kb0 + it) * kBK, b * C + cb * kBM, kEvictFirst);
        tma_load_2d(a + kABytes, &tmW, &tail->full[st], (kb0 + it) * kBK, b * Rp, kEvictLast);
      }
    }
  } else if (warp == 1) {
    // whole warp walks the loop, one elected lane issues: descriptors stay in uniform registers (see sim_umma.cu)
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(kBM, Rp);
    const uint32_t s_base = smem_u32(base);
    int st = 0;
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&tail->full[st], ph);
      tc_fence_after();
      if (leader) {
        const uint32_t a_addr = s_base + (uint32_t)(st * stage_bytes);
        const uint64_t da = make_desc_sw128(a_addr), db = make_desc_sw128(a_addr + kABytes);
        mma_bf16_ss(tmem_d, da, db, idesc, it != 0);
#pragma unroll
        for (int k = 1; k < kBK / 16; ++k) mma_bf16_ss_acc(tmem_d, da + 2 * k, db + 2 * k, idesc);
        mma_commit(&tail->empty[st]);     // frees the smem stage once these MMAs have read it
      }
      __syncwarp();
      if (++st == kPoolStages) { st = 0; ph ^= 1u; }
    }
    if (leader) mma_commit(&tail->accum);   // accumulator complete
    __syncwarp();
  } else {
    // epilogue: warp w owns TMEM lanes 32*(w%4)..+31 == channels cb*128 + 32*(w%4) + lane
    mbar_wait(&tail->accum, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int c = cb * kBM + q * 32 + lane;
    float* out = part + (((long long)s * B + b) * Rp) * C + c;
    if (iters > 0) {
      for (int col = 0; col < Rp; col += 16) {
        uint32_t v[16];
        tmem_ld_16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)col, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) out[(long long)(col + j) * C] = __uint_as_float(v[j]);   // 128 B coalesced per row
      }
    } else {
      for (int col = 0; col < Rp; ++col) out[(long long)col * C] = 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_d, tmem_cols);
}

static int pool_ksplit(int B, int C, int P) {
  const int tiles = B * (C / kBM), nkb = P / kBK;
  int ks = (2 * sm_count()) / (tiles > 0 ? tiles : 1);
  if (ks > nkb / 2) ks = nkb / 2;
  if (ks > 32) ks = 32;
  if (ks < 1) ks = 1;
  return ks;
}

}  // namespace cor

using namespace cor;

extern "C" int cor_pool_umma_ksplit(int B, int C, int P) { return pool_ksplit(B, C, P); }

extern "C" size_t cor_pool_umma_work_bytes(int B, int C, int P, int Rp) {
  return (size_t)pool_ksplit(B, C, P) * B * Rp * C * sizeof(float);
}

extern "C" int cor_pool_umma_fwd(const void* feat_bf16, const void* wts_bf16, int B, int C, int P, int Rp, float* part,
                                 cor_stream_t stream) {
  COR_REQUIRE(feat_bf16 && wts_bf16 && part, "cor_pool_umma_fwd: null pointer");
  COR_REQUIRE(B > 0 && C % kBM == 0 && P % kBK == 0 && Rp % 16 == 0 && Rp >= 16 && Rp <= 256,
              "cor_pool_umma_fwd: need C %% 128 == 0, P %% 64 == 0, 16 <= Rp <= 256, Rp %% 16 == 0 (C=%d P=%d Rp=%d)", C, P, Rp);
  CUtensorMap tmF, tmW;
  int rc = umma::encode_tmap_bf16_2d(&tmF, feat_bf16, (uint64_t)B * C, (uint64_t)P, kBM, kBK);
  if (rc) return rc;
  rc = umma::encode_tmap_bf16_2d(&tmW, wts_bf16, (uint64_t)B * Rp, (uint64_t)P, (uint32_t)Rp, kBK);
  if (rc) return rc;
  const int ks = pool_ksplit(B, C, P);
  uint32_t cols = 32;
  while ((int)cols < Rp) cols <<= 1;
  const size_t smem = (size_t)kPoolStages * (kABytes + (size_t)Rp * kBK * 2) + sizeof(PoolSmemTail) + 1024;
  COR_CUDA(cudaFuncSetAttribute(pool_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long grid = (long long)B * (C / kBM) * ks;
  COR_REQUIRE(grid < 2147483647LL, "cor_pool_umma_fwd: grid too large");
  pool_umma_kernel<<<(unsigned)grid, 192, smem, as_stream(stream)>>>(tmF, tmW, B, C, P, Rp, ks, cols, part);
  return check_launch("pool_umma_kernel");
}
