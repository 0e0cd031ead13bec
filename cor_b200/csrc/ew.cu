// Small element-wise kernels around the tensor-core GEMM (gemm_umma.cu): operand casts and the backward of the fused
// bias + activation (+ dropout mask) epilogue.  HBM-bound streams; the tensors are a few MB at most.
#include "common.cuh"

namespace cor {

// out[r][0:c0] = a[r][:], out[r][c0:c0+c1] = b[r][:]  (b may be null), f32 -> bf16; 4 elements per thread
__global__ void __launch_bounds__(256) cast_cat_bf16_kernel(const float* __restrict__ a, int c0, const float* __restrict__ b, int c1,
                                                           long long rows, bf16* __restrict__ out) {
  const int Ct = c0 + c1;
  const long long total = rows * Ct;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / Ct;
    const int c = (int)(i % Ct);
    out[i] = __float2bfloat16_rn(c < c0 ? a[r * c0 + c] : b[r * c1 + (c - c0)]);
  }
}

// dz[m][n] = dy[m][n] * emul[m][n] * colscale[n] * act'(.)  (bf16, the A operand of dX = dZ W and dW = dZ^T X);
// db[n] = sum_m dz (f32); dcs[n] = sum_m dy[m][n] * emul * act(pre)[m][n] (gradient of the column scale, ConvNeXt's gamma).
// y = the activation's OUTPUT before mask / scale (relu, sigmoid) or pre = its pre-activation (gelu; also the un-scaled
// linear output when a column scale is used), saved bf16 by the GEMM epilogue.  grid = (column blocks, row chunks): a thread
// walks its column over the chunk's rows (coalesced across the warp), per-chunk partial sums are folded in chunk order.
__global__ void __launch_bounds__(128) act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y_f32,
                                                     const bf16* __restrict__ pre_bf16, const float* __restrict__ emul,
                                                     const float* __restrict__ colscale, int act, long long M, int N, long long rows_per_chunk,
                                                     bf16* __restrict__ dz, float* __restrict__ db_part, float* __restrict__ dcs_part) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const long long m0 = (long long)blockIdx.y * rows_per_chunk, m1 = min(M, m0 + rows_per_chunk);
  const float cs = colscale ? colscale[n] : 1.f;
  float acc = 0.f, accs = 0.f;
  for (long long m = m0; m < m1; ++m) {
    const long long i = m * N + n;
    float g = dy[i];
    if (emul) g *= emul[i];
    if (dcs_part) accs = fmaf(g, __bfloat162float(pre_bf16[i]), accs);      // act == none with a column scale
    g *= cs;
    if (act == COR_ACT_RELU) {
      g = y_f32[i] > 0.f ? g : 0.f;
    } else if (act == COR_ACT_SIGMOID) {
      const float s = y_f32[i];
      g *= s * (1.f - s);
    } else if (act == COR_ACT_GELU) {
      const float x = __bfloat162float(pre_bf16[i]);
      const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
      g *= cdf + x * 0.3989422804014327f * __expf(-0.5f * x * x);
    }
    dz[i] = __float2bfloat16_rn(g);
    acc += g;
  }
  if (db_part) db_part[(long long)blockIdx.y * N + n] = acc;
  if (dcs_part) dcs_part[(long long)blockIdx.y * N + n] = accs;
}

__global__ void colsum_fold_kernel(const float* __restrict__ part, int nparts, int N, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int k = 0; k < nparts; ++k) s += part[(long long)k * N + n];
  out[n] = s;
}

static int act_chunks(long long M, int N) {
  const long long colblocks = (N + 127) / 128;
  long long want = ((long long)sm_count() * 8 + colblocks - 1) / colblocks;     // ~8 CTAs per SM in total
  long long maxc = (M + 63) / 64;                                                // >= 64 rows per chunk
  if (want > maxc) want = maxc;
  return (int)(want < 1 ? 1 : want);
}

}  // namespace cor

using namespace cor;

extern "C" int cor_cast_cat_bf16(const float* a, int c0, const float* b, int c1, long long rows, void* out_bf16, cor_stream_t stream) {
  COR_REQUIRE(a && out_bf16 && rows > 0 && c0 > 0 && c1 >= 0 && (c1 == 0 || b), "cor_cast_cat_bf16: bad arguments");
  const long long total = rows * (c0 + c1);
  const int blocks = (int)((total + 255) / 256 < (long long)sm_count() * 8 ? (total + 255) / 256 : (long long)sm_count() * 8);
  cast_cat_bf16_kernel<<<blocks, 256, 0, as_stream(stream)>>>(a, c0, b, c1, rows, reinterpret_cast<bf16*>(out_bf16));
  return check_launch("cast_cat_bf16_kernel");
}

extern "C" size_t cor_act_bwd_work_bytes(long long M, int N) { return (size_t)act_chunks(M, N) * N * 2 * sizeof(float) + 16; }

extern "C" int cor_act_bwd(const float* dy, const float* y_f32, const void* pre_bf16, const float* emul, const float* colscale, int act,
                           long long M, int N, void* dz_bf16, float* db, float* dcolscale, void* work, cor_stream_t stream) {
  COR_REQUIRE(dy && dz_bf16 && M > 0 && N > 0, "cor_act_bwd: bad arguments");
  COR_REQUIRE(act >= COR_ACT_NONE && act <= COR_ACT_SIGMOID, "cor_act_bwd: act %d", act);
  COR_REQUIRE(!(act == COR_ACT_RELU || act == COR_ACT_SIGMOID) || y_f32, "cor_act_bwd: relu / sigmoid need the activation output");
  COR_REQUIRE(act != COR_ACT_GELU || pre_bf16, "cor_act_bwd: gelu needs the saved pre-activation");
  COR_REQUIRE(!dcolscale || (colscale && pre_bf16 && act == COR_ACT_NONE), "cor_act_bwd: a column-scale gradient needs colscale, the saved linear output and no activation");
  COR_REQUIRE(!(db || dcolscale) || work, "cor_act_bwd: column sums need the work buffer");
  const int chunks = act_chunks(M, N);
  const long long per = (M + chunks - 1) / chunks;
  float* dbp = db ? reinterpret_cast<float*>(work) : nullptr;
  float* dcp = dcolscale ? reinterpret_cast<float*>(work) + (size_t)chunks * N : nullptr;
  cudaStream_t st = as_stream(stream);
  act_bwd_kernel<<<dim3((N + 127) / 128, chunks), 128, 0, st>>>(dy, y_f32, reinterpret_cast<const bf16*>(pre_bf16), emul, colscale, act, M, N,
                                                                 per, reinterpret_cast<bf16*>(dz_bf16), dbp, dcp);
  int rc = check_launch("act_bwd_kernel");
  if (rc) return rc;
  if (db) { colsum_fold_kernel<<<(N + 127) / 128, 128, 0, st>>>(dbp, chunks, N, db); rc = check_launch("colsum_fold_kernel"); }
  if (!rc && dcolscale) { colsum_fold_kernel<<<(N + 127) / 128, 128, 0, st>>>(dcp, chunks, N, dcolscale); rc = check_launch("colsum_fold_kernel"); }
  return rc;
}
