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

// the same with 4 adjacent columns per thread (c0 % 4 == 0, c1 % 4 == 0, 16-byte aligned rows): 128-bit loads, 64-bit stores
__global__ void __launch_bounds__(256) cast_cat_bf16_vec_kernel(const float* __restrict__ a, int c0, const float* __restrict__ b, int c1,
                                                               long long rows, bf16* __restrict__ out) {
  const int Cq = (c0 + c1) / 4;
  const long long total = rows * Cq;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r;
    int c;
    if (c1 == 0) { r = 0; c = 0; }
    else { r = i / Cq; c = (int)(i - r * Cq) * 4; }
    const float4 v = c1 == 0 ? reinterpret_cast<const float4*>(a)[i]
                             : (c < c0 ? *reinterpret_cast<const float4*>(a + r * c0 + c) : *reinterpret_cast<const float4*>(b + r * c1 + (c - c0)));
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    reinterpret_cast<uint2*>(out)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
}

// out[b][r][:] = bf16(src[b][r][:]) for r < rows, 0 for rows <= r < rows_padded: a K-padded MN-major GEMM operand (the zero rows
// make whatever the other operand holds at those k indices irrelevant).  4 columns per thread.
__global__ void __launch_bounds__(256) cast_pad_rows_bf16_kernel(const float* __restrict__ src, int rows, int rows_padded, int cols4,
                                                                long long total, bf16* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / cols4;
    const int c = (int)(i - row * cols4);
    const long long b = row / rows_padded;
    const int r = (int)(row - b * rows_padded);
    uint2 o = make_uint2(0u, 0u);
    if (r < rows) {
      const float4 v = reinterpret_cast<const float4*>(src)[(b * rows + r) * cols4 + c];
      __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
      o = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
    reinterpret_cast<uint2*>(out)[i] = o;
  }
}

// dz[m][n] = dy[m][n] * emul[m][n] * colscale[n] * act'(.)  (bf16, the A operand of dX = dZ W and dW = dZ^T X);
// db[n] = sum_m dz (f32); dcs[n] = sum_m dy[m][n] * emul * act(pre)[m][n] (gradient of the column scale, ConvNeXt's gamma).
// y = the activation's OUTPUT before mask / scale (relu, sigmoid) or pre = its pre-activation (gelu; also the un-scaled
// linear output when a column scale is used), saved bf16 by the GEMM epilogue.  grid = (column blocks, row chunks): a thread
// walks its column over the chunk's rows (coalesced across the warp), per-chunk partial sums are folded in chunk order.
template <typename TD>
__global__ void __launch_bounds__(128) act_bwd_kernel(const TD* __restrict__ dy, const float* __restrict__ y_f32,
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
    float g = to_f<TD>(dy[i]);
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
      g *= gelu_grad_fast(x);
    }
    dz[i] = __float2bfloat16_rn(g);
    acc += g;
  }
  if (db_part) db_part[(long long)blockIdx.y * N + n] = acc;
  if (dcs_part) dcs_part[(long long)blockIdx.y * N + n] = accs;
}

template <typename TD>
__device__ __forceinline__ void load4(const TD* p, float (&v)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void load4<bf16>(const bf16* p, float (&v)[4]) {
  const uint2 t = *reinterpret_cast<const uint2*>(p);
  v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
  v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
}

// The same with 4 adjacent columns per thread and four rows in flight (N % 4 == 0, 16-byte aligned tensors): 128-bit /
// 64-bit accesses, the activation a template parameter.  Per-column sums are per thread, so the summation order over rows
// is the scalar kernel's.
template <typename TD, int ACT>
__global__ void __launch_bounds__(128) act_bwd_vec_kernel(const TD* __restrict__ dy, const float* __restrict__ y_f32,
                                                         const bf16* __restrict__ pre_bf16, const float* __restrict__ emul,
                                                         const float* __restrict__ colscale, long long M, int N, long long rows_per_chunk,
                                                         bf16* __restrict__ dz, float* __restrict__ db_part, float* __restrict__ dcs_part) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (n >= N) return;
  const long long m0 = (long long)blockIdx.y * rows_per_chunk, m1 = min(M, m0 + rows_per_chunk);
  float cs[4] = {1.f, 1.f, 1.f, 1.f};
  if (colscale) load4<float>(colscale + n, cs);
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, accs[4] = {0.f, 0.f, 0.f, 0.f};
  // four rows in flight per thread (A/B on B200: a two-stage software pipeline of 2 + 2 rows measured 5 % slower, 292 vs 278 us
  // at 147 456 x 1024 -- with ~6 warps per scheduler the other warps already cover a warp's load phase)
  constexpr int U = 4;
  for (long long m = m0; m < m1; m += U) {
    float g[U][4], x[U][4], e[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (m + u < m1) {
        const long long i = (m + u) * N + n;
        load4<TD>(dy + i, g[u]);
        if (ACT == COR_ACT_GELU || dcs_part) load4<bf16>(pre_bf16 + i, x[u]);
        if (ACT == COR_ACT_RELU || ACT == COR_ACT_SIGMOID) load4<float>(y_f32 + i, x[u]);
        if (emul) load4<float>(emul + i, e[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (m + u < m1) {
        const long long i = (m + u) * N + n;
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float gv = g[u][k];
          if (emul) gv *= e[u][k];
          if (dcs_part) accs[k] = fmaf(gv, x[u][k], accs[k]);
          gv *= cs[k];
          if (ACT == COR_ACT_RELU) gv = x[u][k] > 0.f ? gv : 0.f;
          else if (ACT == COR_ACT_SIGMOID) gv *= x[u][k] * (1.f - x[u][k]);
          else if (ACT == COR_ACT_GELU) gv *= gelu_grad_fast(x[u][k]);
          o[k] = gv;
          acc[k] += gv;
        }
        __nv_bfloat162 lo = __floats2bfloat162_rn(o[0], o[1]), hi = __floats2bfloat162_rn(o[2], o[3]);
        *reinterpret_cast<uint2*>(dz + i) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
      }
    }
  }
  if (db_part) *reinterpret_cast<float4*>(db_part + (long long)blockIdx.y * N + n) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  if (dcs_part) *reinterpret_cast<float4*>(dcs_part + (long long)blockIdx.y * N + n) = make_float4(accs[0], accs[1], accs[2], accs[3]);
}

// block (32, kFoldTy), grid = ceil(N / 32)
__global__ void colsum_fold_kernel(const float* __restrict__ part, int nparts, int N, float* __restrict__ out) {
  __shared__ float sm[kFoldTy][32];
  const int n = blockIdx.x * 32 + threadIdx.x;
  const float t = fold_parts(part, nparts, N, n, n < N, sm);
  if (threadIdx.y == 0 && n < N) out[n] = t;
}

static bool act_vec_ok(int N) { return N % 4 == 0; }
static int act_chunks(long long M, int N) {
  const long long colblocks = act_vec_ok(N) ? (N / 4 + 127) / 128 : (N + 127) / 128;
  long long want = ((long long)sm_count() * 12 + colblocks - 1) / colblocks;    // ~12 CTAs per SM in total
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
  const bool vec = c0 % 4 == 0 && c1 % 4 == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out_bf16)) & 15) == 0;
  if (vec) {
    const long long tv = total / 4;
    const int vb = (int)((tv + 255) / 256 < (long long)sm_count() * 16 ? (tv + 255) / 256 : (long long)sm_count() * 16);
    cast_cat_bf16_vec_kernel<<<vb, 256, 0, as_stream(stream)>>>(a, c0, b, c1, rows, reinterpret_cast<bf16*>(out_bf16));
    return check_launch("cast_cat_bf16_vec_kernel");
  }
  cast_cat_bf16_kernel<<<blocks, 256, 0, as_stream(stream)>>>(a, c0, b, c1, rows, reinterpret_cast<bf16*>(out_bf16));
  return check_launch("cast_cat_bf16_kernel");
}

extern "C" int cor_cast_pad_rows_bf16(const float* src, int batch, int rows, int rows_padded, int cols, void* out_bf16, cor_stream_t stream) {
  COR_REQUIRE(src && out_bf16, "cor_cast_pad_rows_bf16: null pointer");
  COR_REQUIRE(batch > 0 && rows > 0 && rows_padded >= rows && cols > 0 && cols % 4 == 0, "cor_cast_pad_rows_bf16: bad shape (rows=%d padded=%d cols=%d)",
              rows, rows_padded, cols);
  COR_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(out_bf16)) & 15) == 0, "cor_cast_pad_rows_bf16: 16-byte alignment required");
  const long long total = (long long)batch * rows_padded * (cols / 4);
  const int blocks = (int)((total + 255) / 256 < (long long)sm_count() * 8 ? (total + 255) / 256 : (long long)sm_count() * 8);
  cast_pad_rows_bf16_kernel<<<blocks, 256, 0, as_stream(stream)>>>(src, rows, rows_padded, cols / 4, total, reinterpret_cast<bf16*>(out_bf16));
  return check_launch("cast_pad_rows_bf16_kernel");
}

extern "C" size_t cor_act_bwd_work_bytes(long long M, int N) { return (size_t)act_chunks(M, N) * N * 2 * sizeof(float) + 16; }

extern "C" int cor_act_bwd(const void* dy, int dy_dtype, const float* y_f32, const void* pre_bf16, const float* emul, const float* colscale, int act,
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
  COR_REQUIRE(dy_dtype == COR_F32 || dy_dtype == COR_BF16, "cor_act_bwd: gradient dtype %d", dy_dtype);
  const bf16* pre = reinterpret_cast<const bf16*>(pre_bf16);
  bf16* dz = reinterpret_cast<bf16*>(dz_bf16);
  const uintptr_t al = reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(y_f32) | reinterpret_cast<uintptr_t>(pre_bf16) |
                       reinterpret_cast<uintptr_t>(emul) | reinterpret_cast<uintptr_t>(colscale) | reinterpret_cast<uintptr_t>(dz_bf16) |
                       reinterpret_cast<uintptr_t>(work);
  if (act_vec_ok(N) && (al & 15) == 0) {
    const dim3 grid((N / 4 + 127) / 128, chunks);
#define COR_ACT_VEC(TD, ACT) \
  act_bwd_vec_kernel<TD, ACT><<<grid, 128, 0, st>>>(reinterpret_cast<const TD*>(dy), y_f32, pre, emul, colscale, M, N, per, dz, dbp, dcp)
#define COR_ACT_VEC_T(TD)                                            \
  switch (act) {                                                     \
    case COR_ACT_RELU: COR_ACT_VEC(TD, COR_ACT_RELU); break;         \
    case COR_ACT_GELU: COR_ACT_VEC(TD, COR_ACT_GELU); break;         \
    case COR_ACT_SIGMOID: COR_ACT_VEC(TD, COR_ACT_SIGMOID); break;   \
    default: COR_ACT_VEC(TD, COR_ACT_NONE); break;                   \
  }
    if (dy_dtype == COR_F32) { COR_ACT_VEC_T(float) } else { COR_ACT_VEC_T(bf16) }
#undef COR_ACT_VEC_T
#undef COR_ACT_VEC
  } else {
    const dim3 grid((N + 127) / 128, chunks);
    if (dy_dtype == COR_F32)
      act_bwd_kernel<float><<<grid, 128, 0, st>>>(reinterpret_cast<const float*>(dy), y_f32, pre, emul, colscale, act, M, N, per, dz, dbp, dcp);
    else
      act_bwd_kernel<bf16><<<grid, 128, 0, st>>>(reinterpret_cast<const bf16*>(dy), y_f32, pre, emul, colscale, act, M, N, per, dz, dbp, dcp);
  }
  int rc = check_launch("act_bwd_kernel");
  if (rc) return rc;
  if (db) { colsum_fold_kernel<<<(N + 31) / 32, dim3(32, kFoldTy), 0, st>>>(dbp, chunks, N, db); rc = check_launch("colsum_fold_kernel"); }
  if (!rc && dcolscale) { colsum_fold_kernel<<<(N + 31) / 32, dim3(32, kFoldTy), 0, st>>>(dcp, chunks, N, dcolscale); rc = check_launch("colsum_fold_kernel"); }
  return rc;
}
