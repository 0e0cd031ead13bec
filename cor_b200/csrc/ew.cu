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

// dz[m][n] = dy[m][n] * emul[m][n] * act'(.)  (bf16, the A operand of dX = dZ W and dW = dZ^T X); db[n] = sum_m dz (f32,
// summed in row order: deterministic).  y = the activation's OUTPUT before the dropout mask (relu / sigmoid) or its
// pre-activation (gelu, saved bf16 by the GEMM epilogue).  One thread per column walks the rows.
__global__ void __launch_bounds__(128) act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y_f32,
                                                     const bf16* __restrict__ pre_bf16, const float* __restrict__ emul, int act, int M,
                                                     int N, bf16* __restrict__ dz, float* __restrict__ db) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float acc = 0.f;
  for (int m = 0; m < M; ++m) {
    const long long i = (long long)m * N + n;
    float g = dy[i];
    if (emul) g *= emul[i];
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
    const bf16 q = __float2bfloat16_rn(g);
    dz[i] = q;
    acc += g;
  }
  if (db) db[n] = acc;
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

extern "C" int cor_act_bwd(const float* dy, const float* y_f32, const void* pre_bf16, const float* emul, int act, int M, int N,
                           void* dz_bf16, float* db, cor_stream_t stream) {
  COR_REQUIRE(dy && dz_bf16 && M > 0 && N > 0, "cor_act_bwd: bad arguments");
  COR_REQUIRE(act >= COR_ACT_NONE && act <= COR_ACT_SIGMOID, "cor_act_bwd: act %d", act);
  COR_REQUIRE(!(act == COR_ACT_RELU || act == COR_ACT_SIGMOID) || y_f32, "cor_act_bwd: relu / sigmoid need the activation output");
  COR_REQUIRE(act != COR_ACT_GELU || pre_bf16, "cor_act_bwd: gelu needs the saved pre-activation");
  act_bwd_kernel<<<(N + 127) / 128, 128, 0, as_stream(stream)>>>(dy, y_f32, reinterpret_cast<const bf16*>(pre_bf16), emul, act, M, N,
                                                               reinterpret_cast<bf16*>(dz_bf16), db);
  return check_launch("act_bwd_kernel");
}
