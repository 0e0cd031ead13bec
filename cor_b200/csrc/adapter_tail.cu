// Pooling tail of MaskAdapterPooling on the tensor cores (lib/support_model/mask_adapter.py:62-79):
//   out[b,q,:] = mean_{j<G} softmax_p(logsigmoid(maps[b,qG+j,:])) @ feat[b]^T ,  softmax(logsigmoid(x)) = sigmoid(x) / sum_p sigmoid(x)
// The mean over the G maps of a mask and the per-map normalisation are linear in the pooled sums, so they fold into ONE
// weight row per mask,
//   Wq[b,q,p] = (1/G) sum_j s[b,qG+j,p] / den[b,qG+j],   s = sigmoid(maps), den = sum_p s,
// and the pooling becomes the batched GEMM out[b] = Wq[b] feat[b]^T on csrc/gemm_umma.cu (bf16 operands, fp32 accumulate) --
// G times fewer rows than pooling every map, and no CUDA-core contraction (the streaming kernel is compute-bound above ~8
// rows).  This file holds the two element-wise ends: forming Wq (forward) and taking d Wq back to d maps (backward).
#include "common.cuh"

namespace cor {

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + __expf(-x)); }

// grid = (Q, B), 256 threads; wq [B][Qp][P] bf16 (rows q >= Q are the caller's zeros), den [B*R]
__global__ void __launch_bounds__(256) adapter_tail_weights_kernel(const float* __restrict__ maps, int R, int P, int G, int Qp,
                                                                  bf16* __restrict__ wq, float* __restrict__ den) {
  __shared__ float scratch[32];
  __shared__ float inv[64];
  const int q = blockIdx.x, b = blockIdx.y;
  const float* m0 = maps + ((long long)b * R + (long long)q * G) * P;
  for (int j = 0; j < G; ++j) {
    float s[1] = {0.f};
    for (int p = threadIdx.x; p < P; p += blockDim.x) s[0] += sigmoid_f(m0[(long long)j * P + p]);
    block_sum(s, scratch);
    if (threadIdx.x == 0) {
      den[(long long)b * R + q * G + j] = s[0];
      inv[j] = 1.f / ((float)G * s[0]);
    }
  }
  __syncthreads();
  bf16* o = wq + ((long long)b * Qp + q) * P;
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < G; ++j) acc = fmaf(sigmoid_f(m0[(long long)j * P + p]), inv[j], acc);
    o[p] = __float2bfloat16_rn(acc);
  }
}

// d maps[b,r,p] = (gwq[b,q,p] - dot_r / den_r) / (G den_r) * s (1 - s),  dot_r = sum_p gwq[b,q,p] s[b,r,p];  grid = (R, B)
__global__ void __launch_bounds__(256) adapter_tail_bwd_kernel(const float* __restrict__ maps, const float* __restrict__ den,
                                                              const float* __restrict__ gwq, int R, int P, int G, int Q,
                                                              float* __restrict__ gmaps) {
  __shared__ float scratch[32];
  const int r = blockIdx.x, b = blockIdx.y, q = r / G;
  const float* m = maps + ((long long)b * R + r) * P;
  const float* g = gwq + ((long long)b * Q + q) * P;
  float d[1] = {0.f};
  for (int p = threadIdx.x; p < P; p += blockDim.x) d[0] = fmaf(g[p], sigmoid_f(m[p]), d[0]);
  block_sum(d, scratch);
  const float dn = den[(long long)b * R + r];
  const float k = 1.f / ((float)G * dn), mean = d[0] / dn;
  float* o = gmaps + ((long long)b * R + r) * P;
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const float s = sigmoid_f(m[p]);
    o[p] = (g[p] - mean) * k * s * (1.f - s);
  }
}

}  // namespace cor

using namespace cor;

extern "C" int cor_adapter_tail_weights(const float* maps, int B, int R, int P, int G, int Qp, void* wq_bf16, float* den, cor_stream_t stream) {
  COR_REQUIRE(maps && wq_bf16 && den, "cor_adapter_tail_weights: null pointer");
  COR_REQUIRE(B > 0 && R > 0 && P > 0 && G > 0 && G <= 64 && R % G == 0 && Qp >= R / G, "cor_adapter_tail_weights: bad shape (R=%d G=%d Qp=%d)", R, G, Qp);
  adapter_tail_weights_kernel<<<dim3(R / G, B), 256, 0, as_stream(stream)>>>(maps, R, P, G, Qp, reinterpret_cast<bf16*>(wq_bf16), den);
  return check_launch("adapter_tail_weights_kernel");
}

extern "C" int cor_adapter_tail_bwd(const float* maps, const float* den, const float* gwq, int B, int R, int P, int G, float* gmaps,
                                    cor_stream_t stream) {
  COR_REQUIRE(maps && den && gwq && gmaps, "cor_adapter_tail_bwd: null pointer");
  COR_REQUIRE(B > 0 && R > 0 && P > 0 && G > 0 && R % G == 0, "cor_adapter_tail_bwd: bad shape (R=%d G=%d)", R, G);
  adapter_tail_bwd_kernel<<<dim3(R, B), 256, 0, as_stream(stream)>>>(maps, den, gwq, R, P, G, R / G, gmaps);
  return check_launch("adapter_tail_bwd_kernel");
}
