// Validation post-process (utils/trainer_v3_g.py:226-231, utils/vailder.py:427-430,473) and the soft
// metrics of utils/trainer_v3_g.py:381-443, fused:
//   pass 1: bilinear-upsample the logits (align_corners=False), sigmoid, per-sample min / max
//   pass 2: recompute, min-max stretch (+1e-8), write the fp32 map and/or the binarised (>0.5)*255 map,
//           accumulate the metric sums against the ground truth.
// The 256x256 logits stay L2-resident across both passes; the only HBM-scale traffic is the compulsory
// write of the 1024x1024 outputs and the read of the ground-truth mask.
#include "common.cuh"

namespace cor {

template <typename TP>
__device__ __forceinline__ float upsampled_sigmoid(const TP* __restrict__ p, int H, int W, float sh, float sw, int oy, int ox,
                                                   bool same) {
  float z;
  if (same) {
    z = to_f<TP>(__ldg(p + (long long)oy * W + ox));
  } else {
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    src_index(sh, oy, H, y0, y1, ly0, ly1);
    src_index(sw, ox, W, x0, x1, lx0, lx1);
    const TP* r0 = p + (long long)y0 * W;
    const TP* r1 = p + (long long)y1 * W;
    z = ly0 * (lx0 * to_f<TP>(__ldg(r0 + x0)) + lx1 * to_f<TP>(__ldg(r0 + x1))) +
        ly1 * (lx0 * to_f<TP>(__ldg(r1 + x0)) + lx1 * to_f<TP>(__ldg(r1 + x1)));
  }
  return sigmoid_acc(z);
}

// grid = (chunks, N); minmax_part [N][chunks][2]
template <typename TP>
__global__ void __launch_bounds__(256) val_minmax_kernel(const TP* __restrict__ pred, int H, int W, int Ho, int Wo,
                                                         float* __restrict__ part) {
  __shared__ float smn[8], smx[8];
  const int n = blockIdx.y;
  const TP* p = pred + (long long)n * H * W;
  const float sh = (float)H / (float)Ho, sw = (float)W / (float)Wo;
  const bool same = H == Ho && W == Wo;
  float mn = INFINITY, mx = -INFINITY;
  const long long total = (long long)Ho * Wo;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const float v = upsampled_sigmoid<TP>(p, H, W, sh, sw, (int)(o / Wo), (int)(o % Wo), same);
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  mn = warp_min(mn);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { mn = fminf(mn, smn[w]); mx = fmaxf(mx, smx[w]); }
    float* o = part + ((long long)n * gridDim.x + blockIdx.x) * 2;
    o[0] = mn; o[1] = mx;
  }
}

// metric partial sums per CTA: {sum p*g, sum p, sum g, sum |p-g|}
template <typename TP, typename TG>
__global__ void __launch_bounds__(256) val_write_kernel(const TP* __restrict__ pred, int H, int W, int Ho, int Wo,
                                                        const float* __restrict__ mm_part, int chunks, float* __restrict__ post,
                                                        uint8_t* __restrict__ hard, const TG* __restrict__ gt, float gscale,
                                                        double* __restrict__ met_part) {
  __shared__ double scratch[4 * 32];
  __shared__ float s_mn, s_mx;
  const int n = blockIdx.y;
  if (threadIdx.x == 0) {
    float mn = INFINITY, mx = -INFINITY;
    for (int c = 0; c < chunks; ++c) {
      mn = fminf(mn, mm_part[((long long)n * chunks + c) * 2]);
      mx = fmaxf(mx, mm_part[((long long)n * chunks + c) * 2 + 1]);
    }
    s_mn = mn; s_mx = mx;
  }
  __syncthreads();
  const float mn = s_mn, den = s_mx - s_mn + 1e-8f;
  const TP* p = pred + (long long)n * H * W;
  const float sh = (float)H / (float)Ho, sw = (float)W / (float)Wo;
  const bool same = H == Ho && W == Wo;
  const long long total = (long long)Ho * Wo;
  double acc[4] = {0, 0, 0, 0};
  float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const float v = (upsampled_sigmoid<TP>(p, H, W, sh, sw, (int)(o / Wo), (int)(o % Wo), same) - mn) / den;
    const long long g_o = (long long)n * total + o;
    if (post) post[g_o] = v;
    if (hard) hard[g_o] = v > 0.5f ? 255 : 0;
    if (gt) {
      const float g = to_f<TG>(gt[g_o]) * gscale;
      f0 = fmaf(v, g, f0); f1 += v; f2 += g; f3 += fabsf(v - g);
    }
  }
  if (gt) {
    acc[0] = f0; acc[1] = f1; acc[2] = f2; acc[3] = f3;
    block_sum<4>(acc, scratch);
    if (threadIdx.x == 0) {
      double* o = met_part + ((long long)n * gridDim.x + blockIdx.x) * 4;
      o[0] = acc[0]; o[1] = acc[1]; o[2] = acc[2]; o[3] = acc[3];
    }
  }
}

// metrics [N][5] = {dice, mae, iou, mdice, miou}, smooth = 1e-5 (trainer_v3_g.py:381-443)
__global__ void val_metrics_kernel(const double* __restrict__ met_part, int N, int chunks, double total, float* __restrict__ metrics) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double pg = 0, sp = 0, sg = 0, ad = 0;
  for (int c = 0; c < chunks; ++c) {
    const double* o = met_part + ((long long)n * chunks + c) * 4;
    pg += o[0]; sp += o[1]; sg += o[2]; ad += o[3];
  }
  const double s = 1e-5;
  const double bp = total - sp, bg = total - sg, bpg = total - sp - sg + pg;   // sums of (1-p), (1-g), (1-p)(1-g)
  const double dice = (2 * pg + s) / (sp + sg + s), dice_b = (2 * bpg + s) / (bp + bg + s);
  const double iou = (pg + s) / (sp + sg - pg + s), iou_b = (bpg + s) / (bp + bg - bpg + s);
  float* m = metrics + (long long)n * 5;
  m[0] = (float)dice; m[1] = (float)(ad / total); m[2] = (float)iou; m[3] = (float)((dice + dice_b) / 2); m[4] = (float)((iou + iou_b) / 2);
}

static int val_chunks(int N, long long total) {
  long long c = ((long long)sm_count() * 8 + N - 1) / N;
  const long long cap = (total + 1023) / 1024;
  if (c > cap) c = cap;
  if (c < 1) c = 1;
  if (c > 1024) c = 1024;
  return (int)c;
}

}  // namespace cor

using namespace cor;

extern "C" size_t cor_val_post_work_bytes(int N, int Ho, int Wo) {
  (void)Ho; (void)Wo;
  return (size_t)N * 1024 * (2 * sizeof(float) + 4 * sizeof(double)) + 64;
}

extern "C" int cor_val_post(const void* pred, int pred_dtype, int N, int H, int W, int Ho, int Wo, float* post, uint8_t* hard,
                            const void* gt, int gt_dtype, float gt_scale, float* metrics, void* work, cor_stream_t stream) {
  COR_REQUIRE(pred && work && (post || hard || (gt && metrics)), "cor_val_post: nothing to do / null pointer");
  COR_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0 && Ho > 0 && Wo > 0, "cor_val_post: bad shape");
  COR_REQUIRE(!gt || metrics, "cor_val_post: gt given without a metrics buffer");
  cudaStream_t st = as_stream(stream);
  const long long total = (long long)Ho * Wo;
  const int chunks = val_chunks(N, total);
  double* met_part = reinterpret_cast<double*>(work);
  float* mm_part = reinterpret_cast<float*>(met_part + (size_t)N * 1024 * 4);
  dim3 grid(chunks, N);
  if (pred_dtype == COR_F32) val_minmax_kernel<float><<<grid, 256, 0, st>>>((const float*)pred, H, W, Ho, Wo, mm_part);
  else if (pred_dtype == COR_BF16) val_minmax_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)pred, H, W, Ho, Wo, mm_part);
  else COR_REQUIRE(false, "cor_val_post: unsupported pred dtype %d", pred_dtype);
  int rc = check_launch("val_minmax_kernel");
  if (rc) return rc;
#define COR_VW(TP, TG) val_write_kernel<TP, TG><<<grid, 256, 0, st>>>((const TP*)pred, H, W, Ho, Wo, mm_part, chunks, post, hard, (const TG*)gt, gt_scale, met_part)
  if (pred_dtype == COR_F32 && (!gt || gt_dtype == COR_F32)) COR_VW(float, float);
  else if (pred_dtype == COR_BF16 && (!gt || gt_dtype == COR_F32)) COR_VW(bf16, float);
  else if (pred_dtype == COR_F32 && gt_dtype == COR_U8) COR_VW(float, uint8_t);
  else if (pred_dtype == COR_BF16 && gt_dtype == COR_U8) COR_VW(bf16, uint8_t);
  else COR_REQUIRE(false, "cor_val_post: unsupported gt dtype %d", gt_dtype);
#undef COR_VW
  rc = check_launch("val_write_kernel");
  if (rc || !gt) return rc;
  val_metrics_kernel<<<ceil_div(N, 128), 128, 0, st>>>(met_part, N, chunks, (double)total, metrics);
  return check_launch("val_metrics_kernel");
}
