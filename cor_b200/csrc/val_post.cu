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

// vailder.py:427-430,466: sigmoid + min-max at the LOGIT resolution first, then bilinear resize of the
// normalised map (cv2.resize INTER_LINEAR == align_corners=False sampling).  Since the tap weights sum to 1,
// resize((s - mn) / den) == (resize(s) - mn) / den: interpolate the sigmoids, normalise afterwards.
template <typename TP>
__device__ __forceinline__ float resized_sigmoid(const TP* __restrict__ p, int H, int W, float sh, float sw, int oy, int ox) {
  int y0, y1, x0, x1;
  float ly0, ly1, lx0, lx1;
  src_index(sh, oy, H, y0, y1, ly0, ly1);
  src_index(sw, ox, W, x0, x1, lx0, lx1);
  const TP* r0 = p + (long long)y0 * W;
  const TP* r1 = p + (long long)y1 * W;
  return ly0 * (lx0 * sigmoid_acc(to_f<TP>(__ldg(r0 + x0))) + lx1 * sigmoid_acc(to_f<TP>(__ldg(r0 + x1)))) +
         ly1 * (lx0 * sigmoid_acc(to_f<TP>(__ldg(r1 + x0))) + lx1 * sigmoid_acc(to_f<TP>(__ldg(r1 + x1))));
}

// All threads of the block reduce the per-chunk (min,max) partials of one sample (a serial loop in one
// thread would put ~chunks L2 round trips in front of every CTA).
__device__ __forceinline__ void block_minmax(const float* __restrict__ part, int chunks, float& s_mn, float& s_mx) {
  __shared__ float wmn[32], wmx[32];
  float mn = INFINITY, mx = -INFINITY;
  for (int c = threadIdx.x; c < chunks; c += blockDim.x) {
    mn = fminf(mn, part[2 * c]);
    mx = fmaxf(mx, part[2 * c + 1]);
  }
  mn = warp_min(mn);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) { wmn[threadIdx.x >> 5] = mn; wmx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int w = 1; w < nw; ++w) { mn = fminf(mn, wmn[w]); mx = fmaxf(mx, wmx[w]); }
    s_mn = mn; s_mx = mx;
  }
  __syncthreads();
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
__global__ void __launch_bounds__(256) val_write_kernel(const TP* __restrict__ pred, int H, int W, int Ho, int Wo, int post_first,
                                                        const float* __restrict__ mm_part, int chunks, float* __restrict__ post,
                                                        uint8_t* __restrict__ hard, const TG* __restrict__ gt, float gscale,
                                                        double* __restrict__ met_part) {
  __shared__ double scratch[4 * 32];
  __shared__ float s_mn, s_mx;
  const int n = blockIdx.y;
  block_minmax(mm_part + (long long)n * chunks * 2, chunks, s_mn, s_mx);
  const float mn = s_mn, den = s_mx - s_mn + 1e-8f;
  const TP* p = pred + (long long)n * H * W;
  const float sh = (float)H / (float)Ho, sw = (float)W / (float)Wo;
  const bool same = H == Ho && W == Wo;
  const long long total = (long long)Ho * Wo;
  double acc[4] = {0, 0, 0, 0};
  float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const int oy = (int)(o / Wo), ox = (int)(o % Wo);
    const float sg = (post_first && !same) ? resized_sigmoid<TP>(p, H, W, sh, sw, oy, ox) : upsampled_sigmoid<TP>(p, H, W, sh, sw, oy, ox, same);
    const float v = (sg - mn) / den;
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
__global__ void val_metrics_kernel(const double* __restrict__ met_part, int N, int chunks, double total, double s,
                                   float* __restrict__ metrics) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double pg = 0, sp = 0, sg = 0, ad = 0;
  for (int c = 0; c < chunks; ++c) {
    const double* o = met_part + ((long long)n * chunks + c) * 4;
    pg += o[0]; sp += o[1]; sg += o[2]; ad += o[3];
  }
  const double bp = total - sp, bg = total - sg, bpg = total - sp - sg + pg;   // sums of (1-p), (1-g), (1-p)(1-g)
  const double dice = (2 * pg + s) / (sp + sg + s), dice_b = (2 * bpg + s) / (bp + bg + s);
  const double iou = (pg + s) / (sp + sg - pg + s), iou_b = (bpg + s) / (bp + bg - bpg + s);
  float* m = metrics + (long long)n * 5;
  m[0] = (float)dice; m[1] = (float)(ad / total); m[2] = (float)iou; m[3] = (float)((dice + dice_b) / 2); m[4] = (float)((iou + iou_b) / 2);
}

// ---- exact 4x up-sampling fast path (the shipped 256^2 -> 1024^2 case) ------------------------------------
// For an integer factor 4 the align_corners=False source coordinate of output 4*sx+j is sx + (j+.5)/4 - .5:
// outputs j=0,1 blend source columns (sx-1, sx) with weights (.375,.625) / (.125,.875), outputs j=2,3 blend
// (sx, sx+1) with (.875,.125) / (.625,.375); index clamping at the borders reproduces ATen's src<0 -> 0 and
// x1 = x0 rules exactly.  One thread turns a 3x3 source neighbourhood into a 4x4 output block: 9 loads, 16
// results, 16-byte coalesced stores.
// 1/(1+exp(-z)) on the SFU (MUFU.EX2 + MUFU.RCP): ~2 ulp, monotone, used for BOTH the min/max and the map
__device__ __forceinline__ float sigmoid_fast(float z) { return __fdividef(1.f, 1.f + __expf(-z)); }

template <typename TP>
__device__ __forceinline__ void up4_block(const TP* __restrict__ p, int H, int W, int sy, int sx, float (&o)[4][4]) {
  const int ym = max(sy - 1, 0), yp = min(sy + 1, H - 1), xm = max(sx - 1, 0), xp = min(sx + 1, W - 1);
  const int ys[3] = {ym, sy, yp};
  float hrow[3][4];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const TP* q = p + (long long)ys[r] * W;
    const float a = to_f<TP>(__ldg(q + xm)), b = to_f<TP>(__ldg(q + sx)), c = to_f<TP>(__ldg(q + xp));
    hrow[r][0] = 0.625f * b + 0.375f * a;     // lambda0 * v[x0] + lambda1 * v[x1] with x0 = sx-1: (1-.625)*a + .625*b
    hrow[r][1] = 0.875f * b + 0.125f * a;
    hrow[r][2] = 0.875f * b + 0.125f * c;
    hrow[r][3] = 0.625f * b + 0.375f * c;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    o[0][j] = 0.625f * hrow[1][j] + 0.375f * hrow[0][j];
    o[1][j] = 0.875f * hrow[1][j] + 0.125f * hrow[0][j];
    o[2][j] = 0.875f * hrow[1][j] + 0.125f * hrow[2][j];
    o[3][j] = 0.625f * hrow[1][j] + 0.375f * hrow[2][j];
  }
}

template <typename TP>
__global__ void __launch_bounds__(256) val_up4_minmax_kernel(const TP* __restrict__ pred, int H, int W, float* __restrict__ part) {
  __shared__ float smn[8], smx[8];
  const int n = blockIdx.y;
  const TP* p = pred + (long long)n * H * W;
  float mn = INFINITY, mx = -INFINITY;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
    float o[4][4];
    up4_block<TP>(p, H, W, i / W, i % W, o);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) { mn = fminf(mn, o[a][b]); mx = fmaxf(mx, o[a][b]); }
  }
  mn = warp_min(mn);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { mn = fminf(mn, smn[w]); mx = fmaxf(mx, smx[w]); }
    float* o = part + ((long long)n * gridDim.x + blockIdx.x) * 2;
    // sigmoid is monotone: min / max of sigmoid(z) are sigmoid(min z) / sigmoid(max z)
    o[0] = sigmoid_fast(mn); o[1] = sigmoid_fast(mx);
  }
}

template <typename TP, typename TG>
__global__ void __launch_bounds__(256) val_up4_write_kernel(const TP* __restrict__ pred, int H, int W, const float* __restrict__ mm_part,
                                                            int chunks, float* __restrict__ post, uint8_t* __restrict__ hard,
                                                            const TG* __restrict__ gt, float gscale, double* __restrict__ met_part) {
  __shared__ double scratch[4 * 32];
  __shared__ float s_mn, s_mx;
  const int n = blockIdx.y;
  block_minmax(mm_part + (long long)n * chunks * 2, chunks, s_mn, s_mx);
  const float mn = s_mn, inv_den = 1.f / (s_mx - s_mn + 1e-8f);
  const TP* p = pred + (long long)n * H * W;
  const int Wo = 4 * W;
  const long long obase = (long long)n * 16 * H * W;
  float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
    const int sy = i / W, sx = i % W;
    float o[4][4];
    up4_block<TP>(p, H, W, sy, sx, o);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      float v[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) v[b] = (sigmoid_fast(o[a][b]) - mn) * inv_den;
      const long long off = obase + (long long)(4 * sy + a) * Wo + 4 * sx;
      if (post) *reinterpret_cast<float4*>(post + off) = make_float4(v[0], v[1], v[2], v[3]);
      if (hard) *reinterpret_cast<uchar4*>(hard + off) = make_uchar4(v[0] > 0.5f ? 255 : 0, v[1] > 0.5f ? 255 : 0, v[2] > 0.5f ? 255 : 0, v[3] > 0.5f ? 255 : 0);
      if (gt) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const float g = to_f<TG>(gt[off + b]) * gscale;
          f0 = fmaf(v[b], g, f0); f1 += v[b]; f2 += g; f3 += fabsf(v[b] - g);
        }
      }
    }
  }
  if (gt) {
    double acc[4] = {f0, f1, f2, f3};
    block_sum<4>(acc, scratch);
    if (threadIdx.x == 0) {
      double* o = met_part + ((long long)n * gridDim.x + blockIdx.x) * 4;
      o[0] = acc[0]; o[1] = acc[1]; o[2] = acc[2]; o[3] = acc[3];
    }
  }
}

// ---- stand-alone soft metrics (compute_dice / mae / iou / mdice / miou, utils/trainer_v3_g.py:381-443) ----------
// One pass over (pred, gt): {sum p*g, sum p, sum g, sum |p-g|} per sample -> the five metrics.  HBM-bound.
template <typename TG>
__global__ void __launch_bounds__(256) soft_metrics_kernel(const float* __restrict__ pred, const TG* __restrict__ gt, float gscale,
                                                           long long total, double* __restrict__ met_part) {
  __shared__ double scratch[4 * 32];
  const int n = blockIdx.y;
  const float* p = pred + (long long)n * total;
  const TG* g = gt + (long long)n * total;
  float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const float v = p[o], t = to_f<TG>(g[o]) * gscale;
    f0 = fmaf(v, t, f0); f1 += v; f2 += t; f3 += fabsf(v - t);
  }
  double acc[4] = {f0, f1, f2, f3};
  block_sum<4>(acc, scratch);
  if (threadIdx.x == 0) {
    double* o = met_part + ((long long)n * gridDim.x + blockIdx.x) * 4;
    o[0] = acc[0]; o[1] = acc[1]; o[2] = acc[2]; o[3] = acc[3];
  }
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

extern "C" int cor_val_post(const void* pred, int pred_dtype, int N, int H, int W, int Ho, int Wo, int post_first, float* post, uint8_t* hard,
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
  const bool up4 = !post_first && Ho == 4 * H && Wo == 4 * W && (!post || (((uintptr_t)post) & 15) == 0) && (!hard || (((uintptr_t)hard) & 3) == 0);
  if (up4) {
    int c4 = ceil_div((long long)H * W, 256);
    if (c4 > 1024) c4 = 1024;
    dim3 g4(c4, N);
    if (pred_dtype == COR_F32) val_up4_minmax_kernel<float><<<g4, 256, 0, st>>>((const float*)pred, H, W, mm_part);
    else if (pred_dtype == COR_BF16) val_up4_minmax_kernel<bf16><<<g4, 256, 0, st>>>((const bf16*)pred, H, W, mm_part);
    else COR_REQUIRE(false, "cor_val_post: unsupported pred dtype %d", pred_dtype);
    int rc4 = check_launch("val_up4_minmax_kernel");
    if (rc4) return rc4;
#define COR_VW4(TP, TG) val_up4_write_kernel<TP, TG><<<g4, 256, 0, st>>>((const TP*)pred, H, W, mm_part, c4, post, hard, (const TG*)gt, gt_scale, met_part)
    if (pred_dtype == COR_F32 && (!gt || gt_dtype == COR_F32)) COR_VW4(float, float);
    else if (pred_dtype == COR_BF16 && (!gt || gt_dtype == COR_F32)) COR_VW4(bf16, float);
    else if (pred_dtype == COR_F32 && gt_dtype == COR_U8) COR_VW4(float, uint8_t);
    else if (pred_dtype == COR_BF16 && gt_dtype == COR_U8) COR_VW4(bf16, uint8_t);
    else COR_REQUIRE(false, "cor_val_post: unsupported gt dtype %d", gt_dtype);
#undef COR_VW4
    rc4 = check_launch("val_up4_write_kernel");
    if (rc4 || !gt) return rc4;
    val_metrics_kernel<<<ceil_div(N, 128), 128, 0, st>>>(met_part, N, c4, (double)total, 1e-5, metrics);
    return check_launch("val_metrics_kernel");
  }
  // post_first: the min / max are those of the sigmoid at the logit resolution
  const int Hmm = post_first ? H : Ho, Wmm = post_first ? W : Wo;
  if (pred_dtype == COR_F32) val_minmax_kernel<float><<<grid, 256, 0, st>>>((const float*)pred, H, W, Hmm, Wmm, mm_part);
  else if (pred_dtype == COR_BF16) val_minmax_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)pred, H, W, Hmm, Wmm, mm_part);
  else COR_REQUIRE(false, "cor_val_post: unsupported pred dtype %d", pred_dtype);
  int rc = check_launch("val_minmax_kernel");
  if (rc) return rc;
#define COR_VW(TP, TG) val_write_kernel<TP, TG><<<grid, 256, 0, st>>>((const TP*)pred, H, W, Ho, Wo, post_first, mm_part, chunks, post, hard, (const TG*)gt, gt_scale, met_part)
  if (pred_dtype == COR_F32 && (!gt || gt_dtype == COR_F32)) COR_VW(float, float);
  else if (pred_dtype == COR_BF16 && (!gt || gt_dtype == COR_F32)) COR_VW(bf16, float);
  else if (pred_dtype == COR_F32 && gt_dtype == COR_U8) COR_VW(float, uint8_t);
  else if (pred_dtype == COR_BF16 && gt_dtype == COR_U8) COR_VW(bf16, uint8_t);
  else COR_REQUIRE(false, "cor_val_post: unsupported gt dtype %d", gt_dtype);
#undef COR_VW
  rc = check_launch("val_write_kernel");
  if (rc || !gt) return rc;
  val_metrics_kernel<<<ceil_div(N, 128), 128, 0, st>>>(met_part, N, chunks, (double)total, 1e-5, metrics);
  return check_launch("val_metrics_kernel");
}

extern "C" int cor_soft_metrics(const float* pred, const void* gt, int gt_dtype, float gt_scale, int N, long long total, float smooth,
                                float* metrics, void* work, cor_stream_t stream) {
  COR_REQUIRE(pred && gt && metrics && work, "cor_soft_metrics: null pointer");
  COR_REQUIRE(N > 0 && N <= 65535 && total > 0, "cor_soft_metrics: bad shape");
  cudaStream_t st = as_stream(stream);
  const int chunks = val_chunks(N, total);
  double* met_part = reinterpret_cast<double*>(work);
  dim3 grid(chunks, N);
  if (gt_dtype == COR_F32) soft_metrics_kernel<float><<<grid, 256, 0, st>>>(pred, (const float*)gt, gt_scale, total, met_part);
  else if (gt_dtype == COR_U8) soft_metrics_kernel<uint8_t><<<grid, 256, 0, st>>>(pred, (const uint8_t*)gt, gt_scale, total, met_part);
  else COR_REQUIRE(false, "cor_soft_metrics: unsupported gt dtype %d", gt_dtype);
  int rc = check_launch("soft_metrics_kernel");
  if (rc) return rc;
  val_metrics_kernel<<<ceil_div(N, 128), 128, 0, st>>>(met_part, N, chunks, (double)total, (double)smooth, metrics);
  return check_launch("val_metrics_kernel");
}
