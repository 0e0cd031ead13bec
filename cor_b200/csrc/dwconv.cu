// Depth-wise 7x7 convolution on channels-last maps, forward and backward: the spatial mixing step of the ConvNeXt blocks
// in GenerateMaskAdapterMap (lib/support_model/mask_adapter.py:196-199, nn.Conv2d(dim, dim, 7, padding=3, groups=dim)).
//   out[n,y,x,c] = bias[c] + sum_{dy,dx} w[c][dy][dx] * in[n, y+dy-3, x+dx-3, c]          (zero padding)
// Channels-last ([n][h][w][C]) because every other step of the block works on [pixels][C] rows (LayerNorm, the point-wise
// GEMMs): no layout change anywhere.  A CTA owns 32 channels of a band of rows of one image: the band + a 3-pixel zero
// halo is staged in shared memory once ([pixel][32 channels]: a warp's 32 lanes = 32 channels = 32 banks, conflict-free,
// and 128-byte coalesced global reads), each lane keeps its channel's 49 weights in registers and produces 4 adjacent
// outputs per pass (10 shared loads per kernel row instead of 28).
//   d in  = the same kernel on d out with the weights flipped;
//   d w[c][k], d bias[c] = per-(image, band) partial sums from a second kernel (both tiles in shared memory), folded in
//   fixed order.
#include "common.cuh"

namespace cor {

constexpr int kDwK = 7, kDwR = 3, kDwT = kDwK * kDwK;
constexpr int kDwCG = 32;                 // channels per CTA
constexpr int kDwThreads = 256;
constexpr int kDwFwdBudget = 56 * 1024;    // shared memory per CTA: four forward CTAs / two weight-gradient CTAs per SM, so that
constexpr int kDwWgBudget = 104 * 1024;    // one CTA's staging (pure load latency) runs under another's arithmetic

// Tile rows are padded on the right to a whole number of output groups (4 outputs forward, 8 in the weight gradient) so
// the inner loops load without bounds predicates; the padding holds zeros.
__host__ __device__ inline int dw_pad(int w, int blk) { return (w + blk - 1) / blk * blk; }
__host__ __device__ inline int dw_band_rows(int h, int w, int wgrad) {
  // rows per band: forward needs (rows + 6) x (pad4(w) + 6) x 32 floats, the weight gradient (rows + 6) x (pad8(w) + 6)
  // + rows x pad8(w)
  int r;
  if (wgrad) r = (kDwWgBudget / (kDwCG * 4) - 2 * kDwR * (dw_pad(w, 8) + 2 * kDwR)) / (2 * dw_pad(w, 8) + 2 * kDwR);
  else r = kDwFwdBudget / ((dw_pad(w, 4) + 2 * kDwR) * kDwCG * 4) - 2 * kDwR;
  if (r > h) r = h;
  return r < 1 ? 1 : r;
}

// Stage rows [y0 - pad, y0 + rows + pad) x [-pad, w + pad) of one image's 32-channel slice into shared memory
// ([pixel][32]), zeros outside the image.  All 256 threads take part: 8 threads move one pixel's 128 bytes as float4s,
// 32 pixels per pass, 4 passes in flight (16 KB per CTA): the staging is pure load latency.
__device__ __forceinline__ void stage_tile(float* __restrict__ tile, const float* __restrict__ src, int C, int c0, int h, int w, int y0,
                                           int rows, int pad, int tw) {      // tw = tile row stride in pixels, >= w + 2 pad
  const int total = (rows + 2 * pad) * tw;
  const int sub = threadIdx.x >> 3, ch4 = (threadIdx.x & 7) * 4;
  constexpr int U = 4, PP = kDwThreads / 8;
  for (int i0 = sub; i0 < total; i0 += U * PP) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * PP;
      const int ty = i / tw, tx = i - ty * tw;
      const int gy = y0 + ty - pad, gx = tx - pad;
      v[u] = (i < total && gy >= 0 && gy < h && gx >= 0 && gx < w)
                 ? __ldg(reinterpret_cast<const float4*>(src + ((long long)gy * w + gx) * C + c0 + ch4))
                 : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * PP;
      if (i < total) *reinterpret_cast<float4*>(tile + (size_t)i * kDwCG + ch4) = v[u];
    }
  }
}

// grid = (bands, C / 32, n)
template <bool FLIP>
__global__ void __launch_bounds__(kDwThreads) dwconv_cl_kernel(const float* __restrict__ in, const float* __restrict__ wt,
                                                               const float* __restrict__ bias, float* __restrict__ out, int h, int w, int C,
                                                               int band) {
  extern __shared__ float tile[];                       // [(rows + 6)][(w + 6)][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.y * kDwCG + lane;
  const int n = blockIdx.z;
  const int y0 = blockIdx.x * band, rows = min(band, h - y0);
  const int xg = (w + 3) / 4;                            // groups of 4 adjacent outputs per row
  const int tw = xg * 4 + 2 * kDwR;
  const float* src = in + (long long)n * h * w * C;
  stage_tile(tile, src, C, blockIdx.y * kDwCG, h, w, y0, rows, kDwR, tw);
  float wreg[kDwT];
#pragma unroll
  for (int k = 0; k < kDwT; ++k) wreg[k] = wt[(long long)c * kDwT + (FLIP ? kDwT - 1 - k : k)];
  const float b = bias ? bias[c] : 0.f;
  __syncthreads();
  float* dst = out + ((long long)n * h + y0) * w * C + c;
  const int rstride = tw * kDwCG;
  int y = 0, xi = warp;
  while (xi >= xg) { xi -= xg; ++y; }
  while (y < rows) {
    const int x0 = xi * 4;
    float acc[4] = {b, b, b, b};
    const float* row = tile + (y * tw + x0) * kDwCG + lane;
#pragma unroll
    for (int dy = 0; dy < kDwK; ++dy, row += rstride) {
      float v[10];
#pragma unroll
      for (int j = 0; j < 10; ++j) v[j] = row[j * kDwCG];
#pragma unroll
      for (int dx = 0; dx < kDwK; ++dx) {
        const float ww = wreg[dy * kDwK + dx];
#pragma unroll
        for (int o = 0; o < 4; ++o) acc[o] = fmaf(ww, v[o + dx], acc[o]);
      }
    }
    float* d = dst + ((long long)y * w + x0) * C;
#pragma unroll
    for (int o = 0; o < 4; ++o)
      if (x0 + o < w) d[(long long)o * C] = acc[o];
    xi += kDwThreads / 32;
    while (xi >= xg) { xi -= xg; ++y; }
  }
}

// grid = (bands, C / 32, n): partial[(n * bands + band)][C][50] = {d w[49], d bias}
// Warp w takes output rows y = w, w + 8, ...; lane = channel keeps all 49 tap sums (+ the bias sum) in registers.  For a
// chunk of 8 adjacent outputs the 8 gradients and, per kernel row, 14 inputs are loaded once and feed 7 x 8 FMAs each
// (3.7 FMAs per shared load; one tap per warp with two loads per FMA made the kernel shared-memory bound).  The 8 warps'
// sums are folded in fixed order through shared memory.
__global__ void __launch_bounds__(kDwThreads) dwconv_cl_wgrad_kernel(const float* __restrict__ in, const float* __restrict__ dout,
                                                                     float* __restrict__ part, int h, int w, int C, int band) {
  extern __shared__ float sm[];                         // X tile [(rows + 6)][(w + 6)][32] then dY tile [rows][w][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.y * kDwCG + lane;
  const int n = blockIdx.z;
  const int y0 = blockIdx.x * band, rows = min(band, h - y0);
  const int wp = dw_pad(w, 8), tw = wp + 2 * kDwR;
  float* xt = sm;
  float* gt = sm + (size_t)(band + 2 * kDwR) * tw * kDwCG;
  stage_tile(xt, in + (long long)n * h * w * C, C, blockIdx.y * kDwCG, h, w, y0, rows, kDwR, tw);
  stage_tile(gt, dout + (long long)n * h * w * C, C, blockIdx.y * kDwCG, h, w, y0, rows, 0, wp);
  __syncthreads();
  float acc[kDwT + 1];
#pragma unroll
  for (int k = 0; k <= kDwT; ++k) acc[k] = 0.f;
  // (row, 8-output chunk) items round-robin over the warps: balanced for any band height
  const int xc = wp / 8, rstride = tw * kDwCG;
  for (int item = warp; item < rows * xc; item += kDwThreads / 32) {
    {
      const int y = item / xc, x0 = (item - y * xc) * 8;
      float g[8];
      const float* gr = gt + ((size_t)y * wp + x0) * kDwCG + lane;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        g[j] = gr[j * kDwCG];
        acc[kDwT] += g[j];
      }
      const float* xr = xt + ((size_t)y * tw + x0) * kDwCG + lane;
#pragma unroll
      for (int dy = 0; dy < kDwK; ++dy, xr += rstride) {
        float v[14];
#pragma unroll
        for (int j = 0; j < 14; ++j) v[j] = xr[j * kDwCG];
#pragma unroll
        for (int dx = 0; dx < kDwK; ++dx)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[dy * kDwK + dx] = fmaf(g[j], v[j + dx], acc[dy * kDwK + dx]);
      }
    }
  }
  __syncthreads();                                       // tiles consumed: their memory becomes the fold scratch
  float* red = sm;                                       // [8 warps][50][32]
#pragma unroll
  for (int k = 0; k <= kDwT; ++k) red[(warp * (kDwT + 1) + k) * kDwCG + lane] = acc[k];
  __syncthreads();
  float* o = part + (((long long)n * gridDim.x + blockIdx.x) * C + c) * (kDwT + 1);
  for (int k = warp; k <= kDwT; k += kDwThreads / 32) {
    float sacc = 0.f;
    for (int wv = 0; wv < kDwThreads / 32; ++wv) sacc += red[(wv * (kDwT + 1) + k) * kDwCG + lane];
    o[k] = sacc;
  }
}

// d w[c][k] = sum over (image, band) partials in order; d bias likewise
__global__ void dwconv_fold_kernel(const float* __restrict__ part, int nparts, int C, float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float sm[kFoldTy][32];
  const int total = C * (kDwT + 1);
  const int i = blockIdx.x * 32 + threadIdx.x;
  const float s = fold_parts(part, nparts, total, i, i < total, sm);
  if (threadIdx.y != 0 || i >= total) return;
  const int c = i / (kDwT + 1), k = i % (kDwT + 1);
  if (k < kDwT) dw[c * kDwT + k] = s;
  else if (db) db[c] = s;
}

}  // namespace cor

using namespace cor;

extern "C" size_t cor_dwconv7_work_bytes(int n, int h, int w, int C) {
  const int bands = ceil_div(h, dw_band_rows(h, w, 1));
  return (size_t)n * bands * C * (kDwT + 1) * sizeof(float) + 16;
}

extern "C" int cor_dwconv7_cl(const float* in, const float* weight, const float* bias, float* out, int n, int h, int w, int C, int flip,
                              cor_stream_t stream) {
  COR_REQUIRE(in && weight && out, "cor_dwconv7_cl: null pointer");
  COR_REQUIRE(n > 0 && h > 0 && w > 0 && C > 0 && C % kDwCG == 0, "cor_dwconv7_cl: need C %% 32 == 0 (C=%d)", C);
  const int band = dw_band_rows(h, w, 0);
  const size_t smem = (size_t)(band + 2 * kDwR) * (dw_pad(w, 4) + 2 * kDwR) * kDwCG * sizeof(float);
  COR_REQUIRE(smem <= 200 * 1024, "cor_dwconv7_cl: map too wide (w=%d)", w);
  const dim3 grid(ceil_div(h, band), C / kDwCG, n);
  cudaStream_t st = as_stream(stream);
  if (flip) {
    COR_CUDA(cudaFuncSetAttribute(dwconv_cl_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dwconv_cl_kernel<true><<<grid, kDwThreads, smem, st>>>(in, weight, bias, out, h, w, C, band);
  } else {
    COR_CUDA(cudaFuncSetAttribute(dwconv_cl_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dwconv_cl_kernel<false><<<grid, kDwThreads, smem, st>>>(in, weight, bias, out, h, w, C, band);
  }
  return check_launch("dwconv_cl_kernel");
}

extern "C" int cor_dwconv7_cl_wgrad(const float* in, const float* dout, float* dweight, float* dbias, int n, int h, int w, int C, void* work,
                                    cor_stream_t stream) {
  COR_REQUIRE(in && dout && dweight && work, "cor_dwconv7_cl_wgrad: null pointer");
  COR_REQUIRE(n > 0 && h > 0 && w > 0 && C > 0 && C % kDwCG == 0, "cor_dwconv7_cl_wgrad: need C %% 32 == 0 (C=%d)", C);
  const int band = dw_band_rows(h, w, 1);
  const size_t smem = ((size_t)(band + 2 * kDwR) * (dw_pad(w, 8) + 2 * kDwR) + (size_t)band * dw_pad(w, 8)) * kDwCG * sizeof(float);
  COR_REQUIRE(smem <= 220 * 1024, "cor_dwconv7_cl_wgrad: map too wide (w=%d)", w);
  const int bands = ceil_div(h, band);
  const dim3 grid(bands, C / kDwCG, n);
  cudaStream_t st = as_stream(stream);
  float* part = reinterpret_cast<float*>(work);
  COR_CUDA(cudaFuncSetAttribute(dwconv_cl_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dwconv_cl_wgrad_kernel<<<grid, kDwThreads, smem, st>>>(in, dout, part, h, w, C, band);
  int rc = check_launch("dwconv_cl_wgrad_kernel");
  if (rc) return rc;
  const int total = C * (kDwT + 1);
  dwconv_fold_kernel<<<(total + 31) / 32, dim3(32, kFoldTy), 0, st>>>(part, n * bands, C, dweight, dbias);
  return check_launch("dwconv_fold_kernel");
}
