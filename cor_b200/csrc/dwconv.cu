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
constexpr int kDwSmemBudget = 160 * 1024;

__host__ __device__ inline int dw_band_rows(int h, int w, int tiles) {
  // rows per band so that `tiles` padded tiles ((rows + 6) x (w + 6) x 32 floats each) fit the budget
  const int per_row = (w + 2 * kDwR) * kDwCG * 4 * tiles;
  int r = kDwSmemBudget / per_row - 2 * kDwR;
  if (r > h) r = h;
  return r < 1 ? 1 : r;
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
  const int tw = w + 2 * kDwR, th = rows + 2 * kDwR;
  const float* src = in + (long long)n * h * w * C;
  for (int i = warp; i < th * tw; i += kDwThreads / 32) {
    const int ty = i / tw, tx = i % tw;
    const int gy = y0 + ty - kDwR, gx = tx - kDwR;
    tile[i * kDwCG + lane] = (gy >= 0 && gy < h && gx >= 0 && gx < w) ? src[((long long)gy * w + gx) * C + c] : 0.f;
  }
  float wreg[kDwT];
#pragma unroll
  for (int k = 0; k < kDwT; ++k) wreg[k] = wt[(long long)c * kDwT + (FLIP ? kDwT - 1 - k : k)];
  const float b = bias ? bias[c] : 0.f;
  __syncthreads();
  const int xg = (w + 3) / 4;                            // groups of 4 adjacent outputs per row
  float* dst = out + (long long)n * h * w * C;
  for (int item = warp; item < rows * xg; item += kDwThreads / 32) {
    const int y = item / xg, x0 = (item % xg) * 4;
    float acc[4] = {b, b, b, b};
#pragma unroll
    for (int dy = 0; dy < kDwK; ++dy) {
      const float* row = tile + ((y + dy) * tw + x0) * kDwCG + lane;
      float v[10];
#pragma unroll
      for (int j = 0; j < 10; ++j) v[j] = (x0 + j < tw) ? row[j * kDwCG] : 0.f;
#pragma unroll
      for (int dx = 0; dx < kDwK; ++dx) {
        const float ww = wreg[dy * kDwK + dx];
#pragma unroll
        for (int o = 0; o < 4; ++o) acc[o] = fmaf(ww, v[o + dx], acc[o]);
      }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o)
      if (x0 + o < w) dst[((long long)(y0 + y) * w + x0 + o) * C + c] = acc[o];
  }
}

// grid = (bands, C / 32, n): partial[(n * bands + band)][C][50] = {d w[49], d bias}
__global__ void __launch_bounds__(kDwThreads) dwconv_cl_wgrad_kernel(const float* __restrict__ in, const float* __restrict__ dout,
                                                                     float* __restrict__ part, int h, int w, int C, int band) {
  extern __shared__ float sm[];                         // X tile [(rows + 6)][(w + 6)][32] then dY tile [rows][w][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.y * kDwCG + lane;
  const int n = blockIdx.z;
  const int y0 = blockIdx.x * band, rows = min(band, h - y0);
  const int tw = w + 2 * kDwR, th = rows + 2 * kDwR;
  float* xt = sm;
  float* gt = sm + (size_t)(band + 2 * kDwR) * tw * kDwCG;
  const float* src = in + (long long)n * h * w * C;
  const float* gsrc = dout + (long long)n * h * w * C;
  for (int i = warp; i < th * tw; i += kDwThreads / 32) {
    const int ty = i / tw, tx = i % tw;
    const int gy = y0 + ty - kDwR, gx = tx - kDwR;
    xt[i * kDwCG + lane] = (gy >= 0 && gy < h && gx >= 0 && gx < w) ? src[((long long)gy * w + gx) * C + c] : 0.f;
  }
  for (int i = warp; i < rows * w; i += kDwThreads / 32) gt[i * kDwCG + lane] = gsrc[((long long)(y0 + i / w) * w + i % w) * C + c];
  __syncthreads();
  float* o = part + (((long long)n * gridDim.x + blockIdx.x) * C + c) * (kDwT + 1);
  // warp w takes taps w, w + 8, ...; tap 49 is the bias (sum of d out)
  for (int k = warp; k <= kDwT; k += kDwThreads / 32) {
    float acc = 0.f;
    if (k < kDwT) {
      const int dy = k / kDwK, dx = k % kDwK;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;      // four independent chains: the loop is FMA-latency bound otherwise
      for (int y = 0; y < rows; ++y) {
        const float* gr = gt + (size_t)y * w * kDwCG + lane;
        const float* xr = xt + ((size_t)(y + dy) * tw + dx) * kDwCG + lane;
        int x = 0;
        for (; x + 4 <= w; x += 4) {
          a0 = fmaf(gr[(x + 0) * kDwCG], xr[(x + 0) * kDwCG], a0);
          a1 = fmaf(gr[(x + 1) * kDwCG], xr[(x + 1) * kDwCG], a1);
          a2 = fmaf(gr[(x + 2) * kDwCG], xr[(x + 2) * kDwCG], a2);
          a3 = fmaf(gr[(x + 3) * kDwCG], xr[(x + 3) * kDwCG], a3);
        }
        for (; x < w; ++x) a0 = fmaf(gr[x * kDwCG], xr[x * kDwCG], a0);
      }
      acc = (a0 + a1) + (a2 + a3);
    } else {
      for (int i = 0; i < rows * w; ++i) acc += gt[i * kDwCG + lane];
    }
    o[k] = acc;
  }
}

// d w[c][k] = sum over (image, band) partials in order; d bias likewise
__global__ void dwconv_fold_kernel(const float* __restrict__ part, int nparts, int C, float* __restrict__ dw, float* __restrict__ db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * (kDwT + 1)) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += part[(long long)p * C * (kDwT + 1) + i];
  const int c = i / (kDwT + 1), k = i % (kDwT + 1);
  if (k < kDwT) dw[c * kDwT + k] = s;
  else if (db) db[c] = s;
}

}  // namespace cor

using namespace cor;

extern "C" size_t cor_dwconv7_work_bytes(int n, int h, int w, int C) {
  const int bands = ceil_div(h, dw_band_rows(h, w, 2));
  return (size_t)n * bands * C * (kDwT + 1) * sizeof(float) + 16;
}

extern "C" int cor_dwconv7_cl(const float* in, const float* weight, const float* bias, float* out, int n, int h, int w, int C, int flip,
                              cor_stream_t stream) {
  COR_REQUIRE(in && weight && out, "cor_dwconv7_cl: null pointer");
  COR_REQUIRE(n > 0 && h > 0 && w > 0 && C > 0 && C % kDwCG == 0, "cor_dwconv7_cl: need C %% 32 == 0 (C=%d)", C);
  const int band = dw_band_rows(h, w, 1);
  const size_t smem = (size_t)(band + 2 * kDwR) * (w + 2 * kDwR) * kDwCG * sizeof(float);
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
  const int band = dw_band_rows(h, w, 2);
  const size_t smem = ((size_t)(band + 2 * kDwR) * (w + 2 * kDwR) + (size_t)band * w) * kDwCG * sizeof(float);
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
  dwconv_fold_kernel<<<(total + 255) / 256, 256, 0, st>>>(part, n * bands, C, dweight, dbias);
  return check_launch("dwconv_fold_kernel");
}
