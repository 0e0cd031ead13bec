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
constexpr int kDwFwdBudget = 54 * 1024;    // shared memory per forward STAGE (two stages per CTA, two CTAs per SM)
constexpr int kDwWgBudget = 104 * 1024;    // weight gradient: per STAGE (two stages, one 512-thread CTA per SM)

// Tile rows are padded on the right to a whole number of output groups (4 outputs forward, 8 in the weight gradient) so
// the inner loops load without bounds predicates; the padding holds zeros.
__host__ __device__ inline int dw_pad(int w, int blk) { return (w + blk - 1) / blk * blk; }
__host__ __device__ inline int dw_band_rows(int h, int w, int wgrad) {
  // rows per band: a forward stage holds (rows + 6) x (pad4(w) + 6) x 32 floats; a weight-gradient stage
  // (rows + 6) x (pad8(w) + 6) + rows x pad8(w)
  int r;
  if (wgrad) r = (kDwWgBudget / (kDwCG * 4) - 2 * kDwR * (dw_pad(w, 8) + 2 * kDwR)) / (2 * dw_pad(w, 8) + 2 * kDwR);
  else r = kDwFwdBudget / ((dw_pad(w, 4) + 2 * kDwR) * kDwCG * 4) - 2 * kDwR;
  if (r > h) r = h;
  if (r < 1) r = 1;
  if (wgrad) {
    // the (row, 8-output chunk) pairs of a band are dealt to 16 warps: prefer the band height (down to half the maximum)
    // that wastes the fewest warp slots, e.g. 24-wide maps: 5 rows x 3 chunks = 15 of 16
    const int xc = dw_pad(w, 8) / 8, warps = 16;
    int best = r;
    float best_eff = 0.f;
    for (int b = r; b >= (r + 1) / 2 && b >= 1; --b) {
      const int full = h / b, last = h - full * b;
      const int slots = full * ((b * xc + warps - 1) / warps) + (last ? (last * xc + warps - 1) / warps : 0);
      const float eff = (float)(h * xc) / (float)(slots * warps);
      if (eff > best_eff + 1e-6f) { best_eff = eff; best = b; }
    }
    r = best;
  }
  return r;
}

// Stage rows [y0 - pad, y0 + rows + pad) x [-pad, tw - pad) of one image's 32-channel slice into shared memory
// ([pixel][32]), zeros outside the image: 8 threads move one pixel's 128 bytes as 16-byte cp.async copies (zero-filled
// outside the image through src-size 0).  Returns at once; the caller commits the group and waits for it one tile later, so
// the loads of tile k+1 fly under the arithmetic of tile k.
template <int THREADS>
__device__ __forceinline__ void stage_tile_async(float* __restrict__ tile, const float* __restrict__ src, int C, int c0, int h, int w, int y0,
                                                 int rows, int pad, int tw) {
  const int total = (rows + 2 * pad) * tw;
  const int sub = threadIdx.x >> 3, ch4 = (threadIdx.x & 7) * 4;
  // (ty, tx) advance incrementally: a division per 16-byte copy made the staging a third of the kernel's instructions
  constexpr int STEP = THREADS / 8;
  const int dty = STEP / tw, dtx = STEP - dty * tw;
  int ty = sub / tw, tx = sub - ty * tw;
  const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(tile + ch4);
  for (int i = sub; i < total; i += STEP) {
    const int gy = y0 + ty - pad, gx = tx - pad;
    const bool in = gy >= 0 && gy < h && gx >= 0 && gx < w;
    const float* g = src + ((long long)(in ? gy : 0) * w + (in ? gx : 0)) * C + c0 + ch4;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + (uint32_t)i * (kDwCG * 4)), "l"(g), "r"(in ? 16 : 0) : "memory");
    tx += dtx;
    ty += dty;
    if (tx >= tw) { tx -= tw; ++ty; }
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// grid = (CTAs per channel group, C / 32): persistent, each CTA walks (image, band) items of its 32 channels with a two-stage
// cp.async ring (the staging was pure exposed load latency when every CTA staged, synchronised, then computed)
template <bool FLIP>
__global__ void __launch_bounds__(kDwThreads) dwconv_cl_kernel(const float* __restrict__ in, const float* __restrict__ wt,
                                                               const float* __restrict__ bias, float* __restrict__ out, int n_img, int h,
                                                               int w, int C, int band) {
  extern __shared__ float tile[];                       // 2 x [(band + 6)][(pad4(w) + 6)][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.y * kDwCG, c = c0 + lane;
  const int xg = (w + 3) / 4;                            // groups of 4 adjacent outputs per row
  const int tw = xg * 4 + 2 * kDwR;
  const int bands = (h + band - 1) / band;
  const int items = n_img * bands;
  const size_t stage_floats = (size_t)(band + 2 * kDwR) * tw * kDwCG;
  const long long img = (long long)h * w * C;
  float wreg[kDwT];
#pragma unroll
  for (int k = 0; k < kDwT; ++k) wreg[k] = wt[(long long)c * kDwT + (FLIP ? kDwT - 1 - k : k)];
  const float b = bias ? bias[c] : 0.f;
  const int rstride = tw * kDwCG;

  int item = blockIdx.x;
  if (item < items) {
    const int n = item / bands, y0 = (item - n * bands) * band;
    stage_tile_async<kDwThreads>(tile, in + n * img, C, c0, h, w, y0, min(band, h - y0), kDwR, tw);
  }
  cp_async_commit();
  for (int k = 0; item < items; ++k, item += gridDim.x) {
    const int nxt = item + gridDim.x;
    if (nxt < items) {
      const int n = nxt / bands, y0 = (nxt - n * bands) * band;
      stage_tile_async<kDwThreads>(tile + ((k + 1) & 1) * stage_floats, in + n * img, C, c0, h, w, y0, min(band, h - y0), kDwR, tw);
    }
    cp_async_commit();
    cp_async_wait<1>();                                  // this item's tile has landed (the next one may still be in flight)
    __syncthreads();
    const float* cur = tile + (k & 1) * stage_floats;
    const int n = item / bands, y0 = (item - n * bands) * band, rows = min(band, h - y0);
    float* dst = out + n * img + (long long)y0 * w * C + c;
    int y = 0, xi = warp;
    while (xi >= xg) { xi -= xg; ++y; }
    while (y < rows) {
      const int x0 = xi * 4;
      float acc[4] = {b, b, b, b};
      const float* row = cur + (y * tw + x0) * kDwCG + lane;
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
    __syncthreads();                                     // every warp is done with this buffer before the next-but-one item refills it
  }
  cp_async_wait<0>();
}

// grid = (CTAs per channel group, C / 32), 512 threads, persistent: partial[blockIdx.x][C][50] = {d w[49], d bias}
// Each CTA walks (image, band) items of its 32 channels through a two-stage cp.async ring (both tiles of item k+1 load under
// the arithmetic of item k) and keeps the 49 tap sums (+ the bias sum) of its lane's channel in registers across ALL its
// items.  Inside an item the (row, 8-output chunk) pairs go round-robin over the 16 warps; for a chunk the 8 gradients and,
// per kernel row, 14 inputs are loaded once and feed 7 x 8 FMAs (3.7 FMAs per shared load).  The warps' sums are folded in
// fixed order through shared memory once, at the end.
constexpr int kDwWgThreads = 512;
__global__ void __launch_bounds__(kDwWgThreads, 1) dwconv_cl_wgrad_kernel(const float* __restrict__ in, const float* __restrict__ dout,
                                                                          float* __restrict__ part, int n_img, int h, int w, int C, int band) {
  extern __shared__ float sm[];                         // 2 x { X tile [(band + 6)][(pad8(w) + 6)][32], dY tile [band][pad8(w)][32] }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.y * kDwCG, c = c0 + lane;
  const int wp = dw_pad(w, 8), tw = wp + 2 * kDwR;
  const int bands = (h + band - 1) / band, items = n_img * bands;
  const size_t x_floats = (size_t)(band + 2 * kDwR) * tw * kDwCG, stage_floats = x_floats + (size_t)band * wp * kDwCG;
  const long long img = (long long)h * w * C;
  float acc[kDwT + 1];
#pragma unroll
  for (int k = 0; k <= kDwT; ++k) acc[k] = 0.f;
  const int xc = wp / 8, rstride = tw * kDwCG;

  auto stage = [&](int item, int buf) {
    const int n = item / bands, y0 = (item - n * bands) * band, rows = min(band, h - y0);
    float* xt = sm + buf * stage_floats;
    stage_tile_async<kDwWgThreads>(xt, in + n * img, C, c0, h, w, y0, rows, kDwR, tw);
    stage_tile_async<kDwWgThreads>(xt + x_floats, dout + n * img, C, c0, h, w, y0, rows, 0, wp);
  };
  int item = blockIdx.x;
  if (item < items) stage(item, 0);
  cp_async_commit();
  for (int k = 0; item < items; ++k, item += gridDim.x) {
    if (item + (int)gridDim.x < items) stage(item + gridDim.x, (k + 1) & 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const float* xt = sm + (k & 1) * stage_floats;
    const float* gt = xt + x_floats;
    const int n = item / bands, y0 = (item - n * bands) * band, rows = min(band, h - y0);
    (void)n;
    for (int wi = warp; wi < rows * xc; wi += kDwWgThreads / 32) {
      const int y = wi / xc, x0 = (wi - y * xc) * 8;
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
    __syncthreads();                                     // the buffer is free for the next-but-one item
  }
  cp_async_wait<0>();
  __syncthreads();
  float* red = sm;                                       // [16 warps][50][32]: the ring's memory becomes the fold scratch
#pragma unroll
  for (int k = 0; k <= kDwT; ++k) red[(warp * (kDwT + 1) + k) * kDwCG + lane] = acc[k];
  __syncthreads();
  float* o = part + ((long long)blockIdx.x * C + c) * (kDwT + 1);
  for (int k = warp; k <= kDwT; k += kDwWgThreads / 32) {
    float sacc = 0.f;
    for (int wv = 0; wv < kDwWgThreads / 32; ++wv) sacc += red[(wv * (kDwT + 1) + k) * kDwCG + lane];
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

static int dw_wgrad_ctas(int n, int h, int w, int C) {       // persistent CTAs per channel group
  const int items = n * ceil_div(h, dw_band_rows(h, w, 1));
  int ctas = sm_count() / (C / kDwCG);
  if (ctas < 1) ctas = 1;
  return ctas < items ? ctas : items;
}

extern "C" size_t cor_dwconv7_work_bytes(int n, int h, int w, int C) {
  if (C <= 0 || C % kDwCG) return 16;
  return (size_t)dw_wgrad_ctas(n, h, w, C) * C * (kDwT + 1) * sizeof(float) + 16;
}

extern "C" int cor_dwconv7_cl(const float* in, const float* weight, const float* bias, float* out, int n, int h, int w, int C, int flip,
                              cor_stream_t stream) {
  COR_REQUIRE(in && weight && out, "cor_dwconv7_cl: null pointer");
  COR_REQUIRE(n > 0 && h > 0 && w > 0 && C > 0 && C % kDwCG == 0, "cor_dwconv7_cl: need C %% 32 == 0 (C=%d)", C);
  const int band = dw_band_rows(h, w, 0);
  const size_t smem = (size_t)2 * (band + 2 * kDwR) * (dw_pad(w, 4) + 2 * kDwR) * kDwCG * sizeof(float);
  COR_REQUIRE(smem <= 220 * 1024, "cor_dwconv7_cl: map too wide (w=%d)", w);
  const int groups = C / kDwCG, items = n * ceil_div(h, band);
  int per_sm = (int)(220 * 1024 / smem);
  if (per_sm > 3) per_sm = 3;
  if (per_sm < 1) per_sm = 1;
  int ctas = ceil_div(sm_count() * per_sm, groups);
  if (ctas > items) ctas = items;
  const dim3 grid(ctas, groups);
  cudaStream_t st = as_stream(stream);
  if (flip) {
    COR_CUDA(cudaFuncSetAttribute(dwconv_cl_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dwconv_cl_kernel<true><<<grid, kDwThreads, smem, st>>>(in, weight, bias, out, n, h, w, C, band);
  } else {
    COR_CUDA(cudaFuncSetAttribute(dwconv_cl_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dwconv_cl_kernel<false><<<grid, kDwThreads, smem, st>>>(in, weight, bias, out, n, h, w, C, band);
  }
  return check_launch("dwconv_cl_kernel");
}

extern "C" int cor_dwconv7_cl_wgrad(const float* in, const float* dout, float* dweight, float* dbias, int n, int h, int w, int C, void* work,
                                    cor_stream_t stream) {
  COR_REQUIRE(in && dout && dweight && work, "cor_dwconv7_cl_wgrad: null pointer");
  COR_REQUIRE(n > 0 && h > 0 && w > 0 && C > 0 && C % kDwCG == 0, "cor_dwconv7_cl_wgrad: need C %% 32 == 0 (C=%d)", C);
  const int band = dw_band_rows(h, w, 1);
  const size_t stage = ((size_t)(band + 2 * kDwR) * (dw_pad(w, 8) + 2 * kDwR) + (size_t)band * dw_pad(w, 8)) * kDwCG * sizeof(float);
  const size_t fold = (size_t)(kDwWgThreads / 32) * (kDwT + 1) * kDwCG * sizeof(float);
  const size_t smem = 2 * stage > fold ? 2 * stage : fold;
  COR_REQUIRE(smem <= 220 * 1024, "cor_dwconv7_cl_wgrad: map too wide (w=%d)", w);
  const int ctas = dw_wgrad_ctas(n, h, w, C);
  const dim3 grid(ctas, C / kDwCG);
  cudaStream_t st = as_stream(stream);
  float* part = reinterpret_cast<float*>(work);
  COR_CUDA(cudaFuncSetAttribute(dwconv_cl_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dwconv_cl_wgrad_kernel<<<grid, kDwWgThreads, smem, st>>>(in, dout, part, n, h, w, C, band);
  int rc = check_launch("dwconv_cl_wgrad_kernel");
  if (rc) return rc;
  const int total = C * (kDwT + 1);
  dwconv_fold_kernel<<<(total + 31) / 32, dim3(32, kFoldTy), 0, st>>>(part, ctas, C, dweight, dbias);
  return check_launch("dwconv_fold_kernel");
}
