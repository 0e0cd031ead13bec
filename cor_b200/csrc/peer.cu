// Region-row exchange between the GPUs of one node over NVLink peer memory.
//
// One process per GPU.  Every rank owns a "symmetric" allocation (plain cudaMalloc, exported with CUDA IPC and mapped
// by its peers), holding
//     pub   [n_local, C] bf16   the rank's normalised region rows      (written by cor_rows_finalize)
//     gall  [world*n_local, C] f32   d loss / d (all gathered rows)    (written by cor_infonce_bwd)
//     flags [2 channels][2 phases][kPeerMaxWorld] u32                  (written by the PEERS, polled locally)
// and two kernels replace the NCCL pair of the step:
//     peer_gather_kernel   all[p*n : (p+1)*n] = pub of rank p                 (all-gather, pulled over NVLink)
//     peer_reduce_kernel   out = sum_p gall_p[rank*n : (rank+1)*n], p ascending  (reduce-scatter, fixed order -> the
//                          result is bit-identical on every run)
// Both are "enter barrier -> pull -> exit barrier" on monotonically increasing epochs kept in device memory, so the
// launches carry no per-step host state and replay inside a CUDA graph.  enter(e): "what I produced for step e is
// in my buffer" (stream order put the producer kernel before this one); exit(e): "I have finished reading yours" -
// a rank leaves the kernel only after every peer has signalled exit(e), hence the next step's producer kernel may
// overwrite the buffer.  Waits are bounded by COR_PEER_TIMEOUT_S of globaltimer (default 600 s, 0 = wait for ever):
// on expiry the kernel does NOT trap (that would take the whole CUDA context with it, where the NCCL pair this
// replaces would simply have blocked through a slow rank's checkpoint save) -- it records the epoch in the local
// error word (state[kErrWord]) and carries on; the host raises CorError from PeerExchange.check().
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace cor {

constexpr int kPeerMaxWorld = 16;
constexpr int kPeerThreads = 512;
constexpr int kPeerCtas = 64;
constexpr int kStateWords = 4;
constexpr int kErrWord = 2 * kStateWords;       // state[8]: 0, or 0x80000000 | epoch of the first wait that expired

// flags block of one rank: [channel][phase][source rank]
__device__ __forceinline__ unsigned* flag_slot(unsigned* flags, int channel, int phase, int src) {
  return flags + ((channel * 2 + phase) * kPeerMaxWorld + src);
}

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_peer16(const void* p) {
  uint4 r;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

struct PeerCtl {
  unsigned* const* flags;   // device array [world]: every rank's flag block (entry `rank` is the local one)
  unsigned* state;          // local: [channel] {epoch done, CTA counter, epoch signalled, -}, then the error word
  int rank, world, channel;
  unsigned long long timeout_ns;   // 0 = no limit
};

__device__ __forceinline__ void wait_flag(const unsigned* p, unsigned epoch, const PeerCtl& c) {
  const unsigned long long t0 = now_ns();
  while ((int)(ld_acquire_sys(p) - epoch) < 0) {
    __nanosleep(64);
    if (c.timeout_ns && now_ns() - t0 > c.timeout_ns) {
      atomicCAS(&c.state[kErrWord], 0u, 0x80000000u | epoch);     // first failure wins; the host reports it
      return;
    }
  }
}

// "What I produced for my next epoch on this channel is in my buffer": one tiny CTA right after the producer kernel,
// so that whatever the stream runs between this and the exchange kernel hides the skew between the ranks.
__global__ void peer_signal_kernel(PeerCtl c) {
  unsigned* st = c.state + kStateWords * c.channel;
  const unsigned epoch = st[2] + 1u;
  if (threadIdx.x < c.world) {
    __threadfence_system();
    st_release_sys(flag_slot(c.flags[threadIdx.x], c.channel, 0, c.rank), epoch);
  }
  __syncthreads();
  if (threadIdx.x == 0) st[2] = epoch;
}

// Before the producer overwrites the buffer: every peer has finished reading what the last exchange published.
__global__ void peer_wait_exit_kernel(PeerCtl c) {
  const unsigned done = c.state[kStateWords * c.channel];
  if (threadIdx.x < c.world) wait_flag(flag_slot(c.flags[c.rank], c.channel, 1, threadIdx.x), done, c);
}

// Every CTA waits until all peers have signalled the epoch this launch serves.
__device__ __forceinline__ unsigned peer_enter(const PeerCtl& c) {
  const unsigned epoch = c.state[kStateWords * c.channel] + 1u;
  if (threadIdx.x < c.world) wait_flag(flag_slot(c.flags[c.rank], c.channel, 0, threadIdx.x), epoch, c);
  __syncthreads();
  return epoch;
}

// The last CTA to get here tells the peers "rank has finished reading" (nobody waits for it here: see
// peer_wait_exit_kernel) and publishes the new epoch for the next launch.
__device__ __forceinline__ void peer_exit(const PeerCtl& c, unsigned epoch) {
  __shared__ int last;
  unsigned* st = c.state + kStateWords * c.channel;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(&st[1], 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  if (threadIdx.x < c.world) {
    __threadfence_system();
    st_release_sys(flag_slot(c.flags[threadIdx.x], c.channel, 1, c.rank), epoch);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    st[1] = 0u;
    st[0] = epoch;
  }
}

// all[p] = src[p] for every rank p; `vecs` 16-byte vectors per rank.
__global__ void __launch_bounds__(kPeerThreads) peer_gather_kernel(const uint4* const* __restrict__ src, uint4* __restrict__ all,
                                                                   long long vecs, PeerCtl c) {
  const unsigned epoch = peer_enter(c);
  const long long total = vecs * c.world, stride = (long long)gridDim.x * blockDim.x;
  constexpr int U = 4;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += stride * U) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < total) {
        // start with the rank's own slice and walk the ring, so the ranks do not all pull from rank 0 first
        const int k = (int)(i / vecs);
        const int p = (c.rank + k) % c.world;
        v[u] = ld_peer16(src[p] + (i - (long long)k * vecs));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < total) {
        const int k = (int)(i / vecs);
        const int p = (c.rank + k) % c.world;
        all[(long long)p * vecs + (i - (long long)k * vecs)] = v[u];
      }
    }
  }
  peer_exit(c, epoch);
}

// out[i] = sum over ranks p (ascending) of src[p][rank*vecs + i]; float4 vectors.
template <int W>
__device__ __forceinline__ void reduce_body(const uint4* const* __restrict__ src, float4* __restrict__ out, long long vecs, int rank,
                                            int world) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < vecs; i += stride) {
    uint4 v[W];
#pragma unroll
    for (int p = 0; p < W; ++p)
      if (p < world) v[p] = ld_peer16(src[p] + (long long)rank * vecs + i);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p = 0; p < W; ++p)
      if (p < world) {
        a.x += __uint_as_float(v[p].x);
        a.y += __uint_as_float(v[p].y);
        a.z += __uint_as_float(v[p].z);
        a.w += __uint_as_float(v[p].w);
      }
    out[i] = a;
  }
}

__global__ void __launch_bounds__(kPeerThreads) peer_reduce_kernel(const uint4* const* __restrict__ src, float4* __restrict__ out,
                                                                   long long vecs, PeerCtl c) {
  const unsigned epoch = peer_enter(c);
  if (c.world <= 2) reduce_body<2>(src, out, vecs, c.rank, c.world);
  else if (c.world <= 4) reduce_body<4>(src, out, vecs, c.rank, c.world);
  else if (c.world <= 8) reduce_body<8>(src, out, vecs, c.rank, c.world);
  else reduce_body<kPeerMaxWorld>(src, out, vecs, c.rank, c.world);
  peer_exit(c, epoch);
}

static int peer_grid(long long vecs) {
  long long g = (vecs + kPeerThreads - 1) / kPeerThreads;
  if (g < 1) g = 1;
  return (int)(g < kPeerCtas ? g : kPeerCtas);
}

}  // namespace cor

using namespace cor;

extern "C" {

int cor_peer_max_world(void) { return kPeerMaxWorld; }
size_t cor_peer_flag_bytes(void) { return sizeof(unsigned) * 2 * 2 * kPeerMaxWorld; }
size_t cor_peer_state_bytes(void) { return sizeof(unsigned) * (2 * kStateWords + 4); }
int cor_peer_error_word(void) { return kErrWord; }

int cor_peer_alloc(int device, size_t bytes, void** ptr) {
  COR_REQUIRE(ptr && bytes > 0, "cor_peer_alloc: bad arguments");
  int prev = 0;
  COR_CUDA(cudaGetDevice(&prev));
  COR_CUDA(cudaSetDevice(device));
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e == cudaSuccess) e = cudaMemset(*ptr, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaSetDevice(prev);
  if (e != cudaSuccess) {
    set_error("cor_peer_alloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
    return COR_ECUDA;
  }
  return COR_OK;
}

int cor_peer_free(void* ptr) {
  if (ptr) COR_CUDA(cudaFree(ptr));
  return COR_OK;
}

int cor_peer_export(void* ptr, unsigned char* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  COR_REQUIRE(ptr && handle64, "cor_peer_export: bad arguments");
  cudaIpcMemHandle_t h;
  COR_CUDA(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle64, &h, 64);
  return COR_OK;
}

int cor_peer_open(int device, const unsigned char* handle64, void** ptr) {
  COR_REQUIRE(ptr && handle64, "cor_peer_open: bad arguments");
  int prev = 0;
  COR_CUDA(cudaGetDevice(&prev));
  COR_CUDA(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  cudaSetDevice(prev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("cor_peer_open: %s", cudaGetErrorString(e));
    return COR_ECUDA;
  }
  return COR_OK;
}

int cor_peer_close(void* ptr) {
  if (ptr) COR_CUDA(cudaIpcCloseMemHandle(ptr));
  return COR_OK;
}

// COR_PEER_TIMEOUT_S: seconds a wait may last before the error word is set (default 600; 0 = unbounded)
static unsigned long long peer_timeout_ns() {
  double s = 600.0;                                   // read per launch: a getenv is nothing beside a kernel launch
  if (const char* e = getenv("COR_PEER_TIMEOUT_S")) s = atof(e);
  return s <= 0.0 ? 0ull : (unsigned long long)(s * 1e9);
}

static int peer_ctl(const char* what, void* const* peer_flags, void* state, int rank, int world, int channel, PeerCtl* c) {
  COR_REQUIRE(peer_flags && state, "%s: null pointer", what);
  COR_REQUIRE(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "%s: rank %d / world %d (max %d)", what, rank, world,
              kPeerMaxWorld);
  COR_REQUIRE(channel == 0 || channel == 1, "%s: channel %d", what, channel);
  *c = PeerCtl{reinterpret_cast<unsigned* const*>(peer_flags), reinterpret_cast<unsigned*>(state), rank, world, channel, peer_timeout_ns()};
  return COR_OK;
}

int cor_peer_signal(void* const* peer_flags, void* state, int rank, int world, int channel, cor_stream_t stream) {
  PeerCtl c;
  int rc = peer_ctl("cor_peer_signal", peer_flags, state, rank, world, channel, &c);
  if (rc) return rc;
  peer_signal_kernel<<<1, 32, 0, as_stream(stream)>>>(c);
  return check_launch("peer_signal_kernel");
}

int cor_peer_wait_exit(void* const* peer_flags, void* state, int rank, int world, int channel, cor_stream_t stream) {
  PeerCtl c;
  int rc = peer_ctl("cor_peer_wait_exit", peer_flags, state, rank, world, channel, &c);
  if (rc) return rc;
  peer_wait_exit_kernel<<<1, 32, 0, as_stream(stream)>>>(c);
  return check_launch("peer_wait_exit_kernel");
}

int cor_peer_gather_rows(const void* const* peer_src, void* all, long long bytes_per_rank, void* const* peer_flags, void* state,
                         int rank, int world, int channel, cor_stream_t stream) {
  COR_REQUIRE(peer_src && all && peer_flags && state, "cor_peer_gather_rows: null pointer");
  COR_REQUIRE(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "cor_peer_gather_rows: rank %d / world %d (max %d)", rank,
              world, kPeerMaxWorld);
  COR_REQUIRE(channel == 0 || channel == 1, "cor_peer_gather_rows: channel %d", channel);
  COR_REQUIRE(bytes_per_rank > 0 && bytes_per_rank % 16 == 0 && (uintptr_t)all % 16 == 0, "cor_peer_gather_rows: %lld bytes per rank must be a positive multiple of 16, 16-byte aligned", bytes_per_rank);
  PeerCtl c{reinterpret_cast<unsigned* const*>(peer_flags), reinterpret_cast<unsigned*>(state), rank, world, channel, peer_timeout_ns()};
  const long long vecs = bytes_per_rank / 16;
  peer_gather_kernel<<<peer_grid(vecs * world), kPeerThreads, 0, as_stream(stream)>>>(reinterpret_cast<const uint4* const*>(peer_src),
                                                                                     reinterpret_cast<uint4*>(all), vecs, c);
  return check_launch("peer_gather_kernel");
}

int cor_peer_reduce_rows(const void* const* peer_src, float* out, long long floats_per_rank, void* const* peer_flags, void* state,
                         int rank, int world, int channel, cor_stream_t stream) {
  COR_REQUIRE(peer_src && out && peer_flags && state, "cor_peer_reduce_rows: null pointer");
  COR_REQUIRE(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "cor_peer_reduce_rows: rank %d / world %d (max %d)", rank,
              world, kPeerMaxWorld);
  COR_REQUIRE(channel == 0 || channel == 1, "cor_peer_reduce_rows: channel %d", channel);
  COR_REQUIRE(floats_per_rank > 0 && floats_per_rank % 4 == 0 && (uintptr_t)out % 16 == 0, "cor_peer_reduce_rows: %lld floats per rank must be a positive multiple of 4, 16-byte aligned", floats_per_rank);
  PeerCtl c{reinterpret_cast<unsigned* const*>(peer_flags), reinterpret_cast<unsigned*>(state), rank, world, channel, peer_timeout_ns()};
  const long long vecs = floats_per_rank / 4;
  peer_reduce_kernel<<<peer_grid(vecs), kPeerThreads, 0, as_stream(stream)>>>(reinterpret_cast<const uint4* const*>(peer_src),
                                                                             reinterpret_cast<float4*>(out), vecs, c);
  return check_launch("peer_reduce_kernel");
}

}  // extern "C"
