// TEMPORARY: entry points of the tcgen05 kernels until pool_umma.cu / sim_umma.cu land.
#include "common.cuh"
using namespace cor;
extern "C" size_t cor_pool_umma_work_bytes(int, int, int, int) { return 16; }
extern "C" int cor_pool_umma_fwd(const void*, const void*, int, int, int, int, float*, void*, cor_stream_t) {
  set_error("cor_pool_umma_fwd: tcgen05 pooling kernel not built in this revision");
  return COR_EUNSUP;
}
extern "C" int cor_sim_umma_fwd(const void*, const void*, int, int, int, float, float*, float*, void*, cor_stream_t) {
  set_error("cor_sim_umma_fwd: tcgen05 similarity kernel not built in this revision");
  return COR_EUNSUP;
}
