/* A host program in plain C that binds include/cor_b200.h the way any FFI would: link against libcor_b200.so, ask
 * for the ABI version and the scratch sizes, and see a bad call refused with COR_EINVAL and a message.  No GPU work:
 * this is the "does the boundary hold in C" probe run by tests/test_abi_cpu.py.
 *   gcc -std=c99 -Wall -Werror -Iinclude examples/c_abi_probe.c -Lcor_b200 -l:libcor_b200.so -Wl,-rpath,$PWD/cor_b200 */
#include <stdio.h>
#include <string.h>

#include "cor_b200.h"

int main(void) {
  if (cor_abi_version() != COR_ABI_VERSION) {
    fprintf(stderr, "ABI mismatch: library %d, header %d\n", cor_abi_version(), COR_ABI_VERSION);
    return 1;
  }
  /* scratch sizes for BASELINE config 2: 1024 masks of 1024^2 -> 64^2, pooling of 16 x 256 x 4096 with 80 rows */
  size_t prep = cor_mask_prep_work_bytes(1024, 1024, 1024, 64, 64);
  size_t pool = cor_pool_umma_work_bytes(16, 256, 4096, 80);
  size_t seg = cor_seg_loss_work_bytes(16, 256, 256);
  if (prep == 0 || pool == 0 || seg == 0) return 2;
  /* a null pointer must be refused before anything is launched */
  int rc = cor_topk(NULL, NULL, NULL, 8, 4, 64, 2, NULL, NULL, NULL);
  if (rc != COR_EINVAL || strlen(cor_last_error()) == 0) return 3;
  printf("cor_b200 ABI v%d: mask_prep work %zu B, pool partials %zu B, seg_loss work %zu B; bad call -> %d (%s)\n",
         cor_abi_version(), prep, pool, seg, rc, cor_last_error());
  return 0;
}
