// Composite-likelihood fit for large state spaces (the 400 x 400 co-evolution model).
// Placeholder until the batched DMMA GEMM chain lands: every entry point fails loudly.
#include "fit_internal.cuh"

namespace cherry {

int fit_large_workspace_bytes(int S, int, int, size_t*) {
  return fail(CHERRY_ELIMIT, "fit: S=%d > %d is not implemented yet", S, kSmallFitMaxS);
}
int fit_large_expm(const cherry_fit_args& a, cudaStream_t) {
  return fail(CHERRY_ELIMIT, "fit: S=%d > %d is not implemented yet", a.S, kSmallFitMaxS);
}
int fit_large_update(const cherry_fit_args& a, int, cudaStream_t) {
  return fail(CHERRY_ELIMIT, "fit: S=%d > %d is not implemented yet", a.S, kSmallFitMaxS);
}

}  // namespace cherry
