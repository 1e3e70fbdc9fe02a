// Instantiations of the deep tensor-tile kernel (sdf_tc_deep.cuh) for H = 128, M = 2 hidden matrices: one translation unit per
// shape so that the nine activation pairs of each compile in parallel with the others.
#include "sdf_tc_deep.cuh"

int nlo_sdf_tc_deep_launch_h128m2(nlo_sdf_model* m, const float* x, const float* y, const float* sbar, size_t n, float* s, float* jx, float* jy,
                                    cudaStream_t st) {
  return dispatch_deep<128, 2>(m, x, y, sbar, n, s, jx, jy, st);
}
