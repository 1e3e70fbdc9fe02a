// K1 (tcgen05 form) - placeholder until the tensor-tile kernel lands: reports "unsupported" so every
// model resolves to the FP32 SIMT path.
#include "nlo_common.cuh"

bool nlo_sdf_tc_supported(const nlo_sdf_desc*) { return false; }
int nlo_sdf_tc_prepare(nlo_sdf_model*, const float*) { return 0; }
int nlo_sdf_tc_launch(nlo_sdf_model*, const float*, const float*, const float*, size_t, float*, float*, float*, cudaStream_t) {
  return nlo_fail("tensor-tile path not built");
}
