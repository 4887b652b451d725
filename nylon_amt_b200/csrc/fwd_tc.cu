// Tensor-core (tcgen05 / TMEM / TMA) forward of Model_SPEC2MIDI -- HFT_PREC_BF16 / HFT_PREC_F16.
#include "common.cuh"
#include "model.h"

namespace hft {

int tc_prepare_weights(Model* m, cudaStream_t s) { (void)m; (void)s; return HFT_OK; }
void tc_destroy(Model* m) { (void)m; }

int forward_tc(Model* m, int precision, const float* spec, long long sb, long long sbin, long long st, int B, const hft_outputs* o, cudaStream_t s) {
  (void)m; (void)spec; (void)sb; (void)sbin; (void)st; (void)B; (void)o; (void)s;
  set_error("hft_forward: precision %d (tensor-core path) is not built yet", precision);
  return HFT_ERR_UNSUPPORTED;
}

}  // namespace hft
