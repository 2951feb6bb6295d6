// pbx_stubs.cu -- entry points declared in pbx.h whose kernels are not in yet.
// They fail loudly (no CPU fallback); each is removed as its kernel lands.
#include "pbx_common.cuh"

#define PBX_STUB(name)                                            \
  pbx_set_error(name ": kernel not implemented in this build");   \
  return PBX_ERR_UNSUPPORTED

extern "C" {
int pbx_gibbs_mvn_run(pbx_ctx*, const pbx_gibbs_mvn_params*) { PBX_STUB("pbx_gibbs_mvn_run"); }
int pbx_mvn_logpdf(pbx_ctx*, const double*, int32_t, int64_t, const double*, const double*, double,
                   int32_t, double*) {
  PBX_STUB("pbx_mvn_logpdf");
}
}
