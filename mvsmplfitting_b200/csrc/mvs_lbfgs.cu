#include "mvs_internal.cuh"
using namespace mvs;
extern "C" {
int mvs_lbfgs_run(mvs_ctx* ctx, float*, float*, const mvs_lbfgs_config*, mvs_lbfgs_stats*, void*) {
    return set_error(ctx, MVS_ERR_UNSUPPORTED, "lbfgs not built yet");
}
int mvs_fit_host(mvs_ctx* ctx, float*, const float*, const float*, const float*, int, const mvs_loss_config*,
                 const mvs_lbfgs_config*, float*, mvs_lbfgs_stats*, void*) {
    return set_error(ctx, MVS_ERR_UNSUPPORTED, "fit_host not built yet");
}
}
