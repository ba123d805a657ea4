// placeholder (filled in below in this round): SDF interpenetration term
#include "mvs_internal.cuh"
namespace mvs {
int launch_sdf_terms(mvs_ctx* ctx, const float*, cudaStream_t) {
    return set_error(ctx, MVS_ERR_UNSUPPORTED, "SDF term not built yet");
}
int sdf_grid_launch(mvs_ctx* ctx, float*, const int*, int, const float*, int, int, int, cudaStream_t) {
    return set_error(ctx, MVS_ERR_UNSUPPORTED, "SDF grid not built yet");
}
}
