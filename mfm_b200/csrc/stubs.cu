// TEMPORARY: entry points not implemented yet return MFM_ERR_UNSUPPORTED (removed as they land).
#include "internal.h"
#define NYI { mfm_set_last_error_msg("not implemented yet"); return MFM_ERR_UNSUPPORTED; }
extern "C" {
size_t mfm_ode_workspace_bytes(const mfm_field_t*, const mfm_target_t*, const mfm_ode_opts_t*, int) { return 0; }
int mfm_ode_flow(const mfm_field_t*, const mfm_target_t*, const mfm_ode_opts_t*, int, int, const uint32_t*, const float*, float*, float*, int*, void*, size_t, mfm_stream_t) NYI
int mfm_field_eval(const mfm_field_t*, const mfm_target_t*, const mfm_ode_opts_t*, int, const float*, const float*, const float*, float*, float*, void*, size_t, mfm_stream_t) NYI
size_t mfm_flow_mh_workspace_bytes(const mfm_field_t*, const mfm_target_t*, const mfm_ode_opts_t*, int) { return 0; }
int mfm_flow_mh_step(const mfm_field_t*, const mfm_target_t*, const mfm_ode_opts_t*, int, const uint32_t*, int, int, int, int, float*, float*, float*, float*, uint8_t*, float*, float*, int*, void*, size_t, mfm_stream_t) NYI
size_t mfm_fm_workspace_bytes(const mfm_field_t*, const mfm_target_t*, int) { return 0; }
int mfm_fm_loss_grad(const mfm_field_t*, const mfm_target_t*, const uint32_t*, int, int, int, float, const float*, float*, float*, void*, size_t, mfm_stream_t) NYI
int mfm_fm_loss_grad_from_batch(const mfm_field_t*, const mfm_target_t*, int, const float*, const float*, const float*, float*, float*, void*, size_t, mfm_stream_t) NYI
int mfm_adamw_step(float*, const float*, float*, float*, const uint8_t*, long long, int*, float, int, float, float, float, float, float, int, mfm_stream_t) NYI
}
