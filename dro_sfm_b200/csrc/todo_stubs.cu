// Entry points declared in include/drosfm_b200.h whose kernels have not landed yet.
// Each returns DROSFM_ENOTSUP loudly; the file shrinks as the kernels arrive.
#include "common.cuh"
using namespace drosfm;
#define NOT_YET(name) do { set_error(name ": not implemented in this build"); return DROSFM_ENOTSUP; } while (0)
extern "C" {
int drosfm_automask_fwd(const float*, const float* const*, int, const drosfm_photo_opts_t*, float*, int, int, int, drosfm_stream_t) { NOT_YET("automask_fwd"); }
int drosfm_photometric_fwd(const float*, const float* const*, int, const float* const*, int, int, const drosfm_cams_t*, const float* const*, const float*, const drosfm_photo_opts_t*, uint8_t*, float*, void*, int, int, int, drosfm_stream_t) { NOT_YET("photometric_fwd"); }
int drosfm_photometric_bwd(const float*, const float*, const float* const*, int, const float* const*, int, int, const drosfm_cams_t*, const float* const*, const uint8_t*, const drosfm_photo_opts_t*, float* const*, float* const*, void*, int, int, int, drosfm_stream_t) { NOT_YET("photometric_bwd"); }
int drosfm_smoothness_fwd(const float*, const float* const*, int, float, float*, float*, void*, int, int, int, drosfm_stream_t) { NOT_YET("smoothness_fwd"); }
int drosfm_smoothness_bwd(const float*, const float*, const float* const*, int, float, const float*, float* const*, void*, int, int, int, drosfm_stream_t) { NOT_YET("smoothness_bwd"); }
int drosfm_reproj_loss_fwd(const float*, int, const drosfm_cams_t*, const float* const*, const float* const*, int, int, float, float, float, float*, void*, int, int, int, drosfm_stream_t) { NOT_YET("reproj_loss_fwd"); }
int drosfm_reproj_loss_bwd(const float*, const float*, int, const drosfm_cams_t*, const float* const*, const float* const*, int, int, float, float, float, float* const*, void*, int, int, int, drosfm_stream_t) { NOT_YET("reproj_loss_bwd"); }
}
