// libsofa_b200.so, translation unit: analyze half (gpu_profile, comm_profile, top-k, cpu_profile, net_profile,
// nvsmi_profile, window reduction of concurrency_breakdown).
#include "ctx.cuh"
#include "kernels_analyze.cuh"
#include "kernels_sort.cuh"
#include "kernels_window.cuh"
#include "kernels_profiles.cuh"
#include "kernels_side2.cuh"

extern "C" {

#include "capi_analyze.inc"
#include "capi_window.inc"
#include "capi_profiles.inc"
#include "capi_side2.inc"

}  // extern "C"
