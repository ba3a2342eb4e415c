// libsofa_b200.so, translation unit: recorder side files (mpstat / vmstat / nvidia-smi / netstat).
#define SOFA_K_SIDE
#include "ctx.cuh"

extern "C" {

#include "capi_side.inc"

}  // extern "C"
