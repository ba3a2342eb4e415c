// libsofa_b200.so, translation unit: synthetic trace text (bench / test infrastructure).
#include "ctx.cuh"
#include "synth.cuh"

extern "C" {

#include "capi_synth.inc"

}  // extern "C"
