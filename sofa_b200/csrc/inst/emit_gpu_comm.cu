#define INST_NAME gpu_comm
#define INST_KIND SRC_GPU
#define INST_FMT FMT_COMM
#include "emit_inst.inc"
