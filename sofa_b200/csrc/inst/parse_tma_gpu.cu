#define INST_NAME gpu
#define INST_KIND SRC_GPU
#include "parse_tma_inst.inc"
