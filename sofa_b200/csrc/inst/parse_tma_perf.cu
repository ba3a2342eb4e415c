#define INST_NAME perf
#define INST_KIND SRC_PERF
#include "parse_tma_inst.inc"
