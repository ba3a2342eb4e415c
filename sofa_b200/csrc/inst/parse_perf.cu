#define INST_NAME perf
#define INST_KIND SRC_PERF
#include "parse_inst.inc"
