#define INST_NAME perf
#define INST_KIND SRC_PERF
#include "parse_csv_inst.inc"
