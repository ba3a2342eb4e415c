#define INST_NAME perf_csv
#define INST_KIND SRC_PERF
#define INST_FMT FMT_CSV
#include "emit_inst.inc"
