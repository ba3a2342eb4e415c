#define INST_NAME perf_json
#define INST_KIND SRC_PERF
#define INST_FMT FMT_JSON
#include "emit_inst.inc"
