#define INST_NAME gpu_json
#define INST_KIND SRC_GPU
#define INST_FMT FMT_JSON
#include "emit_inst.inc"
