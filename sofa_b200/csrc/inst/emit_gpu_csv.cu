#define INST_NAME gpu_csv
#define INST_KIND SRC_GPU
#define INST_FMT FMT_CSV
#include "emit_inst.inc"
