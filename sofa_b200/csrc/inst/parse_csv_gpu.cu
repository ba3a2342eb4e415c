#define INST_NAME gpu
#define INST_KIND SRC_GPU
#include "parse_csv_inst.inc"
