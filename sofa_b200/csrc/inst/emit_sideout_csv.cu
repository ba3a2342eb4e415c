#define INST_NAME sideout_csv
#define INST_KIND SRC_SIDEOUT
#define INST_FMT FMT_CSV
#include "emit_inst.inc"
