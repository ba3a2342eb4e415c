#define INST_NAME sideout_aux
#define INST_KIND SRC_SIDEOUT
#define INST_FMT FMT_AUX
#include "emit_inst.inc"
