#define INST_NAME sideout_json
#define INST_KIND SRC_SIDEOUT
#define INST_FMT FMT_JSON
#include "emit_inst.inc"
