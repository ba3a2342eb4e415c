#define INST_NAME net_json
#define INST_KIND SRC_NET
#define INST_FMT FMT_JSON
#include "emit_inst.inc"
