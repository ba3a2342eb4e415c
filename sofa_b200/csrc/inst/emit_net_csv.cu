#define INST_NAME net_csv
#define INST_KIND SRC_NET
#define INST_FMT FMT_CSV
#include "emit_inst.inc"
