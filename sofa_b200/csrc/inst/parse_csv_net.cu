#define INST_NAME net
#define INST_KIND SRC_NET
#include "parse_csv_inst.inc"
