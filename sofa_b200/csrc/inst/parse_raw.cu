#define INST_NAME raw
#define INST_KIND SRC_RAW
#include "parse_inst.inc"
