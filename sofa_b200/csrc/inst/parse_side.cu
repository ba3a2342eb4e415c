#define INST_NAME side
#define INST_KIND SRC_SIDE
#include "parse_inst.inc"
