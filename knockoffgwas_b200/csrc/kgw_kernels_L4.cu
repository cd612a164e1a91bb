#define KGW_L 4
#include "kgw_kernels.inc"
