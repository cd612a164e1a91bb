#define KGW_L 8
#include "kgw_kernels.inc"
