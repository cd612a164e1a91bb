#define KGW_L 2
#include "kgw_kernels.inc"
