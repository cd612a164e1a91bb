#define KGW_L 1
#include "kgw_kernels.inc"
