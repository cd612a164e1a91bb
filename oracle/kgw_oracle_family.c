/* oracle/kgw_oracle_family.c -- TEST INFRASTRUCTURE ONLY.
 * IBD-family branch (family_bp.cpp, family_knockoffs.cpp); filled in below. */
#include "kgw_oracle.h"
