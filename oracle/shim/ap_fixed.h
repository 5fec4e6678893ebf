// ap_fixed.h -- placeholder: the embedding kernels include it but use no fixed-point type.
// TEST INFRASTRUCTURE ONLY (see ap_int.h).
#ifndef FR_SHIM_AP_FIXED_H
#define FR_SHIM_AP_FIXED_H
#include "ap_int.h"
#endif
