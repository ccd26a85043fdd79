// tu_wei_p384.cu
#define ECB_TU_CURVE CurveP384
#define ECB_TU_FN dev_wei_mul_p384
#include "tu_wei.inc"
