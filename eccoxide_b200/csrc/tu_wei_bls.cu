// tu_wei_bls.cu
#define ECB_TU_CURVE CurveBLSG1
#define ECB_TU_FN dev_wei_mul_bls
#include "tu_wei.inc"
