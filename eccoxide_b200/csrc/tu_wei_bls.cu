// tu_wei_bls.cu
#define ECB_TU_CURVE CurveBLSG1
#define ECB_TU_FN dev_wei_mul_bls
#define ECB_TU_CURVE_INDEX 2
#define ECB_TU_TABLE_FN dev_wei_table_bls
#define ECB_TU_BASE_FN dev_wei_mul_base_bls
#include "tu_wei.inc"
