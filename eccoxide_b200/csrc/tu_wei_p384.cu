// tu_wei_p384.cu
#define ECB_TU_CURVE CurveP384
#define ECB_TU_FN dev_wei_mul_p384
#define ECB_TU_CURVE_INDEX 1
#define ECB_TU_TABLE_FN dev_wei_table_p384
#define ECB_TU_BASE_FN dev_wei_mul_base_p384
#define ECB_TU_BASE_CT_FN dev_wei_mul_base_ct_p384
#define ECB_TU_DECOMP_FN dev_wei_decompress_p384
#include "tu_wei.inc"
