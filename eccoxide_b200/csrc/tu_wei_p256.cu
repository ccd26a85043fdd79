// tu_wei_p256.cu
#define ECB_TU_CURVE CurveP256
#define ECB_TU_FN dev_wei_mul_p256
#define ECB_TU_CURVE_INDEX 0
#define ECB_TU_TABLE_FN dev_wei_table_p256
#define ECB_TU_BASE_FN dev_wei_mul_base_p256
#define ECB_TU_BASE_CT_FN dev_wei_mul_base_ct_p256
#define ECB_TU_DECOMP_FN dev_wei_decompress_p256
#include "tu_wei.inc"
