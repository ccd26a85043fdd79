// tu_wei_k256.cu — secp256k1 (the reference's p256k1): the a = 0 kernels of tu_wei.inc on generic 8-limb fields
#define ECB_TU_CURVE CurveK256
#define ECB_TU_FN dev_wei_mul_k256
#define ECB_TU_CURVE_INDEX 3
#define ECB_TU_TABLE_FN dev_wei_table_k256
#define ECB_TU_BASE_FN dev_wei_mul_base_k256
#define ECB_TU_DECOMP_FN dev_wei_decompress_k256
#define ECB_TU_MSM_FN dev_wei_msm_k256
#define ECB_TU_MSM_FINISH_FN dev_wei_msm_finish_k256
#include "tu_wei.inc"
