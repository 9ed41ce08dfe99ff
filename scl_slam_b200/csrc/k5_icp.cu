// k5_icp.cu — K5 geometric verification (placeholder until the ICP kernels land).
#include "../../include/scl_engine.h"
extern "C" int scl_icp(scl_engine*, const void*, int, const void*, int, int, const scl_icp_params*, float*, float*, int*, int*)
{
    return SCL_ERR_UNSUPPORTED;
}
