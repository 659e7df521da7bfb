// fp32 instantiation of the C-ABI (the reported reduced-precision variant; FMA contraction on).
#define SCCAV_REAL float
#define SCCAV_SUFFIX f32
#include "capi_impl.cuh"
