// fp64 instantiation of the C-ABI (the default, parity-checked precision).
// Compiled with -fmad=false: products and sums round separately, in the oracle's order.
#define SCCAV_REAL double
#define SCCAV_SUFFIX f64
#include "capi_impl.cuh"
