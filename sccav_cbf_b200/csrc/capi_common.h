// capi_common.h -- host-side helpers shared by the per-precision C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/sccav_cbf.h"

namespace sccav {
void set_error(const char* fmt, ...);
void count_launch();
int sm_count();          // SMs of the current device (cached per device)
int max_smem_optin();    // max opt-in dynamic shared memory per block of the current device
int max_smem_per_sm();   // shared memory of one SM
bool k12_coop_enabled(int spec);   // SCCAV_K12_QP=thread|coop overrides the staged kernel's QP form
bool k12_staged_enabled(int spec); // SCCAV_K12_PIPE=0|1 overrides the choice between the direct-load and the staged filter-step kernel
cudaError_t pool_alloc(void** p, size_t bytes, cudaStream_t st);   // stream-ordered allocation from the library's own pool
}  // namespace sccav

#define SCCAV_CUDA_CHECK(expr)                                                              \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            sccav::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return SCCAV_ECUDA;                                                             \
        }                                                                                   \
    } while (0)
