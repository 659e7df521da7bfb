// capi_common.cu -- precision-independent part of the C-ABI: version, errors, defaults, device
// attribute cache, launch counter and the FMA-peak microbenchmark used by bench.py.
#include <cstdlib>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include <algorithm>
#include <mutex>
#include <vector>

#include "capi_common.h"
#include "course_index.cuh"

namespace sccav {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

struct DevAttr {
    int sms = 0;
    int smem = 0;
    int smem_sm = 0;
};
static DevAttr g_attr[64];

static DevAttr& attr() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    DevAttr& a = g_attr[dev];
    if (a.sms == 0) {
        int sms = 0, smem = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        a.smem = smem;
        int smem_sm = 0;
        cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
        a.smem_sm = smem_sm;
        a.sms = sms > 0 ? sms : 148;
    }
    return a;
}

// The entry points take their scratch and the host-variant buffers from a stream-ordered pool OF THEIR OWN (one per
// device, created on first use): freed blocks stay in the pool for the next call -- mapping tens of MB again at every
// call costs milliseconds -- but only up to a bounded threshold (default 1 GiB, SCCAV_POOL_KEEP_MB), and the device's
// default pool, which other libraries in the process use, is left alone.  sccav_trim_pool() gives everything back.
static cudaMemPool_t g_pool[64] = {};
static std::mutex g_pool_mu;

cudaMemPool_t lib_pool() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    std::lock_guard<std::mutex> lock(g_pool_mu);
    if (!g_pool[dev]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t pool = nullptr;
        if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) {
            cudaGetLastError();
            cudaDeviceGetDefaultMemPool(&pool, dev);            // (very old drivers: fall back to the default pool, untouched)
        } else {
            uint64_t keep = 1024ull << 20;
            if (const char* e = getenv("SCCAV_POOL_KEEP_MB")) keep = (uint64_t)strtoull(e, nullptr, 10) << 20;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        g_pool[dev] = pool;
    }
    return g_pool[dev];
}

cudaError_t pool_alloc(void** p, size_t bytes, cudaStream_t st) {
    return cudaMallocFromPoolAsync(p, bytes ? bytes : 1, lib_pool(), st);
}

int sm_count() { return attr().sms; }
int max_smem_optin() { return attr().smem; }
int max_smem_per_sm() { return attr().smem_sm; }
// Developer knobs for A/B measurements (read per call: tests flip them between launches).
// SCCAV_K12_PIPE=0 selects the direct-load filter-step kernel, =1 the staged one wherever it fits; default:
// staged for prepared slots, direct loads for canonical ellipses (measured).
bool k12_staged_enabled(int spec) {
    const char* e = getenv("SCCAV_K12_PIPE");
    if (e && e[0] == '0') return false;
    if (e && e[0] == '1') return true;
    return spec == 2;
}
// SCCAV_K12_QP=thread|coop overrides the QP form of the filter-step kernels; default: warp-cooperative
// (shortcut + cooperative enumeration) in the staged kernel on prepared slots, one thread per problem elsewhere.
bool k12_coop_enabled(int spec) {
    const char* e = getenv("SCCAV_K12_QP");
    if (e && e[0] == 't') return false;
    if (e && e[0] == 'c') return true;
    return spec == 2;
}

// ---- FMA peak: NCHAIN independent dependent-FMA chains per thread, fully unrolled.
template <typename T, int NCHAIN>
__global__ void __launch_bounds__(256) fma_peak_kernel(T* out, int iters, T a, T b) {
    T acc[NCHAIN];
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) acc[i] = T(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i) acc[i] = fma(acc[i], a, b);
    }
    T s = T(0);
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) s += acc[i];
    if (s == T(-1.2345)) out[0] = s;   // never true; keeps the chains alive
}

template <typename T> static int measure(double* tflops) {
    const int NCHAIN = 8, iters = 4096;
    const int blocks = sm_count() * 8, threads = 256;
    T* d = nullptr;
    SCCAV_CUDA_CHECK(cudaMalloc(&d, sizeof(T)));
    cudaEvent_t e0, e1;
    SCCAV_CUDA_CHECK(cudaEventCreate(&e0));
    SCCAV_CUDA_CHECK(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        SCCAV_CUDA_CHECK(cudaEventRecord(e0, 0));
        fma_peak_kernel<T, NCHAIN><<<blocks, threads>>>(d, iters, T(1.0000001), T(1e-7));
        count_launch();
        SCCAV_CUDA_CHECK(cudaEventRecord(e1, 0));
        SCCAV_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0.f;
        SCCAV_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * NCHAIN * (double)iters * blocks * threads;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return SCCAV_OK;
}

}  // namespace sccav

extern "C" {

int sccav_version(void) { return SCCAV_VERSION; }

const char* sccav_last_error(void) { return sccav::g_err; }

int sccav_device_ok(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n > 0 ? 1 : 0;
}

void sccav_default_params(sccav_params* p) {
    memset(p, 0, sizeof(*p));
    p->model = SCCAV_MODEL_DBM;
    p->nominal = SCCAV_NOMINAL_STANLEY;
    p->alpha = 1.0;                                  // cbf.py:128
    p->L = 2.9;                                      // sce.py:55
    p->lr = 2.9 / 2;                                 // sce.py:56
    p->lf = 2.9 - 2.9 / 2;                           // sce.py:57
    p->max_steer = 30.0 * (3.141592653589793 / 180.0);   // np.radians(30.0), sce.py:58
    p->dt = 0.1;                                     // sce.py:54
    p->k_stanley = 0.5;                              // sce.py:52
    p->ks_stanley = 0.0;
    p->Kp = 1.0;                                     // sce.py:53
    p->target_speed = 30.0 / 3.6;                    // sce.py:590
    p->t_max = 30.0;                                 // sce.py:592
    p->R[0] = 1.0; p->R[1] = 0.0; p->R[2] = 0.0; p->R[3] = 1.0;   // cbf.py:134
    p->seeker_k = 0.2;                               // rdo.py:193
    p->seeker_vmin = 3.0;
}

int sccav_measure_fma_peak(int32_t dtype, double* tflops_out) {
    if (!tflops_out) { sccav::set_error("tflops_out is NULL"); return SCCAV_EINVAL; }
    if (dtype == 64) return sccav::measure<double>(tflops_out);
    if (dtype == 32) return sccav::measure<float>(tflops_out);
    sccav::set_error("dtype must be 32 or 64");
    return SCCAV_EINVAL;
}

int64_t sccav_launch_count(void) { return sccav::g_launches.load(); }

int sccav_trim_pool(void) {
    cudaMemPool_t pool = sccav::lib_pool();
    if (pool) SCCAV_CUDA_CHECK(cudaMemPoolTrimTo(pool, 0));
    return SCCAV_OK;
}

}  // extern "C"

namespace sccav {
template <typename T, typename T2>
static void debug_course_index(const double* cx, const double* cy, int P, const double* fx, const double* fy,
                               const int32_t* hint, int64_t nq, int32_t* idx, int32_t* idx_full, int64_t* evals) {
    std::vector<T2> xy(course_nslot(P));
    for (int i = 0; i < P; ++i) { xy[course_slot(i)].x = (T)cx[i]; xy[course_slot(i)].y = (T)cy[i]; }
    for (int i = P; i < course_nleaf(P) * SCCAV_LEAF; ++i) xy[course_slot(i)] = xy[course_slot(P - 1)];
    // same construction as course_stage (kernels.cuh): origin = the middle point of the course
    T org[2] = {xy[course_slot(P / 2)].x, xy[course_slot(P / 2)].y};
    double e = 0.0;
    for (int i = 0; i < P; ++i)
        e = std::max(e, std::max(fabs((double)xy[course_slot(i)].x - (double)org[0]), fabs((double)xy[course_slot(i)].y - (double)org[1])));
    const float ext = course_ext_inflate(e);
    int lev[2 * SCCAV_MAX_LEVELS], units;
    const int nlev = course_levels(P, lev, &units);
    std::vector<float4> nodes(units);
    std::vector<float2> irs(units);
    CourseIndex<T, T2> ci;
    ci.xy = xy.data();
    ci.chord = nodes.data(); ci.ir = irs.data(); ci.lev = lev; ci.org = org; ci.ext = &ext;
    ci.np = P; ci.nleaf = course_nleaf(P); ci.nlev = nlev;
    for (int k = 0; k < nlev; ++k)
        for (int j = 0; j < lev[2 * k + 1]; ++j) {
            float4 c;
            float2 r;
            capsule_build<T, T2>(xy.data(), P, k, j, (double)org[0], (double)org[1], ext, c, r);
            nodes[lev[2 * k] + j] = c;
            irs[lev[2 * k] + j] = r;
        }
    // the cover table and the dummy node, as course_stage builds them
    const int kc = cover_kc(nlev), dummy = units - 1;
    std::vector<uint16_t> cov((size_t)ci.nleaf * kc);
    std::vector<uint8_t> ncov(ci.nleaf);
    dummy_node(nodes.data(), irs.data(), dummy);
    for (int w = 0; w < ci.nleaf; ++w) ncov[w] = (uint8_t)cover_row(w, nlev, lev, kc, dummy, cov.data() + (size_t)w * kc);
    ci.ncover = ncov.data(); ci.cov = cov.data(); ci.kc = kc; ci.dummy = dummy;
    for (int64_t k = 0; k < nq; ++k) {
        int ne = 0;
        idx[k] = course_nearest<T, T2>(ci, (T)fx[k], (T)fy[k], hint ? hint[k] : 0, &ne);
        if (idx_full) idx_full[k] = course_nearest_full<T, T2>(xy.data(), P, (T)fx[k], (T)fy[k]);
        if (evals) evals[k] = ne;
    }
}
}  // namespace sccav

extern "C" {

int sccav_debug_cover_host(int32_t P, int32_t* nleaf_out, int32_t* nlev_out, int32_t* kc_out, int32_t* count_out,
                           int32_t* level_out, int32_t* index_out) {
    using namespace sccav;
    if (P < 1 || !nleaf_out || !nlev_out || !kc_out) { set_error("bad argument"); return SCCAV_EINVAL; }
    int lev[2 * SCCAV_MAX_LEVELS], units;
    const int nlev = course_levels(P, lev, &units);
    const int nleaf = course_nleaf(P), kc = cover_kc(nlev), dummy = units - 1;
    *nleaf_out = nleaf; *nlev_out = nlev; *kc_out = kc;
    if (!level_out || !index_out || !count_out) return SCCAV_OK;
    if (units > 65535) { set_error("course too long for 16-bit node ids"); return SCCAV_EINVAL; }
    std::vector<uint16_t> row(kc);
    for (int w = 0; w < nleaf; ++w) {
        count_out[w] = cover_row(w, nlev, lev, kc, dummy, row.data());
        for (int k = 0; k < kc; ++k) {
            int h = -1, j = -1;
            if (row[k] != dummy) unit_node(lev, nlev, row[k], h, j);
            level_out[(size_t)w * kc + k] = h;
            index_out[(size_t)w * kc + k] = j;
        }
    }
    return SCCAV_OK;
}

int sccav_debug_course_index_host(const double* cx, const double* cy, int32_t P, const double* fx, const double* fy,
                                  const int32_t* hint, int64_t nq, int32_t dtype, int32_t* idx_out,
                                  int32_t* idx_full_out, int64_t* evals_out) {
    if (!cx || !cy || P < 1 || !fx || !fy || nq < 0 || !idx_out) { sccav::set_error("bad argument"); return SCCAV_EINVAL; }
    if (dtype == 64) sccav::debug_course_index<double, double2>(cx, cy, P, fx, fy, hint, nq, idx_out, idx_full_out, evals_out);
    else if (dtype == 32) sccav::debug_course_index<float, float2>(cx, cy, P, fx, fy, hint, nq, idx_out, idx_full_out, evals_out);
    else { sccav::set_error("dtype must be 32 or 64"); return SCCAV_EINVAL; }
    return SCCAV_OK;
}

}  // extern "C"
