// capi_impl.cuh -- the extern "C" entry points for ONE precision.  Included by capi_f64.cu
// (SCCAV_REAL = double, compiled with -fmad=false) and capi_f32.cu (SCCAV_REAL = float).
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "capi_common.h"
#include "kernels.cuh"

#ifndef SCCAV_REAL
#error "define SCCAV_REAL and SCCAV_SUFFIX before including capi_impl.cuh"
#endif

#define SCCAV_CAT_(a, b) a##b
#define SCCAV_CAT(a, b) SCCAV_CAT_(a, b)
#define SCCAV_FN(name) SCCAV_CAT(name, SCCAV_SUFFIX)

namespace sccav {
namespace {

typedef SCCAV_REAL real;

int check_common(const sccav_params* p, const uint8_t* slot_desc, int32_t M, int64_t N, bool allow_m0) {
    if (!p) { set_error("params is NULL"); return SCCAV_EINVAL; }
    if (N < 0) { set_error("N < 0"); return SCCAV_EINVAL; }
    if (M < 0 || M > SCCAV_MAX_ROWS) { set_error("M must be in [0, %d], got %d", SCCAV_MAX_ROWS, M); return SCCAV_EINVAL; }
    if (M == 0 && !allow_m0) {
        // DBM_CBF_2DS.solve_cbf raises ValueError on an empty obstacle list (cbf.py:177-180)
        set_error("Cannot solve CBF for an empty obstacle list (M = 0)");
        return SCCAV_EINVAL;
    }
    if (M > 0 && !slot_desc) { set_error("slot_desc is NULL"); return SCCAV_EINVAL; }
    for (int m = 0; m < M; ++m) {
        int t = slot_desc[m] & SCCAV_SLOT_TYPE_MASK;
        if (t > SCCAV_SLOT_LANE_SQRT) { set_error("slot %d: unknown type %d", m, t); return SCCAV_EINVAL; }
    }
    if (p->model < 0 || p->model > SCCAV_MODEL_SADBM) { set_error("unknown model %d", p->model); return SCCAV_EINVAL; }
    if ((p->flags & SCCAV_FLAG_BETA_IO) && p->model != SCCAV_MODEL_DBM) {
        set_error("SCCAV_FLAG_BETA_IO: model DBM only (got model %d)", p->model);
        return SCCAV_EINVAL;
    }
    double det = p->R[0] * p->R[3] - p->R[1] * p->R[2];
    if (!(p->R[0] > 0.0) || !(det > 0.0)) {
        // set_qp_cost_weight (cbf.py:154-157) expects a symmetric positive definite 2x2
        set_error("R must be symmetric positive definite 2x2");
        return SCCAV_EINVAL;
    }
    return SCCAV_OK;
}

Params<real> convert(const sccav_params* p) {
    Params<real> q;
    q.model = p->model; q.nominal = p->nominal; q.terminate = p->terminate; q.seeker = p->seeker;
    q.kbm_driver_delta = p->kbm_driver_delta; q.record_stride = p->record_stride; q.flags = p->flags;
    q.alpha = (real)p->alpha; q.lr = (real)p->lr; q.lf = (real)p->lf; q.L = (real)p->L;
    q.max_steer = (real)p->max_steer; q.dt = (real)p->dt; q.k_stanley = (real)p->k_stanley;
    q.ks_stanley = (real)p->ks_stanley; q.Kp = (real)p->Kp; q.target_speed = (real)p->target_speed;
    q.t_max = (real)p->t_max;
    for (int i = 0; i < 4; ++i) q.R[i] = (real)p->R[i];
    {   // same operations, same precision as the device's RInv: bit-identical
        const volatile real det = q.R[0] * q.R[3] - q.R[1] * q.R[2];
        q.Ri[0] = q.R[3] / det; q.Ri[1] = (-q.R[1]) / det; q.Ri[2] = (-q.R[2]) / det; q.Ri[3] = q.R[0] / det;
    }
    q.seeker_k = (real)p->seeker_k; q.seeker_vmin = (real)p->seeker_vmin;
    q.uref0 = (real)p->uref0; q.uref1 = (real)p->uref1;
    q.sadbm_dt = (real)(p->sadbm_dt > 0.0 ? p->sadbm_dt : 0.001);              // cbf.py:323: dt = 0.001
    return q;
}

SlotDesc make_desc(const uint8_t* slot_desc, int M) {
    SlotDesc sd;
    memset(&sd, 0, sizeof(sd));
    for (int m = 0; m < M; ++m) sd.d[m] = slot_desc[m];
    return sd;
}

PerVehicle<real> make_pv(const sccav_pervehicle* pv) {
    PerVehicle<real> q{nullptr, nullptr, nullptr, nullptr, nullptr};
    if (pv) {
        q.aug = (real*)pv->aug;
        q.alpha = (const real*)pv->alpha;
        q.R = (const real*)pv->R;
        q.target_speed = (const real*)pv->target_speed;
        q.count = pv->count;
    }
    return q;
}

// grid for the HBM-bound grid-stride kernels: enough CTAs to fill every SM, in multiples of the SM count
int stream_grid(int64_t N, int block) {
    int64_t need = (N + block - 1) / block;
    int64_t cap = (int64_t)sm_count() * 16;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// block size for kernels that keep rows[3*M][block] in shared memory
int rows_block(int M, size_t& smem) {
    int block = 256;
    while (block > 32 && (size_t)3 * M * block * sizeof(real) > 96 * 1024) block >>= 1;
    smem = (size_t)3 * (M > 0 ? M : 1) * block * sizeof(real);
    return block;
}

int do_prepare(const uint8_t* slot_desc, int32_t M, int64_t N, const real* in, real* out, uint8_t* desc_out, cudaStream_t st) {
    sccav_params dp;
    sccav_default_params(&dp);
    int rc = check_common(&dp, slot_desc, M, N, true);
    if (rc) return rc;
    if (M > 0 && !desc_out) { set_error("slot_desc_out is NULL"); return SCCAV_EINVAL; }
    for (int m = 0; m < M; ++m) {
        const int t = slot_desc[m] & SCCAV_SLOT_TYPE_MASK;
        desc_out[m] = t == SCCAV_SLOT_ELLIPSE ? (uint8_t)((slot_desc[m] & ~SCCAV_SLOT_TYPE_MASK) | SCCAV_SLOT_ELLIPSE_PREP) : slot_desc[m];
    }
    if (N == 0 || M == 0) return SCCAV_OK;
    if (!in || !out) { set_error("NULL array argument"); return SCCAV_EINVAL; }
    PrepareArgs<real> a;
    a.sd = make_desc(slot_desc, M); a.M = M; a.N = N; a.in = in; a.out = out;
    const int block = 256;
    const size_t vec = 2 * sizeof(real);
    if ((N & 1) == 0 && ((uintptr_t)in % vec) == 0 && ((uintptr_t)out % vec) == 0)
        prepare_obstacles_vec2_kernel<real><<<stream_grid((int64_t)M * (N / 2), block), block, 0, st>>>(a);
    else
        prepare_obstacles_kernel<real><<<stream_grid((int64_t)M * N, block), block, 0, st>>>(a);
    count_launch();
    SCCAV_CUDA_CHECK(cudaGetLastError());
    return SCCAV_OK;
}

int do_ingest(int32_t type, int32_t mode, double buffer, int32_t M, int32_t K, int64_t N, const int32_t* box_id, const real* box,
              int32_t* slot_id, real* obst, int32_t* count, int32_t* dropped, cudaStream_t st) {
    if (type != SCCAV_SLOT_ELLIPSE && type != SCCAV_SLOT_CONE) {
        // ObstacleList2D.update_by_bounding_box makes Ellipse2D or CollisionCone2D entries (obstacles.py:843-846)
        set_error("obs_type must be SCCAV_SLOT_ELLIPSE or SCCAV_SLOT_CONE, got %d", type);
        return SCCAV_EINVAL;
    }
    if (mode != SCCAV_INGEST_UPDATE && mode != SCCAV_INGEST_REBUILD) { set_error("unknown ingest mode %d", mode); return SCCAV_EINVAL; }
    if (M < 1 || M > SCCAV_MAX_ROWS) { set_error("M must be in [1, %d], got %d", SCCAV_MAX_ROWS, M); return SCCAV_EINVAL; }
    if (K < 0 || K > 32) { set_error("K must be in [0, 32], got %d", K); return SCCAV_EINVAL; }
    if (N < 0) { set_error("N < 0"); return SCCAV_EINVAL; }
    if (N == 0) return SCCAV_OK;
    if ((K > 0 && (!box_id || !box)) || !slot_id || !obst || !count) { set_error("NULL array argument"); return SCCAV_EINVAL; }
    IngestArgs<real> a;
    a.type = type; a.mode = mode; a.M = M; a.K = K; a.N = N; a.buffer = (real)buffer;
    a.box_id = box_id; a.box = box; a.slot_id = slot_id; a.obst = obst; a.count = count; a.dropped = dropped;
    const int block = 128;
    const size_t smem = (size_t)(K > 0 ? K : 1) * block * sizeof(int32_t);
    ingest_boxes_kernel<real><<<stream_grid(N, block), block, smem, st>>>(a);
    count_launch();
    SCCAV_CUDA_CHECK(cudaGetLastError());
    return SCCAV_OK;
}

int do_barrier_rows(const sccav_params* p, const uint8_t* slot_desc, int32_t M, int64_t N, const real* state,
                    const real* obst, const sccav_pervehicle* pv, real* A, real* b, real* h, cudaStream_t st) {
    int rc = check_common(p, slot_desc, M, N, false);
    if (rc) return rc;
    if (p->model == SCCAV_MODEL_SADBM) { set_error("model SADBM: rows depend on the carried beta -- use sccav_filter_step_*"); return SCCAV_EINVAL; }
    if (N == 0) return SCCAV_OK;
    if (!state || !obst || !A || !b) { set_error("NULL array argument"); return SCCAV_EINVAL; }
    RowsArgs<real> a;
    a.P = convert(p); a.sd = make_desc(slot_desc, M); a.M = M; a.N = N;
    a.state = state; a.obst = obst; a.pv = make_pv(pv); a.A = A; a.b = b; a.h = h;
    const int block = 256;
    barrier_rows_kernel<real><<<stream_grid(N, block), block, 0, st>>>(a);
    count_launch();
    SCCAV_CUDA_CHECK(cudaGetLastError());
    return SCCAV_OK;
}

int do_barrier_partials(const uint8_t* slot_desc, int32_t M, int64_t N, const real* state, const real* obst, real* out,
                        cudaStream_t st) {
    sccav_params dp;
    sccav_default_params(&dp);
    int rc = check_common(&dp, slot_desc, M, N, false);
    if (rc) return rc;
    if (N == 0) return SCCAV_OK;
    if (!state || !obst || !out) { set_error("NULL array argument"); return SCCAV_EINVAL; }
    PartialsArgs<real> a;
    a.sd = make_desc(slot_desc, M); a.M = M; a.N = N; a.state = state; a.obst = obst; a.out = out; a.count = nullptr;
    const int block = 256;
    barrier_partials_kernel<real><<<stream_grid(N, block), block, 0, st>>>(a);
    count_launch();
    SCCAV_CUDA_CHECK(cudaGetLastError());
    return SCCAV_OK;
}

int do_stanley(const sccav_params* p, int64_t N, const real* state, const real* front, const real* cx, const real* cy,
               const real* cyaw, int32_t P, int32_t* target_idx, real* delta, real* err, cudaStream_t st) {
    if (!p) { set_error("params is NULL"); return SCCAV_EINVAL; }
    if (N < 0) { set_error("N < 0"); return SCCAV_EINVAL; }
    if (P < 1 || !cx || !cy || !cyaw) { set_error("Stanley control needs a trajectory (P >= 1)"); return SCCAV_EINVAL; }
    if (N == 0) return SCCAV_OK;
    if (!state || !target_idx || !delta) { set_error("NULL array argument"); return SCCAV_EINVAL; }
    const RolloutSmem<real> lay(P, true);
    const int block = 256;
    size_t smem = lay.off_cyaw + (size_t)block * sizeof(real);
    if (smem > (size_t)max_smem_optin()) { set_error("trajectory of %d points does not fit in shared memory", P); return SCCAV_EINVAL; }
    StanleyArgs<real> a;
    a.P = convert(p); a.N = N; a.np = P; a.state = state; a.front = front; a.cx = cx; a.cy = cy; a.cyaw = cyaw;
    a.target_idx = target_idx; a.delta = delta; a.err = err;
    int64_t need = (N + block - 1) / block;
    int64_t cap = (int64_t)sm_count() * 4;
    SCCAV_CUDA_CHECK(cudaFuncSetAttribute(stanley_kernel<real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stanley_kernel<real><<<(int)(need < cap ? need : cap), block, smem, st>>>(a);
    count_launch();
    SCCAV_CUDA_CHECK(cudaGetLastError());
    return SCCAV_OK;
}

int do_qp2(const sccav_params* p, int32_t M, int64_t N, const real* A, const real* b, const real* r,
           const sccav_pervehicle* pv, real* u, uint32_t* mask, uint8_t* status, int warp, cudaStream_t st) {
    uint8_t dummy[SCCAV_MAX_ROWS] = {0};
    int rc = check_common(p, dummy, M, N, false);
    if (rc) return rc;
    if (N == 0) return SCCAV_OK;
    if (!A || !b || !r || !u) { set_error("NULL array argument"); return SCCAV_EINVAL; }
    QpArgs<real> a;
    a.P = convert(p); a.M = M; a.N = N; a.A = A; a.b = b; a.r = r; a.pv = make_pv(pv);
    a.u = u; a.mask = mask; a.status = status;
    if (warp) {
        const int block = 256;
        int64_t need = (N * 32 + block - 1) / block;
        int64_t cap = (int64_t)sm_count() * 16;
        qp2_warp_kernel<real><<<(int)(need < cap ? need : cap), block, 0, st>>>(a);
    } else {
        size_t smem;
        const int block = rows_block(M, smem);
        SCCAV_CUDA_CHECK(cudaFuncSetAttribute(qp2_kernel<real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        qp2_kernel<real><<<stream_grid(N, block), block, smem, st>>>(a);
    }
    count_launch();
    SCCAV_CUDA_CHECK(cudaGetLastError());
    return SCCAV_OK;
}

int do_filter_step(const sccav_params* p, const uint8_t* slot_desc, int32_t M, int64_t N, const real* state,
                   const real* obst, const real* u_ref, const sccav_pervehicle* pv, real* u, uint32_t* mask,
                   uint8_t* status, real* h_min, cudaStream_t st) {
    int rc = check_common(p, slot_desc, M, N, false);
    if (rc) return rc;
    if (p->model == SCCAV_MODEL_NONE) { set_error("model NONE has no filter step"); return SCCAV_EINVAL; }
    const bool sadbm = p->model == SCCAV_MODEL_SADBM;
    if (sadbm && !(pv && pv->aug) && N > 0) {
        set_error("model SADBM is stateful: sccav_pervehicle.aug [2][N] (beta, last beta_ref) is required");
        return SCCAV_EINVAL;
    }
    if (N == 0) return SCCAV_OK;
    if (!state || !obst || !u_ref || !u) { set_error("NULL array argument"); return SCCAV_EINVAL; }
    FilterArgs<real> a;
    a.P = convert(p); a.sd = make_desc(slot_desc, M); a.M = M; a.N = N;
    a.state = state; a.obst = obst; a.u_ref = u_ref; a.pv = make_pv(pv);
    a.u = u; a.mask = mask; a.status = status; a.h_min = h_min;
    size_t smem;
    const int block = rows_block(M, smem);
    typedef void (*filter_fn)(FilterArgs<real>);
    const int spec = sadbm ? SCCAV_SPEC_GENERIC : choose_spec(slot_desc, M);      // SADBM: generic slot loop only
    const int nf = spec == SCCAV_SPEC_ELLIPSE ? 7 : ((slot_desc[0] & SCCAV_SLOT_STATIC) ? 6 : 8);
    const size_t staged = (size_t)nf * M * 256 * sizeof(real);
    if (k12_staged_enabled(spec) && spec != SCCAV_SPEC_GENERIC && nf < 8 &&
        2 * (staged + 1024) <= (size_t)max_smem_per_sm()) {
        // staged kernel: every field of a vehicle requested at once (cp.async), rows overlay the staged slots; two CTAs per SM
        typedef void (*staged_fn)(FilterArgs<real>);
        const bool coop = k12_coop_enabled(spec);
        const bool dbm = p->model == SCCAV_MODEL_DBM;                 // the common model gets a dispatch-free slot loop
        const staged_fn sk = spec == SCCAV_SPEC_ELLIPSE
            ? (coop ? filter_step_staged_kernel<real, SCCAV_SPEC_ELLIPSE, 7, true, -1> : filter_step_staged_kernel<real, SCCAV_SPEC_ELLIPSE, 7, false, -1>)
            : coop ? (dbm ? filter_step_staged_kernel<real, SCCAV_SPEC_ELLIPSE_PREP, 6, true, SCCAV_MODEL_DBM>
                          : filter_step_staged_kernel<real, SCCAV_SPEC_ELLIPSE_PREP, 6, true, -1>)
                   : filter_step_staged_kernel<real, SCCAV_SPEC_ELLIPSE_PREP, 6, false, -1>;
        SCCAV_CUDA_CHECK(cudaFuncSetAttribute((const void*)sk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged));
        sk<<<(int)std::min<int64_t>((N + 255) / 256, (int64_t)sm_count() * 2), 256, staged, st>>>(a);
        count_launch();
        SCCAV_CUDA_CHECK(cudaGetLastError());
        return SCCAV_OK;
    }
    const bool dcoop = k12_coop_enabled(-1);
    const filter_fn kern = spec == SCCAV_SPEC_ELLIPSE
        ? (dcoop ? filter_step_kernel<real, SCCAV_SPEC_ELLIPSE, true> : filter_step_kernel<real, SCCAV_SPEC_ELLIPSE, false>)
        : spec == SCCAV_SPEC_ELLIPSE_PREP
        ? (dcoop ? filter_step_kernel<real, SCCAV_SPEC_ELLIPSE_PREP, true> : filter_step_kernel<real, SCCAV_SPEC_ELLIPSE_PREP, false>)
        : (dcoop ? filter_step_kernel<real, SCCAV_SPEC_GENERIC, true> : filter_step_kernel<real, SCCAV_SPEC_GENERIC, false>);
    SCCAV_CUDA_CHECK(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(int)std::min<int64_t>((N + block - 1) / block, (int64_t)sm_count() * 2 * (256 / block)), block, smem, st>>>(a);
    count_launch();
    SCCAV_CUDA_CHECK(cudaGetLastError());
    return SCCAV_OK;
}

// Launch geometry of the persistent rollout: one vehicle per thread, ONE CTA per SM per wave.
// For N <= 148*512 the block is sized so that a single balanced wave covers the batch
// (e.g. N = 65,536 -> 147 CTAs of 448 threads); larger batches run 512-thread CTAs in many waves.
void rollout_geometry(int64_t N, int& grid, int& block) {
    const int sms = sm_count();
    // balanced waves: as few waves of one CTA per SM as SCCAV_ROLLOUT_MAXB allows, every SM the same share of each
    // (131,072 vehicles: 2 waves of 443 per SM -> 448-thread CTAs, not a full wave of 512 and a ragged one)
    const int64_t waves = (N + (int64_t)sms * SCCAV_ROLLOUT_MAXB - 1) / ((int64_t)sms * SCCAV_ROLLOUT_MAXB);
    const int64_t w1 = waves < 1 ? 1 : waves;
    const int64_t per_cta = (N + w1 * sms - 1) / (w1 * sms);
    block = (int)((per_cta + 31) / 32 * 32);
    if (block < 32) block = 32;
    if (block > SCCAV_ROLLOUT_MAXB) block = SCCAV_ROLLOUT_MAXB;
    grid = (int)((N + block - 1) / block);
}

// the six instances of the persistent kernel: course in shared memory or not x slot specialisation
typedef void (*rollout_fn)(RolloutArgs<real>);
rollout_fn rollout_instance(bool course_smem, int spec, bool fast = false, bool fused = false) {
    // + the compile-time (DBM, Stanley, no seekers; fused steering or not) instances of the two ellipse specialisations
    if (fast && course_smem && spec == SCCAV_SPEC_ELLIPSE_PREP)
        return fused ? rollout_kernel<real, true, SCCAV_SPEC_ELLIPSE_PREP, true, 1> : rollout_kernel<real, true, SCCAV_SPEC_ELLIPSE_PREP, true, 0>;
    if (fast && course_smem && spec == SCCAV_SPEC_ELLIPSE)
        return fused ? rollout_kernel<real, true, SCCAV_SPEC_ELLIPSE, true, 1> : rollout_kernel<real, true, SCCAV_SPEC_ELLIPSE, true, 0>;
    if (fast && course_smem && spec == SCCAV_SPEC_GENERIC)
        return fused ? rollout_kernel<real, true, SCCAV_SPEC_GENERIC, true, 1> : rollout_kernel<real, true, SCCAV_SPEC_GENERIC, true, 0>;
    if (spec == SCCAV_SPEC_ELLIPSE)
        return course_smem ? rollout_kernel<real, true, SCCAV_SPEC_ELLIPSE> : rollout_kernel<real, false, SCCAV_SPEC_ELLIPSE>;
    if (spec == SCCAV_SPEC_ELLIPSE_PREP)
        return course_smem ? rollout_kernel<real, true, SCCAV_SPEC_ELLIPSE_PREP> : rollout_kernel<real, false, SCCAV_SPEC_ELLIPSE_PREP>;
    return course_smem ? rollout_kernel<real, true, SCCAV_SPEC_GENERIC> : rollout_kernel<real, false, SCCAV_SPEC_GENERIC>;
}

// grid / block / dynamic shared memory of a rollout launch; the block is capped by what the
// kernel's register count allows (cudaFuncGetAttributes), the course goes to shared memory if it fits
int rollout_launch_shape(int M, int64_t N, int np, bool stan, int spec, int& grid, int& block, size_t& smem, bool& course_smem, bool trig = false) {
    rollout_geometry(N, grid, block);
    cudaFuncAttributes fa;
    SCCAV_CUDA_CHECK(cudaFuncGetAttributes(&fa, (const void*)rollout_instance(true, spec)));
    const int max_block = fa.maxThreadsPerBlock / 32 * 32;
    if (block > max_block) { block = max_block; grid = (int)((N + block - 1) / block); }
    const size_t cap = (size_t)max_smem_optin();
    const RolloutSmem<real> lay(np, true, trig);
    size_t rows_b = (size_t)3 * (M > 0 ? M : 1) * block * sizeof(real);
    course_smem = stan && (lay.course_bytes + rows_b <= cap);
    smem = rows_b + (course_smem ? lay.course_bytes : 0);
    while (smem > cap && block > 32) {            // huge M * block: shrink the CTA
        block >>= 1;
        grid = (int)((N + block - 1) / block);
        rows_b = (size_t)3 * (M > 0 ? M : 1) * block * sizeof(real);
        course_smem = stan && (lay.course_bytes + rows_b <= cap);
        smem = rows_b + (course_smem ? lay.course_bytes : 0);
    }
    return SCCAV_OK;
}

// Several roads in one launch: choose the block so that the CTAs of all roads fill the SMs in as few waves as possible
// (a road's vehicles never share a CTA with another road's: the course lives in the CTA's shared memory).
int roads_geometry(const void* kern, int M, int np, int n_roads, int64_t group, int& grid, int& block, size_t& smem, int& ctas_per_road, bool trig = false) {
    const RolloutSmem<real> lay(np, true, trig);
    const size_t cap = (size_t)max_smem_optin();
    cudaFuncAttributes fa;
    SCCAV_CUDA_CHECK(cudaFuncGetAttributes(&fa, kern));
    const int max_block = std::min(fa.maxThreadsPerBlock / 32 * 32, SCCAV_ROLLOUT_MAXB);
    double best_cost = 1e300;
    block = 0;
    for (int b = max_block; b >= 64; b -= 32) {
        const size_t sm = lay.course_bytes + (size_t)3 * (M > 0 ? M : 1) * b * sizeof(real);
        if (sm > cap) continue;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) { cudaGetLastError(); continue; }
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, b, sm) != cudaSuccess || occ < 1) { cudaGetLastError(); continue; }
        const int64_t cpr = (group + b - 1) / b;
        const int64_t ctas = cpr * n_roads;
        const int64_t waves = (ctas + (int64_t)occ * sm_count() - 1) / ((int64_t)occ * sm_count());
        // a wave of a latency-bound kernel lasts about as long as its resident warps take turns: waves x resident threads --
        // but never less than what ~10 warps take (below that an SM idles between the dependent instructions of its few
        // warps: 1,024 CTAs of 64 threads in 7 waves were measured at 27 ms against 11 ms for 128 CTAs of 512 in one)
        const int64_t resident = std::min<int64_t>((int64_t)occ * b, (ctas + sm_count() - 1) / sm_count() * b);
        const double cost = (double)waves * (double)std::max<int64_t>(resident, 320);
        if (cost < best_cost) { best_cost = cost; block = b; smem = sm; ctas_per_road = (int)cpr; grid = (int)ctas; }
    }
    if (!block) { set_error("a road of %d points does not fit in shared memory", np); return SCCAV_EINVAL; }
    return SCCAV_OK;
}

int do_rollout(const sccav_params* p, const uint8_t* slot_desc, int32_t M, int64_t N, int32_t T, const real* state,
               real* obst, const real* cx, const real* cy, const real* cyaw, int32_t P, const sccav_pervehicle* pv,
               const sccav_rollout_out* out, cudaStream_t st, int32_t n_roads = 0, const int32_t* road_np = nullptr) {
    int rc = check_common(p, slot_desc, M, N, true);
    if (rc) return rc;
    if (T < 0) { set_error("T < 0"); return SCCAV_EINVAL; }
    if (p->record_stride < 0) { set_error("record_stride < 0"); return SCCAV_EINVAL; }
    if (p->model == SCCAV_MODEL_DUM) { set_error("model DUM has no closed loop (the reference has no unicycle plant on this path)"); return SCCAV_EINVAL; }
    if (p->model == SCCAV_MODEL_SADBM) { set_error("model SADBM has no closed loop here (filter entry points only)"); return SCCAV_EINVAL; }
    if (p->flags & SCCAV_FLAG_BETA_IO) { set_error("SCCAV_FLAG_BETA_IO: filter entry points only (the rollout's nominal controller produces delta)"); return SCCAV_EINVAL; }
    if (N == 0) return SCCAV_OK;
    if (!state || !out || !out->state) { set_error("state / out->state is NULL"); return SCCAV_EINVAL; }
    if (M > 0 && !obst) { set_error("obst is NULL"); return SCCAV_EINVAL; }
    const bool stan = p->nominal == SCCAV_NOMINAL_STANLEY;
    if (stan && (P < 1 || !cx || !cy || !cyaw)) { set_error("Stanley nominal control needs a course (P >= 1)"); return SCCAV_EINVAL; }
    if (n_roads > 0) {
        if (!stan) { set_error("several roads need the Stanley nominal controller"); return SCCAV_EINVAL; }
        if (!road_np) { set_error("road_np is NULL"); return SCCAV_EINVAL; }
        if (N % n_roads != 0) { set_error("N = %lld vehicles do not split evenly over %d roads", (long long)N, n_roads); return SCCAV_EINVAL; }
    }
    RolloutArgs<real> a;
    a.n_roads = n_roads; a.road_stride = P; a.road_np = road_np; a.group = n_roads > 0 ? N / n_roads : 0; a.ctas_per_road = 0;
    a.P = convert(p); a.sd = make_desc(slot_desc, M); a.M = M; a.N = N; a.T_steps = T; a.np = stan ? P : 0;
    a.state = state; a.obst = obst; a.cx = cx; a.cy = cy; a.cyaw = cyaw; a.pv = make_pv(pv);
    a.o_state = (real*)out->state; a.o_steps = out->steps; a.o_tidx = out->target_idx; a.o_nact = out->n_active;
    a.o_ninf = out->n_infeasible; a.o_hmin = (real*)out->h_min; a.o_bmin = (real*)out->beta_min;
    a.o_bmax = (real*)out->beta_max; a.o_bint = (real*)out->beta_int; a.o_traj = (real*)out->traj;
    a.o_tridx = out->traj_idx; a.o_trmask = out->traj_mask;
    a.o_evals = out->n_evals;
    int grid, block;
    size_t smem;
    bool course_smem;
    // SCCAV_FLAG_PREPARED_ROWS: ingest the ELLIPSE slots once (KP) into a stream-ordered scratch and
    // run the closed loop on their prepared form
    const bool filt = M > 0 && p->model != SCCAV_MODEL_NONE;
    uint8_t desc2[SCCAV_MAX_ROWS];
    for (int m = 0; m < M; ++m) desc2[m] = slot_desc[m];
    void* prep = nullptr;
    bool any_ellipse = false;
    for (int m = 0; m < M; ++m) any_ellipse |= (slot_desc[m] & SCCAV_SLOT_TYPE_MASK) == SCCAV_SLOT_ELLIPSE;
    if (filt && any_ellipse && (p->flags & SCCAV_FLAG_PREPARED_ROWS) && !p->seeker) {
        SCCAV_CUDA_CHECK(pool_alloc(&prep, (size_t)M * SCCAV_NFIELD * (size_t)N * sizeof(real), st));
        rc = do_prepare(slot_desc, M, N, obst, (real*)prep, desc2, st);
        if (rc) { cudaFreeAsync(prep, st); return rc; }
        a.obst = (real*)prep;
        a.sd = make_desc(desc2, M);
        any_ellipse = false;
    }
    const int spec = filt ? choose_spec(desc2, M) : SCCAV_SPEC_GENERIC;
    // (the compile-time ELLIPSE instances read the weights and the target speed from the launch parameters: per-vehicle ones
    // take the general instances)
    const bool uniform_w = !pv || (!pv->alpha && !pv->R && !pv->target_speed);
    const bool fast = p->model == SCCAV_MODEL_DBM && stan && !p->seeker && (spec == SCCAV_SPEC_GENERIC || uniform_w);
    // (the fused-steer compile-time instances stage (sin, cos) of the course yaws: twice the bytes of that table)
    const bool trig = fast && (p->flags & SCCAV_FLAG_FUSED_STEER) != 0;
    rc = rollout_launch_shape(M, N, a.np, stan, spec, grid, block, smem, course_smem, trig);
    // (a course that fits only without that table runs the general instance, which reads the flag at run time)
    if (!rc && trig && !course_smem && n_roads == 0) rc = rollout_launch_shape(M, N, a.np, stan, spec, grid, block, smem, course_smem, false);
    if (rc) { if (prep) cudaFreeAsync(prep, st); return rc; }
    // scratch for the loop-invariant terms of canonical ellipses: stream-ordered, lives for this launch
    a.pre = nullptr;
    // (+ the reciprocal half axes of RADIAL slots when the rows are prepared; those launches get it for one step as well)
    bool radial_prep = false;
    if (p->flags & SCCAV_FLAG_PREPARED_ROWS)
        for (int m = 0; m < M; ++m) radial_prep |= (desc2[m] & SCCAV_SLOT_TYPE_MASK) == SCCAV_SLOT_RADIAL;
    if (((any_ellipse && T >= 2) || (radial_prep && T >= 1)) && filt) {
        void* scratch = nullptr;
        cudaError_t me = pool_alloc(&scratch, (size_t)M * SCCAV_NPRE * (size_t)N * sizeof(real), st);
        if (me != cudaSuccess) { if (prep) cudaFreeAsync(prep, st); SCCAV_CUDA_CHECK(me); }
        a.pre = (real*)scratch;
    }
    rollout_fn kern = rollout_instance(course_smem, spec, fast, (p->flags & SCCAV_FLAG_FUSED_STEER) != 0);
    // compile-time prepared instance on static ellipses: the loop reads the packed symmetric form (pack_sym_kernel)
    a.sym = nullptr;
    void* symbuf = nullptr;
    if (prep && fast && (course_smem || n_roads > 0) && spec == SCCAV_SPEC_ELLIPSE_PREP && (desc2[0] & SCCAV_SLOT_STATIC) && !getenv("SCCAV_NO_SYM")) {
        cudaError_t me = pool_alloc(&symbuf, (size_t)M * 6 * (size_t)N * sizeof(real), st);
        if (me != cudaSuccess) { if (a.pre) cudaFreeAsync(a.pre, st); cudaFreeAsync(prep, st); SCCAV_CUDA_CHECK(me); }
        SymArgs<real> sa;
        sa.M = M; sa.N = N; sa.prep = (const real*)prep; sa.sym = (Real<real>::T2*)symbuf;
        pack_sym_kernel<real><<<stream_grid((int64_t)M * N, 256), 256, 0, st>>>(sa);
        count_launch();
        a.sym = symbuf;
    }
    // compile-time canonical-ellipse instance: room for the paired copy of the fields and hoisted terms (filled by the kernel)
    if (!symbuf && a.pre && fast && (course_smem || n_roads > 0) && spec == SCCAV_SPEC_ELLIPSE && !getenv("SCCAV_NO_SYM")) {
        cudaError_t me = pool_alloc(&symbuf, (size_t)M * 10 * (size_t)N * sizeof(real), st);
        if (me != cudaSuccess) { cudaFreeAsync(a.pre, st); if (prep) cudaFreeAsync(prep, st); SCCAV_CUDA_CHECK(me); }
        a.sym = symbuf;
    }
    if (n_roads > 0) {
        kern = rollout_instance(true, spec, fast, (p->flags & SCCAV_FLAG_FUSED_STEER) != 0);
        rc = roads_geometry((const void*)kern, M, a.np, n_roads, a.group, grid, block, smem, a.ctas_per_road, trig);
        if (rc) { if (a.pre) cudaFreeAsync(a.pre, st); if (prep) cudaFreeAsync(prep, st); if (symbuf) cudaFreeAsync(symbuf, st); return rc; }
    }
    if (getenv("SCCAV_DEBUG_LAUNCH")) fprintf(stderr, "[sccav] rollout launch: grid %d block %d smem %zu roads %d ctas_per_road %d spec %d fast %d\n", grid, block, smem, n_roads, a.ctas_per_road, spec, (int)fast);
    cudaError_t le = cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (le == cudaSuccess) { kern<<<grid, block, smem, st>>>(a); le = cudaGetLastError(); }
    count_launch();
    if (a.pre) cudaFreeAsync(a.pre, st);
    if (prep) cudaFreeAsync(prep, st);
    if (symbuf) cudaFreeAsync(symbuf, st);
    SCCAV_CUDA_CHECK(le);
    return SCCAV_OK;
}

// ---- host-buffer variants: stream-ordered device allocations, H2D, kernel, D2H, sync
struct DevBuf {
    void* p = nullptr;
    cudaStream_t st;
    explicit DevBuf(cudaStream_t s) : st(s) {}
    ~DevBuf() { if (p) cudaFreeAsync(p, st); }
    cudaError_t alloc(size_t bytes) { return pool_alloc(&p, bytes, st); }
    cudaError_t upload(const void* src, size_t bytes) {
        cudaError_t e = alloc(bytes);
        if (e != cudaSuccess) return e;
        return cudaMemcpyAsync(p, src, bytes, cudaMemcpyHostToDevice, st);
    }
};

// ---- pipelined host API: device buffers, streams and events owned by a handle; the upload of submission i + 1 and
// the download of submission i - 1 run beside the kernel of submission i
struct PipeSlot {
    void* in[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};        // state, obst, alpha, R, target_speed, count
    void* out[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t h2d_done = nullptr, kernel_done = nullptr, d2h_done = nullptr;
};

struct Pipeline {
    uint32_t magic = 0x53434356u + sizeof(real);
    sccav_params prm;
    uint8_t desc[SCCAV_MAX_ROWS];
    int M = 0, T = 0, P = 0, depth = 0, device = 0;
    int64_t N = 0;
    void *cx = nullptr, *cy = nullptr, *cyaw = nullptr, *obst_resident = nullptr;
    cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
    PipeSlot* slots = nullptr;
    int64_t next = 0;
};

void pipeline_free(Pipeline* pl) {
    if (!pl) return;
    if (pl->s_in) cudaStreamSynchronize(pl->s_in);
    if (pl->s_run) cudaStreamSynchronize(pl->s_run);
    if (pl->s_out) cudaStreamSynchronize(pl->s_out);
    for (int i = 0; pl->slots && i < pl->depth; ++i) {
        PipeSlot& s = pl->slots[i];
        for (void* p : s.in) if (p) cudaFree(p);
        for (void* p : s.out) if (p) cudaFree(p);
        if (s.h2d_done) cudaEventDestroy(s.h2d_done);
        if (s.kernel_done) cudaEventDestroy(s.kernel_done);
        if (s.d2h_done) cudaEventDestroy(s.d2h_done);
    }
    delete[] pl->slots;
    for (void* p : {pl->cx, pl->cy, pl->cyaw, pl->obst_resident}) if (p) cudaFree(p);
    if (pl->s_in) cudaStreamDestroy(pl->s_in);
    if (pl->s_run) cudaStreamDestroy(pl->s_run);
    if (pl->s_out) cudaStreamDestroy(pl->s_out);
    pl->magic = 0;
    delete pl;
}

}  // namespace
}  // namespace sccav

extern "C" {

int SCCAV_FN(sccav_pipeline_create_)(const sccav_params* p, const uint8_t* slot_desc, int32_t M, int64_t N, int32_t T,
                                     const SCCAV_REAL* course_x, const SCCAV_REAL* course_y, const SCCAV_REAL* course_yaw, int32_t P,
                                     const SCCAV_REAL* obst_resident, int32_t depth, void** handle_out) {
    using namespace sccav;
    typedef SCCAV_REAL real;
    if (!handle_out) { set_error("handle_out is NULL"); return SCCAV_EINVAL; }
    *handle_out = nullptr;
    int rc = check_common(p, slot_desc, M, N, true);
    if (rc) return rc;
    if (N < 1 || T < 0) { set_error("N < 1 or T < 0"); return SCCAV_EINVAL; }
    if (depth < 1 || depth > 8) { set_error("depth must be in [1, 8], got %d", depth); return SCCAV_EINVAL; }
    if (p->record_stride != 0) { set_error("the pipelined API returns the per-vehicle summaries only (record_stride must be 0)"); return SCCAV_EINVAL; }
    const bool stan = p->nominal == SCCAV_NOMINAL_STANLEY;
    if (stan && (P < 1 || !course_x || !course_y || !course_yaw)) { set_error("Stanley nominal control needs a course"); return SCCAV_EINVAL; }
    if (obst_resident && p->seeker) { set_error("moving obstacles (params.seeker) cannot be resident: pass them with every submission"); return SCCAV_EINVAL; }
    Pipeline* pl = new (std::nothrow) Pipeline();
    if (!pl) { set_error("out of host memory"); return SCCAV_ENOMEM; }
    pl->prm = *p; pl->M = M; pl->N = N; pl->T = T; pl->P = stan ? P : 0; pl->depth = depth;
    memset(pl->desc, 0, sizeof(pl->desc));
    for (int m = 0; m < M; ++m) pl->desc[m] = slot_desc[m];
    cudaGetDevice(&pl->device);
    const size_t n = (size_t)N;
    const size_t obst_b = (size_t)M * SCCAV_NFIELD * n * sizeof(real);
#define SCCAV_PL_CHECK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_error("%s failed: %s", #expr, cudaGetErrorString(_e)); pipeline_free(pl); return SCCAV_ECUDA; } } while (0)
    SCCAV_PL_CHECK(cudaStreamCreateWithFlags(&pl->s_in, cudaStreamNonBlocking));
    SCCAV_PL_CHECK(cudaStreamCreateWithFlags(&pl->s_run, cudaStreamNonBlocking));
    SCCAV_PL_CHECK(cudaStreamCreateWithFlags(&pl->s_out, cudaStreamNonBlocking));
    if (pl->P > 0) {
        const size_t cb = (size_t)P * sizeof(real);
        SCCAV_PL_CHECK(cudaMalloc(&pl->cx, cb)); SCCAV_PL_CHECK(cudaMalloc(&pl->cy, cb)); SCCAV_PL_CHECK(cudaMalloc(&pl->cyaw, cb));
        SCCAV_PL_CHECK(cudaMemcpyAsync(pl->cx, course_x, cb, cudaMemcpyHostToDevice, pl->s_in));
        SCCAV_PL_CHECK(cudaMemcpyAsync(pl->cy, course_y, cb, cudaMemcpyHostToDevice, pl->s_in));
        SCCAV_PL_CHECK(cudaMemcpyAsync(pl->cyaw, course_yaw, cb, cudaMemcpyHostToDevice, pl->s_in));
    }
    if (obst_resident && M > 0) {
        SCCAV_PL_CHECK(cudaMalloc(&pl->obst_resident, obst_b));
        SCCAV_PL_CHECK(cudaMemcpyAsync(pl->obst_resident, obst_resident, obst_b, cudaMemcpyHostToDevice, pl->s_in));
    }
    pl->slots = new (std::nothrow) PipeSlot[depth];
    if (!pl->slots) { pipeline_free(pl); set_error("out of host memory"); return SCCAV_ENOMEM; }
    const size_t in_b[6] = {4 * n * sizeof(real), pl->obst_resident ? 0 : obst_b, n * sizeof(real), 4 * n * sizeof(real), n * sizeof(real), n * 4};
    const size_t out_b[10] = {4 * n * sizeof(real), n * 4, n * 4, n * 4, n * 4, n * sizeof(real), n * sizeof(real), n * sizeof(real), n * sizeof(real), n * 4};
    for (int i = 0; i < depth; ++i) {
        PipeSlot& s = pl->slots[i];
        for (int k = 0; k < 6; ++k) if (in_b[k]) SCCAV_PL_CHECK(cudaMalloc(&s.in[k], in_b[k]));
        for (int k = 0; k < 10; ++k) SCCAV_PL_CHECK(cudaMalloc(&s.out[k], out_b[k]));
        SCCAV_PL_CHECK(cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming));
        SCCAV_PL_CHECK(cudaEventCreateWithFlags(&s.kernel_done, cudaEventDisableTiming));
        SCCAV_PL_CHECK(cudaEventCreateWithFlags(&s.d2h_done, cudaEventDisableTiming));
    }
    SCCAV_PL_CHECK(cudaStreamSynchronize(pl->s_in));
#undef SCCAV_PL_CHECK
    *handle_out = pl;
    return SCCAV_OK;
}

int SCCAV_FN(sccav_pipeline_submit_)(void* handle, const SCCAV_REAL* state, const SCCAV_REAL* obst, const sccav_pervehicle* pv,
                                     const sccav_rollout_out* out, int64_t* ticket_out) {
    using namespace sccav;
    typedef SCCAV_REAL real;
    Pipeline* pl = (Pipeline*)handle;
    if (!pl || pl->magic != 0x53434356u + sizeof(real)) { set_error("not a pipeline handle of this precision"); return SCCAV_EINVAL; }
    if (!state || !out || !out->state) { set_error("state / out->state is NULL"); return SCCAV_EINVAL; }
    if (pl->M > 0 && !obst && !pl->obst_resident) { set_error("obst is NULL and the pipeline holds no resident obstacles"); return SCCAV_EINVAL; }
    if (out->traj || out->traj_idx || out->traj_mask) { set_error("the pipelined API returns the per-vehicle summaries only"); return SCCAV_EINVAL; }
    PipeSlot& s = pl->slots[pl->next % pl->depth];
    const size_t n = (size_t)pl->N;
    const size_t obst_b = (size_t)pl->M * SCCAV_NFIELD * n * sizeof(real);
    // inputs of this slot were last read by the kernel of submission (next - depth)
    SCCAV_CUDA_CHECK(cudaStreamWaitEvent(pl->s_in, s.kernel_done, 0));
    SCCAV_CUDA_CHECK(cudaMemcpyAsync(s.in[0], state, 4 * n * sizeof(real), cudaMemcpyHostToDevice, pl->s_in));
    real* d_obst = (real*)pl->obst_resident;
    if (pl->M > 0 && obst && !pl->obst_resident) {
        SCCAV_CUDA_CHECK(cudaMemcpyAsync(s.in[1], obst, obst_b, cudaMemcpyHostToDevice, pl->s_in));
        d_obst = (real*)s.in[1];
    }
    sccav_pervehicle dpv = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (pv && pv->alpha) { SCCAV_CUDA_CHECK(cudaMemcpyAsync(s.in[2], pv->alpha, n * sizeof(real), cudaMemcpyHostToDevice, pl->s_in)); dpv.alpha = s.in[2]; }
    if (pv && pv->R) { SCCAV_CUDA_CHECK(cudaMemcpyAsync(s.in[3], pv->R, 4 * n * sizeof(real), cudaMemcpyHostToDevice, pl->s_in)); dpv.R = s.in[3]; }
    if (pv && pv->target_speed) { SCCAV_CUDA_CHECK(cudaMemcpyAsync(s.in[4], pv->target_speed, n * sizeof(real), cudaMemcpyHostToDevice, pl->s_in)); dpv.target_speed = s.in[4]; }
    if (pv && pv->count) { SCCAV_CUDA_CHECK(cudaMemcpyAsync(s.in[5], pv->count, n * 4, cudaMemcpyHostToDevice, pl->s_in)); dpv.count = (const int32_t*)s.in[5]; }
    SCCAV_CUDA_CHECK(cudaEventRecord(s.h2d_done, pl->s_in));
    // kernel: after this slot's upload, and after the download that last read this slot's outputs
    SCCAV_CUDA_CHECK(cudaStreamWaitEvent(pl->s_run, s.h2d_done, 0));
    SCCAV_CUDA_CHECK(cudaStreamWaitEvent(pl->s_run, s.d2h_done, 0));
    sccav_rollout_out dout;
    memset(&dout, 0, sizeof(dout));
    dout.state = s.out[0];
    if (out->steps) dout.steps = (int32_t*)s.out[1];
    if (out->target_idx) dout.target_idx = (int32_t*)s.out[2];
    if (out->n_active) dout.n_active = (int32_t*)s.out[3];
    if (out->n_infeasible) dout.n_infeasible = (int32_t*)s.out[4];
    if (out->h_min) dout.h_min = s.out[5];
    if (out->beta_min) dout.beta_min = s.out[6];
    if (out->beta_max) dout.beta_max = s.out[7];
    if (out->beta_int) dout.beta_int = s.out[8];
    if (out->n_evals) dout.n_evals = (int32_t*)s.out[9];
    int rc = do_rollout(&pl->prm, pl->desc, pl->M, pl->N, pl->T, (const real*)s.in[0], d_obst, (const real*)pl->cx, (const real*)pl->cy,
                        (const real*)pl->cyaw, pl->P, &dpv, &dout, pl->s_run);
    if (rc) return rc;
    SCCAV_CUDA_CHECK(cudaEventRecord(s.kernel_done, pl->s_run));
    // download
    SCCAV_CUDA_CHECK(cudaStreamWaitEvent(pl->s_out, s.kernel_done, 0));
    SCCAV_CUDA_CHECK(cudaMemcpyAsync(out->state, s.out[0], 4 * n * sizeof(real), cudaMemcpyDeviceToHost, pl->s_out));
#define SCCAV_PL_BACK(field, k, bytes) if (out->field) SCCAV_CUDA_CHECK(cudaMemcpyAsync(out->field, s.out[k], bytes, cudaMemcpyDeviceToHost, pl->s_out));
    SCCAV_PL_BACK(steps, 1, n * 4)
    SCCAV_PL_BACK(target_idx, 2, n * 4)
    SCCAV_PL_BACK(n_active, 3, n * 4)
    SCCAV_PL_BACK(n_infeasible, 4, n * 4)
    SCCAV_PL_BACK(h_min, 5, n * sizeof(real))
    SCCAV_PL_BACK(beta_min, 6, n * sizeof(real))
    SCCAV_PL_BACK(beta_max, 7, n * sizeof(real))
    SCCAV_PL_BACK(beta_int, 8, n * sizeof(real))
    SCCAV_PL_BACK(n_evals, 9, n * 4)
#undef SCCAV_PL_BACK
    if (pl->M > 0 && pl->prm.seeker && obst) SCCAV_CUDA_CHECK(cudaMemcpyAsync((void*)obst, s.in[1], obst_b, cudaMemcpyDeviceToHost, pl->s_out));
    SCCAV_CUDA_CHECK(cudaEventRecord(s.d2h_done, pl->s_out));
    if (ticket_out) *ticket_out = pl->next;
    ++pl->next;
    return SCCAV_OK;
}

int SCCAV_FN(sccav_pipeline_wait_)(void* handle, int64_t ticket) {
    using namespace sccav;
    Pipeline* pl = (Pipeline*)handle;
    if (!pl || pl->magic != 0x53434356u + sizeof(SCCAV_REAL)) { set_error("not a pipeline handle of this precision"); return SCCAV_EINVAL; }
    if (ticket < 0 || ticket >= pl->next || ticket < pl->next - pl->depth) {
        set_error("ticket %lld is not one of the last %d submissions", (long long)ticket, pl->depth);
        return SCCAV_EINVAL;
    }
    SCCAV_CUDA_CHECK(cudaEventSynchronize(pl->slots[ticket % pl->depth].d2h_done));
    return SCCAV_OK;
}

int SCCAV_FN(sccav_pipeline_destroy_)(void* handle) {
    using namespace sccav;
    Pipeline* pl = (Pipeline*)handle;
    if (!pl) return SCCAV_OK;
    if (pl->magic != 0x53434356u + sizeof(SCCAV_REAL)) { set_error("not a pipeline handle of this precision"); return SCCAV_EINVAL; }
    pipeline_free(pl);
    return SCCAV_OK;
}

int SCCAV_FN(sccav_barrier_rows_)(const sccav_params* p, const uint8_t* slot_desc, int32_t M, int64_t N,
                                  const SCCAV_REAL* state, const SCCAV_REAL* obst, const sccav_pervehicle* pv,
                                  SCCAV_REAL* A_out, SCCAV_REAL* b_out, SCCAV_REAL* h_out, void* stream) {
    return sccav::do_barrier_rows(p, slot_desc, M, N, state, obst, pv, A_out, b_out, h_out, (cudaStream_t)stream);
}

int SCCAV_FN(sccav_prepare_obstacles_)(const uint8_t* slot_desc, int32_t M, int64_t N, const SCCAV_REAL* obst_in,
                                        SCCAV_REAL* obst_out, uint8_t* slot_desc_out, void* stream) {
    return sccav::do_prepare(slot_desc, M, N, obst_in, obst_out, slot_desc_out, (cudaStream_t)stream);
}

int SCCAV_FN(sccav_ingest_boxes_)(int32_t obs_type, int32_t mode, double buffer, int32_t M, int32_t K, int64_t N,
                                   const int32_t* box_id, const SCCAV_REAL* box, int32_t* slot_id, SCCAV_REAL* obst,
                                   int32_t* count, int32_t* dropped, void* stream) {
    return sccav::do_ingest(obs_type, mode, buffer, M, K, N, box_id, box, slot_id, obst, count, dropped, (cudaStream_t)stream);
}

int SCCAV_FN(sccav_actuator_shaping_)(int64_t N, const SCCAV_REAL* u, double max_steer, double rate, int32_t flags,
                                       SCCAV_REAL* throttle_prev, SCCAV_REAL* brake_prev, SCCAV_REAL* throttle_out,
                                       SCCAV_REAL* brake_out, SCCAV_REAL* steer_out, void* stream) {
    using namespace sccav;
    if (N < 0) { set_error("N < 0"); return SCCAV_EINVAL; }
    if (N == 0) return SCCAV_OK;
    if (!u || !throttle_prev || !brake_prev) { set_error("NULL array argument"); return SCCAV_EINVAL; }
    ActuatorArgs<SCCAV_REAL> a;
    a.N = N; a.u = u; a.max_steer = (SCCAV_REAL)max_steer; a.rate = (SCCAV_REAL)rate; a.flags = flags;
    a.thr_prev = throttle_prev; a.brk_prev = brake_prev; a.thr = throttle_out; a.brk = brake_out; a.steer = steer_out;
    const int block = 256;
    actuator_kernel<SCCAV_REAL><<<stream_grid(N, block), block, 0, (cudaStream_t)stream>>>(a);
    count_launch();
    SCCAV_CUDA_CHECK(cudaGetLastError());
    return SCCAV_OK;
}

int SCCAV_FN(sccav_spline_course_)(int32_t C, int32_t K, const SCCAV_REAL* wx, const SCCAV_REAL* wy, double ds, int32_t P_max,
                                    SCCAV_REAL* cx, SCCAV_REAL* cy, SCCAV_REAL* cyaw, SCCAV_REAL* ck, int32_t* np_out, void* stream) {
    using namespace sccav;
    if (C < 0 || P_max < 1) { set_error("C < 0 or P_max < 1"); return SCCAV_EINVAL; }
    if (K < 2 || K > SCCAV_MAX_KNOTS) { set_error("a course needs 2 .. %d way-points, got %d", SCCAV_MAX_KNOTS, K); return SCCAV_EINVAL; }
    if (!(ds > 0.0)) { set_error("ds must be positive"); return SCCAV_EINVAL; }
    if (C == 0) return SCCAV_OK;
    if (C > 65535) { set_error("at most 65535 courses per call"); return SCCAV_EINVAL; }
    if (!wx || !wy || !cx || !cy || !cyaw || !np_out) { set_error("NULL array argument"); return SCCAV_EINVAL; }
    CourseArgs<SCCAV_REAL> a;
    a.C = C; a.K = K; a.P_max = P_max; a.ds = (SCCAV_REAL)ds; a.wx = wx; a.wy = wy;
    a.cx = cx; a.cy = cy; a.cyaw = cyaw; a.ck = ck; a.np_out = np_out;
    dim3 grid((unsigned)((P_max + 255) / 256), (unsigned)C);
    spline_course_kernel<SCCAV_REAL><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    count_launch();
    SCCAV_CUDA_CHECK(cudaGetLastError());
    return SCCAV_OK;
}

int SCCAV_FN(sccav_fit_lanes_)(int32_t C, int32_t K, const SCCAV_REAL* x, const SCCAV_REAL* y, const SCCAV_REAL* sigma,
                                const int32_t* count, int32_t degree, SCCAV_REAL* coeffs, int32_t* status, void* stream) {
    using namespace sccav;
    if (C < 0 || K < 0) { set_error("C < 0 or K < 0"); return SCCAV_EINVAL; }
    if (degree < 1 || degree > 5) { set_error("degree must be in [1, 5] (a LANE slot holds 6 coefficients), got %d", degree); return SCCAV_EINVAL; }
    if (C == 0) return SCCAV_OK;
    if (!x || !y || !coeffs) { set_error("NULL array argument"); return SCCAV_EINVAL; }
    LaneFitArgs<SCCAV_REAL> a;
    a.C = C; a.K = K; a.degree = degree; a.x = x; a.y = y; a.sigma = sigma; a.count = count; a.coeffs = coeffs; a.status = status;
    const int block = 128;
    lane_fit_kernel<SCCAV_REAL><<<stream_grid(C, block), block, 0, (cudaStream_t)stream>>>(a);
    count_launch();
    SCCAV_CUDA_CHECK(cudaGetLastError());
    return SCCAV_OK;
}

int SCCAV_FN(sccav_barrier_partials_)(const uint8_t* slot_desc, int32_t M, int64_t N, const SCCAV_REAL* state,
                                      const SCCAV_REAL* obst, SCCAV_REAL* out, void* stream) {
    return sccav::do_barrier_partials(slot_desc, M, N, state, obst, out, (cudaStream_t)stream);
}

int SCCAV_FN(sccav_stanley_control_)(const sccav_params* p, int64_t N, const SCCAV_REAL* state, const SCCAV_REAL* front,
                                     const SCCAV_REAL* course_x, const SCCAV_REAL* course_y, const SCCAV_REAL* course_yaw,
                                     int32_t P, int32_t* target_idx, SCCAV_REAL* delta_out, SCCAV_REAL* err_out, void* stream) {
    return sccav::do_stanley(p, N, state, front, course_x, course_y, course_yaw, P, target_idx, delta_out, err_out,
                             (cudaStream_t)stream);
}

int SCCAV_FN(sccav_qp2_solve_)(const sccav_params* p, int32_t M, int64_t N, const SCCAV_REAL* A, const SCCAV_REAL* b,
                               const SCCAV_REAL* r, const sccav_pervehicle* pv, SCCAV_REAL* u_out, uint32_t* active_out,
                               uint8_t* status_out, int32_t warp_per_problem, void* stream) {
    return sccav::do_qp2(p, M, N, A, b, r, pv, u_out, active_out, status_out, warp_per_problem, (cudaStream_t)stream);
}

int SCCAV_FN(sccav_filter_step_)(const sccav_params* p, const uint8_t* slot_desc, int32_t M, int64_t N,
                                 const SCCAV_REAL* state, const SCCAV_REAL* obst, const SCCAV_REAL* u_ref,
                                 const sccav_pervehicle* pv, SCCAV_REAL* u_out, uint32_t* active_out,
                                 uint8_t* status_out, SCCAV_REAL* h_min_out, void* stream) {
    return sccav::do_filter_step(p, slot_desc, M, N, state, obst, u_ref, pv, u_out, active_out, status_out, h_min_out,
                                 (cudaStream_t)stream);
}

int SCCAV_FN(sccav_rollout_)(const sccav_params* p, const uint8_t* slot_desc, int32_t M, int64_t N, int32_t T,
                             const SCCAV_REAL* state, SCCAV_REAL* obst, const SCCAV_REAL* course_x,
                             const SCCAV_REAL* course_y, const SCCAV_REAL* course_yaw, int32_t P,
                             const sccav_pervehicle* pv, const sccav_rollout_out* out, void* stream) {
    return sccav::do_rollout(p, slot_desc, M, N, T, state, obst, course_x, course_y, course_yaw, P, pv, out,
                             (cudaStream_t)stream);
}

int SCCAV_FN(sccav_rollout_roads_)(const sccav_params* p, const uint8_t* slot_desc, int32_t M, int64_t N, int32_t T,
                                   const SCCAV_REAL* state, SCCAV_REAL* obst, const SCCAV_REAL* course_x,
                                   const SCCAV_REAL* course_y, const SCCAV_REAL* course_yaw, int32_t P_max, int32_t C,
                                   const int32_t* course_np, const sccav_pervehicle* pv, const sccav_rollout_out* out, void* stream) {
    if (C < 1) { sccav::set_error("C < 1"); return SCCAV_EINVAL; }
    return sccav::do_rollout(p, slot_desc, M, N, T, state, obst, course_x, course_y, course_yaw, P_max, pv, out,
                             (cudaStream_t)stream, C, course_np);
}

int SCCAV_FN(sccav_drive_ticks_)(const sccav_params* p, const sccav_drive_params* dp, const uint8_t* slot_desc, int32_t M, int32_t n_fixed,
                                 int64_t N, int32_t T, const SCCAV_REAL* state0, const SCCAV_REAL* ego, int32_t K, const int32_t* box_id,
                                 const SCCAV_REAL* box, const SCCAV_REAL* dt, SCCAV_REAL* obst, const SCCAV_REAL* traj_x,
                                 const SCCAV_REAL* traj_y, const SCCAV_REAL* traj_yaw, const SCCAV_REAL* traj_v, int32_t P,
                                 const sccav_pervehicle* pv, int32_t* target_idx, SCCAV_REAL* carry, SCCAV_REAL* act_out,
                                 SCCAV_REAL* u_out, uint32_t* active_out, int32_t* target_idx_out, SCCAV_REAL* state_out, void* stream) {
    using namespace sccav;
    typedef SCCAV_REAL real;
    int rc = check_common(p, slot_desc, M, N, true);
    if (rc) return rc;
    if (!dp) { set_error("drive params is NULL"); return SCCAV_EINVAL; }
    if (p->model != SCCAV_MODEL_DBM) { set_error("the driver tick uses DBM_CBF_2DS (model DBM)"); return SCCAV_EINVAL; }
    if (p->flags & SCCAV_FLAG_BETA_IO) { set_error("SCCAV_FLAG_BETA_IO: filter entry points only"); return SCCAV_EINVAL; }
    if (T < 0 || K < 0 || n_fixed < 0 || n_fixed > M) { set_error("T < 0, K < 0 or n_fixed outside [0, M]"); return SCCAV_EINVAL; }
    if (P < 1 || !traj_x || !traj_y || !traj_yaw || !traj_v) { set_error("a trajectory (x, y, yaw, v)[P >= 1] is required"); return SCCAV_EINVAL; }
    if (N == 0 || T == 0) return SCCAV_OK;
    if (!state0 && !ego) { set_error("either the ego stream or the initial state of the stand-in plant is required"); return SCCAV_EINVAL; }
    if (K > 0 && (!box_id || !box)) { set_error("box_id / box is NULL"); return SCCAV_EINVAL; }
    if ((M > 0 && !obst) || !target_idx || !carry || !act_out) { set_error("NULL array argument"); return SCCAV_EINVAL; }
    for (int m = n_fixed; m < M && K > 0; ++m)
        if ((slot_desc[m] & SCCAV_SLOT_TYPE_MASK) != SCCAV_SLOT_CONE || (slot_desc[m] & SCCAV_SLOT_SHARED)) {
            set_error("slot %d is rebuilt from the boxes every tick: it must be a per-vehicle CONE slot", m);
            return SCCAV_EINVAL;
        }
    DriveArgs<real> a;
    a.P = convert(p); a.sd = make_desc(slot_desc, M); a.M = M; a.n_fixed = n_fixed; a.K = K; a.T_ticks = T; a.np = P; a.N = N;
    a.kp = (real)dp->kp; a.ki = (real)dp->ki; a.kd = (real)dp->kd; a.rad_to_steer = (real)dp->rad_to_steer;
    a.max_steer_cmd = (real)dp->max_steer_cmd; a.rate = (real)dp->rate; a.cone_buffer = (real)dp->cone_buffer; a.act_flags = dp->act_flags;
    a.state0 = state0; a.ego = ego; a.box_id = K > 0 ? box_id : nullptr; a.box = box; a.dt = dt; a.obst = obst;
    a.cx = traj_x; a.cy = traj_y; a.cyaw = traj_yaw; a.cv = traj_v; a.pv = make_pv(pv);
    a.target_idx = target_idx; a.carry = carry; a.act = act_out; a.u = u_out; a.mask = active_out; a.tidx = target_idx_out; a.o_state = state_out;
    const RolloutSmem<real> lay(P, true);
    int block = 256;
    size_t smem = lay.course_bytes + (size_t)3 * (M > 0 ? M : 1) * block * sizeof(real);
    while (smem > (size_t)max_smem_optin() && block > 32) { block >>= 1; smem = lay.course_bytes + (size_t)3 * (M > 0 ? M : 1) * block * sizeof(real); }
    if (smem > (size_t)max_smem_optin()) { set_error("trajectory of %d points does not fit in shared memory", P); return SCCAV_EINVAL; }
    SCCAV_CUDA_CHECK(cudaFuncSetAttribute(drive_ticks_kernel<real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    drive_ticks_kernel<real><<<(int)((N + block - 1) / block), block, smem, (cudaStream_t)stream>>>(a);
    count_launch();
    SCCAV_CUDA_CHECK(cudaGetLastError());
    return SCCAV_OK;
}

int SCCAV_FN(sccav_rollout_launch_info_)(const uint8_t* slot_desc, int32_t M, int64_t N, int32_t P, int32_t* info) {
    using namespace sccav;
    if (!info || N < 1 || M < 0 || M > SCCAV_MAX_ROWS || (M > 0 && !slot_desc)) { set_error("bad argument"); return SCCAV_EINVAL; }
    int grid, block;
    size_t smem;
    bool course_smem;
    const int spec = choose_spec(slot_desc, M);
    int rc = rollout_launch_shape(M, N, P, P > 0, spec, grid, block, smem, course_smem);
    if (rc) return rc;
    // (reported for the default parameters -- model DBM, Stanley nominal, no seekers -- i.e. the compile-time instance)
    const void* kern = (const void*)rollout_instance(course_smem, spec, P > 0);
    cudaFuncAttributes fa;
    SCCAV_CUDA_CHECK(cudaFuncGetAttributes(&fa, kern));
    int occ = 0;
    SCCAV_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SCCAV_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, block, smem));
    info[0] = grid; info[1] = block; info[2] = (int32_t)smem; info[3] = fa.numRegs; info[4] = fa.maxThreadsPerBlock;
    info[5] = occ; info[6] = course_smem ? 1 : 0; info[7] = sm_count();
    return SCCAV_OK;
}

int SCCAV_FN(sccav_filter_step_host_)(const sccav_params* p, const uint8_t* slot_desc, int32_t M, int64_t N,
                                      const SCCAV_REAL* state, const SCCAV_REAL* obst, const SCCAV_REAL* u_ref,
                                      const sccav_pervehicle* pv, SCCAV_REAL* u_out, uint32_t* active_out,
                                      uint8_t* status_out, SCCAV_REAL* h_min_out, void* stream) {
    using namespace sccav;
    typedef SCCAV_REAL real;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_common(p, slot_desc, M, N, false);
    if (rc) return rc;
    if (N == 0) return SCCAV_OK;
    if (!state || !obst || !u_ref || !u_out) { set_error("NULL array argument"); return SCCAV_EINVAL; }
    const size_t n = (size_t)N;
    DevBuf d_state(st), d_obst(st), d_uref(st), d_alpha(st), d_R(st), d_cnt(st), d_u(st), d_mask(st), d_status(st), d_hmin(st);
    SCCAV_CUDA_CHECK(d_state.upload(state, 4 * n * sizeof(real)));
    SCCAV_CUDA_CHECK(d_obst.upload(obst, (size_t)M * SCCAV_NFIELD * n * sizeof(real)));
    SCCAV_CUDA_CHECK(d_uref.upload(u_ref, 2 * n * sizeof(real)));
    sccav_pervehicle dpv = {nullptr, nullptr, nullptr, nullptr, nullptr};
    DevBuf d_aug(st);
    if (pv && pv->alpha) { SCCAV_CUDA_CHECK(d_alpha.upload(pv->alpha, n * sizeof(real))); dpv.alpha = d_alpha.p; }
    if (pv && pv->R) { SCCAV_CUDA_CHECK(d_R.upload(pv->R, 4 * n * sizeof(real))); dpv.R = d_R.p; }
    if (pv && pv->count) { SCCAV_CUDA_CHECK(d_cnt.upload(pv->count, n * sizeof(int32_t))); dpv.count = (const int32_t*)d_cnt.p; }
    if (pv && pv->aug) { SCCAV_CUDA_CHECK(d_aug.upload(pv->aug, 2 * n * sizeof(real))); dpv.aug = d_aug.p; }
    SCCAV_CUDA_CHECK(d_u.alloc(2 * n * sizeof(real)));
    if (active_out) SCCAV_CUDA_CHECK(d_mask.alloc(n * sizeof(uint32_t)));
    if (status_out) SCCAV_CUDA_CHECK(d_status.alloc(n));
    if (h_min_out) SCCAV_CUDA_CHECK(d_hmin.alloc(n * sizeof(real)));
    rc = do_filter_step(p, slot_desc, M, N, (const real*)d_state.p, (const real*)d_obst.p, (const real*)d_uref.p, &dpv,
                        (real*)d_u.p, (uint32_t*)d_mask.p, (uint8_t*)d_status.p, (real*)d_hmin.p, st);
    if (rc) return rc;
    SCCAV_CUDA_CHECK(cudaMemcpyAsync(u_out, d_u.p, 2 * n * sizeof(real), cudaMemcpyDeviceToHost, st));
    if (active_out) SCCAV_CUDA_CHECK(cudaMemcpyAsync(active_out, d_mask.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    if (status_out) SCCAV_CUDA_CHECK(cudaMemcpyAsync(status_out, d_status.p, n, cudaMemcpyDeviceToHost, st));
    if (h_min_out) SCCAV_CUDA_CHECK(cudaMemcpyAsync(h_min_out, d_hmin.p, n * sizeof(real), cudaMemcpyDeviceToHost, st));
    if (pv && pv->aug) SCCAV_CUDA_CHECK(cudaMemcpyAsync(pv->aug, d_aug.p, 2 * n * sizeof(real), cudaMemcpyDeviceToHost, st));
    SCCAV_CUDA_CHECK(cudaStreamSynchronize(st));
    return SCCAV_OK;
}

int SCCAV_FN(sccav_rollout_host_)(const sccav_params* p, const uint8_t* slot_desc, int32_t M, int64_t N, int32_t T,
                                  const SCCAV_REAL* state, SCCAV_REAL* obst, const SCCAV_REAL* course_x,
                                  const SCCAV_REAL* course_y, const SCCAV_REAL* course_yaw, int32_t P,
                                  const sccav_pervehicle* pv, const sccav_rollout_out* out, void* stream) {
    using namespace sccav;
    typedef SCCAV_REAL real;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_common(p, slot_desc, M, N, true);
    if (rc) return rc;
    if (T < 0 || p->record_stride < 0) { set_error("T < 0 or record_stride < 0"); return SCCAV_EINVAL; }
    if (N == 0) return SCCAV_OK;
    if (!state || !out || !out->state) { set_error("state / out->state is NULL"); return SCCAV_EINVAL; }
    if (M > 0 && !obst) { set_error("obst is NULL"); return SCCAV_EINVAL; }
    const size_t n = (size_t)N;
    const bool stan = p->nominal == SCCAV_NOMINAL_STANLEY;
    if (stan && (P < 1 || !course_x || !course_y || !course_yaw)) { set_error("Stanley nominal control needs a course"); return SCCAV_EINVAL; }
    const size_t trec = p->record_stride > 0 ? ((size_t)T + p->record_stride - 1) / p->record_stride : 0;
    DevBuf d_state(st), d_obst(st), d_cx(st), d_cy(st), d_cyaw(st), d_alpha(st), d_R(st), d_ts(st), d_cnt(st);
    DevBuf o_state(st), o_steps(st), o_tidx(st), o_nact(st), o_ninf(st), o_hmin(st), o_bmin(st), o_bmax(st), o_bint(st),
        o_traj(st), o_tridx(st), o_trmask(st), o_evals(st);
    SCCAV_CUDA_CHECK(d_state.upload(state, 4 * n * sizeof(real)));
    const size_t obst_b = (size_t)M * SCCAV_NFIELD * n * sizeof(real);
    if (M > 0) SCCAV_CUDA_CHECK(d_obst.upload(obst, obst_b));
    if (stan) {
        SCCAV_CUDA_CHECK(d_cx.upload(course_x, (size_t)P * sizeof(real)));
        SCCAV_CUDA_CHECK(d_cy.upload(course_y, (size_t)P * sizeof(real)));
        SCCAV_CUDA_CHECK(d_cyaw.upload(course_yaw, (size_t)P * sizeof(real)));
    }
    sccav_pervehicle dpv = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (pv && pv->alpha) { SCCAV_CUDA_CHECK(d_alpha.upload(pv->alpha, n * sizeof(real))); dpv.alpha = d_alpha.p; }
    if (pv && pv->R) { SCCAV_CUDA_CHECK(d_R.upload(pv->R, 4 * n * sizeof(real))); dpv.R = d_R.p; }
    if (pv && pv->count) { SCCAV_CUDA_CHECK(d_cnt.upload(pv->count, n * sizeof(int32_t))); dpv.count = (const int32_t*)d_cnt.p; }
    if (pv && pv->target_speed) { SCCAV_CUDA_CHECK(d_ts.upload(pv->target_speed, n * sizeof(real))); dpv.target_speed = d_ts.p; }
    sccav_rollout_out dout;
    memset(&dout, 0, sizeof(dout));
    SCCAV_CUDA_CHECK(o_state.alloc(4 * n * sizeof(real)));
    dout.state = o_state.p;
#define SCCAV_OUT(field, buf, bytes)                         \
    if (out->field) {                                        \
        SCCAV_CUDA_CHECK(buf.alloc(bytes));                  \
        dout.field = (decltype(dout.field))buf.p;            \
    }
    SCCAV_OUT(steps, o_steps, n * 4)
    SCCAV_OUT(target_idx, o_tidx, n * 4)
    SCCAV_OUT(n_active, o_nact, n * 4)
    SCCAV_OUT(n_infeasible, o_ninf, n * 4)
    SCCAV_OUT(h_min, o_hmin, n * sizeof(real))
    SCCAV_OUT(beta_min, o_bmin, n * sizeof(real))
    SCCAV_OUT(beta_max, o_bmax, n * sizeof(real))
    SCCAV_OUT(beta_int, o_bint, n * sizeof(real))
    SCCAV_OUT(n_evals, o_evals, n * 4)
    if (trec) {
        SCCAV_OUT(traj, o_traj, trec * SCCAV_TRAJ_FIELDS * n * sizeof(real))
        SCCAV_OUT(traj_idx, o_tridx, trec * n * 4)
        SCCAV_OUT(traj_mask, o_trmask, trec * n * 4)
        // rows past a vehicle's last step keep the caller's initial contents
        if (out->traj) SCCAV_CUDA_CHECK(cudaMemcpyAsync(o_traj.p, out->traj, trec * SCCAV_TRAJ_FIELDS * n * sizeof(real), cudaMemcpyHostToDevice, st));
        if (out->traj_idx) SCCAV_CUDA_CHECK(cudaMemcpyAsync(o_tridx.p, out->traj_idx, trec * n * 4, cudaMemcpyHostToDevice, st));
        if (out->traj_mask) SCCAV_CUDA_CHECK(cudaMemcpyAsync(o_trmask.p, out->traj_mask, trec * n * 4, cudaMemcpyHostToDevice, st));
    }
#undef SCCAV_OUT
    rc = do_rollout(p, slot_desc, M, N, T, (const real*)d_state.p, (real*)d_obst.p, (const real*)d_cx.p, (const real*)d_cy.p,
                    (const real*)d_cyaw.p, P, &dpv, &dout, st);
    if (rc) return rc;
#define SCCAV_BACK(field, buf, bytes) \
    if (out->field) SCCAV_CUDA_CHECK(cudaMemcpyAsync(out->field, buf.p, bytes, cudaMemcpyDeviceToHost, st));
    SCCAV_BACK(state, o_state, 4 * n * sizeof(real))
    SCCAV_BACK(steps, o_steps, n * 4)
    SCCAV_BACK(target_idx, o_tidx, n * 4)
    SCCAV_BACK(n_active, o_nact, n * 4)
    SCCAV_BACK(n_infeasible, o_ninf, n * 4)
    SCCAV_BACK(h_min, o_hmin, n * sizeof(real))
    SCCAV_BACK(beta_min, o_bmin, n * sizeof(real))
    SCCAV_BACK(beta_max, o_bmax, n * sizeof(real))
    SCCAV_BACK(beta_int, o_bint, n * sizeof(real))
    SCCAV_BACK(n_evals, o_evals, n * 4)
    if (trec) {
        SCCAV_BACK(traj, o_traj, trec * SCCAV_TRAJ_FIELDS * n * sizeof(real))
        SCCAV_BACK(traj_idx, o_tridx, trec * n * 4)
        SCCAV_BACK(traj_mask, o_trmask, trec * n * 4)
    }
#undef SCCAV_BACK
    if (M > 0 && p->seeker) SCCAV_CUDA_CHECK(cudaMemcpyAsync(obst, d_obst.p, obst_b, cudaMemcpyDeviceToHost, st));
    SCCAV_CUDA_CHECK(cudaStreamSynchronize(st));
    return SCCAV_OK;
}

}  // extern "C"
