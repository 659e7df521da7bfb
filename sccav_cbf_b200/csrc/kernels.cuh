// kernels.cuh -- the sm_100a kernels of the path (launchers: capi_impl.cuh).
//
//   K0  barrier_partials_kernel      h and its partials per (vehicle, slot)                        HBM-bound
//   K1  barrier_rows_kernel          reads state + obstacle SoA, writes rows                       HBM-bound   (regime i)
//   K2  qp2_kernel / qp2_warp_kernel reads rows, writes u / active set / status                    HBM-bound   (regime i)
//   K12 filter_step_kernel           K1+K2 fused, direct loads (rows never leave the SM)                       (regime i)
//       filter_step_staged_kernel    the same with all fields of a vehicle staged by cp.async, rows overlaying them
//   K3  rollout_kernel               persistent closed loop, fp64 issue / latency bound                        (regime ii)
//   KS  stanley_kernel, KP prepare_obstacles_kernel, KB ingest_boxes_kernel, KA actuator_kernel,
//   KC  spline_course_kernel, KL lane_fit_kernel        the steps either side of the path
//
// Layout: thread == vehicle; all global accesses are SoA with the vehicle index fastest, so a warp
// reads/writes 32 consecutive elements (256 B for fp64) per request.  Rows live in shared memory
// as rows[3*M][blockDim] (conflict-free: consecutive threads -> consecutive banks).  The course
// (cx, cy interleaved, + cyaw) is staged once per CTA into shared memory and read with broadcast
// LDS.128 by the argmin loop.
#pragma once
#include "path.cuh"

namespace sccav {

template <typename T> struct PerVehicle {
    const T* alpha;
    const T* R;
    const T* target_speed;
    const int32_t* count;    // obstacles of vehicle n = its first count[n] slots (sccav_pervehicle.count)
    T* aug;                  // [2][N] SADBM: beta, beta_ref_last (read and written)
};

template <typename T>
__device__ __forceinline__ int slot_count(const PerVehicle<T>& pv, int M, int64_t n) {
    if (!pv.count) return M;
    const int c = pv.count[n];
    return c < 0 ? 0 : (c > M ? M : c);
}

template <typename T>
__device__ __forceinline__ void load_weights(const Params<T>& P, const PerVehicle<T>& pv, int64_t N, int64_t n,
                                             T& alpha, T& R00, T& R01, T& R10, T& R11) {
    alpha = pv.alpha ? pv.alpha[n] : P.alpha;
    if (pv.R) { R00 = pv.R[n]; R01 = pv.R[N + n]; R10 = pv.R[2 * N + n]; R11 = pv.R[3 * N + n]; }
    else { R00 = P.R[0]; R01 = P.R[1]; R10 = P.R[2]; R11 = P.R[3]; }
}

// inverse of the cost weight of vehicle n: the launch-wide one was formed on the host with the same IEEE
// operations (convert(), capi_impl.cuh); per-vehicle weights are inverted here, once per vehicle
template <typename T>
__device__ __forceinline__ RInv<T> load_rinv(const Params<T>& P, const PerVehicle<T>& pv, T R00, T R01, T R10, T R11) {
    if (pv.R) return RInv<T>(R00, R01, R10, R11);
    RInv<T> q;
    q.i00 = P.Ri[0]; q.i01 = P.Ri[1]; q.i10 = P.Ri[2]; q.i11 = P.Ri[3];
    return q;
}

// ------------------------------------------------------------------------------------------ KP
// Obstacle ingest: ELLIPSE slots -> ELLIPSE_PREP (include/sccav_cbf.h), other slots copied.
// HBM-bound: 64 B read + 64 B written per (vehicle, slot).  in == out is allowed (every thread
// reads its 8 fields before it writes them).
template <typename T> struct PrepareArgs {
    SlotDesc sd;
    int M;
    int64_t N;
    const T* in;
    T* out;
};

// Two adjacent vehicles per thread: every field is one 16-byte (fp64) / 8-byte (fp32) load and store -- LDG.E.128 /
// STG.E.128 on the SoA rows, twice the bytes in flight per thread.  Used when N is even and the buffers are aligned.
template <typename T>
__global__ void __launch_bounds__(256) prepare_obstacles_vec2_kernel(const __grid_constant__ PrepareArgs<T> a) {
    typedef typename Real<T>::T2 T2;
    const int64_t N = a.N, H = N >> 1;
    const int64_t total = (int64_t)a.M * H;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(i / H);
        const int64_t n2 = i - (int64_t)m * H;
        const T2* f = reinterpret_cast<const T2*>(a.in + (int64_t)m * SCCAV_NFIELD * N) + n2;
        T2* o = reinterpret_cast<T2*>(a.out + (int64_t)m * SCCAV_NFIELD * N) + n2;
        T2 v[SCCAV_NFIELD];
#pragma unroll
        for (int k = 0; k < SCCAV_NFIELD; ++k) v[k] = f[k * H];
        if ((a.sd.d[m] & SCCAV_SLOT_TYPE_MASK) == SCCAV_SLOT_ELLIPSE) {
            T m00, m01, m10, m11, wx, wy;
            ellipse_prepare<T>(v[2].x, v[3].x, v[4].x, v[5].x, v[6].x, m00, m01, m10, m11, wx, wy);
            v[2].x = m00; v[3].x = m01; v[4].x = m10; v[5].x = m11; v[6].x = wx; v[7].x = wy;
            ellipse_prepare<T>(v[2].y, v[3].y, v[4].y, v[5].y, v[6].y, m00, m01, m10, m11, wx, wy);
            v[2].y = m00; v[3].y = m01; v[4].y = m10; v[5].y = m11; v[6].y = wx; v[7].y = wy;
        }
#pragma unroll
        for (int k = 0; k < SCCAV_NFIELD; ++k) o[k * H] = v[k];
    }
}

template <typename T>
__global__ void __launch_bounds__(256) prepare_obstacles_kernel(const __grid_constant__ PrepareArgs<T> a) {
    const int64_t N = a.N;
    const int64_t total = (int64_t)a.M * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(i / N);
        const int64_t n = i - (int64_t)m * N;
        const T* f = a.in + (int64_t)m * SCCAV_NFIELD * N + n;
        T* o = a.out + (int64_t)m * SCCAV_NFIELD * N + n;
        T v[SCCAV_NFIELD];
#pragma unroll
        for (int k = 0; k < SCCAV_NFIELD; ++k) v[k] = f[k * N];
        if ((a.sd.d[m] & SCCAV_SLOT_TYPE_MASK) == SCCAV_SLOT_ELLIPSE) {
            T m00, m01, m10, m11, wx, wy;
            ellipse_prepare<T>(v[2], v[3], v[4], v[5], v[6], m00, m01, m10, m11, wx, wy);
            v[2] = m00; v[3] = m01; v[4] = m10; v[5] = m11; v[6] = wx; v[7] = wy;
        }
#pragma unroll
        for (int k = 0; k < SCCAV_NFIELD; ++k) o[k * N] = v[k];
    }
}

// Rollout-private form of static prepared ellipses: the symmetric matrix S = M^T M instead of M, in 16-byte pairs --
//   sym[m][0][n] = (cx, cy),  sym[m][1][n] = (s00, s01),  sym[m][2][n] = (s11, 0)
// so that  h = d^T S d - 1,  grad h = 2 S d:  three 16-byte loads and eight flops per row where the public ELLIPSE_PREP
// slot takes six 8-byte loads and twelve.  Same functions, a few ulp from the M form (SCCAV_FLAG_PREPARED_ROWS).
template <typename T> struct SymArgs {
    int M;
    int64_t N;
    const T* prep;       // [M][8][N] ELLIPSE_PREP slots (cx, cy, m00, m01, m10, m11, ., .)
    typename Real<T>::T2* sym;
};
template <typename T>
__global__ void __launch_bounds__(256) pack_sym_kernel(const __grid_constant__ SymArgs<T> a) {
    typedef Real<T> R;
    const int64_t N = a.N, total = (int64_t)a.M * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / N, n = i - m * N;
        const T* f = a.prep + m * SCCAV_NFIELD * N + n;
        const T m00 = f[2 * N], m01 = f[3 * N], m10 = f[4 * N], m11 = f[5 * N];
        typename R::T2* o = a.sym + m * 3 * N + n;
        o[0] = R::make2(f[0], f[N]);
        o[N] = R::make2(fma(m00, m00, m10 * m10), fma(m00, m01, m10 * m11));
        o[2 * N] = R::make2(fma(m01, m01, m11 * m11), T(0));
    }
}

// ------------------------------------------------------------------------------------------ KB
// Batched ObstacleList2D.update_by_bounding_box (include/sccav_cbf.h): one thread = one vehicle's list.
// Integer bookkeeping (which id sits in which slot) is exact; the only arithmetic is a + buffer and
// hypot(extent).  HBM-bound: reads K (4 + 48) + M (4 + 64) + 4, writes up to M (4 + 64) + 8 bytes per vehicle.
template <typename T> struct IngestArgs {
    int type, mode, M, K;
    int64_t N;
    T buffer;
    const int32_t* box_id;
    const T* box;
    int32_t* slot_id;
    T* obst;
    int32_t* count;
    int32_t* dropped;
};

// fields of one obstacle from its bounding box; create = from_bounding_box, else update_by_bounding_box
template <typename T>
__device__ __forceinline__ void box_to_slot(int type, bool create, T buffer, const T* __restrict__ bx, int64_t N, T (&f)[SCCAV_NFIELD]) {
    typedef Real<T> R;
    const T ex = bx[0], ey = bx[N], lx = bx[2 * N], ly = bx[3 * N], yaw = bx[4 * N], sp = bx[5 * N];
    if (type == SCCAV_SLOT_ELLIPSE) {
        f[0] = lx; f[1] = ly; f[4] = yaw;                                   // obstacles.py:298-302,326-330
        f[2] = create ? ex + buffer : ex;                                   // Ellipse2D.__init__ :159-160 / update(a=, b=)
        f[3] = create ? ey + buffer : ey;
        if (create) { f[5] = T(0); f[6] = T(0); f[7] = T(0); }              // vel = Vector2()   :158
    } else {
        const T a = R::hypot_(ex, ey);                                      // obstacles.py:528,541
        f[0] = lx; f[1] = ly; f[2] = T(0); f[3] = sp;                       // s_obs = [x, y, 0.0, velocity]   :529,542
        f[4] = create ? a + buffer : a;                                     // CollisionCone2D.__init__ :357 / self.a = hypot
        if (create) { f[5] = T(0); f[6] = T(0); f[7] = T(0); }              // beta = 0   :352
    }
}

template <typename T>
__global__ void __launch_bounds__(128) ingest_boxes_kernel(const __grid_constant__ IngestArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int32_t* s_bid = reinterpret_cast<int32_t*>(smem_raw) + threadIdx.x;     // [K][blockDim]
    const int B = blockDim.x, M = a.M, K = a.K;
    const int64_t N = a.N;
    for (int64_t n = (int64_t)blockIdx.x * B + threadIdx.x; n < N; n += (int64_t)gridDim.x * B) {
        for (int k = 0; k < K; ++k) s_bid[k * B] = a.box_id[(int64_t)k * N + n];
        int cnt = a.mode == SCCAV_INGEST_REBUILD ? 0 : a.count[n];
        cnt = cnt < 0 ? 0 : (cnt > M ? M : cnt);
        uint32_t matched = 0u;                       // boxes whose id was found among the held entries
        int w = 0;                                   // write cursor of the stable compaction
        T f[SCCAV_NFIELD];
        for (int j = 0; j < cnt; ++j) {
            const int32_t id = a.slot_id[(int64_t)j * N + n];
            int kf = -1;
            if (id >= 0)
                for (int k = 0; k < K; ++k)
                    if (s_bid[k * B] == id) { if (kf < 0) kf = k; matched |= 1u << k; }
            if (kf < 0) continue;                    // id left the scene: entry removed   obstacles.py:851-853
            const T* src = a.obst + (int64_t)j * SCCAV_NFIELD * N + n;
#pragma unroll
            for (int q = 0; q < SCCAV_NFIELD; ++q) f[q] = src[q * N];
            box_to_slot<T>(a.type, false, a.buffer, a.box + (int64_t)kf * SCCAV_BOX_FIELDS * N + n, N, f);   // :840-841
            T* dst = a.obst + (int64_t)w * SCCAV_NFIELD * N + n;
#pragma unroll
            for (int q = 0; q < SCCAV_NFIELD; ++q) dst[q * N] = f[q];
            a.slot_id[(int64_t)w * N + n] = id;
            ++w;
        }
        int drop = 0;
        for (int k = 0; k < K; ++k) {
            const int32_t id = s_bid[k * B];
            if (id < 0 || ((matched >> k) & 1u)) continue;
            bool dup = false;                        // a repeated new id: the first box wins
            for (int k2 = 0; k2 < k; ++k2) dup |= s_bid[k2 * B] == id;
            if (dup) continue;
            if (w >= M) { ++drop; continue; }
            box_to_slot<T>(a.type, true, a.buffer, a.box + (int64_t)k * SCCAV_BOX_FIELDS * N + n, N, f);     // :843-846
            T* dst = a.obst + (int64_t)w * SCCAV_NFIELD * N + n;
#pragma unroll
            for (int q = 0; q < SCCAV_NFIELD; ++q) dst[q * N] = f[q];
            a.slot_id[(int64_t)w * N + n] = id;
            ++w;
        }
        const int old = a.mode == SCCAV_INGEST_REBUILD ? M : cnt;
        for (int j = w; j < old; ++j) a.slot_id[(int64_t)j * N + n] = -1;
        a.count[n] = w;
        if (a.dropped) a.dropped[n] = drop;
    }
}

// ------------------------------------------------------------------------------------------ KA
// Actuator shaping after the filter (multi_obstacle_CBF_local_with_lanes.py:955-980), element-wise.
template <typename T> struct ActuatorArgs {
    int64_t N;
    const T* u;
    T max_steer, rate;
    int flags;
    T* thr_prev;
    T* brk_prev;
    T* thr;
    T* brk;
    T* steer;
};

// (a, delta) -> (throttle, brake, steer); tp / bp = the previous tick's throttle / brake (read, then replaced)
template <typename T>
__device__ __forceinline__ void actuator_step(T ua, T d, T max_steer, T rate, int flags, T& tp, T& bp, T& throttle, T& brake, T& steer) {
    typedef Real<T> R;
    brake = bp;                                                   // `brake` is the driver's variable of the last tick
    if (ua > T(0)) {
        throttle = R::tanh_(ua);
        throttle = fmax(T(0), fmin(T(1), throttle));              // :959-960
        if (throttle - tp > rate) throttle = tp + rate;           // :961-962
        if (flags & SCCAV_ACT_RESET_BRAKE) brake = T(0);
    } else {
        throttle = T(0);                                          // :964
        brake = -R::tanh_(ua);
        brake = fmax(T(0), fmin(T(1), brake));                    // :965-966
        if (brake - bp > rate) brake = bp + rate;                 // :967-968
    }
    if (d > T(0)) d = fmax(T(0), fmin(d, max_steer));             // :973-976
    else d = fmax(-max_steer, fmin(d, T(0)));
    tp = throttle;                                                // :970-971
    bp = brake;
    steer = d;
}

template <typename T>
__global__ void __launch_bounds__(256) actuator_kernel(const __grid_constant__ ActuatorArgs<T> a) {
    const int64_t N = a.N;
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
        T tp = a.thr_prev[n], bp = a.brk_prev[n];
        T throttle, brake, d;
        actuator_step<T>(a.u[n], a.u[N + n], a.max_steer, a.rate, a.flags, tp, bp, throttle, brake, d);
        a.thr_prev[n] = tp;
        a.brk_prev[n] = bp;
        if (a.thr) a.thr[n] = throttle;
        if (a.brk) a.brk[n] = brake;
        if (a.steer) a.steer[n] = d;
    }
}

// ------------------------------------------------------------------------------------------ KC
// Course generation (include/sccav_cbf.h): CTA (chunk, c) samples 256 points of course c.  Every CTA
// first rebuilds the spline coefficients of its course in shared memory (K <= 64 knots: a few hundred
// flops by one thread) -- cheaper than a second kernel and a round trip through HBM.
template <typename T> struct CourseArgs {
    int C, K, P_max;
    T ds;
    const T* wx;
    const T* wy;
    T* cx;
    T* cy;
    T* cyaw;
    T* ck;
    int32_t* np_out;
};

// natural cubic spline through (s_k, a_k): coefficients b, c, d per segment (cubic_spline_planner.py:17-42,95-115)
template <typename T>
__device__ void natural_spline(int K, const T* s, const T* a, T* b, T* c, T* d, T* cp, T* dp) {
    // interior rows i = 1 .. K-2:  h[i-1] c[i-1] + 2 (h[i-1] + h[i]) c[i] + h[i] c[i+1] = rhs_i ;  c[0] = c[K-1] = 0
    cp[0] = T(0); dp[0] = T(0);
    for (int i = 1; i + 1 < K; ++i) {
        const T h0 = s[i] - s[i - 1], h1 = s[i + 1] - s[i];
        const T rhs = T(3) * (a[i + 1] - a[i]) / h1 - T(3) * (a[i] - a[i - 1]) / h0;
        const T m = T(2) * (h0 + h1) - h0 * cp[i - 1];
        cp[i] = h1 / m;
        dp[i] = (rhs - h0 * dp[i - 1]) / m;
    }
    c[K - 1] = T(0);
    for (int i = K - 2; i >= 1; --i) c[i] = dp[i] - cp[i] * c[i + 1];
    c[0] = T(0);
    for (int i = 0; i + 1 < K; ++i) {
        const T h = s[i + 1] - s[i];
        d[i] = (c[i + 1] - c[i]) / (T(3) * h);
        b[i] = (a[i + 1] - a[i]) / h - h * (c[i + 1] + T(2) * c[i]) / T(3);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) spline_course_kernel(const __grid_constant__ CourseArgs<T> a) {
    typedef Real<T> R;
    __shared__ T s[SCCAV_MAX_KNOTS], ax[SCCAV_MAX_KNOTS], ay[SCCAV_MAX_KNOTS];
    __shared__ T bx[SCCAV_MAX_KNOTS], cx_[SCCAV_MAX_KNOTS], dx_[SCCAV_MAX_KNOTS];
    __shared__ T by[SCCAV_MAX_KNOTS], cy_[SCCAV_MAX_KNOTS], dy_[SCCAV_MAX_KNOTS];
    __shared__ T cp[SCCAV_MAX_KNOTS], dp[SCCAV_MAX_KNOTS];
    __shared__ int s_np;
    const int c = blockIdx.y, K = a.K;
    if (threadIdx.x < K) {
        ax[threadIdx.x] = a.wx[(int64_t)c * K + threadIdx.x];
        ay[threadIdx.x] = a.wy[(int64_t)c * K + threadIdx.x];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        s[0] = T(0);                                                             // Spline2D.__calc_s  :129-136
        for (int k = 0; k + 1 < K; ++k) s[k + 1] = s[k] + R::hypot_(ax[k + 1] - ax[k], ay[k + 1] - ay[k]);
        s_np = (int)R::ceil_(s[K - 1] / a.ds);                                   // len(np.arange(0, s[-1], ds))  :180
    }
    __syncthreads();
    if (threadIdx.x == 0) natural_spline<T>(K, s, ax, bx, cx_, dx_, cp, dp);
    if (threadIdx.x == 32) {
        __shared__ T cp2[SCCAV_MAX_KNOTS], dp2[SCCAV_MAX_KNOTS];
        natural_spline<T>(K, s, ay, by, cy_, dy_, cp2, dp2);
    }
    __syncthreads();
    const int np = s_np;
    if (blockIdx.x == 0 && threadIdx.x == 0) a.np_out[c] = np;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= np || j >= a.P_max) return;
    const T t = (T)j * a.ds;                                                     // np.arange: start + j * step
    int i = 0;
    while (i + 2 < K && t >= s[i + 1]) ++i;                                       // bisect.bisect(x, t) - 1   :88-92
    const T u = t - s[i];
    const T u2 = u * u, u3 = u2 * u;
    const int64_t o = (int64_t)c * a.P_max + j;
    a.cx[o] = ax[i] + bx[i] * u + cx_[i] * u2 + dx_[i] * u3;                      // Spline.calc  :44-60
    a.cy[o] = ay[i] + by[i] * u + cy_[i] * u2 + dy_[i] * u3;
    const T gx = bx[i] + T(2) * cx_[i] * u + T(3) * dx_[i] * u2;                  // Spline.calcd :62-76
    const T gy = by[i] + T(2) * cy_[i] * u + T(3) * dy_[i] * u2;
    a.cyaw[o] = R::atan2_(gy, gx);                                               // calc_yaw :167-173
    if (a.ck) {
        const T hx = T(2) * cx_[i] + T(6) * dx_[i] * u;                           // Spline.calcdd :78-90
        const T hy = T(2) * cy_[i] + T(6) * dy_[i] * u;
        a.ck[o] = (hy * gx - hx * gy) / R::pow_(gx * gx + gy * gy, T(1.5));       // calc_curvature :156-165
    }
}

// ------------------------------------------------------------------------------------------ KL
// Weighted polynomial lane fit (include/sccav_cbf.h): one thread = one lane; points are read [K][C] (coalesced
// over lanes, L2-resident across the degree + 2 passes).  Range of x, then orthogonal polynomials in
// t = (x - mid) / half over the weighted points, then the expansion into powers of x.
template <typename T> struct LaneFitArgs {
    int C, K, degree;
    const T* x;
    const T* y;
    const T* sigma;
    const int32_t* count;
    T* coeffs;
    int32_t* status;
};

template <typename T>
__global__ void __launch_bounds__(128) lane_fit_kernel(const __grid_constant__ LaneFitArgs<T> a) {
    // arithmetic in double for both storage types: powers of x about x = 0 cancel badly in fp32 from degree 4 on,
    // and the kernel is a few hundred flops per lane
    typedef double W;
    typedef Real<W> R;
    const int C = a.C;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
        const int n = a.degree;
        int K = a.count ? a.count[c] : a.K;
        K = K < 0 ? 0 : (K > a.K ? a.K : K);
        W out[6] = {W(0), W(0), W(0), W(0), W(0), W(0)};
        bool ok = K > n;
        W lo = R::inf(), hi = -R::inf();
        for (int k = 0; k < K; ++k) {
            const W xv = a.x[(int64_t)k * C + c];
            lo = xv < lo ? xv : lo;
            hi = xv > hi ? xv : hi;
        }
        const W mid = W(0.5) * (lo + hi);
        W half = W(0.5) * (hi - lo);
        if (!(half > W(0))) { half = W(1); ok = false; }           // all abscissae equal (or NaN)
        const W inv = W(1) / half;
        // Forsythe's method: polynomials p_0..p_n orthogonal over the weighted points, built by the three-term
        // recurrence p_{j+1} = (t - al_j) p_j - be_j p_{j-1}; the fit is sum d_j p_j with d_j = <y, p_j> / <p_j, p_j>.
        // No matrix is formed (the normal equations lose half the digits when the sigmas span 10 .. 0.01);
        // pass j re-runs the recurrence per point, O(K n^2) flops in all.
        W al[6], be[6], d[6], nrm[6];
        for (int j = 0; j <= n && ok; ++j) {
            W sN = W(0), sA = W(0), sD = W(0);
            for (int k = 0; k < K; ++k) {
                const W t = (a.x[(int64_t)k * C + c] - mid) * inv;
                const W sg = a.sigma ? a.sigma[(int64_t)k * C + c] : W(10);
                const W w = W(1) / (sg * sg);
                W pm = W(0), p = W(1);
                for (int q = 0; q < j; ++q) {
                    const W pn = (t - al[q]) * p - be[q] * pm;
                    pm = p; p = pn;
                }
                const W wp = w * p;
                sN += wp * p;
                sA += wp * p * t;
                sD += wp * a.y[(int64_t)k * C + c];
            }
            if (!(sN > W(0)) || !(sN < R::inf())) { ok = false; break; }      // fewer distinct abscissae than coefficients
            nrm[j] = sN;
            al[j] = sA / sN;
            be[j] = j > 0 ? sN / nrm[j - 1] : W(0);
            d[j] = sD / sN;
        }
        if (ok && n > 0 && !(nrm[n] > nrm[0] * R::feas_eps() * R::feas_eps())) ok = false;
        if (ok) {
            // monomial coefficients (in t) of sum d_j p_j through the same recurrence on coefficient vectors
            W cm[6] = {W(0), W(0), W(0), W(0), W(0), W(0)}, cp[6] = {W(1), W(0), W(0), W(0), W(0), W(0)};
            W at[6] = {d[0], W(0), W(0), W(0), W(0), W(0)};
            for (int j = 0; j < n; ++j) {
                W cn[6];
                for (int q = 0; q < 6; ++q) cn[q] = (q > 0 ? cp[q - 1] : W(0)) - al[j] * cp[q] - be[j] * cm[q];
                for (int q = 0; q < 6; ++q) { cm[q] = cp[q]; cp[q] = cn[q]; at[q] += d[j + 1] * cn[q]; }
            }
            // p(t) = sum at_j t^j with t = x inv - mid inv: Horner in polynomials of x
            const W sh = -mid * inv;
            for (int j = n; j >= 0; --j) {
                for (int q = n; q >= 1; --q) out[q] = out[q] * sh + out[q - 1] * inv;
                out[0] = out[0] * sh + at[j];
            }
        } else {
            const W nan = R::inf() - R::inf();
            for (int q = 0; q < 6; ++q) out[q] = nan;
        }
        for (int q = 0; q < 6; ++q) a.coeffs[(int64_t)q * C + c] = (T)out[q];
        if (a.status) a.status[c] = ok ? 0 : 1;
    }
}

// ------------------------------------------------------------------------------------------ K1
template <typename T> struct RowsArgs {
    Params<T> P;
    SlotDesc sd;
    int M;
    int64_t N;
    const T* state;
    const T* obst;
    PerVehicle<T> pv;
    T* A;
    T* b;
    T* h;
};

template <typename T>
__global__ void __launch_bounds__(256) barrier_rows_kernel(const __grid_constant__ RowsArgs<T> a) {
    typedef Real<T> R;
    const int64_t N = a.N;
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
        T x = a.state[n], y = a.state[N + n], th = a.state[2 * N + n], v = a.state[3 * N + n];
        T alpha = a.pv.alpha ? a.pv.alpha[n] : a.P.alpha;
        T sth, cth;
        R::sincos_(th, &sth, &cth);
        const T vlr = v / a.P.lr;
        const int Mv = slot_count<T>(a.pv, a.M, n);
        for (int m = 0; m < a.M; ++m) {
            T A0 = T(0), A1 = T(0), b = -R::inf();             // empty slot: a vacuous row
            Partials<T> p;
            p.h = R::inf();
            if (m < Mv) {
                const int desc = a.sd.d[m];
                const int64_t nn = (desc & SCCAV_SLOT_SHARED) ? 0 : n;
                const T* f = a.obst + (int64_t)m * SCCAV_NFIELD * N + nn;
                p = slot_partials<T>(desc, f, N, x, y, th, v, sth, cth);
                model_row<T>(a.P, p, sth, cth, v, alpha, vlr, A0, A1, b);
            }
            a.A[(int64_t)m * N + n] = A0;
            a.A[((int64_t)a.M + m) * N + n] = A1;
            a.b[(int64_t)m * N + n] = b;
            if (a.h) a.h[(int64_t)m * N + n] = p.h;
        }
    }
}

// ------------------------------------------------------------------------------------------ K0
// Barrier values and partials only (what the obstacle objects of the class API return):
// out[m][6][N] = h, h_x, h_y, h_theta, h_v, h_t   -- ObstacleList2D.f/dx/dy/dtheta/dv/dt
template <typename T> struct PartialsArgs {
    SlotDesc sd;
    int M;
    int64_t N;
    const T* state;
    const T* obst;
    const int32_t* count;
    T* out;
};

template <typename T>
__global__ void __launch_bounds__(256) barrier_partials_kernel(const __grid_constant__ PartialsArgs<T> a) {
    typedef Real<T> R;
    const int64_t N = a.N;
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
        T x = a.state[n], y = a.state[N + n], th = a.state[2 * N + n], v = a.state[3 * N + n];
        T sth, cth;
        R::sincos_(th, &sth, &cth);
        int Mv = a.M;
        if (a.count) { Mv = a.count[n]; Mv = Mv < 0 ? 0 : (Mv > a.M ? a.M : Mv); }
        for (int m = 0; m < a.M; ++m) {
            Partials<T> p;
            p.h = R::inf(); p.hx = p.hy = p.hth = p.hv = p.ht = T(0);      // empty slot
            if (m < Mv) {
                const int desc = a.sd.d[m];
                const int64_t nn = (desc & SCCAV_SLOT_SHARED) ? 0 : n;
                const T* f = a.obst + (int64_t)m * SCCAV_NFIELD * N + nn;
                p = slot_partials<T>(desc, f, N, x, y, th, v, sth, cth);
            }
            T* o = a.out + (int64_t)m * 6 * N + n;
            o[0] = p.h; o[N] = p.hx; o[2 * N] = p.hy; o[3 * N] = p.hth; o[4 * N] = p.hv; o[5 * N] = p.ht;
        }
    }
}

// ------------------------------------------------------------------------------------------ K2
template <typename T> struct QpArgs {
    Params<T> P;
    int M;
    int64_t N;
    const T* A;
    const T* b;
    const T* r;
    PerVehicle<T> pv;
    T* u;
    uint32_t* mask;
    uint8_t* status;
};

template <typename T>
__global__ void __launch_bounds__(256) qp2_kernel(const __grid_constant__ QpArgs<T> a) {
    typedef Real<T> R;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    T* rows = reinterpret_cast<T*>(smem_raw) + threadIdx.x;
    const T* warp_rows = rows - lane;
    const int stride = blockDim.x;
    const int64_t N = a.N;
    const bool enumerate = (a.P.flags & SCCAV_FLAG_QP_ENUMERATE) != 0;
    // warp-uniform trip count (the warp's first vehicle decides): the active solve below is warp-cooperative
    for (int64_t nw = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x - lane); nw < N; nw += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = nw + lane;
        const bool valid = n < N;
        int Mv = 0;
        T r0 = T(0), r1 = T(0), worst = -R::inf();
        // (lanes past N take part in the cooperative enumeration of their warp's problems with the launch-wide weight)
        T alpha, R00 = a.P.R[0], R01 = a.P.R[1], R10 = a.P.R[2], R11 = a.P.R[3];
        RowNz nz{0u, 0u};
        bool feas = true;
        QpScan<T> scan;
        scan.reset();
        RInv<T> Ri;
        Ri.i00 = a.P.Ri[0]; Ri.i01 = a.P.Ri[1]; Ri.i10 = a.P.Ri[2]; Ri.i11 = a.P.Ri[3];
        if (valid) {
            Mv = slot_count<T>(a.pv, a.M, n);
            load_weights<T>(a.P, a.pv, N, n, alpha, R00, R01, R10, R11);
            Ri = load_rinv<T>(a.P, a.pv, R00, R01, R10, R11);
            r0 = a.r[n]; r1 = a.r[N + n];
            for (int m = 0; m < Mv; ++m) {
                const T a0 = a.A[(int64_t)m * N + n], a1 = a.A[((int64_t)a.M + m) * N + n], bk = a.b[(int64_t)m * N + n];
                rows[(3 * m + 0) * stride] = a0;
                rows[(3 * m + 1) * stride] = a1;
                rows[(3 * m + 2) * stride] = bk;
                if (a0 != T(0)) nz.nz0 |= 1u << m;
                if (a1 != T(0)) nz.nz1 |= 1u << m;
                T t0 = a0 * r0, t1 = a1 * r1;
                T rk = (t0 + t1) - bk;
                if (-rk > worst) worst = -rk;
                T tol = R::feas_eps() * ((R::abs_(t0) + R::abs_(t1)) + R::abs_(bk));
                if (!(rk >= -tol)) feas = false;
                scan.row(m, a0, a1, rk, Ri);
            }
        }
        T u0 = r0, u1 = r1;
        uint32_t mask = 0u;
        const int st = qp2_solve_active_warp<T>(!feas, warp_rows, stride, lane, Mv, nz, r0, r1, R00, R01, R10, R11, Ri,
                                                a.pv.R == nullptr, worst, scan, u0, u1, mask, enumerate);
        __syncwarp();                          // the next problem's rows overwrite the columns read above
        if (valid) {
            a.u[n] = u0;
            a.u[N + n] = u1;
            if (a.mask) a.mask[n] = mask;
            if (a.status) a.status[n] = (uint8_t)st;
        }
    }
}

// Warp-per-problem variant: lane k owns row k (M <= 32).  Candidates are generated by the owning
// lanes and checked by all lanes at once (one ballot per candidate); the enumeration order and
// acceptance rule are those of qp2_solve, so results are identical.  Meant for small batches of
// large problems (latency), the thread-per-problem kernel for throughput.
template <typename T>
__global__ void __launch_bounds__(256) qp2_warp_kernel(const __grid_constant__ QpArgs<T> a) {
    typedef Real<T> R;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t N = a.N;
    for (int64_t n = warp; n < N; n += nwarps) {
        const int M = slot_count<T>(a.pv, a.M, n);
        const bool own = lane < M;
        // lanes >= M hold a vacuous row 0*u >= -inf
        T a0 = own ? a.A[(int64_t)lane * N + n] : T(0);
        T a1 = own ? a.A[((int64_t)a.M + lane) * N + n] : T(0);
        T bk = own ? a.b[(int64_t)lane * N + n] : -R::inf();
        T alpha, R00, R01, R10, R11;
        load_weights<T>(a.P, a.pv, N, n, alpha, R00, R01, R10, R11);
        const T r0 = a.r[n], r1 = a.r[N + n];
        T u0 = r0, u1 = r1, worst;
        uint32_t mask = 0u;
        int status = SCCAV_STATUS_INACTIVE;
        if (!qp_check_coop<T>(lane, own, a0, a1, bk, r0, r1, 0u, worst)) {
            RowNz nz;
            nz.nz0 = __ballot_sync(0xffffffffu, own && a0 != T(0));
            nz.nz1 = __ballot_sync(0xffffffffu, own && a1 != T(0));
            const RInv<T> Ri = load_rinv<T>(a.P, a.pv, R00, R01, R10, R11);
            status = qp2_coop_active<T>(lane, M, a0, a1, bk, nz, r0, r1, R00, R01, R10, R11, Ri, worst, u0, u1, mask);
        }
        if (lane == 0) {
            a.u[n] = u0;
            a.u[N + n] = u1;
            if (a.mask) a.mask[n] = mask;
            if (a.status) a.status[n] = (uint8_t)status;
        }
    }
}

// ------------------------------------------------------------------------------------------ K12
template <typename T> struct FilterArgs {
    Params<T> P;
    SlotDesc sd;
    int M;
    int64_t N;
    const T* state;
    const T* obst;
    const T* u_ref;
    PerVehicle<T> pv;
    T* u;
    uint32_t* mask;
    uint8_t* status;
    T* h_min;
};

#ifndef SCCAV_K12_MINB
#define SCCAV_K12_MINB 2
#endif
template <typename T, int SPEC, bool COOP>
__global__ void __launch_bounds__(256, SCCAV_K12_MINB) filter_step_kernel(const __grid_constant__ FilterArgs<T> a) {
    typedef Real<T> R;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    T* rows = reinterpret_cast<T*>(smem_raw) + threadIdx.x;
    const T* warp_rows = rows - lane;
    const int stride = blockDim.x;
    const int64_t N = a.N;
    const bool enumerate = (a.P.flags & SCCAV_FLAG_QP_ENUMERATE) != 0;
    // warp-uniform trip count (the warp's first vehicle decides): the active solve is warp-cooperative
    for (int64_t nw = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x - lane); nw < N; nw += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = nw + lane;
        const bool valid = n < N;
        T ur0 = T(0), ur1 = T(0), hmin = R::inf();
        // (lanes past N take part in the cooperative enumeration of their warp's problems with the launch-wide weight)
        T alpha = T(0), R00 = a.P.R[0], R01 = a.P.R[1], R10 = a.P.R[2], R11 = a.P.R[3];
        int Mv = 0;
        T augv[2] = {T(0), T(0)};
        RowPhase<T> ph;
        ph.r0 = T(0); ph.r1 = T(0); ph.worst = -R::inf(); ph.nz.nz0 = 0u; ph.nz.nz1 = 0u; ph.feas = true;
        ph.scan.reset();
        RInv<T> Ri;
        Ri.i00 = a.P.Ri[0]; Ri.i01 = a.P.Ri[1]; Ri.i10 = a.P.Ri[2]; Ri.i11 = a.P.Ri[3];
        if (valid) {
            T x = a.state[n], y = a.state[N + n], th = a.state[2 * N + n], v = a.state[3 * N + n];
            ur0 = a.u_ref[n]; ur1 = a.u_ref[N + n];
            load_weights<T>(a.P, a.pv, N, n, alpha, R00, R01, R10, R11);
            Ri = load_rinv<T>(a.P, a.pv, R00, R01, R10, R11);
            Mv = slot_count<T>(a.pv, a.M, n);
            if (Mv > 0) {
                T sth, cth;
                R::sincos_(th, &sth, &cth);
                if (a.P.model == SCCAV_MODEL_SADBM) { augv[0] = a.pv.aug[n]; augv[1] = a.pv.aug[N + n]; }
                ph = filter_rows<T, SPEC, COOP, -1, true>(a.P, a.sd, Mv, N, n, a.obst, x, y, th, v, sth, cth, alpha, ur0, ur1, rows, stride,
                                                hmin, nullptr, 0xffffffffu, &Ri, augv);
            }
        }
        T q0 = ph.r0, q1 = ph.r1;
        uint32_t mask = 0u;
        int st = SCCAV_STATUS_INACTIVE;
        if (COOP) {
            st = qp2_solve_active_warp<T>(!ph.feas, warp_rows, stride, lane, Mv, ph.nz, ph.r0, ph.r1, R00, R01, R10, R11, Ri,
                                          a.pv.R == nullptr, ph.worst, ph.scan, q0, q1, mask, enumerate);
            __syncwarp();                      // the next vehicle's rows overwrite the columns read above
        } else if (!ph.feas) {
            // one thread, one problem: plain enumeration
            const RowView<T> rv{rows, stride};
            st = qp2_solve_active<T>(rv, Mv, ph.nz, ph.r0, ph.r1, R00, R01, R10, R11, Ri, ph.worst, q0, q1, mask);
        }
        if (valid) {
            T u0 = ur0, u1 = ur1;                                           // empty obstacle list: u = u_ref
            if (Mv > 0 && a.P.model == SCCAV_MODEL_SADBM) {
                // cbf.py:419-429: integrate the solved rate, hand the new beta and this call's beta_ref to the next call
                const T beta_new = augv[0] + q1 * a.P.sadbm_dt;
                u0 = q0;
                u1 = R::atan2_((a.P.lf + a.P.lr) * R::tan_(beta_new), a.P.lr);
                a.pv.aug[n] = beta_new;
                a.pv.aug[N + n] = R::atan2_(a.P.lr * R::tan_(ur1), a.P.lf + a.P.lr);
            } else
            if (Mv > 0) { u0 = q0; u1 = filter_convert<T, -1, true>(a.P, q0, q1, ph.r0); }
            a.u[n] = u0;
            a.u[N + n] = u1;
            if (a.mask) a.mask[n] = mask;
            if (a.status) a.status[n] = (uint8_t)st;
            if (a.h_min) a.h_min[n] = hmin;
        }
    }
}

// K12, staged (all per-vehicle ELLIPSE or static ELLIPSE_PREP slots, M * NF * 8 B <= 448 B per vehicle).
// The direct-load kernel above stalls on HBM latency: its 16 resident warps (fp64 pairs: 128 registers)
// have no registers left to hold more than two slots of loads in flight.  Here a thread requests ALL
// fields of its vehicle at once with cp.async (LDGSTS, one coalesced 256 B request per warp and field,
// no register until the value is used) into its own column of shared memory, converts u_ref while they
// fly (sincos, tan, atan2: ~400 instructions), waits once, and evaluates the slots out of shared memory.
// Row m then overwrites the first three staged fields of slot m, which nobody reads again, so the
// staging area doubles as the row store: M * NF * 256 * 8 B = 96 KB (NF = 6) per CTA, two CTAs per SM.
// A thread only reads what it requested itself: no barrier beyond cp.async.wait_group; the warp-level
// sync at the end of the iteration is the one the cooperative QP needs anyway.  Arithmetic = filter_vehicle.
template <typename T> __device__ __forceinline__ void cp_async_elem(T* smem_dst, const T* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(s), "l"(gmem_src), "n"((int)sizeof(T)) : "memory");
}

template <typename T, int SPEC, int NF, bool COOP, int MODEL>
__global__ void __launch_bounds__(256, SCCAV_K12_MINB) filter_step_staged_kernel(const __grid_constant__ FilterArgs<T> a) {
    typedef Real<T> R;
    static_assert(NF >= 3 && NF <= SCCAV_NFIELD, "row m lives in the first three staged fields of slot m");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int B = blockDim.x;
    const int M = a.M;
    const int lane = threadIdx.x & 31;
    T* stage = reinterpret_cast<T*>(smem_raw) + threadIdx.x;       // stage[(m * NF + f) * B]; rows: f = 0, 1, 2
    const T* warp_rows = stage - lane;
    const int64_t N = a.N;
    const Params<T>& P = a.P;
    const bool enumerate = (P.flags & SCCAV_FLAG_QP_ENUMERATE) != 0;
    // warp-uniform trip count (the warp's first vehicle decides): the active solve is warp-cooperative
    for (int64_t nw = (int64_t)blockIdx.x * B + (threadIdx.x - lane); nw < N; nw += (int64_t)gridDim.x * B) {
        const int64_t n = nw + lane;
        const bool valid = n < N;
        T x = T(0), y = T(0), th = T(0), v = T(0), ur0 = T(0), ur1 = T(0);
        // (lanes past N take part in the cooperative enumeration of their warp's problems with the launch-wide weight)
        T alpha = T(0), R00 = P.R[0], R01 = P.R[1], R10 = P.R[2], R11 = P.R[3];
        RInv<T> Ri;
        Ri.i00 = P.Ri[0]; Ri.i01 = P.Ri[1]; Ri.i10 = P.Ri[2]; Ri.i11 = P.Ri[3];
        int Mv = 0;
        if (valid) {
            // the vehicle's own six values first: they are needed first (sincos, tan), and the asynchronous copies
            // below would otherwise queue in front of them (the volatile asm pins the program order)
            x = a.state[n]; y = a.state[N + n]; th = a.state[2 * N + n]; v = a.state[3 * N + n];
            ur0 = a.u_ref[n]; ur1 = a.u_ref[N + n];
            Mv = slot_count<T>(a.pv, M, n);
            const T* src = a.obst + n;
            T* dst = stage;
            for (int m = 0; m < Mv; ++m, src += (int64_t)SCCAV_NFIELD * N, dst += NF * B) {
#pragma unroll
                for (int f = 0; f < NF; ++f) cp_async_elem<T>(dst + f * B, src + (int64_t)f * N);
            }
            load_weights<T>(P, a.pv, N, n, alpha, R00, R01, R10, R11);
            Ri = load_rinv<T>(P, a.pv, R00, R01, R10, R11);
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        T sth, cth;
        R::sincos_(th, &sth, &cth);
        const T vlr = v / P.lr;
        const int model = MODEL >= 0 ? MODEL : P.model;
        T r0 = ur0, r1;
        if (model == SCCAV_MODEL_KBM) r1 = (ur0 * R::tan_(ur1)) / P.L;                      // cbf.py:75
        else if (model == SCCAV_MODEL_DUM) r1 = ur1;                                         // cbf.py:253
        else if (P.flags & SCCAV_FLAG_BETA_IO) r1 = ur1;                                     // the caller holds beta
        else r1 = R::atan2_(P.lr * R::tan_(ur1), P.lf + P.lr);                               // cbf.py:175
        T hmin = R::inf(), worst = -R::inf();
        bool feas = true;
        RowNz nz{0u, 0u};
        QpScan<T> scan;
        scan.reset();
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        {
            const T* f = stage;
            for (int m = 0; m < Mv; ++m, f += NF * B) {
                T g[NF];
#pragma unroll
                for (int i = 0; i < NF; ++i) g[i] = f[i * B];
                Partials<T> p;
                if (SPEC == SCCAV_SPEC_ELLIPSE) {
                    const bool st_ = (a.sd.d[0] & SCCAV_SLOT_STATIC) != 0;
                    p = ellipse_partials<T>(x, y, g[0], g[1], g[2], g[3], g[4], st_ ? T(0) : g[5], st_ ? T(0) : g[NF - 1]);
                }
                else if (NF >= 8) p = ellipse_prep_partials<T>(x, y, g[0], g[1], g[2], g[3], g[4], g[5], g[NF - 2], g[NF - 1]);
                else p = ellipse_prep_partials<T>(x, y, g[0], g[1], g[2], g[3], g[4], g[5], T(0), T(0), true);
                put_row<T, COOP, NF, MODEL, MODEL == SCCAV_MODEL_DBM>(P, p, sth, cth, v, alpha, vlr, r0, r1, stage, B, m, hmin, worst, feas, nz, &scan, &Ri);
            }
        }
        T q0 = r0, q1 = r1;
        uint32_t mask = 0u;
        int st = SCCAV_STATUS_INACTIVE;
        if (COOP) {
            st = qp2_solve_active_warp<T>(!feas, warp_rows, B, lane, Mv, nz, r0, r1, R00, R01, R10, R11, Ri,
                                          a.pv.R == nullptr, worst, scan, q0, q1, mask, enumerate, NF);
            __syncwarp();                      // the next vehicle's fields overwrite the columns read above
        } else if (!feas) {
            // one thread, one problem: plain enumeration (what pays when the rows themselves are the cost)
            const RowView<T> rv{stage, B, NF};
            st = qp2_solve_active<T>(rv, Mv, nz, r0, r1, R00, R01, R10, R11, Ri, worst, q0, q1, mask);
        }
        if (valid) {
            T u0 = ur0, u1 = ur1;                                           // empty obstacle list: u = u_ref
            if (Mv > 0) { u0 = q0; u1 = filter_convert<T, -1, true>(P, q0, q1, r0); }
            a.u[n] = u0;
            a.u[N + n] = u1;
            if (a.mask) a.mask[n] = mask;
            if (a.status) a.status[n] = (uint8_t)st;
            if (a.h_min) a.h_min[n] = hmin;
        }
    }
}

// ------------------------------------------------------------------------------------------ K3
template <typename T> struct RolloutArgs {
    Params<T> P;
    SlotDesc sd;
    int M;
    int64_t N;
    int T_steps;
    int np;              // course points
    const T* state;
    T* obst;
    T* pre;              // scratch [M][SCCAV_NPRE][N] for loop-invariant obstacle terms (may be NULL)
    const void* sym;     // rollout-private 16-byte-paired rows: static prepared ellipses in symmetric form [M][3][N] pairs (pack_sym_kernel),
                         // or canonical ellipses with their hoisted terms [M][5][N] pairs (written by the rollout kernel itself); or NULL
    const T* cx;
    const T* cy;
    const T* cyaw;
    // several roads in one launch (sccav_rollout_roads_*): road c = course arrays + c * road_stride with road_np[c]
    // points; its vehicles are [c * group, (c + 1) * group), served by ctas_per_road consecutive CTAs.  n_roads = 0: one road.
    int n_roads, road_stride, ctas_per_road;
    int64_t group;
    const int32_t* road_np;
    PerVehicle<T> pv;
    // outputs
    T* o_state;
    int32_t* o_steps;
    int32_t* o_tidx;
    int32_t* o_nact;
    int32_t* o_ninf;
    T* o_hmin;
    T* o_bmin;
    T* o_bmax;
    T* o_bint;
    T* o_traj;
    int32_t* o_tridx;
    uint32_t* o_trmask;
    int32_t* o_evals;
};

// RadialObstacleSpawner.update_seekers -- radial_dynamic_obstacles.py:193-239
// direct (SCCAV_FLAG_SEEKER_DIRECT): (cos, sin) of the heading as the normalised offset -- the same unit vector, a few ulp from
// sincos(atan2(dy, dx)); a seeker exactly on the ego heads along +x like atan2(0, 0) = 0
template <typename T>
__device__ __forceinline__ void seeker_update(T* f, int64_t fs, T ex, T ey, T dt, T k, T vmin, bool direct = false) {
    typedef Real<T> R;
    T cx = f[0], cy = f[fs];
    const T hyp = R::hypot_(ex - cx, ey - cy);
    T vmag = k * hyp;
    if (vmag < vmin) vmag = vmin;
    T s, c;
    if (direct) {
        if (hyp > T(0)) { c = (ex - cx) / hyp; s = (ey - cy) / hyp; }
        else { c = T(1); s = T(0); }
    } else {
        T yaw = R::atan2_(ey - cy, ex - cx);
        R::sincos_(yaw, &s, &c);
    }
    T vx = vmag * c, vy = vmag * s;
    f[5 * fs] = vx;
    f[6 * fs] = vy;
    f[0] = cx + vx * dt;
    f[fs] = cy + vy * dt;
}

// Up to 512 threads = 16 warps per CTA, one CTA per SM: the register file is per SM sub-partition (16,384 registers
// each, warps dealt round-robin), so a CTA of more than 12 warps caps the kernel at 128 registers per thread -- the same
// cap for 14 warps (65,536 vehicles / 148 SMs = 443 vehicles per SM: config 2 runs 448-thread CTAs in one wave) and for
// 16 (batches of many waves, and 1,024-vehicle roads as two CTAs).
#define SCCAV_ROLLOUT_MAXB 512

// Shared-memory layout of the rollout kernel (bytes), shared by the launcher and the kernel:
//   [ course xy : T2 x nslot (leaf-padded) ][ node chords : 16 B x nodes ][ node (1 / len^2, radius) : 8 B x nodes ][ header : level table int x 32, origin T x 2, extent float ]
//   [ cover table : 2 B x nleaf x kc ][ cover counts : 1 B x nleaf, padded to 16 ][ cyaw : T x np_pad ][ rows : T x 3 M block ]
template <typename T> struct RolloutSmem {
    int nslot, units, np_pad;
    int kc;
    size_t off_node, off_ir, off_hdr, off_cov, off_ncov, off_cyaw, off_rows, course_bytes;
    // trig: the cyaw region holds (sin, cos) of every course yaw instead of the yaw (the fused-steer instances)
    __host__ __device__ RolloutSmem(int np, bool course_smem, bool trig = false) {
        nslot = course_smem ? course_nslot(np) : 0;
        units = course_smem ? course_node_units(np) : 0;
        np_pad = course_smem ? ((np + 1) & ~1) : 0;
        size_t o = (size_t)nslot * 2 * sizeof(T);
        o = (o + 15) & ~(size_t)15;
        off_node = o; o += (size_t)units * 16;
        off_ir = o; o += ((size_t)units * 8 + 15) & ~(size_t)15;
        off_hdr = o; o += course_smem ? 160 : 0;       // 2 x SCCAV_MAX_LEVELS ints | 2 T (at +128) | float (at +144)
        kc = course_smem ? cover_kc(course_levels(np, nullptr, nullptr)) : 0;
        off_cov = o; o += ((size_t)course_nleaf(np) * kc * sizeof(uint16_t) + 15) & ~(size_t)15;
        off_ncov = o; o += course_smem ? (((size_t)course_nleaf(np) + 15) & ~(size_t)15) : 0;
        off_cyaw = o; o += (size_t)np_pad * sizeof(T) * (trig ? 2 : 1);
        off_rows = o;
        course_bytes = o;
    }
};

// Stage one course into shared memory (leaf-padded points, optionally cyaw) and build the capsules of its tree
// (course_index.cuh).  Called by every thread of the CTA; `scratch` = at least 32 doubles of shared memory that are
// free during the build.  Nodes are built warp-cooperatively: the lanes of a warp share the points of one node and
// reduce the largest chord distance (a maximum: the same radius as the serial capsule_build).
template <typename T, typename T2, bool TRIG = false>
__device__ __forceinline__ CourseIndex<T, T2> course_stage(unsigned char* smem, const RolloutSmem<T>& lay, int np,
                                                           const T* __restrict__ cx, const T* __restrict__ cy,
                                                           const T* __restrict__ cyaw, T* s_cyaw, double* scratch) {
    typedef Real<T> R;
    T2* s_cxy = reinterpret_cast<T2*>(smem);
    int* s_lev = reinterpret_cast<int*>(smem + lay.off_hdr);
    T* s_org = reinterpret_cast<T*>(smem + lay.off_hdr + 128);
    float* s_ext = reinterpret_cast<float*>(smem + lay.off_hdr + 144);
    CourseIndex<T, T2> ci;
    ci.xy = s_cxy;
    ci.chord = reinterpret_cast<float4*>(smem + lay.off_node);
    ci.ir = reinterpret_cast<float2*>(smem + lay.off_ir);
    ci.lev = s_lev; ci.org = s_org; ci.ext = s_ext;
    ci.np = np; ci.nleaf = course_nleaf(np);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    // origin of the fp32 frame: the middle point of the course
    const T ox = cx[np / 2], oy = cy[np / 2];
    double ext = 0.0;
    for (int i = threadIdx.x; i < np; i += blockDim.x) {
        T px = cx[i], py = cy[i];
        s_cxy[course_slot(i)] = R::make2(px, py);
        if (s_cyaw) {
            if (TRIG) { T sc, cc; R::sincos_(cyaw[i], &sc, &cc); s_cyaw[2 * i] = sc; s_cyaw[2 * i + 1] = cc; }
            else s_cyaw[i] = cyaw[i];
        }
        ext = fmax(ext, fmax(fabs((double)px - (double)ox), fabs((double)py - (double)oy)));
    }
    // the tail of the last leaf: copies of the last point (course_nslot)
    for (int i = np + threadIdx.x; i < ci.nleaf * SCCAV_LEAF; i += blockDim.x) s_cxy[course_slot(i)] = R::make2(cx[np - 1], cy[np - 1]);
    // extent of the course around the origin: CTA-wide max
    ext = warp_max<double>(ext);
    if (lane == 0) scratch[warp] = ext;
    __syncthreads();
    ext = 0.0;
    for (int w = 0; w < nwarp; ++w) ext = fmax(ext, scratch[w]);
    const float extf = course_ext_inflate(ext);
    if (threadIdx.x == 0) { s_org[0] = ox; s_org[1] = oy; s_ext[0] = extf; }
    // level by level (same recurrence as course_levels)
    int base = 0, cnt = ci.nleaf, k = 0;
    for (;;) {
        if (threadIdx.x == 0) { s_lev[2 * k] = base; s_lev[2 * k + 1] = cnt; }
        for (int j = warp; j < cnt; j += nwarp) {
            int lo, hi;
            node_range(np, k, j, lo, hi);
            const T2 pa = s_cxy[course_slot(lo)], pb = s_cxy[course_slot(hi - 1)];
            float4 c;
            double inv;
            capsule_chord((double)pa.x - (double)ox, (double)pa.y - (double)oy, (double)pb.x - (double)ox, (double)pb.y - (double)oy, c, inv);
            double m = 0.0;
            for (int i = lo + lane; i < hi; i += 32) {
                const T2 p = s_cxy[course_slot(i)];
                const double d2 = chord_dist2(c, inv, (double)p.x - (double)ox, (double)p.y - (double)oy);
                m = d2 > m ? d2 : m;
            }
            m = warp_max<double>(m);
            if (lane == 0) {
                const float2 r = capsule_ir(inv, m, extf);
                ci.chord[base + j] = c;
                ci.ir[base + j] = r;
            }
        }
        base += cnt;
        ++k;
        if (cnt <= 2 || k >= SCCAV_MAX_LEVELS) break;
        cnt = (cnt + 1) >> 1;
    }
    ci.nlev = k;
    __syncthreads();
    // the cover of every window leaf (course_index.cuh), and the dummy node its padding points at
    uint8_t* s_ncov = reinterpret_cast<uint8_t*>(smem + lay.off_ncov);
    uint16_t* s_cov = reinterpret_cast<uint16_t*>(smem + lay.off_cov);
    const int dummy = base;                  // the first id after the last level
    if (threadIdx.x == 0) dummy_node(ci.chord, ci.ir, dummy);
    for (int w = threadIdx.x; w < ci.nleaf; w += blockDim.x) s_ncov[w] = (uint8_t)cover_row(w, k, s_lev, lay.kc, dummy, s_cov + (size_t)w * lay.kc);
    ci.ncover = s_ncov; ci.cov = s_cov; ci.kc = lay.kc; ci.dummy = dummy;
    __syncthreads();
    return ci;
}

// FAST = the launch is known to be (model DBM, Stanley nominal, no seekers): those three become compile-time
// constants, which removes the other plants, the seeker loop and the per-row model dispatch from the instance the
// headline configurations run (a smaller loop body: fewer instructions and fewer instruction-cache misses).
// FUSED: SCCAV_FLAG_FUSED_STEER known at compile time (1 / 0; the FAST instances) or read from the flags (-1).
template <typename T, bool COURSE_SMEM, int SPEC, bool FAST = false, int FUSED = -1>
__global__ void __launch_bounds__(SCCAV_ROLLOUT_MAXB, 1) rollout_kernel(const __grid_constant__ RolloutArgs<T> a) {
    typedef Real<T> R;
    typedef typename R::T2 T2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // the road of this CTA (the shared-memory layout is sized for the longest road, a.np)
    const int road = a.n_roads > 0 ? (int)(blockIdx.x / a.ctas_per_road) : 0;
    int np = a.np;
    if (a.n_roads > 0) {
        np = a.road_np[road];
        np = np < 1 ? 1 : (np > a.np ? a.np : np);
    }
    const T* __restrict__ g_cx = a.cx + (a.n_roads > 0 ? (int64_t)road * a.road_stride : 0);
    const T* __restrict__ g_cy = a.cy + (a.n_roads > 0 ? (int64_t)road * a.road_stride : 0);
    const T* __restrict__ g_cyaw = a.cyaw + (a.n_roads > 0 ? (int64_t)road * a.road_stride : 0);
    constexpr bool TRIG = FAST && FUSED == 1;       // (sin, cos) of the course yaws in place of the yaws (stanley_law_beta)
    const RolloutSmem<T> lay(a.np, COURSE_SMEM, TRIG);
    T2* s_cxy = reinterpret_cast<T2*>(smem_raw);
    T* s_cyaw = reinterpret_cast<T*>(smem_raw + lay.off_cyaw);
    T* rows = reinterpret_cast<T*>(smem_raw + lay.off_rows) + threadIdx.x;
    const int stride = blockDim.x;
    const bool stan = FAST ? true : a.P.nominal == SCCAV_NOMINAL_STANLEY;
    const int model = FAST ? SCCAV_MODEL_DBM : a.P.model;
    constexpr int MODEL = FAST ? SCCAV_MODEL_DBM : -1;
    CourseIndex<T, T2> ci;
    ci.xy = s_cxy; ci.chord = nullptr; ci.ir = nullptr; ci.lev = nullptr; ci.org = nullptr; ci.ext = nullptr; ci.ncover = nullptr; ci.cov = nullptr; ci.kc = 0; ci.dummy = 0;
    ci.np = np; ci.nleaf = 0; ci.nlev = 0;
    // stage the course once per CTA (leaf-padded) and build the capsules of its tree
    if (COURSE_SMEM && stan)
        ci = course_stage<T, T2, TRIG>(smem_raw, lay, np, g_cx, g_cy, g_cyaw, s_cyaw, reinterpret_cast<double*>(smem_raw + lay.off_rows));
    const int64_t N = a.N;
    int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a.n_roads > 0) {
        const int64_t in_road = (int64_t)(blockIdx.x - road * a.ctas_per_road) * blockDim.x + threadIdx.x;
        if (in_road >= a.group) return;
        n = (int64_t)road * a.group + in_road;
    }
    if (n >= N) return;

    const Params<T>& P = a.P;
    T x = a.state[n], y = a.state[N + n], yaw = a.state[2 * N + n], v = a.state[3 * N + n];
    // the compile-time ellipse instances are launched with launch-wide weights and target speed only (the launcher
    // sends per-vehicle ones to the general instances): constant-bank operands instead of 12 registers held across the loop
    constexpr bool UW = FAST && SPEC != SCCAV_SPEC_GENERIC;
    T alpha, R00, R01, R10, R11;
    if (UW) { alpha = P.alpha; R00 = P.R[0]; R01 = P.R[1]; R10 = P.R[2]; R11 = P.R[3]; }
    else load_weights<T>(P, a.pv, N, n, alpha, R00, R01, R10, R11);
    const T tspeed = UW ? P.target_speed : (a.pv.target_speed ? a.pv.target_speed[n] : P.target_speed);
    const int last_idx = np - 1;
    const int Mv = slot_count<T>(a.pv, a.M, n);
    const bool filt = Mv > 0 && model != SCCAV_MODEL_NONE;

    // loop-invariant obstacle terms (ellipses): once per (vehicle, slot) into the scratch; `moving`
    // = slots whose ellipse has a velocity (their h_t needs vx, vy, a^2, b^2 every step)
    uint32_t moving = 0u;
    if (a.pre && filt) {
        for (int m = 0; m < Mv; ++m) {
            const int desc = a.sd.d[m];
            if ((desc & SCCAV_SLOT_TYPE_MASK) == SCCAV_SLOT_RADIAL && (P.flags & SCCAV_FLAG_PREPARED_ROWS)) {
                // prepared RADIAL rows: the reciprocals of the (constant) half axes, once per launch
                const int64_t nr = (desc & SCCAV_SLOT_SHARED) ? 0 : n;
                const T* fr = a.obst + (int64_t)m * SCCAV_NFIELD * N + nr;
                T* pr = a.pre + (int64_t)m * SCCAV_NPRE * N + n;
                pr[0] = T(1) / fr[2 * N];
                pr[N] = T(1) / fr[3 * N];
                continue;
            }
            if ((desc & SCCAV_SLOT_TYPE_MASK) != SCCAV_SLOT_ELLIPSE) continue;
            const int64_t nn = (desc & SCCAV_SLOT_SHARED) ? 0 : n;
            const T* f = a.obst + (int64_t)m * SCCAV_NFIELD * N + nn;
            ellipse_precompute<T>(f[2 * N], f[3 * N], f[4 * N], a.pre + (int64_t)m * SCCAV_NPRE * N + n, N);
            if (FAST && SPEC == SCCAV_SPEC_ELLIPSE && a.sym && !(desc & SCCAV_SLOT_SHARED)) {
                // + the paired copy the all-static row loop reads (path.cuh): (cx, cy), (a, b), (ct, st), (k0, k1), (k2, k3)
                const T* q = a.pre + (int64_t)m * SCCAV_NPRE * N + n;
                T2* g = reinterpret_cast<T2*>(const_cast<void*>(a.sym)) + (int64_t)m * 5 * N + n;
                g[0] = R::make2(f[0], f[N]);
                g[N] = R::make2(f[2 * N], f[3 * N]);
                g[2 * N] = R::make2(q[0], q[N]);
                g[3 * N] = R::make2(q[2 * N], q[3 * N]);
                g[4 * N] = R::make2(q[4 * N], q[5 * N]);
            }
            if (!(desc & SCCAV_SLOT_STATIC) && (f[5 * N] != T(0) || f[6 * N] != T(0))) moving |= 1u << m;
        }
    }

    T time = T(0);
    int target_idx = 0, near_idx = 0, adv = 0, evals = 0;
    T syaw, cyw;
    R::sincos_(yaw, &syaw, &cyw);
    if (stan) {
        // target_idx, _ = calc_target_index(state, cx, cy)   sce.py:605
        T fx = x + P.L * cyw, fy = y + P.L * syaw;
        if (COURSE_SMEM) target_idx = course_nearest_full<T, T2>(s_cxy, np, fx, fy);
        else {
            T best = R::inf();
            for (int i = 0; i < np; ++i) {
                T dx = fx - g_cx[i], dy = fy - g_cy[i];
                T d2 = dx * dx + dy * dy;
                if (d2 < best) { best = d2; target_idx = i; }
            }
        }
        near_idx = target_idx;
        evals = np;
    }
    int steps = 0, nact = 0, ninf = 0;
    T hmin_all = R::inf(), bmin = R::inf(), bmax = -R::inf(), bint = T(0);
    const int stride_rec = P.record_stride;
    // SCCAV_FLAG_FUSED_STEER (DBM only): beta(max_steer) with the operations the clipped plant would use
    const bool fused = (FUSED >= 0 ? FUSED != 0 : (P.flags & SCCAV_FLAG_FUSED_STEER) != 0) && model == SCCAV_MODEL_DBM;
    const T beta_clip = R::atan2_(P.lr * R::tan_(P.max_steer), P.lf + P.lr);

    while (steps < a.T_steps) {
        if (P.terminate && !(P.t_max >= time && last_idx > target_idx)) break;          // sce.py:630
        // ---- nominal control
        T ur0, ur1;
        T beta_ref = T(0);                   // (fused-steer instances: the reference point's beta, formed by the Stanley law itself)
        if (stan) {
            T a_ref = P.Kp * (tspeed - v);                                                 // sce.py:135-143
            T d_ref;
            if (COURSE_SMEM && TRIG) {
                // fused steering: beta_ref straight from the Stanley law (path.cuh, stanley_law_beta); delta_ref itself is only
                // needed by a vehicle without obstacles, which takes the literal law on the yaws in global memory
                T fx, fy;
                const int idx = stanley_search<T, T2>(P, ci, x, y, syaw, cyw, near_idx, adv, &evals, fx, fy);
                if (filt) { beta_ref = stanley_law_beta<T, T2>(P, ci.pt(idx), reinterpret_cast<const T2*>(s_cyaw), idx, fx, fy, v, target_idx, syaw, cyw); d_ref = T(0); }
                else d_ref = stanley_law<T, T2, true>(P, ci.pt(idx), g_cyaw, idx, fx, fy, yaw, v, target_idx, syaw, cyw);
            } else
            if (COURSE_SMEM) d_ref = stanley<T, T2, (FUSED == 1)>(P, ci, s_cyaw, x, y, yaw, v, syaw, cyw, target_idx, near_idx, adv, &evals);
            else {
                // global-memory course fallback (P too large for shared memory): exhaustive scan
                T fx = x + P.L * cyw, fy = y + P.L * syaw;
                T best = R::inf();
                int idx = 0;
                for (int i = 0; i < np; ++i) {
                    T dx = fx - g_cx[i], dy = fy - g_cy[i];
                    T d2 = dx * dx + dy * dy;
                    if (d2 < best) { best = d2; idx = i; }
                }
                evals += np;
                d_ref = stanley_law<T, T2>(P, R::make2(g_cx[idx], g_cy[idx]), g_cyaw, idx, fx, fy, yaw, v, target_idx);
            }
            ur0 = (model == SCCAV_MODEL_KBM) ? tspeed : a_ref;                             // sce.py:646-648
            ur1 = d_ref;
        } else {
            ur0 = P.uref0;
            ur1 = P.uref1;
        }
        // ---- filter
        T u0 = ur0, u1 = ur1, u1raw = ur1, hmin = R::inf();
        uint32_t mask = 0u;
        int status = SCCAV_STATUS_INACTIVE;
        if (filt)
            status = filter_vehicle<T, SPEC, MODEL, FAST>(P, a.sd, Mv, N, n, a.obst, x, y, yaw, v, syaw, cyw, alpha, R00, R01, R10, R11, UW || a.pv.R == nullptr,
                                             ur0, ur1, rows, stride, u0, u1, u1raw, mask, hmin, a.pre, moving, !fused, TRIG ? &beta_ref : nullptr,
                                             (FAST && SPEC != SCCAV_SPEC_GENERIC) ? a.sym : nullptr);
        // ---- plant
        T px = x, py = y, pyaw = yaw, pv_ = v;
        T beta = T(0);
        // SCCAV_FLAG_FUSED_STEER: beta straight from the QP's beta* (include/sccav_cbf.h); delta only where it is recorded
        const bool rec_now = stride_rec > 0 && (steps % stride_rec) == 0;
        bool fused_done = false;
        if (fused && filt && R::abs_(u1raw) < T(1.5)) {
            beta = u1raw;
            if (beta < -beta_clip) beta = -beta_clip;
            if (beta > beta_clip) beta = beta_clip;
            if (rec_now) u1 = filter_convert<T, MODEL>(P, u0, u1raw, ur0);
            fused_done = true;
        } else if (fused && filt) {
            u1 = filter_convert<T, MODEL>(P, u0, u1raw, ur0);
        }
        T delta = u1;
        if (delta < -P.max_steer) delta = -P.max_steer;
        if (delta > P.max_steer) delta = P.max_steer;
        if (model == SCCAV_MODEL_DBM) {
            // State.update_com   sce.py:122-131 (no yaw normalisation)
            if (!fused_done) beta = R::atan2_(P.lr * R::tan_(delta), P.lf + P.lr);
            x = x + (v * cyw - (v * syaw) * beta) * P.dt;
            y = y + (v * syaw + (v * cyw) * beta) * P.dt;
            yaw = yaw + ((v * beta) / P.lr) * P.dt;
            v = v + u0 * P.dt;
        } else {
            // State.update / update_by_vel   sce.py:86-120
            x = x + (v * cyw) * P.dt;
            y = y + (v * syaw) * P.dt;
            yaw = yaw + ((v / P.L) * R::tan_(delta)) * P.dt;
            yaw = normalize_angle<T>(yaw);
            if (model == SCCAV_MODEL_KBM) v = u0;
            else v = v + u0 * P.dt;
        }
        R::sincos_(yaw, &syaw, &cyw);
        // ---- moving obstacles
        if (FAST ? false : (P.seeker != 0)) {
            for (int m = 0; m < Mv; ++m) {
                const int desc = a.sd.d[m];
                if ((desc & SCCAV_SLOT_TYPE_MASK) == SCCAV_SLOT_RADIAL && !(desc & SCCAV_SLOT_SHARED))
                    seeker_update<T>(a.obst + (int64_t)m * SCCAV_NFIELD * N + n, N, x, y, P.dt, P.seeker_k, P.seeker_vmin,
                                     (P.flags & SCCAV_FLAG_SEEKER_DIRECT) != 0);
            }
        }
        // ---- bookkeeping
        if (rec_now) {
            const int64_t rec = steps / stride_rec;
            if (a.o_traj) {
                T* tr = a.o_traj + rec * SCCAV_TRAJ_FIELDS * N + n;
                tr[0] = px; tr[N] = py; tr[2 * N] = pyaw; tr[3 * N] = pv_;
                tr[4 * N] = u0; tr[5 * N] = u1; tr[6 * N] = beta;
            }
            if (a.o_tridx) a.o_tridx[rec * N + n] = target_idx;
            if (a.o_trmask) a.o_trmask[rec * N + n] = mask;
        }
        time = time + P.dt;                                                                // sce.py:830
        ++steps;
        nact += (mask != 0u);
        ninf += (status == SCCAV_STATUS_INFEASIBLE);
        if (hmin < hmin_all) hmin_all = hmin;
        if (beta < bmin) bmin = beta;
        if (beta > bmax) bmax = beta;
        bint = bint + beta * P.dt;
    }
    a.o_state[n] = x; a.o_state[N + n] = y; a.o_state[2 * N + n] = yaw; a.o_state[3 * N + n] = v;
    if (a.o_steps) a.o_steps[n] = steps;
    if (a.o_tidx) a.o_tidx[n] = target_idx;
    if (a.o_nact) a.o_nact[n] = nact;
    if (a.o_ninf) a.o_ninf[n] = ninf;
    if (a.o_hmin) a.o_hmin[n] = hmin_all;
    if (a.o_bmin) a.o_bmin[n] = bmin;
    if (a.o_bmax) a.o_bmax[n] = bmax;
    if (a.o_bint) a.o_bint[n] = bint;
    if (a.o_evals) a.o_evals[n] = evals;
}

// ------------------------------------------------------------------------------------------ KS
// One Stanley control call for N vehicles (LateralStanley.control, cbf/controllers.py:104-151 ==
// stanley_control, stanley_controller_ellipse.py:146-169): nearest way-point (exact pruned search
// from the previous target index), front-axle error, monotone index clamp, steering law.
template <typename T> struct StanleyArgs {
    Params<T> P;          // L (front-axle offset), k_stanley, ks_stanley
    int64_t N;
    int np;
    const T* state;       // [4][N]
    const T* front;       // [2][N] externally supplied front-axle coordinates, or NULL
    const T* cx;
    const T* cy;
    const T* cyaw;
    int32_t* target_idx;  // [N] in: last target index, out: new one
    T* delta;             // [N]
    T* err;               // [N] front-axle cross-track error (may be NULL)
};

template <typename T>
__global__ void __launch_bounds__(256) stanley_kernel(const __grid_constant__ StanleyArgs<T> a) {
    typedef Real<T> R;
    typedef typename R::T2 T2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int np = a.np;
    const RolloutSmem<T> lay(np, true);
    // (cyaw stays in global memory here: one read per vehicle; its region is the build's scratch)
    const CourseIndex<T, T2> ci = course_stage<T, T2>(smem_raw, lay, np, a.cx, a.cy, nullptr, nullptr,
                                                      reinterpret_cast<double*>(smem_raw + lay.off_cyaw));
    const int64_t N = a.N;
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
        T x = a.state[n], y = a.state[N + n], yaw = a.state[2 * N + n], v = a.state[3 * N + n];
        T fx, fy;
        if (a.front) { fx = a.front[n]; fy = a.front[N + n]; }           // controllers.py:105-110
        else {
            T syaw, cyw;
            R::sincos_(yaw, &syaw, &cyw);
            fx = x + a.P.L * cyw;
            fy = y + a.P.L * syaw;
        }
        // (a target index past the course -- e.g. kept from a longer one -- is clamped: the reference would raise IndexError)
        int tidx = a.target_idx[n];
        tidx = tidx < 0 ? 0 : (tidx >= np ? np - 1 : tidx);
        const int idx = course_nearest<T, T2>(ci, fx, fy, tidx, nullptr);
        T2 c = ci.pt(idx);
        T s2, c2;
        R::sincos_(yaw + R::pi() / T(2), &s2, &c2);
        T e = (fx - c.x) * (-c2) + (fy - c.y) * (-s2);
        int use = idx;
        if (tidx >= idx) use = tidx;
        T theta_e = normalize_angle<T>(a.cyaw[use] - yaw);
        T theta_d = R::atan2_(a.P.k_stanley * e, v + a.P.ks_stanley);
        a.delta[n] = theta_e + theta_d;
        a.target_idx[n] = use;
        if (a.err) a.err[n] = e;
    }
}

// ------------------------------------------------------------------------------------------ KD
// The per-tick loop of the CARLA driver (carla_scripts/multi_obstacle_CBF_local_with_lanes.py:861-983), T ticks in ONE
// launch, thread = ego vehicle, everything a tick carries over in registers:
//   lateral_stanley.control   cbf/controllers.py:104-151 (front axle at lf, atan2(k e, v + ks), its own last_target_idx)
//   delta *= rad_to_steer     :876-877
//   acc_pid.set_dt / control  cbf/controllers.py:153-180 (kp, ki, kd; the tick's own dt), reference speed = trajectory[idx][3]
//   obstacle list             the shared lane slots + one fresh CollisionCone2D per actor box of the tick (:913-928)
//   solve_cbf                 DBM_CBF_2DS (cbf/cbf.py:166-220); an empty list passes u_ref through (:935-936)
//   throttle / brake / steer  :955-980
// The ego state of every tick comes from the caller's stream (the simulator owns the plant); without a stream a
// stand-in plant -- State.update_com with the filtered (a, delta) and the tick's dt -- closes the loop.
template <typename T> struct DriveArgs {
    Params<T> P;
    SlotDesc sd;
    int M, n_fixed, K, T_ticks, np;
    int64_t N;
    T kp, ki, kd, rad_to_steer, max_steer_cmd, rate, cone_buffer;
    int act_flags;
    const T* state0;       // [4][N] initial ego state (stand-in plant) or NULL
    const T* ego;          // [T][4][N] ego state of every tick or NULL
    const int32_t* box_id; // [T][K][N] or NULL (K = 0: the cone slots of `obst` are used as they are, count from pv)
    const T* box;          // [T][K][6][N]
    const T* dt;           // [T] or NULL (P.dt)
    T* obst;               // [M][8][N]: slots [0, n_fixed) given by the caller, the rest rebuilt from the boxes every tick
    const T* cx; const T* cy; const T* cyaw; const T* cv;      // trajectory (x, y, yaw, v) [np]
    PerVehicle<T> pv;
    int32_t* target_idx;   // [N] in / out: LateralStanley's last target index
    T* carry;              // [4][N] in / out: PID e_prev, PID integral, throttle_prev, brake_prev
    T* act;                // [T][3][N] throttle, brake, steer
    T* u;                  // [T][2][N] filtered (a, delta) or NULL
    uint32_t* mask;        // [T][N] or NULL
    int32_t* tidx;         // [T][N] or NULL
    T* o_state;            // [4][N] final ego state or NULL
};

template <typename T>
__global__ void __launch_bounds__(256) drive_ticks_kernel(const __grid_constant__ DriveArgs<T> a) {
    typedef Real<T> R;
    typedef typename R::T2 T2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int np = a.np;
    const RolloutSmem<T> lay(np, true);
    T* s_cyaw = reinterpret_cast<T*>(smem_raw + lay.off_cyaw);
    T* rows = reinterpret_cast<T*>(smem_raw + lay.off_rows) + threadIdx.x;
    const int stride = blockDim.x;
    const CourseIndex<T, T2> ci = course_stage<T, T2>(smem_raw, lay, np, a.cx, a.cy, a.cyaw, s_cyaw,
                                                      reinterpret_cast<double*>(smem_raw + lay.off_rows));
    const int64_t N = a.N;
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const Params<T>& P = a.P;
    T alpha, R00, R01, R10, R11;
    load_weights<T>(P, a.pv, N, n, alpha, R00, R01, R10, R11);
    T x = T(0), y = T(0), yaw = T(0), v = T(0);
    if (a.state0) { x = a.state0[n]; y = a.state0[N + n]; yaw = a.state0[2 * N + n]; v = a.state0[3 * N + n]; }
    int last_idx = a.target_idx[n], near_idx = last_idx, adv = 0;
    T eprev = a.carry[n], ie = a.carry[N + n], thr_prev = a.carry[2 * N + n], brk_prev = a.carry[3 * N + n];
    int Mv = slot_count<T>(a.pv, a.M, n);
    for (int t = 0; t < a.T_ticks; ++t) {
        if (a.ego) {
            const T* e = a.ego + (int64_t)t * 4 * N + n;
            x = e[0]; y = e[N]; yaw = e[2 * N]; v = e[3 * N];
        }
        const T dt = a.dt ? a.dt[t] : P.dt;
        // ---- LateralStanley.control (class form: front axle at L = lf)
        T syaw, cyw;
        R::sincos_(yaw, &syaw, &cyw);
        const T fx = x + P.L * cyw, fy = y + P.L * syaw;
        const int idx = course_nearest<T, T2>(ci, fx, fy, near_idx + adv, nullptr);
        adv = idx - near_idx;
        adv = adv < -4 * SCCAV_LEAF ? 0 : (adv > 4 * SCCAV_LEAF ? 0 : adv);
        near_idx = idx;
        T delta = stanley_law<T, T2>(P, ci.pt(idx), s_cyaw, idx, fx, fy, yaw, v, last_idx);
        delta = delta * a.rad_to_steer;                                                  // :876-877
        // ---- PID1.control(ego_v, trajectory[target_idx][3])
        const T e = a.cv[last_idx] - v;                                                  // controllers.py:174
        const T de = (e - eprev) / dt;                                                   // :175
        ie = ie + dt * e;                                                                // :176
        const T u_a = (a.kp * e + a.ki * ie) + a.kd * de;                                // :178
        eprev = e;
        // ---- obstacle list of the tick: fixed slots, then one fresh cone per box (from_bounding_box semantics)
        if (a.box_id) {
            int w = a.n_fixed;
            for (int k = 0; k < a.K; ++k) {
                const int32_t id = a.box_id[((int64_t)t * a.K + k) * N + n];
                if (id < 0 || w >= a.M) continue;
                // CollisionCone2D(a_cone = hypot(extent), s, s_obs = [x, y, yaw, |v|]) with the default buffer  (:918-928;
                // the driver builds the cone itself, so the actor's yaw is kept -- from_bounding_box would zero it)
                const T* bx = a.box + ((int64_t)t * a.K + k) * SCCAV_BOX_FIELDS * N + n;
                T* dst = a.obst + (int64_t)w * SCCAV_NFIELD * N + n;
                dst[0] = bx[2 * N]; dst[N] = bx[3 * N]; dst[2 * N] = bx[4 * N]; dst[3 * N] = bx[5 * N];
                dst[4 * N] = R::hypot_(bx[0], bx[N]) + a.cone_buffer;
                dst[5 * N] = T(0); dst[6 * N] = T(0); dst[7 * N] = T(0);
                ++w;
            }
            Mv = w;
        }
        // ---- solve_cbf
        T u0 = u_a, u1 = delta, u1raw = delta, hmin = R::inf();
        uint32_t mask = 0u;
        if (Mv > 0)
            filter_vehicle<T, SCCAV_SPEC_GENERIC, -1>(P, a.sd, Mv, N, n, a.obst, x, y, yaw, v, syaw, cyw, alpha, R00, R01, R10, R11,
                                                       a.pv.R == nullptr, u_a, delta, rows, stride, u0, u1, u1raw, mask, hmin);
        // ---- actuators
        T throttle, brake, steer;
        actuator_step<T>(u0, u1, a.max_steer_cmd, a.rate, a.act_flags, thr_prev, brk_prev, throttle, brake, steer);
        T* o = a.act + (int64_t)t * 3 * N + n;
        o[0] = throttle; o[N] = brake; o[2 * N] = steer;
        if (a.u) { a.u[(int64_t)t * 2 * N + n] = u0; a.u[((int64_t)t * 2 + 1) * N + n] = u1; }
        if (a.mask) a.mask[(int64_t)t * N + n] = mask;
        if (a.tidx) a.tidx[(int64_t)t * N + n] = last_idx;
        // ---- stand-in plant (no ego stream): State.update_com, sce.py:122-131, with the tick's dt
        if (!a.ego) {
            T dl = u1;
            if (dl < -P.max_steer) dl = -P.max_steer;
            if (dl > P.max_steer) dl = P.max_steer;
            const T beta = R::atan2_(P.lr * R::tan_(dl), P.lf + P.lr);
            x = x + (v * cyw - (v * syaw) * beta) * dt;
            y = y + (v * syaw + (v * cyw) * beta) * dt;
            yaw = yaw + ((v * beta) / P.lr) * dt;
            v = v + u0 * dt;
        }
    }
    a.target_idx[n] = last_idx;
    a.carry[n] = eprev; a.carry[N + n] = ie; a.carry[2 * N + n] = thr_prev; a.carry[3 * N + n] = brk_prev;
    if (a.o_state) { a.o_state[n] = x; a.o_state[N + n] = y; a.o_state[2 * N + n] = yaw; a.o_state[3 * N + n] = v; }
}

}  // namespace sccav
