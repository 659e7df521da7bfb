// path.cuh -- device functions of the CBF-QP hot path (barriers, rows, QP, nominal, plant).
//
// Written for sm_100a.  Every function restates the reference arithmetic in the SAME operation
// order as oracle/oracle.py (compiled with -fmad=false for the fp64 build so that a*b+c is not
// contracted: the waypoint index and the active set are compared bit-for-bit with the oracle).
// Citations are file:line under the reference root.
//
// No tensor cores here on purpose: n = 2 unknowns, tens of rows -- nothing is a dense contraction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/sccav_cbf.h"
#include "course_index.cuh"

namespace sccav {

// ------------------------------------------------------------------------------------------
// scalar traits
// ------------------------------------------------------------------------------------------
template <typename T> struct Real;

template <> struct Real<double> {
    typedef double2 T2;
    static __device__ __forceinline__ double pi() { return 3.141592653589793; }       // np.pi
    static __device__ __forceinline__ double zero_tol() { return 1e-3; }              // cbf/utils.py:27
    static __device__ __forceinline__ double feas_eps() { return 1e-12; }
    static __device__ __forceinline__ double par_eps() { return 1e-12; }
    static __device__ __forceinline__ double tie_eps() { return 1e-9; }
    static __device__ __forceinline__ double qp_tie_margin() { return 1e-6; }
    static __device__ __forceinline__ double qp_res_margin() { return 1.25e-7; }
    static __device__ __forceinline__ double lane_xtol() { return 1e-12; }
    static __device__ __forceinline__ double lane_dtol() { return 8 * 0x1p-52; }
    static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
    static __device__ __forceinline__ void sincos_(double x, double* s, double* c) { ::sincos(x, s, c); }
    static __device__ __forceinline__ double tan_(double x) { return ::tan(x); }
    static __device__ __forceinline__ double atan_(double x) { return ::atan(x); }
    static __device__ __forceinline__ double atan2_(double y, double x) { return ::atan2(y, x); }
    static __device__ __forceinline__ double sqrt_(double x) { return ::sqrt(x); }
    static __device__ __forceinline__ double hypot_(double x, double y) { return ::hypot(x, y); }
    static __device__ __forceinline__ double abs_(double x) { return ::fabs(x); }
    static __device__ __forceinline__ double rint_(double x) { return ::rint(x); }
    static __device__ __forceinline__ double tanh_(double x) { return ::tanh(x); }
    static __device__ __forceinline__ double pow_(double x, double y) { return ::pow(x, y); }
    static __device__ __forceinline__ double ceil_(double x) { return ::ceil(x); }
    static __device__ __forceinline__ T2 make2(double a, double b) { return make_double2(a, b); }
};

template <> struct Real<float> {
    typedef float2 T2;
    static __device__ __forceinline__ float pi() { return 3.14159274f; }
    static __device__ __forceinline__ float zero_tol() { return 1e-3f; }
    static __device__ __forceinline__ float feas_eps() { return 1e-5f; }
    static __device__ __forceinline__ float par_eps() { return 1e-5f; }
    static __device__ __forceinline__ float tie_eps() { return 1e-4f; }
    static __device__ __forceinline__ float qp_tie_margin() { return 1e-2f; }
    static __device__ __forceinline__ float qp_res_margin() { return 1.25e-3f; }
    static __device__ __forceinline__ float lane_xtol() { return 1e-6f; }
    static __device__ __forceinline__ float lane_dtol() { return 8 * 0x1p-23f; }
    static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
    static __device__ __forceinline__ void sincos_(float x, float* s, float* c) { ::sincosf(x, s, c); }
    static __device__ __forceinline__ float tan_(float x) { return ::tanf(x); }
    static __device__ __forceinline__ float atan_(float x) { return ::atanf(x); }
    static __device__ __forceinline__ float atan2_(float y, float x) { return ::atan2f(y, x); }
    static __device__ __forceinline__ float sqrt_(float x) { return ::sqrtf(x); }
    static __device__ __forceinline__ float hypot_(float x, float y) { return ::hypotf(x, y); }
    static __device__ __forceinline__ float abs_(float x) { return ::fabsf(x); }
    static __device__ __forceinline__ float rint_(float x) { return ::rintf(x); }
    static __device__ __forceinline__ float tanh_(float x) { return ::tanhf(x); }
    static __device__ __forceinline__ float pow_(float x, float y) { return ::powf(x, y); }
    static __device__ __forceinline__ float ceil_(float x) { return ::ceilf(x); }
    static __device__ __forceinline__ T2 make2(float a, float b) { return make_float2(a, b); }
};

// parameters in the kernel's precision
template <typename T> struct Params {
    int model, nominal, terminate, seeker, kbm_driver_delta, record_stride, flags;
    T alpha, lr, lf, L, max_steer, dt, k_stanley, ks_stanley, Kp, target_speed, t_max;
    T R[4];
    T Ri[4];             // inverse of R, formed on the host with the operations of RInv below
    T seeker_k, seeker_vmin, uref0, uref1;
    T sadbm_dt;          // SADBM_CBF_2DS: the class's fixed dt
};

struct SlotDesc {
    uint8_t d[SCCAV_MAX_ROWS];
};

template <typename T> struct Partials {
    T h, hx, hy, hth, hv, ht;
};

// cbf/utils.py:93-106 == stanley_controller_ellipse.py:172-185
// The reference's loops never terminate for |a| >~ 1e16 or inf; beyond 1e4 (a diverged scenario)
// whole turns are removed first -- same rule as oracle.normalize_angle.
template <typename T> __device__ __forceinline__ T normalize_angle(T a) {
    const T pi = Real<T>::pi();
    if (!(Real<T>::abs_(a) <= T(1e4))) a = a - (T(2.0) * pi) * Real<T>::rint_(a / (T(2.0) * pi));
    while (a > pi) a -= T(2.0) * pi;
    while (a < -pi) a += T(2.0) * pi;
    return a;
}

// ------------------------------------------------------------------------------------------
// barriers
// ------------------------------------------------------------------------------------------
// Ellipse2D.evaluate/dx/dy/dt -- cbf/obstacles.py:183-230,310-317 (h_t ignores theta, as there)
template <typename T>
__device__ __forceinline__ Partials<T> ellipse_partials(T x, T y, T cx, T cy, T a, T b, T th, T vx, T vy) {
    typedef Real<T> R;
    Partials<T> o;
    T dx = x - cx, dy = y - cy, st, ct;
    R::sincos_(th, &st, &ct);
    T p = dx * ct + dy * st;
    T q = (-dx) * st + dy * ct;
    T pa = p / a, qb = q / b;
    o.h = (pa * pa + qb * qb) - T(1);
    T aa = a * a, bb = b * b;
    o.hx = ((T(2) * ct) / aa) * p + ((T(-2) * st) / bb) * q;
    o.hy = ((T(2) * st) / aa) * p + ((T(2) * ct) / bb) * q;
    o.hth = T(0);
    o.hv = T(0);
    // a static obstacle has h_t = -2 (. * 0 + . * 0) = 0: skip the two divisions (the sign of that zero
    // cannot reach u or the active set)
    o.ht = T(0);
    if (vx != T(0) || vy != T(0)) o.ht = T(-2) * ((dx / aa) * vx + (dy / bb) * vy);
    return o;
}

// Loop-invariant part of ellipse_partials for an obstacle whose fields do not change during a
// rollout: cos/sin(theta) and the four gradient coefficients (obstacles.py:218,229), computed ONCE
// with the same operations, so the per-step values below are bit-identical to ellipse_partials.
#define SCCAV_NPRE 6   // ct, st, 2ct/a^2, -2st/b^2, 2st/a^2, 2ct/b^2
template <typename T>
__device__ __forceinline__ void ellipse_precompute(T a, T b, T th, T* pre, int64_t ps) {
    T st, ct;
    Real<T>::sincos_(th, &st, &ct);
    T aa = a * a, bb = b * b;
    pre[0] = ct;
    pre[ps] = st;
    pre[2 * ps] = (T(2) * ct) / aa;
    pre[3 * ps] = (T(-2) * st) / bb;
    pre[4 * ps] = (T(2) * st) / aa;
    pre[5 * ps] = (T(2) * ct) / bb;
}

template <typename T>
__device__ __forceinline__ Partials<T> ellipse_partials_pre(T x, T y, T cx, T cy, T a, T b, T vx, T vy,
                                                            const T* pre, int64_t ps) {
    Partials<T> o;
    const T ct = pre[0], st = pre[ps];
    T dx = x - cx, dy = y - cy;
    T p = dx * ct + dy * st;
    T q = (-dx) * st + dy * ct;
    T pa = p / a, qb = q / b;
    o.h = (pa * pa + qb * qb) - T(1);
    o.hx = pre[2 * ps] * p + pre[3 * ps] * q;
    o.hy = pre[4 * ps] * p + pre[5 * ps] * q;
    o.hth = T(0);
    o.hv = T(0);
    o.ht = T(0);
    if (vx != T(0) || vy != T(0)) {                       // static obstacle: h_t = -2 (x 0 + y 0) = 0
        T aa = a * a, bb = b * b;
        o.ht = T(-2) * ((dx / aa) * vx + (dy / bb) * vy);
    }
    return o;
}

// ELLIPSE_PREP (include/sccav_cbf.h): the ingest-time half of Ellipse2D -- everything that does not
// depend on the vehicle -- and the per-solve half.  Same functions as ellipse_partials
// (cbf/obstacles.py:193,218,229,316), a few ulp apart; written with explicit fma (this is not
// reference operation order, so contraction is welcome).
template <typename T>
__device__ __forceinline__ void ellipse_prepare(T a, T b, T th, T vx, T vy, T& m00, T& m01, T& m10, T& m11, T& wx, T& wy) {
    T st, ct;
    Real<T>::sincos_(th, &st, &ct);
    m00 = ct / a; m01 = st / a;
    m10 = (-st) / b; m11 = ct / b;
    wx = vx / (a * a); wy = vy / (b * b);
}

// (is_static: h_t = -2 (dx 0 + dy 0) is not evaluated -- it is a zero whose sign cannot reach u or the active set)
template <typename T>
__device__ __forceinline__ Partials<T> ellipse_prep_partials(T x, T y, T cx, T cy, T m00, T m01, T m10, T m11, T wx, T wy,
                                                             bool is_static = false) {
    Partials<T> o;
    T dx = x - cx, dy = y - cy;
    T pa = fma(m00, dx, m01 * dy);
    T qb = fma(m10, dx, m11 * dy);
    o.h = fma(pa, pa, fma(qb, qb, T(-1)));
    o.hx = T(2) * fma(m00, pa, m10 * qb);
    o.hy = T(2) * fma(m01, pa, m11 * qb);
    o.hth = T(0);
    o.hv = T(0);
    o.ht = is_static ? T(0) : T(-2) * fma(dx, wx, dy * wy);
    return o;
}

// single_obstacle_CBF1 -- test_scripts/radial_dynamic_obstacles.py:391-405
template <typename T>
__device__ __forceinline__ Partials<T> radial_partials(T x, T y, T v, T cx, T cy, T a, T b, T kv, T vx, T vy) {
    Partials<T> o;
    T dx = x - cx, dy = y - cy;
    T da = dx / a, db = dy / b;
    o.h = ((da * da + db * db) - T(1)) - ((kv * v) / (T(1) + v));
    T aa = a * a, bb = b * b;
    o.hx = (T(2) * dx) / aa;
    o.hy = (T(2) * dy) / bb;
    o.hth = T(0);
    T opv = T(1) + v;
    o.hv = (-kv) / (opv * opv);
    o.ht = T(-2) * ((dx / aa) * vx + (dy / bb) * vy);
    return o;
}

// SCCAV_FLAG_PREPARED_ROWS in the rollout: the same functions with the reciprocals 1 / a, 1 / b from the launch's scratch and
// vo = v / (1 + v), iopv2 = 1 / (1 + v)^2 formed once per step -- no division per row (eight in the reference's order)
template <typename T> struct RadialStep {
    T vo, iopv2;
};
template <typename T>
__device__ __forceinline__ Partials<T> radial_partials_fast(T x, T y, T cx, T cy, T ia, T ib, T kv, T vx, T vy, const RadialStep<T>& rs) {
    Partials<T> o;
    T dx = x - cx, dy = y - cy;
    T da = dx * ia, db = dy * ib;
    o.h = ((da * da + db * db) - T(1)) - kv * rs.vo;
    T iaa = ia * ia, ibb = ib * ib;
    o.hx = (T(2) * dx) * iaa;
    o.hy = (T(2) * dy) * ibb;
    o.hth = T(0);
    o.hv = (-kv) * rs.iopv2;
    o.ht = T(-2) * ((dx * iaa) * vx + (dy * ibb) * vy);
    return o;
}

// D_CBF -- test_scripts/stanley_controller_ellipse.py:251-255
template <typename T>
__device__ __forceinline__ Partials<T> distance_partials(T x, T y, T cx, T cy, T Ds) {
    Partials<T> o;
    T dx = x - cx, dy = y - cy;
    o.h = Real<T>::sqrt_(dx * dx + dy * dy) - Ds;
    o.hx = (T(2) * dx) / (o.h + Ds);
    o.hy = (T(2) * dy) / (o.h + Ds);
    o.hth = o.hv = o.ht = T(0);
    return o;
}

// CollisionCone2D.update/evaluate/dx/dy/dv/dtheta/dt -- cbf/obstacles.py:468-502,401-458
// (sv, cv) = sin/cos of the ego yaw, computed once by the caller.
template <typename T>
__device__ __forceinline__ Partials<T> cone_partials(T x, T y, T th, T v, T sth, T cth, T cx, T cy, T tho, T vo, T a, T beta) {
    typedef Real<T> R;
    const T ZT = R::zero_tol();
    Partials<T> o;
    T s_vx = v * cth, s_vy = v * sth;
    T so, co;
    R::sincos_(tho + beta, &so, &co);
    T o_vx = vo * co, o_vy = vo * so;
    T prx = x - cx, pry = y - cy;
    T vrx = s_vx - o_vx, vry = s_vy - o_vy;
    T dist = R::sqrt_(prx * prx + pry * pry);
    T vrn = R::sqrt_(vrx * vrx + vry * vry);
    T cb = ZT;
    if (R::abs_(dist) > R::abs_(a)) cb = R::sqrt_(dist * dist - a * a) + ZT;
    T cos_phi = T(0);
    if (dist > ZT) cos_phi = cb / dist;
    o.h = (prx * vrx + pry * vry) + ((dist * vrn) * cos_phi);
    o.hx = vrx + (vrn * prx) / (cb + ZT);
    o.hy = vry + (vrn * pry) / (cb + ZT);
    T sb, cbt;
    if (beta == T(0)) { sb = sth; cbt = cth; }          // th + 0 == th exactly
    else R::sincos_(th + beta, &sb, &cbt);
    o.hv = (prx * cbt + pry * sb) + ((vrx * cbt + vry * sb) * cb) / (vrn + ZT);
    o.hth = ((-prx) * s_vy + pry * s_vx) + (((-vrx) * s_vy + vry * s_vx) * cb) / (vrn + ZT);
    o.ht = ((-vrx) * o_vx - vry * o_vy) + (((-vrn) * (prx * o_vx + pry * o_vy)) / (cb + ZT));
    return o;
}

// Horner value / first / second derivative of sum c_i x^i (cbf/obstacles.py:589-592)
template <typename T>
__device__ __forceinline__ void poly3(const T (&c)[6], T x, T& g, T& dg, T& ddg) {
    g = T(0);
#pragma unroll
    for (int i = 5; i >= 0; --i) g = c[i] + g * x;
    dg = T(0);
#pragma unroll
    for (int i = 5; i >= 1; --i) dg = (T(i) * c[i]) + dg * x;
    ddg = T(0);
#pragma unroll
    for (int i = 5; i >= 2; --i) ddg = (T(i - 1) * (T(i) * c[i])) + ddg * x;
}

template <typename T> __device__ __forceinline__ T poly0(const T (&c)[6], T x) {
    T g = T(0);
#pragma unroll
    for (int i = 5; i >= 0; --i) g = c[i] + g * x;
    return g;
}

// PolyLane.get_shortest_distance_x (cbf/obstacles.py:641-679): safeguarded Newton from x0 = px,
// same iteration / stopping rule as oracle.lane_closest_x.
template <typename T> __device__ T lane_closest_x(const T (&c)[6], T px, T py) {
    typedef Real<T> R;
    T x = px;
    for (int it = 0; it < 50; ++it) {
        T g, dg, ddg;
        poly3(c, x, g, dg, ddg);
        T ex = x - px, ey = g - py;
        T grad = ex + ey * dg;
        T hess = (T(1) + dg * dg) + ey * ddg;
        T step = (hess > T(0)) ? (-grad) / hess : -grad;
        T D0 = ex * ex + ey * ey;
        // (a step may raise D by its rounding noise, 8 ulp: near the minimum the full Newton step changes D by less than that,
        // and rejecting it there turns the quadratic convergence into a bisection -- 18 iterations where 4 do, and a warp
        // waits for its slowest lane: 31 % of config 4's samples sat in this loop at 4.6 of 32 lanes active)
        const T Dacc = D0 + R::lane_dtol() * D0;
        T t = T(1), xn = x;
        bool ok = false;
        for (int ls = 0; ls < 30; ++ls) {
            xn = x + t * step;
            T gn = poly0(c, xn);
            T Dn = (xn - px) * (xn - px) + (gn - py) * (gn - py);
            if (Dn <= Dacc) { ok = true; break; }
            t = t * T(0.5);
        }
        if (!ok) break;
        T dxn = R::abs_(xn - x);
        T lim = R::lane_xtol() * (T(1) + R::abs_(x));
        x = xn;
        if (dxn <= lim) break;
    }
    return x;
}

// PolyLane.update/evaluate/dx/dy -- cbf/obstacles.py:620-636,607-612,681-689
template <typename T>
__device__ __forceinline__ Partials<T> lane_partials(T x, T y, const T (&c)[6], T buffer, bool sqrt_form = false) {
    typedef Real<T> R;
    Partials<T> o;
    T cx = lane_closest_x(c, x, y);
    T g, dg, ddg;
    poly3(c, cx, g, dg, ddg);
    T eta = ((T(1) + dg * ddg) + dg * dg) - y * ddg;
    if (R::abs_(eta) < R::zero_tol()) eta = R::zero_tol();
    T ex = cx - x, ey = g - y;
    o.h = (ex * ex + ey * ey) - buffer;
    T te = T(2) / eta;
    o.hx = te * ((x - cx) * (eta - T(1)) - (y - g) * dg);
    o.hy = te * ((-(x - cx)) * dg + (y - g) * (eta - dg * dg));
    o.hth = o.hv = o.ht = T(0);
    if (sqrt_form) {
        // CBF_lane_sqrt -- stanley_controller_ellipse.py:489-492: distance instead of squared distance
        o.h = R::sqrt_(ex * ex + ey * ey) - buffer;
        const T den = T(2) * (o.h + buffer);
        o.hx = o.hx / den;
        o.hy = o.hy / den;
    }
    return o;
}

// dispatch on slot type; fields are read from the SoA obstacle buffer obst[(m*8+f)*N + n]
template <typename T>
__device__ __forceinline__ Partials<T> slot_partials(int desc, const T* __restrict__ f, int64_t fs,
                                                     T x, T y, T th, T v, T sth, T cth,
                                                     const T* pre = nullptr, int64_t ps = 0,
                                                     const T* ego_beta = nullptr, const RadialStep<T>* rs = nullptr) {
    const int type = desc & SCCAV_SLOT_TYPE_MASK;
    const bool is_static = (desc & SCCAV_SLOT_STATIC) != 0;
    // f points at field 0 of this slot for this vehicle; fs = stride between fields (N)
    // pre (optional): this slot's loop-invariant values for this vehicle, stride ps
    switch (type) {
        case SCCAV_SLOT_ELLIPSE: {
            T vx = T(0), vy = T(0);
            if (!is_static) { vx = f[5 * fs]; vy = f[6 * fs]; }
            if (pre) {
                T cx = f[0], cy = f[fs], a = f[2 * fs], b = f[3 * fs];
                return ellipse_partials_pre<T>(x, y, cx, cy, a, b, vx, vy, pre, ps);
            }
            T cx = f[0], cy = f[fs], a = f[2 * fs], b = f[3 * fs], t = f[4 * fs];
            return ellipse_partials<T>(x, y, cx, cy, a, b, t, vx, vy);
        }
        case SCCAV_SLOT_CONE: {
            T cx = f[0], cy = f[fs], to = f[2 * fs], vo = f[3 * fs], a = f[4 * fs];
            // SADBM pushes the vehicle's beta into every cone after each solve (cbf.py:424-426)
            T be = ego_beta ? *ego_beta : f[5 * fs];
            return cone_partials<T>(x, y, th, v, sth, cth, cx, cy, to, vo, a, be);
        }
        case SCCAV_SLOT_LANE:
        case SCCAV_SLOT_LANE_SQRT: {
            T buf = f[0];
            T c[6] = {f[fs], f[2 * fs], f[3 * fs], f[4 * fs], f[5 * fs], f[6 * fs]};
            return lane_partials<T>(x, y, c, buf, type == SCCAV_SLOT_LANE_SQRT);
        }
        case SCCAV_SLOT_RADIAL: {
            if (rs && pre) return radial_partials_fast<T>(x, y, f[0], f[fs], pre[0], pre[ps], f[4 * fs], f[5 * fs], f[6 * fs], *rs);
            T cx = f[0], cy = f[fs], a = f[2 * fs], b = f[3 * fs], kv = f[4 * fs], vx = f[5 * fs], vy = f[6 * fs];
            return radial_partials<T>(x, y, v, cx, cy, a, b, kv, vx, vy);
        }
        case SCCAV_SLOT_ELLIPSE_PREP: {
            T wx = T(0), wy = T(0);
            if (!is_static) { wx = f[6 * fs]; wy = f[7 * fs]; }
            return ellipse_prep_partials<T>(x, y, f[0], f[fs], f[2 * fs], f[3 * fs], f[4 * fs], f[5 * fs], wx, wy, is_static);
        }
        default: {
            T cx = f[0], cy = f[fs], Ds = f[2 * fs];
            return distance_partials<T>(x, y, cx, cy, Ds);
        }
    }
}

// ------------------------------------------------------------------------------------------
// row assembly: A0*u0 + A1*u1 >= b
// ------------------------------------------------------------------------------------------
// DBM_CBF_2DS gc/fc + F -- cbf/cbf.py:159-164,200-207
template <typename T>
__device__ __forceinline__ void dbm_row(const Partials<T>& p, T sth, T cth, T v, T alpha, T vlr, T& A0, T& A1, T& b) {
    // vlr = v / lr, formed once per vehicle by the caller: a division inside the slot loop is not hoisted
    // by the compiler (it cannot fold h_theta * (v / lr) even where h_theta is the constant 0 -- inf, NaN)
    A0 = p.hv;
    A1 = (p.hx * ((-v) * sth) + p.hy * (v * cth)) + p.hth * vlr;
    T Lf = p.hx * (v * cth) + p.hy * (v * sth);
    b = -((Lf + alpha * p.h) + p.ht);
}

// KBM_VC_CBF2D F -- cbf/cbf.py:94-101
template <typename T>
__device__ __forceinline__ void kbm_row(const Partials<T>& p, T sth, T cth, T alpha, T& A0, T& A1, T& b) {
    A0 = p.hx * cth + p.hy * sth;
    A1 = p.hth;
    b = -(alpha * p.h);
}

// DUM_CBF_2DS gc/fc + F -- cbf/cbf.py:237-245,277-286: Lg h = [h_v, h_theta], Lf h = v cos h_x + v sin h_y
template <typename T>
__device__ __forceinline__ void dum_row(const Partials<T>& p, T sth, T cth, T v, T alpha, T& A0, T& A1, T& b) {
    A0 = p.hv;
    A1 = p.hth;
    T Lf = p.hx * (v * cth) + p.hy * (v * sth);
    b = -((Lf + alpha * p.h) + p.ht);
}

// SADBM_CBF_2DS gc/fc + F -- cbf/cbf.py:337-346,386-397.  State (x, y, theta, v, beta), controls (a, d(beta)/dt):
// Lg h = [h_v, h_beta], Lf h = h_x v cos(theta+beta) + h_y v sin(theta+beta) + h_theta v sin(beta) / lr.
// h_beta = h_theta for the cone (obstacles.py:460-466) and 0 for every other obstacle (obstacles.py:124-125),
// which have h_theta = 0 too -- so h_theta serves.  The caller passes sin / cos of theta+beta and v sin(beta) / lr.
template <typename T>
__device__ __forceinline__ void sadbm_row(const Partials<T>& p, T sthb, T cthb, T v, T alpha, T vsb_lr, T& A0, T& A1, T& b) {
    A0 = p.hv;
    A1 = p.hth;
    T Lf = (p.hx * (v * cthb) + p.hy * (v * sthb)) + p.hth * vsb_lr;
    b = -((Lf + alpha * p.h) + p.ht);
}

// row of the configured model
// (MODEL >= 0: the model is known at compile time -- no dispatch inside the slot loop)
template <typename T, int MODEL = -1>
__device__ __forceinline__ void model_row(const Params<T>& P, const Partials<T>& p, T sth, T cth, T v, T alpha, T vlr, T& A0, T& A1, T& b) {
    const int model = MODEL >= 0 ? MODEL : P.model;
    if (model == SCCAV_MODEL_KBM) kbm_row<T>(p, sth, cth, alpha, A0, A1, b);
    else if (model == SCCAV_MODEL_DUM) dum_row<T>(p, sth, cth, v, alpha, A0, A1, b);
    else if (model == SCCAV_MODEL_SADBM) sadbm_row<T>(p, sth, cth, v, alpha, vlr, A0, A1, b);
    else dbm_row<T>(p, sth, cth, v, alpha, vlr, A0, A1, b);
}

// ------------------------------------------------------------------------------------------
// 2-variable QP on rows held in shared memory: row k at rows[(3k + {0,1,2}) * stride]
// Same enumeration order / tolerances as oracle.qp2_exact: {} , singles, pairs; first KKT point.
// ------------------------------------------------------------------------------------------
template <typename T> struct RowView {
    const T* rows;
    int stride;
    int pitch = 3;       // elements between consecutive rows of one problem (3, or NF where rows overlay staged slots)
    __device__ __forceinline__ T A0(int k) const { return rows[(pitch * k + 0) * stride]; }
    __device__ __forceinline__ T A1(int k) const { return rows[(pitch * k + 1) * stride]; }
    __device__ __forceinline__ T b(int k) const { return rows[(pitch * k + 2) * stride]; }
};

template <typename T>
__device__ __forceinline__ bool qp_check(const RowView<T>& rv, int m, T u0, T u1, int skip_a, int skip_b, T& worst) {
    typedef Real<T> R;
    bool feas = true;
    worst = -R::inf();
    for (int k = 0; k < m; ++k) {
        T a0 = rv.A0(k), a1 = rv.A1(k), bk = rv.b(k);
        T t0 = a0 * u0, t1 = a1 * u1;
        T rk = (t0 + t1) - bk;
        if (-rk > worst) worst = -rk;
        T tol = R::feas_eps() * ((R::abs_(t0) + R::abs_(t1)) + R::abs_(bk));
        if (!(rk >= -tol) && k != skip_a && k != skip_b) feas = false;
    }
    return feas;
}

// Which pairs (j, k) can have a non-zero determinant?  det2 = A0_j A1_k - A1_j A0_k is exactly 0 (or NaN)
// -- and oracle.qp2_exact skips the pair at its parallel-rows test -- whenever A0_j = A0_k = 0 or
// A1_j = A1_k = 0.  nz0 / nz1 = bit masks of the rows whose A0 / A1 is not exactly zero (NaN counts as
// non-zero).  Ellipse, lane, radial and distance rows under DBM have A0 = h_v = 0 (cbf/obstacles.py:232-236):
// for an all-ellipse problem the whole pair phase vanishes.
struct RowNz {
    uint32_t nz0, nz1;
    uint32_t viol = 0xffffffffu;   // rows the reference point violates strictly (rk < 0); all ones = not recorded
    __device__ __forceinline__ bool any_pair() const { return nz0 != 0u && nz1 != 0u; }
    __device__ __forceinline__ bool pair(int j, int k) const {
        return (((nz0 >> j) | (nz0 >> k)) & ((nz1 >> j) | (nz1 >> k)) & 1u) != 0u;
    }
};

// inverse of the cost weight, with the operations of oracle.qp2_exact (cbf/cbf.py:182-186 builds P = 2R)
template <typename T> struct RInv {
    T i00, i01, i10, i11;
    __host__ __device__ __forceinline__ RInv() {}
    __host__ __device__ __forceinline__ RInv(T R00, T R01, T R10, T R11) {
        const T det = R00 * R11 - R01 * R10;
        i00 = R11 / det; i01 = (-R01) / det; i10 = (-R10) / det; i11 = R00 / det;
    }
};

// Enumeration with the least-violation bookkeeping of oracle.qp2_exact (every candidate checked against
// every row), minus the pairs that RowNz proves degenerate.
template <typename T, bool VIOL = true>
__device__ __forceinline__ int qp2_solve_active_full(const RowView<T>& rv, int m, RowNz nz, T r0, T r1, T R00, T R01, T R10, T R11,
                                                  const RInv<T>& Ri, T worst0, T& u0o, T& u1o, uint32_t& masko) {
    typedef Real<T> R;
    T worst;
    T fbw = worst0, fb0 = r0, fb1 = r1;
    uint32_t fbm = 0u;
    // singles, in index order.  The rows r violates (rk < 0, the same operations as here) were recorded while the rows
    // were written: each lane walks ITS OWN violated rows, so a warp runs as many iterations as its busiest lane has
    // candidates (one or two) instead of one per distinct row index among its lanes.
    // (VIOL = false walks every row index: where most lanes of a warp are active at once and violate many rows -- the
    // seeker crowds of config 3 -- the plain loop, whose row loads run ahead of the candidates, is the faster one:
    // 119 vs 136 ms)
    uint32_t todo = (VIOL ? nz.viol : 0xffffffffu) & (m >= 32 ? 0xffffffffu : ((1u << m) - 1u));
    while (todo) {
        const int k = __ffs((int)todo) - 1;
        todo &= todo - 1u;
        T a0 = rv.A0(k), a1 = rv.A1(k), bk = rv.b(k);
        T rk = (a0 * r0 + a1 * r1) - bk;
        if (!(rk < T(0))) continue;
        T g0 = Ri.i00 * a0 + Ri.i01 * a1;
        T g1 = Ri.i10 * a0 + Ri.i11 * a1;
        T den = a0 * g0 + a1 * g1;
        if (!(den > T(0))) continue;
        T t = (-rk) / den;
        T u0 = r0 + g0 * t, u1 = r1 + g1 * t;
        if (qp_check(rv, m, u0, u1, k, -1, worst)) {
            u0o = u0; u1o = u1; masko = 1u << k;
            return SCCAV_STATUS_ACTIVE;
        }
        if (worst < fbw - R::tie_eps() * (R::abs_(worst) + R::abs_(fbw))) { fbw = worst; fb0 = u0; fb1 = u1; fbm = 1u << k; }
    }
    if (nz.any_pair()) {
        for (int j = 0; j < m; ++j) {
            T aj0 = rv.A0(j), aj1 = rv.A1(j), bj = rv.b(j);
            for (int k = j + 1; k < m; ++k) {
                if (!nz.pair(j, k)) continue;
                T ak0 = rv.A0(k), ak1 = rv.A1(k), bk = rv.b(k);
                T t1 = aj0 * ak1, t2 = aj1 * ak0;
                T det2 = t1 - t2;
                if (!(R::abs_(det2) > R::par_eps() * (R::abs_(t1) + R::abs_(t2)))) continue;
                T u0 = (bj * ak1 - aj1 * bk) / det2;
                T u1 = (aj0 * bk - bj * ak0) / det2;
                T e0 = u0 - r0, e1 = u1 - r1;
                T w0 = T(2) * (R00 * e0 + R01 * e1);
                T w1 = T(2) * (R10 * e0 + R11 * e1);
                T lj = (w0 * ak1 - ak0 * w1) / det2;
                T lk = (aj0 * w1 - w0 * aj1) / det2;
                bool feas = qp_check(rv, m, u0, u1, j, k, worst);
                if (feas && lj >= T(0) && lk >= T(0)) {
                    u0o = u0; u1o = u1; masko = (1u << j) | (1u << k);
                    return SCCAV_STATUS_ACTIVE;
                }
                if (worst < fbw - R::tie_eps() * (R::abs_(worst) + R::abs_(fbw))) { fbw = worst; fb0 = u0; fb1 = u1; fbm = (1u << j) | (1u << k); }
            }
        }
    }
    u0o = fb0; u1o = fb1; masko = fbm;
    return SCCAV_STATUS_INFEASIBLE;
}

// Shortcut before the enumeration.  If the optimum has ONE active row k, it is the projection of r onto
// that half-plane in the metric of the cost, at distance d_k = -rk / sqrt(A_k R^-1 A_k^T); it satisfies
// every other row j, so d_k >= d_j for every row j that r violates: k is the row with the largest
// rk^2 / den.  One scan finds it (ratios compared by cross-multiplication, no division), its candidate
// is formed with the operations of the enumeration and checked against every row.  The enumeration --
// which returns the FIRST accepted single in index order -- is still run whenever its answer could
// differ: the candidate fails (optimum on a pair, or infeasible rows); another violated row is within
// 1e-6 of the largest ratio (qp_tie_margin: a tie, the earlier row may be accepted within the feasibility
// tolerance); or the winning residual is small against its own rounding scale (qp_res_margin).
// Otherwise a row j != k has d_j < d_k (1 - 5e-7): by Cauchy-Schwarz in the metric of the cost its candidate
// misses row k by at least 5e-7 |rk|, and the guard  |rk| qp_res_margin > feas_eps scale_k  (qp_res_margin =
// 5e-7 / 4) puts that miss a factor 4 above the acceptance tolerance feas_eps scale_k of row k -- so the
// enumeration rejects j and accepts k: same point, same bits.
// tests/test_gpu_parity.py::test_qp_shortcut_equals_enumeration compares the two bit for bit.
// The scan runs inside the row loop (QpScan::row, all lanes converged, predicated) or, in the rollout, over the
// stored rows of the lanes whose reference point is infeasible; finish() forms and checks the winner's candidate
// and returns true (and the solution) if the shortcut answers the problem.
template <typename T> struct QpScan {
    T bn, bd;      // largest ratio  rk^2 / den  as a fraction
    T sn, sd;      // runner-up
    int kb;
    __device__ __forceinline__ void reset() { bn = T(0); bd = T(1); sn = T(0); sd = T(1); kb = -1; }
    // rk = (a0 r0 + a1 r1) - bk of row k, as computed by the feasibility test of the reference point
    __device__ __forceinline__ void row(int k, T a0, T a1, T rk, const RInv<T>& Ri) {
        T g0 = Ri.i00 * a0 + Ri.i01 * a1;
        T g1 = Ri.i10 * a0 + Ri.i11 * a1;
        T den = a0 * g0 + a1 * g1;
        if (rk < T(0) && den > T(0)) {
            T sc = rk * rk;
            if (sc * bd > bn * den) { sn = bn; sd = bd; bn = sc; bd = den; kb = k; }
            else if (sc * sd >= sn * den) { sn = sc; sd = den; }
        }
    }
    __device__ __forceinline__ bool finish(const RowView<T>& rv, int m, T r0, T r1, const RInv<T>& Ri,
                                           T& u0o, T& u1o, uint32_t& masko) const {
        typedef Real<T> R;
        if (!(kb >= 0 && sn * bd < (T(1) - R::qp_tie_margin()) * (bn * sd))) return false;
        T a0 = rv.A0(kb), a1 = rv.A1(kb), bk = rv.b(kb);
        T t0 = a0 * r0, t1 = a1 * r1;
        T rk = (t0 + t1) - bk;
        T g0 = Ri.i00 * a0 + Ri.i01 * a1;
        T g1 = Ri.i10 * a0 + Ri.i11 * a1;
        T den = a0 * g0 + a1 * g1;
        T t = (-rk) / den;
        T u0 = r0 + g0 * t, u1 = r1 + g1 * t;
        if (!(-rk * R::qp_res_margin() > R::feas_eps() * ((R::abs_(t0) + R::abs_(t1)) + R::abs_(bk)))) return false;
        T worst;
        if (!qp_check(rv, m, u0, u1, kb, -1, worst)) return false;
        u0o = u0; u1o = u1; masko = 1u << kb;
        return true;
    }
};

// The reference point r violates at least one row (worst0 = its largest violation): one thread, one problem.
// (The persistent rollout uses this form: plain enumeration, no shortcut -- its warps are not converged.)
template <typename T, bool VIOL = true>
__device__ __forceinline__ int qp2_solve_active(const RowView<T>& rv, int m, RowNz nz, T r0, T r1,
                                                T R00, T R01, T R10, T R11, const RInv<T>& Ri, T worst0,
                                                T& u0o, T& u1o, uint32_t& masko) {
    return qp2_solve_active_full<T, VIOL>(rv, m, nz, r0, r1, R00, R01, R10, R11, Ri, worst0, u0o, u1o, masko);
}

// ------------------------------------------------------------------------------------------
// Warp-cooperative enumeration of ONE problem: lane k owns row k (m <= 32; lanes >= m hold the vacuous
// row 0 u >= -inf).  Every candidate is checked by all lanes at once (one ballot); order, acceptance rule
// and least-violation bookkeeping are those of qp2_solve_active_full, so the results are identical.
// All 32 lanes must call it with the same m, nz, r, R, worst0.
// ------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T shfl(T v, int src);
template <> __device__ __forceinline__ double shfl<double>(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
template <> __device__ __forceinline__ float shfl<float>(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }

template <typename T> __device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T w = __shfl_xor_sync(0xffffffffu, v, o);
        v = (w > v) ? w : v;
    }
    return v;
}

template <typename T>
__device__ __forceinline__ bool qp_check_coop(int lane, bool own, T a0, T a1, T bk, T u0, T u1, uint32_t skip, T& worst) {
    typedef Real<T> R;
    T t0 = a0 * u0, t1 = a1 * u1;
    T rk = (t0 + t1) - bk;
    T tol = R::feas_eps() * ((R::abs_(t0) + R::abs_(t1)) + R::abs_(bk));
    bool bad = own && !((skip >> lane) & 1u) && !(rk >= -tol);
    worst = warp_max<T>(own ? -rk : -R::inf());
    return __ballot_sync(0xffffffffu, bad) == 0u;
}

template <typename T>
__device__ __forceinline__ int qp2_coop_active(int lane, int m, T a0, T a1, T bk, RowNz nz, T r0, T r1,
                                               T R00, T R01, T R10, T R11, const RInv<T>& Ri, T worst0,
                                               T& u0o, T& u1o, uint32_t& masko) {
    typedef Real<T> R;
    const bool own = lane < m;
    int status = -1;
    T worst;
    T fbw = worst0, fb0 = r0, fb1 = r1;
    uint32_t fbm = 0u;
    // singles: each lane prepares its own candidate, then they are tried in index order
    T rk = (a0 * r0 + a1 * r1) - bk;
    T g0 = Ri.i00 * a0 + Ri.i01 * a1, g1 = Ri.i10 * a0 + Ri.i11 * a1;
    T den = a0 * g0 + a1 * g1;
    const bool cand = own && (rk < T(0)) && (den > T(0));
    T t = cand ? (-rk) / den : T(0);
    T c0 = r0 + g0 * t, c1 = r1 + g1 * t;
    uint32_t todo = __ballot_sync(0xffffffffu, cand);
    while (todo && status < 0) {
        const int k = __ffs(todo) - 1;
        todo &= todo - 1;
        T s0 = shfl<T>(c0, k), s1 = shfl<T>(c1, k);
        if (qp_check_coop<T>(lane, own, a0, a1, bk, s0, s1, 1u << k, worst)) { u0o = s0; u1o = s1; masko = 1u << k; status = SCCAV_STATUS_ACTIVE; }
        else if (worst < fbw - R::tie_eps() * (R::abs_(worst) + R::abs_(fbw))) { fbw = worst; fb0 = s0; fb1 = s1; fbm = 1u << k; }
    }
    // pairs, lexicographic (j < k): rows are broadcast from their owners
    if (status < 0 && nz.any_pair()) {
        for (int j = 0; j < m && status < 0; ++j) {
            T aj0 = shfl<T>(a0, j), aj1 = shfl<T>(a1, j), bj = shfl<T>(bk, j);
            for (int k = j + 1; k < m && status < 0; ++k) {
                if (!nz.pair(j, k)) continue;
                T ak0 = shfl<T>(a0, k), ak1 = shfl<T>(a1, k), bkk = shfl<T>(bk, k);
                T t1 = aj0 * ak1, t2 = aj1 * ak0;
                T det2 = t1 - t2;
                if (!(R::abs_(det2) > R::par_eps() * (R::abs_(t1) + R::abs_(t2)))) continue;
                T p0 = (bj * ak1 - aj1 * bkk) / det2;
                T p1 = (aj0 * bkk - bj * ak0) / det2;
                T e0 = p0 - r0, e1 = p1 - r1;
                T w0 = T(2) * (R00 * e0 + R01 * e1);
                T w1 = T(2) * (R10 * e0 + R11 * e1);
                T lj = (w0 * ak1 - ak0 * w1) / det2;
                T lk = (aj0 * w1 - w0 * aj1) / det2;
                const uint32_t pm = (1u << j) | (1u << k);
                const bool feas = qp_check_coop<T>(lane, own, a0, a1, bk, p0, p1, pm, worst);
                if (feas && lj >= T(0) && lk >= T(0)) { u0o = p0; u1o = p1; masko = pm; status = SCCAV_STATUS_ACTIVE; }
                else if (worst < fbw - R::tie_eps() * (R::abs_(worst) + R::abs_(fbw))) { fbw = worst; fb0 = p0; fb1 = p1; fbm = pm; }
            }
        }
    }
    if (status < 0) { u0o = fb0; u1o = fb1; masko = fbm; status = SCCAV_STATUS_INFEASIBLE; }
    return status;
}

// Active solves of a CONVERGED warp whose lanes hold one problem each (rows of lane L in the column
// warp_rows + L of shared memory).  `need` = this lane's reference point violates a row.  Lanes first try
// the shortcut on their own; what is left (pairs, infeasible rows, ties) is enumerated by the whole warp,
// one problem after the other -- 32 lanes on one problem instead of one lane on it and 31 waiting.
template <typename T>
__device__ __forceinline__ int qp2_solve_active_warp(bool need, const T* warp_rows, int stride, int lane, int m, RowNz nz,
                                                     T r0, T r1, T R00, T R01, T R10, T R11, const RInv<T>& Ri, bool uniform_R,
                                                     T worst0, const QpScan<T>& scan, T& u0o, T& u1o, uint32_t& masko, bool enumerate,
                                                     int pitch = 3) {
    int status = SCCAV_STATUS_INACTIVE;
    if (need && !enumerate) {
        const RowView<T> rv{warp_rows + lane, stride, pitch};
        if (scan.finish(rv, m, r0, r1, Ri, u0o, u1o, masko)) { status = SCCAV_STATUS_ACTIVE; need = false; }
    }
    uint32_t todo = __ballot_sync(0xffffffffu, need);
    if (todo) __syncwarp();                    // rows written by their lanes are read by the whole warp below
    while (todo) {
        const int L = __ffs(todo) - 1;
        todo &= todo - 1;
        const int pm = __shfl_sync(0xffffffffu, m, L);
        RowNz pnz;
        pnz.nz0 = __shfl_sync(0xffffffffu, nz.nz0, L);
        pnz.nz1 = __shfl_sync(0xffffffffu, nz.nz1, L);
        const T p0 = shfl<T>(r0, L), p1 = shfl<T>(r1, L), pw = shfl<T>(worst0, L);
        T q00 = R00, q01 = R01, q10 = R10, q11 = R11;
        RInv<T> qi = Ri;
        if (!uniform_R) {
            q00 = shfl<T>(R00, L); q01 = shfl<T>(R01, L); q10 = shfl<T>(R10, L); q11 = shfl<T>(R11, L);
            qi.i00 = shfl<T>(Ri.i00, L); qi.i01 = shfl<T>(Ri.i01, L); qi.i10 = shfl<T>(Ri.i10, L); qi.i11 = shfl<T>(Ri.i11, L);
        }
        const T* col = warp_rows + L;
        const bool own = lane < pm;
        const T a0 = own ? col[(pitch * lane + 0) * stride] : T(0);
        const T a1 = own ? col[(pitch * lane + 1) * stride] : T(0);
        const T bk = own ? col[(pitch * lane + 2) * stride] : -Real<T>::inf();
        T s0, s1;
        uint32_t sm;
        const int st = qp2_coop_active<T>(lane, pm, a0, a1, bk, pnz, p0, p1, q00, q01, q10, q11, qi, pw, s0, s1, sm);
        if (lane == L) { u0o = s0; u1o = s1; masko = sm; status = st; }
    }
    return status;
}

// ------------------------------------------------------------------------------------------
// one solve_cbf for one vehicle: rows of all slots -> smem -> QP -> converted output
// (cbf/cbf.py:166-220 for DBM, :67-110 for KBM).  u_ref = (a|v, delta); returns (a|v, delta).
//
// SPEC selects a compile-time specialisation of the slot loop (identical arithmetic):
//   SPEC_GENERIC  any slot mix, dispatch on the slot descriptor
//   SPEC_ELLIPSE  every slot is a per-vehicle (not SHARED) ELLIPSE -- the BASELINE configs 2 and 5;
//                 no dispatch, and the fields of slot m+1 are in flight while slot m is evaluated
//   SPEC_ELLIPSE_PREP  every slot is a per-vehicle ELLIPSE_PREP with the same STATIC flag: 6 (static)
//                 or 8 loads and ~25 fma per row, nothing else -- the HBM-bound form of the operator
// ------------------------------------------------------------------------------------------
#define SCCAV_SPEC_GENERIC 0
#define SCCAV_SPEC_ELLIPSE 1
#define SCCAV_SPEC_ELLIPSE_PREP 2

// row of one slot into shared memory + running feasibility test of the reference point
// (qp_check(r) of qp2_solve, evaluated on the fly with the same operations)
// ELL: the slot is an ellipse and the model is DBM at compile time (the rollout's compile-time instances): h_theta = h_v = 0,
// so the terms 0 * (v / lr) of A1 and A0 * r0 = 0 of the test are not evaluated (sums with an exact zero: the same values,
// up to the sign of a zero that nothing reads), and v / lr itself is never formed.
template <typename T, bool SCAN = false, int PITCH = 3, int MODEL = -1, bool ELL = false>
__device__ __forceinline__ void put_row(const Params<T>& P, const Partials<T>& p, T sth, T cth, T v, T alpha, T vlr, T r0, T r1,
                                        T* rows, int stride, int m, T& hmin, T& worst, bool& feas, RowNz& nz,
                                        QpScan<T>* scan = nullptr, const RInv<T>* Ri = nullptr) {
    typedef Real<T> R;
    T A0, A1, b;
    if (ELL) {
        A0 = T(0);
        A1 = p.hx * ((-v) * sth) + p.hy * (v * cth);                                     // dbm_row with h_theta = 0
        const T Lf = p.hx * (v * cth) + p.hy * (v * sth);
        b = -((Lf + alpha * p.h) + p.ht);
    } else
    model_row<T, MODEL>(P, p, sth, cth, v, alpha, vlr, A0, A1, b);
    rows[(PITCH * m + 0) * stride] = A0;
    rows[(PITCH * m + 1) * stride] = A1;
    rows[(PITCH * m + 2) * stride] = b;
    if (p.h < hmin) hmin = p.h;
    if (A0 != T(0)) nz.nz0 |= 1u << m;
    if (A1 != T(0)) nz.nz1 |= 1u << m;
    T t0 = ELL ? T(0) : A0 * r0, t1 = A1 * r1;
    T rk = ELL ? t1 - b : (t0 + t1) - b;
    if (-rk > worst) worst = -rk;
    T tol = R::feas_eps() * ((ELL ? R::abs_(t1) : R::abs_(t0) + R::abs_(t1)) + R::abs_(b));
    if (!(rk >= -tol)) feas = false;
    if (rk < T(0)) nz.viol |= 1u << m;
    if (SCAN) scan->row(m, A0, A1, rk, *Ri);
}

// what the row phase of one vehicle leaves behind for the QP
template <typename T> struct RowPhase {
    T r0, r1;        // reference point in QP coordinates (a | v, beta | omega)
    T worst;         // its largest row violation
    RowNz nz;
    bool feas;       // the reference point satisfies every row: u = u_ref, no solve
    QpScan<T> scan;  // most violated row in the metric of R (only filled when filter_rows<.., SCAN = true>)
};

// phase 1: u_ref -> QP coordinates, rows of all slots -> shared memory, feasibility of the reference point
// BIO: honour SCCAV_FLAG_BETA_IO (the filter-step kernels; the closed-loop kernels reject the flag on the host and do
// not even test it -- one more live value in their time loop was measured at 2.7 % of the rollout)
template <typename T, int SPEC, bool SCAN = false, int MODEL = -1, bool BIO = false>
__device__ __forceinline__ RowPhase<T> filter_rows(const Params<T>& P, const SlotDesc& sd, int M, int64_t N, int64_t n,
                                                   const T* __restrict__ obst, T x, T y, T th, T v, T sth, T cth,
                                                   T alpha, T uref0, T uref1, T* rows, int stride, T& hmin,
                                                   const T* __restrict__ pre = nullptr, uint32_t moving = 0xffffffffu,
                                                   const RInv<T>* Ri = nullptr, const T* aug = nullptr, const T* r1_pre = nullptr,
                                                   const void* sym = nullptr) {
    typedef Real<T> R;
    hmin = R::inf();
    T vlr = v / P.lr;                                                                   // cbf.py:160 (g_c[2][1])
    T r0 = uref0, r1;
    T ego_beta = T(0);
    const int model = MODEL >= 0 ? MODEL : P.model;
    const bool sadbm = model == SCCAV_MODEL_SADBM;
    if (sadbm) {
        // aug = (beta, beta_ref_last) of this vehicle.  cbf.py:359-372: u_ref[1] -> d(beta_ref)/dt; the rows use
        // sin / cos(theta + beta) and v sin(beta) / lr where the other models use sin / cos(theta) and v / lr
        ego_beta = aug[0];
        const T beta_ref = R::atan2_(P.lr * R::tan_(uref1), P.lf + P.lr);
        r1 = (beta_ref - aug[1]) / P.sadbm_dt;
        T sb, cb;
        R::sincos_(ego_beta, &sb, &cb);
        vlr = (v * sb) / P.lr;
    } else
    if (model == SCCAV_MODEL_KBM) r1 = (uref0 * R::tan_(uref1)) / P.L;                  // cbf.py:75
    else if (model == SCCAV_MODEL_DUM) r1 = uref1;                                       // cbf.py:253: u_ref as given
    else if (BIO && (P.flags & SCCAV_FLAG_BETA_IO)) r1 = uref1;                          // the caller holds beta
    else if (r1_pre) r1 = *r1_pre;                                                       // (the rollout's fused-steer instances)
    else r1 = R::atan2_(P.lr * R::tan_(uref1), P.lf + P.lr);                             // cbf.py:175
    // rows -> shared memory; an inactive step never re-reads them
    bool feas = true;
    T worst = -R::inf();
    RowNz nz{0u, 0u, 0u};                                      // (violated rows are recorded by put_row)
    QpScan<T> scan;
    scan.reset();
    if (SPEC == SCCAV_SPEC_ELLIPSE) {
        const int64_t ss = (int64_t)SCCAV_NFIELD * N;          // slot stride
        const T* f = obst + n;
        if (pre) {
            // rollout: loop-invariant terms hoisted into `pre`; a static obstacle (moving bit clear)
            // needs neither its velocity nor a^2, b^2 (h_t = -2 (. * 0 + . * 0) = 0)
            const int64_t ps = (int64_t)SCCAV_NPRE * N;
            const T* q = pre + n;
            // (two loops: the per-row test of the moving bit costs predicated instructions in every row)
            if (sym && moving == 0u) {
                // the rollout's paired copy of (cx, cy), (a, b) and the six hoisted terms: five 16-byte loads per row instead
                // of ten 8-byte ones, the same values into the same operations
                typedef typename R::T2 T2;
                const T2* g = reinterpret_cast<const T2*>(sym) + n;
                for (int m = 0; m < M; ++m, g += 5 * N) {
                    const T2 c = g[0], ab = g[N], t = g[2 * N], k0 = g[3 * N], k1 = g[4 * N];
                    const T loc[6] = {t.x, t.y, k0.x, k0.y, k1.x, k1.y};
                    Partials<T> p = ellipse_partials_pre<T>(x, y, c.x, c.y, ab.x, ab.y, T(0), T(0), loc, 1);
                    put_row<T, SCAN, 3, MODEL, MODEL == SCCAV_MODEL_DBM>(P, p, sth, cth, v, alpha, vlr, r0, r1, rows, stride, m, hmin, worst, feas, nz, &scan, Ri);
                }
            } else
            if (moving == 0u) {
                for (int m = 0; m < M; ++m, f += ss, q += ps) {
                    T cx = f[0], cy = f[N], a = f[2 * N], b = f[3 * N];
                    Partials<T> p = ellipse_partials_pre<T>(x, y, cx, cy, a, b, T(0), T(0), q, N);
                    put_row<T, SCAN, 3, MODEL, MODEL == SCCAV_MODEL_DBM>(P, p, sth, cth, v, alpha, vlr, r0, r1, rows, stride, m, hmin, worst, feas, nz, &scan, Ri);
                }
            } else {
                for (int m = 0; m < M; ++m, f += ss, q += ps) {
                    T cx = f[0], cy = f[N], a = f[2 * N], b = f[3 * N];
                    T vx = T(0), vy = T(0);
                    if ((moving >> m) & 1u) { vx = f[5 * N]; vy = f[6 * N]; }
                    Partials<T> p = ellipse_partials_pre<T>(x, y, cx, cy, a, b, vx, vy, q, N);
                    put_row<T, SCAN, 3, MODEL, MODEL == SCCAV_MODEL_DBM>(P, p, sth, cth, v, alpha, vlr, r0, r1, rows, stride, m, hmin, worst, feas, nz, &scan, Ri);
                }
            }
        } else {
            const bool st_ = (sd.d[0] & SCCAV_SLOT_STATIC) != 0;      // static: the velocity fields are not read
            T cx = f[0], cy = f[N], a = f[2 * N], b = f[3 * N], t = f[4 * N], vx = st_ ? T(0) : f[5 * N], vy = st_ ? T(0) : f[6 * N];
            for (int m = 0; m < M; ++m) {
                f += ss;
                T ncx = cx, ncy = cy, na = a, nb = b, nt = t, nvx = vx, nvy = vy;
                if (m + 1 < M) {
                    ncx = f[0]; ncy = f[N]; na = f[2 * N]; nb = f[3 * N]; nt = f[4 * N];
                    if (!st_) { nvx = f[5 * N]; nvy = f[6 * N]; }
                }
                Partials<T> p = ellipse_partials<T>(x, y, cx, cy, a, b, t, vx, vy);
                put_row<T, SCAN, 3, MODEL>(P, p, sth, cth, v, alpha, vlr, r0, r1, rows, stride, m, hmin, worst, feas, nz, &scan, Ri);
                cx = ncx; cy = ncy; a = na; b = nb; t = nt; vx = nvx; vy = nvy;
            }
        }
    } else if (SPEC == SCCAV_SPEC_ELLIPSE_PREP) {
        const int64_t ss = (int64_t)SCCAV_NFIELD * N;
        const T* f = obst + n;
        const bool is_static = (sd.d[0] & SCCAV_SLOT_STATIC) != 0;
#ifndef SCCAV_PREP_UNROLL
#define SCCAV_PREP_UNROLL 2
#endif
        constexpr int kPrepUnroll = SCCAV_PREP_UNROLL;
        // (two loops: a run-time is_static inside one loop costs ten predicated instructions per row)
        if (sym) {
            // the rollout's packed symmetric form (kernels.cuh, pack_sym_kernel): three 16-byte loads per row
            typedef typename R::T2 T2;
            const T2* g = reinterpret_cast<const T2*>(sym) + n;
#pragma unroll kPrepUnroll
            for (int m = 0; m < M; ++m, g += 3 * N) {
                const T2 c = g[0], s0 = g[N], s1 = g[2 * N];
                Partials<T> p;
                const T dx = x - c.x, dy = y - c.y;
                const T gx = fma(s0.x, dx, s0.y * dy), gy = fma(s0.y, dx, s1.x * dy);
                p.h = fma(gx, dx, fma(gy, dy, T(-1)));
                p.hx = T(2) * gx; p.hy = T(2) * gy;
                p.hth = T(0); p.hv = T(0); p.ht = T(0);
                put_row<T, SCAN, 3, MODEL, MODEL == SCCAV_MODEL_DBM>(P, p, sth, cth, v, alpha, vlr, r0, r1, rows, stride, m, hmin, worst, feas, nz, &scan, Ri);
            }
        } else
        if (is_static) {
#pragma unroll kPrepUnroll
            for (int m = 0; m < M; ++m, f += ss) {
                Partials<T> p = ellipse_prep_partials<T>(x, y, f[0], f[N], f[2 * N], f[3 * N], f[4 * N], f[5 * N], T(0), T(0), true);
                put_row<T, SCAN, 3, MODEL, MODEL == SCCAV_MODEL_DBM>(P, p, sth, cth, v, alpha, vlr, r0, r1, rows, stride, m, hmin, worst, feas, nz, &scan, Ri);
            }
        } else {
#pragma unroll kPrepUnroll
            for (int m = 0; m < M; ++m, f += ss) {
                Partials<T> p = ellipse_prep_partials<T>(x, y, f[0], f[N], f[2 * N], f[3 * N], f[4 * N], f[5 * N], f[6 * N], f[7 * N], false);
                put_row<T, SCAN, 3, MODEL, MODEL == SCCAV_MODEL_DBM>(P, p, sth, cth, v, alpha, vlr, r0, r1, rows, stride, m, hmin, worst, feas, nz, &scan, Ri);
            }
        }
    } else {
        T rsth = sth, rcth = cth;                              // trig of the row assembly (theta + beta under SADBM)
        if (sadbm) R::sincos_(th + ego_beta, &rsth, &rcth);
        // prepared RADIAL rows (the rollout with SCCAV_FLAG_PREPARED_ROWS: `pre` then holds 1 / a, 1 / b of those slots)
        RadialStep<T> rstep;
        const bool rfast = pre != nullptr && (P.flags & SCCAV_FLAG_PREPARED_ROWS) != 0;
        if (rfast) { const T opv = T(1) + v; rstep.vo = v / opv; rstep.iopv2 = T(1) / (opv * opv); }
        for (int m = 0; m < M; ++m) {
            const int desc = sd.d[m];
            const int64_t nn = (desc & SCCAV_SLOT_SHARED) ? 0 : n;
            const T* f = obst + (int64_t)m * SCCAV_NFIELD * N + nn;
            const T* pr = pre ? pre + (int64_t)m * SCCAV_NPRE * N + n : nullptr;
            Partials<T> p = slot_partials<T>(desc, f, N, x, y, th, v, sth, cth, pr, N, sadbm ? &ego_beta : nullptr, rfast ? &rstep : nullptr);
            put_row<T, SCAN, 3, MODEL>(P, p, rsth, rcth, v, alpha, vlr, r0, r1, rows, stride, m, hmin, worst, feas, nz, &scan, Ri);
        }
    }
    RowPhase<T> ph;
    ph.r0 = r0; ph.r1 = r1; ph.worst = worst; ph.nz = nz; ph.feas = feas; ph.scan = scan;
    return ph;
}

// phase 3: QP coordinates -> (a | v, delta)
template <typename T, int MODEL = -1, bool BIO = false>
__device__ __forceinline__ T filter_convert(const Params<T>& P, T q0, T q1, T r0) {
    typedef Real<T> R;
    const int model = MODEL >= 0 ? MODEL : P.model;
    if (model == SCCAV_MODEL_KBM) {
        if (P.kbm_driver_delta) return R::atan_((q1 * P.L) / q0);                        // sce.py:652
        return R::atan2_(q1 * P.L, r0);                                                  // cbf.py:109
    }
    if (model == SCCAV_MODEL_DUM) return q1;                                             // cbf.py:293: u as solved
    if (BIO && (P.flags & SCCAV_FLAG_BETA_IO)) return q1;                                // beta out, as it came in
    return R::atan2_((P.lf + P.lr) * R::tan_(q1), P.lr);                                 // cbf.py:216
}

// all three phases by one thread (the persistent rollout kernel; K12 compacts phase 2 across the CTA)
template <typename T, int SPEC, int MODEL = -1, bool VIOL = true>
__device__ __forceinline__ int filter_vehicle(const Params<T>& P, const SlotDesc& sd, int M, int64_t N, int64_t n,
                                              const T* __restrict__ obst, T x, T y, T th, T v, T sth, T cth,
                                              T alpha, T R00, T R01, T R10, T R11, bool uniform_R, T uref0, T uref1,
                                              T* rows, int stride, T& u0, T& u1, T& u1raw, uint32_t& mask, T& hmin,
                                              const T* __restrict__ pre = nullptr, uint32_t moving = 0xffffffffu,
                                              bool convert = true, const T* r1_pre = nullptr, const void* sym = nullptr) {
    const RowPhase<T> ph = filter_rows<T, SPEC, false, MODEL>(P, sd, M, N, n, obst, x, y, th, v, sth, cth, alpha, uref0, uref1,
                                                rows, stride, hmin, pre, moving, nullptr, nullptr, r1_pre, sym);
    T q0 = ph.r0, q1 = ph.r1;
    int status = SCCAV_STATUS_INACTIVE;
    mask = 0u;
    if (!ph.feas) {
        RowView<T> rv{rows, stride};
        // inverse weight: the launch-wide one comes from the host (constant bank, no registers held across
        // the time loop); per-vehicle weights are inverted on the spot
        RInv<T> Ri;
        if (uniform_R) { Ri.i00 = P.Ri[0]; Ri.i01 = P.Ri[1]; Ri.i10 = P.Ri[2]; Ri.i11 = P.Ri[3]; }
        else Ri = RInv<T>(R00, R01, R10, R11);
        // (plain enumeration: the one-scan shortcut of the filter-step kernels was measured here -- the larger loop body
        // costs more in instruction fetch than the scan saves: 14.1 vs 12.9 ms on config 2)
        status = qp2_solve_active<T, VIOL>(rv, M, ph.nz, ph.r0, ph.r1, R00, R01, R10, R11, Ri, ph.worst, q0, q1, mask);
    }
    u0 = q0;
    u1raw = q1;
    if (convert) u1 = filter_convert<T, MODEL>(P, q0, q1, ph.r0);      // (the caller converts later, or never, otherwise)
    return status;
}

// host-side choice of the slot-loop specialisation
__host__ __device__ inline int choose_spec(const uint8_t* d, int M) {
    if (M < 1) return SCCAV_SPEC_GENERIC;
    const int first = d[0];
    const int type = first & SCCAV_SLOT_TYPE_MASK;
    if (first & SCCAV_SLOT_SHARED) return SCCAV_SPEC_GENERIC;
    if (type == SCCAV_SLOT_ELLIPSE) {
        for (int m = 0; m < M; ++m)
            if (d[m] != first) return SCCAV_SPEC_GENERIC;                   // same STATIC flag on every slot
        return SCCAV_SPEC_ELLIPSE;
    }
    if (type == SCCAV_SLOT_ELLIPSE_PREP) {
        for (int m = 0; m < M; ++m)
            if (d[m] != first) return SCCAV_SPEC_GENERIC;
        return SCCAV_SPEC_ELLIPSE_PREP;
    }
    return SCCAV_SPEC_GENERIC;
}

// ------------------------------------------------------------------------------------------
// nominal controller: Stanley (function form) -- stanley_controller_ellipse.py:146-212
// The global argmin compares squared distances dx*dx + dy*dy (monotone in np.hypot; first
// minimum wins, strict <), over ALL P points, like calc_target_index (:202-205).
// ------------------------------------------------------------------------------------------
// cross-track error + steering law once the nearest index is known (sce.py:208-212,159-167)
// NORMAL_IDENTITY (opt-in, the SCCAV_FLAG_FUSED_STEER instances): the front-axle normal
// [-cos(yaw + pi/2), -sin(yaw + pi/2)] of sce.py:208-209 is taken as [sin yaw, -cos yaw] from the sincos the plant
// already holds -- the same vector to an ulp or two, one sincos per tick cheaper.
template <typename T, typename CXY, bool NORMAL_IDENTITY = false>
__device__ __forceinline__ T stanley_law(const Params<T>& P, CXY c, const T* __restrict__ cyaw, int idx,
                                         T fx, T fy, T yaw, T v, int& target_idx, T syaw = T(0), T cyw = T(0)) {
    typedef Real<T> R;
    T e;
    if (NORMAL_IDENTITY) {
        e = (fx - c.x) * syaw + (fy - c.y) * (-cyw);
    } else {
        T s2, c2;
        R::sincos_(yaw + R::pi() / T(2), &s2, &c2);                                      // sce.py:208-209
        e = (fx - c.x) * (-c2) + (fy - c.y) * (-s2);
    }
    if (target_idx >= idx) idx = target_idx;                                             // sce.py:159-160
    T theta_e = normalize_angle<T>(cyaw[idx] - yaw);
    T theta_d = R::atan2_(P.k_stanley * e, v + P.ks_stanley);
    target_idx = idx;
    return theta_e + theta_d;
}

// SCCAV_FLAG_FUSED_STEER, the Stanley law in the QP's own coordinate.  The reference forms delta_ref = theta_e +
// atan2(k e, v + ks) (sce.py:159-167) and the filter turns it into beta_ref = atan2(lr tan(delta_ref), L) (cbf.py:175).
// tan has period pi and tan(atan2(y, x)) = y / x, so with (se, ce) = sin / cos(cyaw - yaw) -- from the sin / cos of the
// course yaws staged once per CTA and the sin / cos of the vehicle's yaw the plant already holds --
//   tan(delta_ref) = (se (v + ks) + ce k e) / (ce (v + ks) - se k e)
// and beta_ref is ONE atan2 of (lr num, L den) folded into the right half plane: the same function of the same state, a
// few ulp from the literal sequence (normalize_angle, atan2, tan, atan2), which is what the flag allows.
template <typename T, typename CXY>
__device__ __forceinline__ T stanley_law_beta(const Params<T>& P, CXY c, const CXY* __restrict__ trig, int idx, T fx, T fy, T v,
                                              int& target_idx, T syaw, T cyw) {
    typedef Real<T> R;
    const T e = (fx - c.x) * syaw + (fy - c.y) * (-cyw);
    if (target_idx >= idx) idx = target_idx;                                             // sce.py:159-160
    target_idx = idx;
    const CXY t = trig[idx];                                                             // (sin, cos) of cyaw[idx]
    const T se = t.x * cyw - t.y * syaw, ce = t.y * cyw + t.x * syaw;
    const T ye = P.k_stanley * e, xe = v + P.ks_stanley;
    T num = se, den = ce;
    if (ye != T(0) || xe != T(0)) {                                                      // atan2(0, 0) = 0
        num = se * xe + ce * ye;
        den = ce * xe - se * ye;
    }
    T yy = P.lr * num, xx = (P.lf + P.lr) * den;
    if (xx < T(0)) { yy = -yy; xx = -xx; }
    return R::atan2_(yy, xx);
}

// the search half of stanley(): front axle, nearest index, hint bookkeeping
template <typename T, typename T2>
__device__ __forceinline__ int stanley_search(const Params<T>& P, const CourseIndex<T, T2>& ci, T x, T y, T syaw, T cyw,
                                              int& near_idx, int& adv, int* evals, T& fx, T& fy) {
    fx = x + P.L * cyw;
    fy = y + P.L * syaw;
    const int idx = course_nearest<T, T2>(ci, fx, fy, near_idx + adv, evals);
    adv = idx - near_idx;
    adv = adv < -4 * SCCAV_LEAF ? 0 : (adv > 4 * SCCAV_LEAF ? 0 : adv);      // a jump is not a trend
    near_idx = idx;
    return idx;
}

// course staged in shared memory with its capsule tree: exact pruned nearest search.  The hint is the previous
// nearest index advanced by the previous tick's advance (a vehicle moves about as far as it did last tick);
// any hint gives the same index, a good one makes the search cheap.
template <typename T, typename T2, bool NORMAL_IDENTITY = false>
__device__ __forceinline__ T stanley(const Params<T>& P, const CourseIndex<T, T2>& ci, const T* __restrict__ cyaw,
                                     T x, T y, T yaw, T v, T syaw, T cyw, int& target_idx, int& near_idx, int& adv, int* evals) {
    T fx = x + P.L * cyw;
    T fy = y + P.L * syaw;
    int idx = course_nearest<T, T2>(ci, fx, fy, near_idx + adv, evals);
    adv = idx - near_idx;
    adv = adv < -4 * SCCAV_LEAF ? 0 : (adv > 4 * SCCAV_LEAF ? 0 : adv);      // a jump is not a trend
    near_idx = idx;
    return stanley_law<T, T2, NORMAL_IDENTITY>(P, ci.pt(idx), cyaw, idx, fx, fy, yaw, v, target_idx, syaw, cyw);
}

}  // namespace sccav
