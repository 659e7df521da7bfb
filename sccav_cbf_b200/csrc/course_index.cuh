// course_index.cuh -- exact nearest-way-point search for the Stanley controller.
//
// The reference scans ALL P course points every tick (calc_target_index,
// test_scripts/stanley_controller_ellipse.py:188-212: np.hypot over the whole course + np.argmin,
// first minimum wins).  At P = 2034 that scan is > 80 % of the arithmetic of a closed-loop step.
// This header returns THE SAME index -- the lexicographic minimum of (d2_i, i), d2_i computed with
// the same operations as the full scan -- while touching a few dozen points:
//
//   * the course is cut into leaves of LEAF consecutive points and supers of SUPER_LEAVES leaves;
//     every leaf / super carries a CAPSULE: the chord from its first to its last point plus a
//     radius rho >= max_i dist(p_i, chord) (a few mm for a leaf of a smooth course);
//   * for every point p_i of a block  |f - p_i| >= dist(f, chord) - rho  (triangle inequality through
//     the chord point closest to p_i), so a block is skipped only when
//         dist(f, chord) > sqrt(best) (1 + eps) + rho (1 + 4 eps) + slack ,
//     with eps and slack (1e-9 and 1e-12 x the course extent in fp64) orders of magnitude above the
//     rounding error of the few operations involved (a mis-rounded chord parameter t only moves the
//     foot point ALONG the chord, a second-order effect covered by slack).  Neither a smaller
//     distance nor an equal one with a smaller index can therefore hide in a skipped block;
//   * the search starts in the leaf of a hint (the previous tick's nearest index), which makes the
//     bound tight immediately; any hint gives the same result, only the cost differs.
//
// Control flow is written so that the lanes of a warp each walk their OWN blocks inside common
// loops (trip count = per-lane count; a warp pays the maximum over its lanes, not the union).
// In shared memory every leaf is padded by one point so that lanes scanning different leaves hit
// different banks.  Functions are __host__ __device__: tests/test_course_index.py runs them on
// the CPU against the exhaustive scan (sccav_debug_course_index_host); the rollout kernel runs
// them on shared memory.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace sccav {

#define SCCAV_LEAF 16
#define SCCAV_LEAF_SHIFT 4
#define SCCAV_SUPER_LEAVES 8

template <typename T> struct IndexEps;
template <> struct IndexEps<double> {
    static __host__ __device__ __forceinline__ double rel() { return 1e-9; }
    static __host__ __device__ __forceinline__ double abs_rel() { return 1e-12; }
    static __host__ __device__ __forceinline__ double tiny() { return 1e-300; }
};
template <> struct IndexEps<float> {
    static __host__ __device__ __forceinline__ float rel() { return 1e-4f; }
    static __host__ __device__ __forceinline__ float abs_rel() { return 1e-5f; }
    static __host__ __device__ __forceinline__ float tiny() { return 1e-30f; }
};

// position of course point i in the padded point array (one spare slot after every leaf)
__host__ __device__ __forceinline__ int course_slot(int i) { return i + (i >> SCCAV_LEAF_SHIFT); }
__host__ __device__ __forceinline__ int course_nleaf(int np) { return (np + SCCAV_LEAF - 1) / SCCAV_LEAF; }
__host__ __device__ __forceinline__ int course_nsup(int np) {
    return (course_nleaf(np) + SCCAV_SUPER_LEAVES - 1) / SCCAV_SUPER_LEAVES;
}
__host__ __device__ __forceinline__ int course_nslot(int np) { return np + course_nleaf(np); }

// capsule = chord a + t ab (t in [0,1]) with radius; stored as three 2-vectors
template <typename T2> struct Capsules {
    T2* a;    // chord start
    T2* ab;   // chord vector
    T2* ir;   // (1 / |ab|^2  or 0, inflated radius)
};

// View of a staged course
template <typename T, typename T2> struct CourseIndex {
    const T2* xy;      // [course_nslot(np)] padded points, point i at course_slot(i)
    Capsules<T2> leaf; // [nleaf]
    Capsules<T2> sup;  // [nsup]
    int np, nleaf, nsup;
    __host__ __device__ __forceinline__ T2 pt(int i) const { return xy[course_slot(i)]; }
};

// squared distance from f to the chord of a capsule (computed >= true up to the slack, see header)
template <typename T, typename T2>
__host__ __device__ __forceinline__ T chord_dist2(T2 a, T2 ab, T inv, T fx, T fy) {
    // explicit fma: this is a conservative bound, not reference arithmetic (the fp64 unit is
    // otherwise compiled with -fmad=false)
    T vx = fx - a.x, vy = fy - a.y;
    T t = fma(vx, ab.x, vy * ab.y) * inv;
    t = t < T(0) ? T(0) : (t > T(1) ? T(1) : t);
    T ex = fma(-t, ab.x, vx), ey = fma(-t, ab.y, vy);
    return fma(ex, ex, ey * ey);
}

// Capsule of points [lo, hi) (indices into the padded array through course_slot); `extent` = an
// upper bound of |coordinates| of the whole course (absolute slack of the radius).
template <typename T, typename T2>
__host__ __device__ inline void capsule_build(const T2* xy, int lo, int hi, T extent, T2& a, T2& ab, T2& ir) {
    a = xy[course_slot(lo)];
    T2 b = xy[course_slot(hi - 1)];
    ab.x = b.x - a.x;
    ab.y = b.y - a.y;
    T l2 = ab.x * ab.x + ab.y * ab.y;
    T inv = l2 > T(0) ? T(1) / l2 : T(0);
    T m = T(0);
    for (int i = lo; i < hi; ++i) {
        T2 p = xy[course_slot(i)];
        T d2 = chord_dist2<T, T2>(a, ab, inv, p.x, p.y);
        m = d2 > m ? d2 : m;
    }
    T rho = (T)sqrt((double)m);
    ir.x = inv;
    ir.y = rho * (T(1) + T(4) * IndexEps<T>::rel()) + IndexEps<T>::abs_rel() * extent + IndexEps<T>::tiny();
}

// true when no point of the capsule can be at squared distance <= best from f  (reach = sqrt(best)(1+eps))
template <typename T, typename T2>
__host__ __device__ __forceinline__ bool capsule_skip(const Capsules<T2>& c, int k, T fx, T fy, T reach) {
    T2 ir = c.ir[k];
    T d2 = chord_dist2<T, T2>(c.a[k], c.ab[k], ir.x, fx, fy);
    T thr = reach + ir.y;
    return d2 > thr * thr;
}

template <typename T> __host__ __device__ __forceinline__ T index_reach(T best) {
    return (T)sqrt((double)best) * (T(1) + IndexEps<T>::rel());
}
template <> __host__ __device__ __forceinline__ float index_reach<float>(float best) {
    return sqrtf(best) * (1.0f + IndexEps<float>::rel());
}

// Scan one leaf in ascending index order (strict <: first minimum inside the leaf), then merge
// lexicographically with the running (best, ib).
template <typename T, typename T2>
__host__ __device__ __forceinline__ void index_scan_leaf(const CourseIndex<T, T2>& ci, int leaf, T fx, T fy, T& best, int& ib) {
    const int lo = leaf << SCCAV_LEAF_SHIFT;
    const T2* p = ci.xy + lo + leaf;                       // course_slot(lo)
    const int cnt = (ci.np - lo < SCCAV_LEAF) ? ci.np - lo : SCCAV_LEAF;
    T lb = (T)INFINITY;
    int li = 0;
    if (cnt == SCCAV_LEAF) {
#pragma unroll
        for (int j = 0; j < SCCAV_LEAF; ++j) {
            T2 q = p[j];
            T dx = fx - q.x, dy = fy - q.y;
            T d2 = dx * dx + dy * dy;
            if (d2 < lb) { lb = d2; li = j; }
        }
    } else {
        for (int j = 0; j < cnt; ++j) {
            T2 q = p[j];
            T dx = fx - q.x, dy = fy - q.y;
            T d2 = dx * dx + dy * dy;
            if (d2 < lb) { lb = d2; li = j; }
        }
    }
    li += lo;
    if (lb < best || (lb == best && li < ib)) { best = lb; ib = li; }
}

// Exhaustive scan (what the reference does): first minimum of d2 over all points.
template <typename T, typename T2>
__host__ __device__ inline int course_nearest_full(const T2* xy, int np, T fx, T fy) {
    T best = (T)INFINITY;
    int ib = 0;
    for (int i = 0; i < np; ++i) {
        T2 p = xy[course_slot(i)];
        T dx = fx - p.x, dy = fy - p.y;
        T d2 = dx * dx + dy * dy;
        if (d2 < best) { best = d2; ib = i; }
    }
    return ib;
}

__host__ __device__ __forceinline__ int lowest_bit(uint32_t m) {
#ifdef __CUDA_ARCH__
    return __ffs((int)m) - 1;
#else
    return __builtin_ctz(m);
#endif
}

// Exact global nearest index (first minimum) of (fx, fy) over the whole course.
// hint: any index (the previous nearest index; clamped into [0, np)); evals (optional) counts
// distance evaluations + capsule tests for the roofline accounting.
// UNR: unroll factor of the two capsule-test loops (2 lets independent tests interleave in the lean ellipse
// instances of the rollout; the larger generic instances are better off with 1 -- registers).
template <typename T, typename T2, int UNR = 2>
__host__ __device__ inline int course_nearest(const CourseIndex<T, T2>& ci, T fx, T fy, int hint, int* evals) {
    if (hint < 0) hint = 0;
    if (hint >= ci.np) hint = ci.np - 1;
    // phase A: the hint's leaf and the next one (a vehicle advances about half a leaf per tick)
    const int leaf0 = hint >> SCCAV_LEAF_SHIFT;
    const int leaf1 = (leaf0 + 1 < ci.nleaf) ? leaf0 + 1 : leaf0;
    T best = (T)INFINITY;
    int ib = ci.np;
    int ne = 2 * SCCAV_LEAF;
#pragma unroll 1
    for (int l = leaf0; l <= leaf1; ++l) index_scan_leaf<T, T2>(ci, l, fx, fy, best, ib);
    // NaN / overflowing query (np.argmin of all-NaN is 0): no usable bound, do what the reference does
    if (!(best < (T)INFINITY)) return course_nearest_full<T, T2>(ci.xy, ci.np, fx, fy);
    T reach = index_reach<T>(best);
    for (int s0 = 0; s0 < ci.nsup; s0 += 32) {
        // phase B: which supers can hold a point at distance <= best?  (same trip count for all lanes)
        const int ns = (ci.nsup - s0 < 32) ? ci.nsup - s0 : 32;
        uint32_t smask = 0u;
#pragma unroll UNR
        for (int j = 0; j < ns; ++j)
            if (!capsule_skip<T, T2>(ci.sup, s0 + j, fx, fy, reach)) smask |= 1u << j;
        ne += ns;
        // phase C: every lane walks its own supers / leaves
        while (smask) {
            const int s = s0 + lowest_bit(smask);
            smask &= smask - 1u;
            if (capsule_skip<T, T2>(ci.sup, s, fx, fy, reach)) { ++ne; continue; }   // bound tightened since phase B
            const int l0 = s * SCCAV_SUPER_LEAVES;
            uint32_t todo = 0u;
#pragma unroll UNR
            for (int j = 0; j < SCCAV_SUPER_LEAVES; ++j) {
                const int l = l0 + j;
                if (l < ci.nleaf && l != leaf0 && l != leaf1 && !capsule_skip<T, T2>(ci.leaf, l, fx, fy, reach)) todo |= 1u << j;
            }
            ne += SCCAV_SUPER_LEAVES;
            while (todo) {
                const int l = l0 + lowest_bit(todo);
                todo &= todo - 1u;
                const T before = best;
                index_scan_leaf<T, T2>(ci, l, fx, fy, best, ib);
                ne += SCCAV_LEAF;
                if (best < before) {
                    // tighter bound: drop the remaining leaves of this super that it now excludes
                    reach = index_reach<T>(best);
                    uint32_t keep = 0u, rest = todo;
                    while (rest) {
                        const int j = lowest_bit(rest);
                        rest &= rest - 1u;
                        if (!capsule_skip<T, T2>(ci.leaf, l0 + j, fx, fy, reach)) keep |= 1u << j;
                        ++ne;
                    }
                    todo = keep;
                }
            }
        }
    }
    if (evals) *evals += ne;
    return ib;
}

}  // namespace sccav
