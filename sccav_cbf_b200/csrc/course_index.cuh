// course_index.cuh -- exact nearest-way-point search for the Stanley controller.
//
// The reference scans ALL P course points every tick (calc_target_index,
// test_scripts/stanley_controller_ellipse.py:188-212: np.hypot over the whole course + np.argmin,
// first minimum wins).  At P = 2034 that scan is > 80 % of the arithmetic of a closed-loop step.
// This header returns THE SAME index (the lexicographic minimum of (d2_i, i), d2_i computed with the
// same operations as the full scan) while touching only a few dozen points:
//
//   * the course is cut into leaves of LEAF consecutive points and supers of SUPER_LEAVES leaves;
//     every leaf / super carries a bounding circle (centre c, radius r >= max |p_i - c|, inflated);
//   * a block is skipped only when  |f - c|^2 > (sqrt(best) (1+eps) + r)^2 , which (triangle
//     inequality + eps >> rounding error) proves every point inside has d2_i > best, so neither a
//     smaller distance nor an equal one with a smaller index can hide in a skipped block;
//   * the search starts in the leaf of a hint (the previous tick's nearest index), which makes the
//     bound tight immediately; any hint gives the same result, only the cost differs.
//
// Control flow is written so that the lanes of a warp each walk their OWN blocks inside common
// loops (trip count = per-lane count; a warp pays the maximum over its lanes, not the union).
// Functions are __host__ __device__: tests/test_course_index.py runs them on the CPU against the
// exhaustive scan (sccav_debug_course_index_host), the rollout kernel runs them on shared memory.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace sccav {

#define SCCAV_LEAF 16
#define SCCAV_SUPER_LEAVES 8

template <typename T> struct IndexEps;
template <> struct IndexEps<double> {
    static __host__ __device__ __forceinline__ double rel() { return 1e-9; }
};
template <> struct IndexEps<float> {
    static __host__ __device__ __forceinline__ float rel() { return 1e-4f; }
};

// View of a staged course: points (x, y) interleaved + bounding circles of leaves and supers.
template <typename T, typename T2> struct CourseIndex {
    const T2* xy;      // [np]
    const T2* leaf_c;  // [nleaf] centre
    const T* leaf_r;   // [nleaf] inflated radius
    const T2* sup_c;   // [nsup]
    const T* sup_r;    // [nsup]
    int np, nleaf, nsup;
};

__host__ __device__ __forceinline__ int course_nleaf(int np) { return (np + SCCAV_LEAF - 1) / SCCAV_LEAF; }
__host__ __device__ __forceinline__ int course_nsup(int np) {
    return (course_nleaf(np) + SCCAV_SUPER_LEAVES - 1) / SCCAV_SUPER_LEAVES;
}

// Bounding circle of points [lo, hi): centre = middle of the bounding box, radius = max distance,
// inflated so that it is a certain upper bound whatever the rounding of the few operations here.
template <typename T, typename T2>
__host__ __device__ inline void bounding_circle(const T2* xy, int lo, int hi, T2& c, T& r) {
    T minx = xy[lo].x, maxx = minx, miny = xy[lo].y, maxy = miny;
    for (int i = lo + 1; i < hi; ++i) {
        T x = xy[i].x, y = xy[i].y;
        minx = x < minx ? x : minx; maxx = x > maxx ? x : maxx;
        miny = y < miny ? y : miny; maxy = y > maxy ? y : maxy;
    }
    c.x = (minx + maxx) * T(0.5);
    c.y = (miny + maxy) * T(0.5);
    T m = T(0);
    for (int i = lo; i < hi; ++i) {
        T dx = xy[i].x - c.x, dy = xy[i].y - c.y;
        T d2 = dx * dx + dy * dy;
        m = d2 > m ? d2 : m;
    }
    T rr = (T)sqrt((double)m);
    r = rr * (T(1) + T(4) * IndexEps<T>::rel()) + (T)1e-30;
}

// One candidate: lexicographic (d2, i) minimum, d2 with the operations of the exhaustive scan.
template <typename T, typename T2>
__host__ __device__ __forceinline__ void index_try(const T2* xy, int i, T fx, T fy, T& best, int& ib) {
    T2 p = xy[i];
    T dx = fx - p.x, dy = fy - p.y;
    T d2 = dx * dx + dy * dy;
    if (d2 < best || (d2 == best && i < ib)) { best = d2; ib = i; }
}

template <typename T, typename T2>
__host__ __device__ __forceinline__ void index_scan_leaf(const CourseIndex<T, T2>& ci, int leaf, T fx, T fy, T& best, int& ib) {
    const int lo = leaf * SCCAV_LEAF;
    if (lo + SCCAV_LEAF <= ci.np) {
#pragma unroll
        for (int j = 0; j < SCCAV_LEAF; ++j) index_try<T, T2>(ci.xy, lo + j, fx, fy, best, ib);
    } else {
        for (int i = lo; i < ci.np; ++i) index_try<T, T2>(ci.xy, i, fx, fy, best, ib);
    }
}

// true when the circle (c, r) cannot contain a point at squared distance <= best from f
template <typename T, typename T2>
__host__ __device__ __forceinline__ bool index_can_skip(T2 c, T r, T fx, T fy, T reach) {
    T dx = fx - c.x, dy = fy - c.y;
    T dc2 = dx * dx + dy * dy;
    T thr = reach + r;
    return dc2 > thr * thr;
}

template <typename T> __host__ __device__ __forceinline__ T index_reach(T best) {
    return (T)sqrt((double)best) * (T(1) + IndexEps<T>::rel());
}
template <> __host__ __device__ __forceinline__ float index_reach<float>(float best) {
    return sqrtf(best) * (1.0f + IndexEps<float>::rel());
}

// Exhaustive scan (what the reference does): first minimum of d2 over all points.
template <typename T, typename T2>
__host__ __device__ inline int course_nearest_full(const T2* xy, int np, T fx, T fy) {
    T best = (T)INFINITY;
    int ib = 0;
    for (int i = 0; i < np; ++i) {
        T2 p = xy[i];
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
// distance evaluations + circle tests for the roofline accounting.
template <typename T, typename T2>
__host__ __device__ inline int course_nearest(const CourseIndex<T, T2>& ci, T fx, T fy, int hint, int* evals) {
    if (hint < 0) hint = 0;
    if (hint >= ci.np) hint = ci.np - 1;
    // phase A: the hint's leaf and the next one (a vehicle advances about half a leaf per tick)
    const int leaf0 = hint / SCCAV_LEAF;
    const int leaf1 = (leaf0 + 1 < ci.nleaf) ? leaf0 + 1 : leaf0;
    T best = (T)INFINITY;
    int ib = ci.np;
    int ne = 2 * SCCAV_LEAF;
    index_scan_leaf<T, T2>(ci, leaf0, fx, fy, best, ib);
    if (leaf1 != leaf0) index_scan_leaf<T, T2>(ci, leaf1, fx, fy, best, ib);
    // NaN / overflowing query (np.argmin of all-NaN is 0): no usable bound, do what the reference does
    if (!(best < (T)INFINITY)) return course_nearest_full<T, T2>(ci.xy, ci.np, fx, fy);
    T reach = index_reach<T>(best);
    for (int s0 = 0; s0 < ci.nsup; s0 += 32) {
        // phase B: which supers can hold a point at distance <= best?  (same trip count for all lanes)
        const int ns = (ci.nsup - s0 < 32) ? ci.nsup - s0 : 32;
        uint32_t smask = 0u;
        for (int j = 0; j < ns; ++j)
            if (!index_can_skip<T, T2>(ci.sup_c[s0 + j], ci.sup_r[s0 + j], fx, fy, reach)) smask |= 1u << j;
        ne += ns;
        // phase C: every lane walks its own supers / leaves
        while (smask) {
            const int s = s0 + lowest_bit(smask);
            smask &= smask - 1u;
            const int l0 = s * SCCAV_SUPER_LEAVES;
            uint32_t todo = 0u;
#pragma unroll
            for (int j = 0; j < SCCAV_SUPER_LEAVES; ++j) {
                const int l = l0 + j;
                if (l < ci.nleaf && l != leaf0 && l != leaf1 &&
                    !index_can_skip<T, T2>(ci.leaf_c[l], ci.leaf_r[l], fx, fy, reach))
                    todo |= 1u << j;
            }
            ne += SCCAV_SUPER_LEAVES;
            while (todo) {
                const int l = l0 + lowest_bit(todo);
                todo &= todo - 1u;
                // the bound may have tightened since the mask was built
                if (index_can_skip<T, T2>(ci.leaf_c[l], ci.leaf_r[l], fx, fy, reach)) continue;
                const T before = best;
                index_scan_leaf<T, T2>(ci, l, fx, fy, best, ib);
                ne += SCCAV_LEAF + 1;
                if (best < before) reach = index_reach<T>(best);
            }
        }
    }
    if (evals) *evals += ne;
    return ib;
}

}  // namespace sccav
