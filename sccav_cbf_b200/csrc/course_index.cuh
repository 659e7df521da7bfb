// course_index.cuh -- exact nearest-way-point search for the Stanley controller.
//
// The reference scans ALL P course points every tick (calc_target_index,
// test_scripts/stanley_controller_ellipse.py:188-212: np.hypot over the whole course + np.argmin,
// first minimum wins).  At P = 2034 that scan is > 80 % of the arithmetic of a closed-loop step.
// This header returns THE SAME index -- the lexicographic minimum of (d2_i, i), d2_i computed with
// the same operations as the full scan -- while evaluating 16 points:
//
//   * the course is cut into leaves of 8 consecutive points; leaves are the bottom level of a tree of
//     fan-out 8 (node j of level k = nodes 8j .. 8j+7 of level k-1); every node carries a CAPSULE: a
//     chord from (about) its first to its last point and a radius rho >= max_i dist(p_i, chord);
//   * for every point p_i of a node  |f - p_i| >= dist(f, chord) - rho  (triangle inequality through the
//     chord point closest to p_i), so a node is skipped only when  dist(f, chord) > sqrt(best) + rho + slack.
//     The test is a BOUND, not reference arithmetic, so it is evaluated in single precision relative to
//     an origin on the course (12 fp32 instructions, 24 bytes per node) with every rounding error charged
//     to the slack: the chord is whatever its fp32 end points say and rho is measured against THAT chord in
//     double precision at build time; the query point's conversion, the subtraction, the (mis-rounded)
//     chord parameter and the final products are covered by  2e-6 (|f - o|_1 + extent)  -- an order of
//     magnitude above their sum (<= 5e-7 of the same quantity) -- and sqrt(best) is rounded up.  Neither a
//     smaller distance nor an equal one with a smaller index can therefore hide in a skipped node;
//   * the search scans the two leaves around a hint (the index predicted from the previous ticks), which
//     makes the bound tight immediately, then CLIMBS: at every level the siblings of the scanned leaves'
//     ancestor are tested in ONE unrolled batch of 8 independent tests (the same code, the same trip count
//     and eight-fold instruction-level parallelism for every lane of a warp).  A sibling that cannot be
//     excluded is opened: its 8 children are tested in another batch, a surviving leaf is scanned.  Any
//     hint gives the same result, only the cost differs.
//
// In shared memory every leaf is padded by one point so that lanes scanning different leaves hit
// different banks.  Functions are __host__ __device__: tests/test_course_index.py runs them on the CPU
// against the exhaustive scan (sccav_debug_course_index_host); the kernels run them on shared memory.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace sccav {

#define SCCAV_LEAF_SHIFT 3
#define SCCAV_LEAF (1 << SCCAV_LEAF_SHIFT)
#define SCCAV_FAN_SHIFT 3
#define SCCAV_FAN (1 << SCCAV_FAN_SHIFT)
#define SCCAV_MAX_LEVELS 7          /* 7 levels of fan-out 8 over leaves of 8 points (16 M points): far more than any course */
#define SCCAV_TOP_MAX SCCAV_FAN     /* the top level holds at most 8 nodes, siblings under a virtual root.  (A flat top of 32 tight
                                       nodes instead of 4 loose ones + their children was measured: 13.1 vs 12.3 ms) */

// position of course point i in the padded point array (one spare slot after every leaf)
__host__ __device__ __forceinline__ int course_slot(int i) { return i + (i >> SCCAV_LEAF_SHIFT); }
__host__ __device__ __forceinline__ int course_nleaf(int np) { return (np + SCCAV_LEAF - 1) / SCCAV_LEAF; }
// (whole leaves: the tail of the last leaf is filled with copies of the last point -- equal distances, later indices:
// they can never win the strict first-minimum comparison -- so that every leaf scan is the same 8 unrolled points)
__host__ __device__ __forceinline__ int course_nslot(int np) { return course_nleaf(np) * (SCCAV_LEAF + 1); }

// Levels of the tree over a course of np points: level 0 = the leaves, level k + 1 has ceil(n_k / 8) nodes; the
// top level has at most SCCAV_TOP_MAX nodes, all siblings under a root that is never tested (long, curved nodes
// have loose capsules: a flat top of tight ones is tested in full instead -- the same loads for every lane).  A node is 32 bytes (chord float4,
// (1/len^2, radius) float2, 8 spare); a group of 8 siblings is followed by 16 spare bytes, so that lanes of a warp
// reading the same member of DIFFERENT groups hit different banks (the stride between groups is 272 B = 17 x 16 B).
// lev[2k] = first 16-byte unit of level k in the node storage, lev[2k + 1] = number of nodes of level k.
// Returns the number of levels (>= 1); *units (optional) = 16-byte units of node storage.  lev may be NULL.
#define SCCAV_NODE_UNITS 2          /* 16-byte units per node */
#define SCCAV_GROUP_UNITS 17        /* 16-byte units per group of 8 nodes (8 x 2 + 1 spare) */
__host__ __device__ inline int course_levels(int np, int* lev, int* units) {
    int cnt = course_nleaf(np), o = 0, k = 0;
    for (;;) {
        if (lev) { lev[2 * k] = o; lev[2 * k + 1] = cnt; }
        o += ((cnt + SCCAV_FAN - 1) >> SCCAV_FAN_SHIFT) * SCCAV_GROUP_UNITS;
        ++k;
        if (cnt <= SCCAV_TOP_MAX || k >= SCCAV_MAX_LEVELS) break;
        cnt = (cnt + SCCAV_FAN - 1) >> SCCAV_FAN_SHIFT;
    }
    if (units) *units = o;
    return k;
}
__host__ __device__ inline int course_node_units(int np) {
    int u;
    course_levels(np, nullptr, &u);
    return u;
}

// View of a staged course: padded points + the fp32 capsules of the tree nodes
template <typename T, typename T2> struct CourseIndex {
    const T2* xy;      // [course_nslot(np)] padded points, point i at course_slot(i)
    float4* node;      // node storage (16-byte units): chord (start x, y, vector z, w) at +0, (1 / |chord|^2 or 0, radius) at +1
    const int* lev;    // [2 nlev] first unit and node count of every level
    const T* org;      // origin (x, y) of the fp32 frame: a point of the course
    const float* ext;  // [1] inflated max-norm extent of the course around the origin
    int np, nleaf, nlev;
    __host__ __device__ __forceinline__ T2 pt(int i) const { return xy[course_slot(i)]; }
    // first unit of node n of the level that starts at unit `base`
    static __host__ __device__ __forceinline__ int unit(int base, int n) {
        return base + (n >> SCCAV_FAN_SHIFT) * SCCAV_GROUP_UNITS + (n & (SCCAV_FAN - 1)) * SCCAV_NODE_UNITS;
    }
};

// point range [lo, hi) of node j of level k
__host__ __device__ __forceinline__ void node_range(int np, int k, int j, int& lo, int& hi) {
    const int sh = SCCAV_FAN_SHIFT * k + SCCAV_LEAF_SHIFT;
    const int64_t l = (int64_t)j << sh, h = (int64_t)(j + 1) << sh;
    lo = (int)(l < np ? l : np);
    hi = (int)(h < np ? h : np);
}

// fp32 chord of a node from its first and last point (double precision, relative to the origin)
__host__ __device__ __forceinline__ void capsule_chord(double ax, double ay, double bx, double by, float4& c, double& inv) {
    c.x = (float)ax; c.y = (float)ay;
    c.z = (float)(bx - ax); c.w = (float)(by - ay);
    const double l2 = (double)c.z * (double)c.z + (double)c.w * (double)c.w;
    inv = l2 > 0.0 ? 1.0 / l2 : 0.0;
}

// exact (double) squared distance of the point (px, py) to the fp32 chord c
__host__ __device__ __forceinline__ double chord_dist2(const float4& c, double inv, double px, double py) {
    const double vx = px - (double)c.x, vy = py - (double)c.y;
    double t = (vx * (double)c.z + vy * (double)c.w) * inv;
    t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
    const double ex = vx - t * (double)c.z, ey = vy - t * (double)c.w;
    return ex * ex + ey * ey;
}

// (1 / |chord|^2, radius) of a node from the largest squared chord distance m of its points; the radius is inflated
// past the double -> float rounding, the sqrt and (for a float course) the rounding of p - o
__host__ __device__ __forceinline__ float2 capsule_ir(double inv, double m, float ext) {
    float2 r;
    r.x = (float)inv;
    r.y = (float)(sqrt(m) * 1.000001 + 2e-7 * (double)ext + 1e-30);
    return r;
}

__host__ __device__ __forceinline__ float course_ext_inflate(double ext) { return (float)(ext * 1.000001 + 1e-30); }

// Capsule of node j of level k, serial form (the kernels build theirs warp-cooperatively with the same functions:
// the radius is a maximum, so the order of the points does not matter).
template <typename T, typename T2>
__host__ __device__ inline void capsule_build(const T2* xy, int np, int k, int j, double ox, double oy, float ext, float4& c, float2& r) {
    int lo, hi;
    node_range(np, k, j, lo, hi);
    const T2 a = xy[course_slot(lo)], b = xy[course_slot(hi - 1)];
    double inv;
    capsule_chord((double)a.x - ox, (double)a.y - oy, (double)b.x - ox, (double)b.y - oy, c, inv);
    double m = 0.0;
    for (int i = lo; i < hi; ++i) {
        const T2 p = xy[course_slot(i)];
        const double d2 = chord_dist2(c, inv, (double)p.x - ox, (double)p.y - oy);
        m = d2 > m ? d2 : m;
    }
    r = capsule_ir(inv, m, ext);
}

// the query in the fp32 frame: point, and (reach + slack) -- everything of the skip test that does not depend on the node
struct IndexQuery {
    float qx, qy, slack, base;
};

template <typename T> __host__ __device__ __forceinline__ float index_reach32(T best) {
    // sqrt(best) rounded up: conversion and sqrtf are each within 6e-8 relative
    return sqrtf((float)best) * 1.000002f;
}

template <typename T, typename T2>
__host__ __device__ __forceinline__ IndexQuery index_query(const CourseIndex<T, T2>& ci, T fx, T fy) {
    IndexQuery q;
    q.qx = (float)(fx - ci.org[0]);
    q.qy = (float)(fy - ci.org[1]);
    q.slack = 2e-6f * ((fabsf(q.qx) + fabsf(q.qy)) + ci.ext[0]);
    q.base = INFINITY;                                  // no incumbent yet: nothing can be skipped
    return q;
}

__host__ __device__ __forceinline__ float sat01(float t) {
#ifdef __CUDA_ARCH__
    return __saturatef(t);
#else
    return fminf(fmaxf(t, 0.0f), 1.0f);
#endif
}

// One batch: the (up to) 8 nodes of group `group` of level `level`, except ex0 / ex1.  Returns the bit mask of the
// nodes that can NOT be excluded (a NaN anywhere keeps the node: conservative).
template <typename T, typename T2>
__host__ __device__ __forceinline__ uint32_t index_test8(const CourseIndex<T, T2>& ci, int level, int group, int ex0, int ex1,
                                                         const IndexQuery& q, int& ne) {
    const int first = group << SCCAV_FAN_SHIFT;
    const int left = ci.lev[2 * level + 1] - first;                         // nodes of the level from `first` on
    // which of the 8 members exist and are wanted (one register -> 8 predicates)
    uint32_t want = left >= SCCAV_FAN ? 0xffu : ((1u << (left > 0 ? left : 0)) - 1u);
    const uint32_t e0 = (uint32_t)(ex0 - first), e1 = (uint32_t)(ex1 - first);
    if (e0 < (uint32_t)SCCAV_FAN) want &= ~(1u << e0);
    if (e1 < (uint32_t)SCCAV_FAN) want &= ~(1u << e1);
    const float4* __restrict__ g = ci.node + ci.lev[2 * level] + group * SCCAV_GROUP_UNITS;
    uint32_t mask = 0u;
#pragma unroll
    for (int j = 0; j < SCCAV_FAN; ++j) {
        if (want & (1u << j)) {
            const float4 c = g[SCCAV_NODE_UNITS * j];
            const float2 r = *reinterpret_cast<const float2*>(g + SCCAV_NODE_UNITS * j + 1);
            const float vx = q.qx - c.x, vy = q.qy - c.y;
            const float t = sat01(fmaf(vx, c.z, vy * c.w) * r.x);
            const float ex = fmaf(-t, c.z, vx), ey = fmaf(-t, c.w, vy);
            const float d2 = fmaf(ex, ex, ey * ey);
            const float thr = q.base + r.y;
            if (!(d2 > thr * thr)) mask |= 1u << j;
        }
    }
#ifdef __CUDA_ARCH__
    ne += __popc(want);
#else
    ne += __builtin_popcount(want);
#endif
    return mask;
}

// Scan one leaf in ascending index order (strict <: first minimum inside the leaf), then merge
// lexicographically with the running (best, ib).  The staged array holds whole leaves (course_nslot).
template <typename T, typename T2>
__host__ __device__ __forceinline__ void index_scan_leaf(const CourseIndex<T, T2>& ci, int leaf, T fx, T fy, T& best, int& ib) {
    const int lo = leaf << SCCAV_LEAF_SHIFT;
    const T2* p = ci.xy + lo + leaf;                       // course_slot(lo)
    T lb = (T)INFINITY;
    int li = 0;
#pragma unroll
    for (int j = 0; j < SCCAV_LEAF; ++j) {
        T2 q = p[j];
        T dx = fx - q.x, dy = fy - q.y;
        T d2 = dx * dx + dy * dy;
        if (d2 < lb) { lb = d2; li = j; }
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

__host__ __device__ __forceinline__ int lowest_bit64(uint64_t m) {
#ifdef __CUDA_ARCH__
    return __ffsll((long long)m) - 1;
#else
    return __builtin_ctzll(m);
#endif
}

// Exact global nearest index (first minimum) of (fx, fy) over the whole course.
// hint: any index (the predicted nearest index; clamped into [0, np)); evals (optional) counts
// distance evaluations + capsule tests for the roofline accounting.
//
// ONE work loop: `pending` holds, per level, the nodes still to be processed in the group of the current position's
// ancestor (8 bits a level): a level-0 entry is a leaf to scan, an entry of level k >= 1 a node to OPEN (its 8 children
// are tested in one batch; those that cannot be excluded become entries one level down).  It starts with the two leaves
// around the hint and their ancestors up to the virtual root -- opening an ancestor tests the siblings of the level below,
// which is the climb -- and always takes the lowest level first, i.e. the nearest work.  Every lane of a warp runs the
// same two pieces of code (one batch of tests, one leaf scan) whatever level or node it is working on.
template <typename T, typename T2>
__host__ __device__ inline int course_nearest(const CourseIndex<T, T2>& ci, T fx, T fy, int hint, int* evals) {
    if (hint < 0) hint = 0;
    if (hint >= ci.np) hint = ci.np - 1;
    // the two leaves around the hint, inside the hint's group of 8 leaves
    const int leaf0 = hint >> SCCAV_LEAF_SHIFT;
    int lo = ((hint & (SCCAV_LEAF - 1)) >= SCCAV_LEAF / 2) ? leaf0 : leaf0 - 1;
    if ((lo & (SCCAV_FAN - 1)) == SCCAV_FAN - 1) lo = (lo == leaf0) ? lo - 1 : lo + 1;
    if (lo > ci.nleaf - 2) lo = ci.nleaf - 2;
    if (lo < 0) lo = 0;
    int hi = (lo + 1 < ci.nleaf) ? lo + 1 : lo;
    if ((hi >> SCCAV_FAN_SHIFT) != (lo >> SCCAV_FAN_SHIFT)) hi = lo;        // (a last group of one leaf)
    T best = (T)INFINITY;
    int ib = ci.np;
    int ne = 0;
    IndexQuery q = index_query<T, T2>(ci, fx, fy);
    uint64_t pending = (1ull << (lo & (SCCAV_FAN - 1))) | (1ull << (hi & (SCCAV_FAN - 1)));
    for (int k = 1; k <= ci.nlev; ++k) pending |= 1ull << (8 * k + ((lo >> (SCCAV_FAN_SHIFT * k)) & (SCCAV_FAN - 1)));
    int cur = lo;
    while (pending) {
        const int bit = lowest_bit64(pending);
        pending &= pending - 1u;
        const int m = bit >> 3;
        const int node = ((cur >> (SCCAV_FAN_SHIFT * (m + 1))) << SCCAV_FAN_SHIFT) + (bit & 7);
        cur = node << (SCCAV_FAN_SHIFT * m);
        if (m == 0) {
            const T before = best;
            index_scan_leaf<T, T2>(ci, node, fx, fy, best, ib);
            ne += SCCAV_LEAF;
            if (best < before) q.base = (index_reach32<T>(best) + q.slack) * 1.000002f;
        } else {
            // children of `node`, except the scanned leaves / their ancestor (which has its own entry)
            const int sh = SCCAV_FAN_SHIFT * (m - 1);
            pending |= (uint64_t)index_test8<T, T2>(ci, m - 1, node, lo >> sh, hi >> sh, q, ne) << (8 * (m - 1));
        }
    }
    // NaN / overflowing query (np.argmin of all-NaN is 0): nothing could be compared, do what the reference does
    if (!(best < (T)INFINITY)) return course_nearest_full<T, T2>(ci.xy, ci.np, fx, fy);
    if (evals) *evals += ne;
    return ib;
}

}  // namespace sccav
