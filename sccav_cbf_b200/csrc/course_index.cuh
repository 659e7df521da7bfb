// course_index.cuh -- exact nearest-way-point search for the Stanley controller.
//
// The reference scans ALL P course points every tick (calc_target_index,
// test_scripts/stanley_controller_ellipse.py:188-212: np.hypot over the whole course + np.argmin,
// first minimum wins).  At P = 2034 that scan is > 80 % of the arithmetic of a closed-loop step.
// This header returns THE SAME index -- the lexicographic minimum of (d2_i, i), d2_i computed with
// the same operations as the full scan -- while evaluating 16 points:
//
//   * the course is cut into leaves of 8 consecutive points; leaves are the bottom level of a BINARY tree
//     (node j of level h = leaves [j 2^h, (j + 1) 2^h)); every node carries a CAPSULE: a
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
//     makes the bound tight immediately, then tests the CANONICAL COVER of everything else: the leaves left
//     of the window are the disjoint union of at most one node per level (the set bits of the window's
//     leaf index), the leaves right of it likewise -- nodes that grow geometrically with their distance
//     from the window, so a node is never large where it is close.  At most two tests per level, the same
//     loop for every lane of a warp, about 8 tests per query.  A cover node that cannot be excluded is
//     OPENED: its two children are tested, the surviving ones opened in turn (depth first, a bit per level
//     as the stack), a surviving leaf is scanned.  Any hint gives the same result, only the cost differs.
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
#define SCCAV_MAX_LEVELS 16         /* binary levels over leaves of 8 points: 2^16 leaves = 524,288 points, far more than any course */

// position of course point i in the padded point array (one spare slot after every leaf)
__host__ __device__ __forceinline__ int course_slot(int i) { return i + (i >> SCCAV_LEAF_SHIFT); }
__host__ __device__ __forceinline__ int course_nleaf(int np) { return (np + SCCAV_LEAF - 1) / SCCAV_LEAF; }
// (whole leaves: the tail of the last leaf is filled with copies of the last point -- equal distances, later indices:
// they can never win the strict first-minimum comparison -- so that every leaf scan is the same 8 unrolled points)
__host__ __device__ __forceinline__ int course_nslot(int np) { return course_nleaf(np) * (SCCAV_LEAF + 1); }

// Levels of the tree over a course of np points: level 0 = the leaves, level h has ceil(nleaf / 2^h) nodes, node j of
// level h = leaves [j 2^h, (j + 1) 2^h) (the last node of a level may be short); the top level is the first with at
// most two nodes -- the root is never tested, a window is never empty.  A node is 24 bytes in two arrays: its chord
// (float4) and (1 / len^2, radius) (float2).
// lev[2h] = id of the first node of level h (ids run through the levels), lev[2h + 1] = number of nodes of level h.
// Returns the number of levels (>= 1); *units (optional) = number of node ids, the dummy node included.  lev may be NULL.
__host__ __device__ inline int course_levels(int np, int* lev, int* units) {
    int cnt = course_nleaf(np), o = 0, h = 0;
    for (;;) {
        if (lev) { lev[2 * h] = o; lev[2 * h + 1] = cnt; }
        o += cnt;
        ++h;
        if (cnt <= 2 || h >= SCCAV_MAX_LEVELS) break;
        cnt = (cnt + 1) >> 1;
    }
    if (units) *units = o + 1;                     // + the dummy node (cover_row)
    return h;
}
// row length of the cover table: a window has at most three cover nodes per level and four at one (cover_node), and the top
// level is never part of a cover; rounded up to the 8 entries one 16-byte load brings
__host__ __device__ __forceinline__ int cover_kc(int nlev) { return ((3 * (nlev - 1) + 1) + 7) & ~7; }

__host__ __device__ inline int course_node_units(int np) {
    int u;
    course_levels(np, nullptr, &u);
    return u;
}

// View of a staged course: padded points + the fp32 capsules of the tree nodes
template <typename T, typename T2> struct CourseIndex {
    const T2* xy;      // [course_nslot(np)] padded points, point i at course_slot(i)
    float4* chord;     // [node id] chord (start x, y, vector z, w)
    float2* ir;        // [node id] (1 / |chord|^2 or 0, radius)
    const int* lev;    // [2 nlev] first unit and node count of every level
    const T* org;      // origin (x, y) of the fp32 frame: a point of the course
    const float* ext;  // [1] inflated max-norm extent of the course around the origin
    const uint8_t* ncover;  // [nleaf] cover_count per window leaf
    const uint16_t* cov;    // [nleaf][kc] the cover of every window leaf: node ids, nearest first,
                            // padded with the id of the DUMMY node (a capsule at infinity: its test never fails)
    int kc, dummy;          // row length of cov (a multiple of 8), id of the dummy node
    int np, nleaf, nlev;
    __host__ __device__ __forceinline__ T2 pt(int i) const { return xy[course_slot(i)]; }
};

// point range [lo, hi) of node j of level h
__host__ __device__ __forceinline__ void node_range(int np, int h, int j, int& lo, int& hi) {
    const int sh = h + SCCAV_LEAF_SHIFT;
    const int64_t l = (int64_t)j << sh, u = (int64_t)(j + 1) << sh;
    lo = (int)(l < np ? l : np);
    hi = (int)(u < np ? u : np);
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

// One capsule test: true = node `id` can NOT be excluded (a NaN anywhere keeps the node: conservative).
__host__ __device__ __forceinline__ bool index_test(const float4* __restrict__ chord, const float2* __restrict__ ir, int id, const IndexQuery& q) {
    const float4 c = chord[id];
    const float2 r = ir[id];
    const float vx = q.qx - c.x, vy = q.qy - c.y;
    const float t = sat01(fmaf(vx, c.z, vy * c.w) * r.x);
    const float ex = fmaf(-t, c.z, vx), ey = fmaf(-t, c.w, vy);
    const float d2 = fmaf(ex, ex, ey * ey);
    const float thr = q.base + r.y;
    return !(d2 > thr * thr);
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

__host__ __device__ __forceinline__ int lowest_bit32(uint32_t m) {
#ifdef __CUDA_ARCH__
    return __ffs((int)m) - 1;
#else
    return __builtin_ctz(m);
#endif
}

__host__ __device__ __forceinline__ int lowest_bit64(uint64_t m) {
#ifdef __CUDA_ARCH__
    return __ffsll((long long)m) - 1;
#else
    return __builtin_ctzll(m);
#endif
}

// The cover of everything outside the window [w, w + 2), in leaf units: at level h (nodes of 2^h leaves)
//   left of the window   a = ((w + 1) >> h) - 1:  node a - 1 if a > 0, and node a - 2 if a is even as well;
//   right of the window  b = (w >> h) + 2:        node b if it exists, and node b + 1 if b is even (and it exists).
// (The boundary of what is covered so far is a multiple of 2^h at level h; one node is taken where that leaves a
// multiple of 2^(h+1), two otherwise -- never none.  So a node of 2^h leaves lies at least 2^h - 1 leaves from the
// window: a node is never large where it is close, which is what makes its capsule test succeed -- canonical
// segment-tree covers put a node of any size right next to the window and were measured to fail 2 - 3 times per query.)
// slot 0, 1 = left, 2, 3 = right.  Returns the node index, or -1 if the slot is empty at this level.
__host__ __device__ __forceinline__ int cover_node(int w, int h, int slot, int cnt) {
    if (slot < 2) {
        const int a = ((w + 1) >> h) - 1;
        if (a <= 0) return -1;
        if (slot == 0) return a - 1;
        return (a & 1) ? -1 : a - 2;
    }
    const int b = (w >> h) + 2;
    if (slot == 2) return b < cnt ? b : -1;
    return (!(b & 1) && b + 1 < cnt) ? b + 1 : -1;
}

// The cover of window leaf w as a row of node ids, nearest (lowest level) first, padded with the dummy id;
// returns the number of cover nodes.  Built once per course (course_stage in kernels.cuh, the host test hook).
__host__ __device__ inline int cover_row(int w, int nlev, const int* lev, int kc, int dummy, uint16_t* row) {
    int c = 0;
    for (int h = 0; h < nlev - 1; ++h)
        for (int s = 0; s < 4; ++s) {
            const int j = cover_node(w, h, s, lev[2 * h + 1]);
            if (j >= 0 && c < kc) row[c++] = (uint16_t)(lev[2 * h] + j);
        }
    for (int k = c; k < kc; ++k) row[k] = (uint16_t)dummy;
    return c;
}
// the dummy node: a capsule so far away that d2 overflows to +inf, which no finite threshold reaches
__host__ __device__ __forceinline__ void dummy_node(float4* chord, float2* ir, int dummy) {
    chord[dummy] = make_float4(3.0e38f, 3.0e38f, 0.f, 0.f);
    ir[dummy] = make_float2(0.f, 0.f);
}
// (level, index) of node `id`
__host__ __device__ inline void unit_node(const int* lev, int nlev, int id, int& h, int& j) {
    h = 0;
    while (h + 1 < nlev && id >= lev[2 * (h + 1)]) ++h;
    j = id - lev[2 * h];
}

// Exact global nearest index (first minimum) of (fx, fy) over the whole course.
// hint: any index (the predicted nearest index; clamped into [0, np)); evals (optional) counts
// distance evaluations + capsule tests for the roofline accounting.
//
// 1. WINDOW: the two leaves [w, w + 2) around the hint are scanned (16 exact distances): the incumbent.
// 2. COVER (cover_node): one or two nodes per level on either side of the window, growing geometrically with their
//    distance from it.  Every lane of a warp runs the same loop over the levels with four independent tests in each;
//    about 18 of them are live per query and (measured on config 2's queries) 0.04 - 0.2 per query fail.
// 3. A cover node that cannot be excluded is OPENED, nearest (lowest level) first: both children are tested; a
//    surviving child is opened in turn (the left one first, one bit per level remembers a surviving right one), a
//    surviving leaf is scanned, which tightens the bound for everything after it.
template <typename T, typename T2>
__host__ __device__ inline int course_nearest(const CourseIndex<T, T2>& ci, T fx, T fy, int hint, int* evals) {
    if (hint < 0) hint = 0;
    if (hint >= ci.np) hint = ci.np - 1;
    const int L = ci.nleaf;
    int w = ((hint + SCCAV_LEAF / 2) >> SCCAV_LEAF_SHIFT) - 1;
    if (w > L - 2) w = L - 2;
    if (w < 0) w = 0;
    T best = (T)INFINITY;
    int ib = ci.np;
    int ne = 2 * SCCAV_LEAF;
    IndexQuery q = index_query<T, T2>(ci, fx, fy);
    index_scan_leaf<T, T2>(ci, w, fx, fy, best, ib);
    if (L > 1) index_scan_leaf<T, T2>(ci, w + 1, fx, fy, best, ib);
    q.base = (index_reach32<T>(best) + q.slack) * 1.000002f;
    // NaN / overflowing query (np.argmin of all-NaN is 0): nothing can be compared, do what the reference does
    if (!(best < (T)INFINITY)) return course_nearest_full<T, T2>(ci.xy, ci.np, fx, fy);
    // ---- cover: the window's row of the cover table, four independent tests per 8-byte load (a padding entry
    // tests the dummy node, which never fails: the warp would run the slot for its other lanes anyway)
    uint64_t fail = 0ull;                   // bit k: cover node k of the row survives
    {
        // (four tests per batch; eight per 16-byte load measured the same, and computing the cover from w at every level
        // instead of reading the table costs 40 more instructions per level for the same time at 128 registers)
        const uint2* __restrict__ row = reinterpret_cast<const uint2*>(ci.cov + (size_t)w * ci.kc);
        const float4* __restrict__ g = ci.chord;
        const float2* __restrict__ gi = ci.ir;
        const int nc = ((int)ci.ncover[w] + 3) >> 2;        // (batches that hold a real node; a warp runs its largest count)
        for (int c = 0; c < nc; ++c) {
            const uint2 e = row[c];
            uint32_t f = 0u;
            f |= index_test(g, gi, (int)(e.x & 0xffffu), q) ? 1u : 0u;
            f |= index_test(g, gi, (int)(e.x >> 16), q) ? 2u : 0u;
            f |= index_test(g, gi, (int)(e.y & 0xffffu), q) ? 4u : 0u;
            f |= index_test(g, gi, (int)(e.y >> 16), q) ? 8u : 0u;
            fail |= (uint64_t)f << (4 * c);
        }
    }
    ne += (int)ci.ncover[w];
    // ---- survivors
    while (fail) {
        const int bit = lowest_bit64(fail);
        fail &= fail - 1ull;
        const int unit = ci.cov[(size_t)w * ci.kc + bit];
        if (unit == ci.dummy) continue;     // (only an overflowing threshold gets here)
        int h, j;
        unit_node(ci.lev, ci.nlev, unit, h, j);
        if (h == 0) {
            const T before = best;
            index_scan_leaf<T, T2>(ci, j, fx, fy, best, ib);
            ne += SCCAV_LEAF;
            if (best < before) q.base = (index_reach32<T>(best) + q.slack) * 1.000002f;
            continue;
        }
        uint32_t pend = 0u;                 // bit k: the right child at level k of the path's ancestor at level k + 1 survives
        for (;;) {
            // open node (h, j), h >= 1
            const int hc = h - 1, c = 2 * j;
            const int id0 = ci.lev[2 * hc] + c;
            const bool has1 = c + 1 < ci.lev[2 * hc + 1];
            bool f0 = index_test(ci.chord, ci.ir, id0, q);
            bool f1 = has1 && index_test(ci.chord, ci.ir, id0 + 1, q);
            ne += 1 + (int)has1;
            if (hc == 0) {
                const T before = best;
                if (f0) { index_scan_leaf<T, T2>(ci, c, fx, fy, best, ib); ne += SCCAV_LEAF; }
                if (f1) { index_scan_leaf<T, T2>(ci, c + 1, fx, fy, best, ib); ne += SCCAV_LEAF; }
                if (best < before) q.base = (index_reach32<T>(best) + q.slack) * 1.000002f;
                f0 = f1 = false;
            }
            if (f0) {
                if (f1) pend |= 1u << hc;
                h = hc; j = c;
            } else if (f1) {
                h = hc; j = c + 1;
            } else {
                if (!pend) break;
                const int k = lowest_bit32(pend);
                pend &= pend - 1u;
                j = ((c >> (k - hc)) | 1);  // the path's ancestor at level k was a left child: its right sibling
                h = k;
            }
        }
    }
    if (evals) *evals += ne;
    return ib;
}

}  // namespace sccav
