/*
 * oracle.c -- CPU restatement (plain C, libm) of the CBF-QP hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may build, load or call this file.  The
 * product (sccav_cbf_b200/, libsccav_cbf.so) never links or loads it.
 *
 * It restates, function by function, the arithmetic of the reference
 * (Safety-Critical-Control-WIRIN/sccav_cbf; citations are file:line under the reference root) in the
 * same operation order as oracle/oracle.py, which is pinned against the reference's golden
 * vectors (tests/test_oracle_golden.py).  Build: `make -C oracle` (gcc -O2 -ffp-contract=off, so
 * no multiply-add is fused; pthreads only parallelise over independent vehicles).
 *
 * Third-party arithmetic that is absent from the reference tree (un-vendored, un-pinned):
 *   cvxopt.solvers.cp  -> exact optimum of the 2-variable QP by working-set enumeration (qp2_exact)
 *   scipy Newton-CG    -> safeguarded Newton with a fixed stopping rule (lane_closest_x)
 * Parity pinning: see the header of oracle/oracle.py.
 *
 * Data layouts are those of include/sccav_cbf.h (only its POD structs / constants are shared).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#include "../include/sccav_cbf.h"

#define ZERO_TOL 1e-3            /* cbf/utils.py:27 */
#define PI_ 3.141592653589793    /* np.pi */
#define QP_FEAS_EPS 1e-12
#define QP_PAR_EPS 1e-12
#define QP_TIE_EPS 1e-9
#define LANE_MAX_IT 50
#define LANE_LS_MAX 30
#define LANE_XTOL 1e-12
#define LANE_DTOL (8 * 0x1p-52)   /* a step may raise D by its rounding noise (oracle.py, lane_closest_x) */

typedef struct { double h, hx, hy, hth, hv, ht; } part_t;

/* cbf/utils.py:93-106 */
static double normalize_angle(double a) {
    /* cbf/utils.py:93-106.  The reference's two while-loops never terminate for |a| >~ 1e16 (a -= 2 pi
     * no longer changes a) or inf; beyond 1e4 (a diverged scenario) whole turns are removed first. */
    if (!(fabs(a) <= 1e4)) a = a - (2.0 * PI_) * rint(a / (2.0 * PI_));
    while (a > PI_) a -= 2.0 * PI_;
    while (a < -PI_) a += 2.0 * PI_;
    return a;
}

/* Ellipse2D.evaluate/dx/dy/dt -- cbf/obstacles.py:183-230,310-317 */
static part_t ellipse_partials(double x, double y, double cx, double cy, double a, double b, double th,
                               double vx, double vy) {
    part_t o;
    double dx = x - cx, dy = y - cy, ct = cos(th), st = sin(th);
    double p = dx * ct + dy * st;
    double q = -dx * st + dy * ct;
    double pa = p / a, qb = q / b;
    o.h = pa * pa + qb * qb - 1;
    o.hx = (2 * ct / (a * a)) * p + (-2 * st / (b * b)) * q;
    o.hy = (2 * st / (a * a)) * p + (2 * ct / (b * b)) * q;
    o.hth = 0.0;
    o.hv = 0.0;
    o.ht = -2 * ((dx / (a * a)) * vx + (dy / (b * b)) * vy);
    return o;
}

/* single_obstacle_CBF1 -- test_scripts/radial_dynamic_obstacles.py:391-405 */
static part_t radial_partials(double x, double y, double v, double cx, double cy, double a, double b,
                              double kv, double vx, double vy) {
    part_t o;
    double da = (x - cx) / a, db = (y - cy) / b;
    o.h = da * da + db * db - 1 - (kv * v / (1 + v));
    o.hx = 2 * (x - cx) / (a * a);
    o.hy = 2 * (y - cy) / (b * b);
    o.hth = 0.0;
    o.hv = -kv / ((1 + v) * (1 + v));
    o.ht = -2 * (((x - cx) / (a * a)) * vx + ((y - cy) / (b * b)) * vy);
    return o;
}

/* D_CBF -- test_scripts/stanley_controller_ellipse.py:251-255 */
static part_t distance_partials(double x, double y, double cx, double cy, double Ds) {
    part_t o;
    o.h = sqrt((x - cx) * (x - cx) + (y - cy) * (y - cy)) - Ds;
    o.hx = 2 * (x - cx) / (o.h + Ds);
    o.hy = 2 * (y - cy) / (o.h + Ds);
    o.hth = o.hv = o.ht = 0.0;
    return o;
}

/* CollisionCone2D.update/evaluate/dx/dy/dv/dtheta/dt -- cbf/obstacles.py:468-502,401-458 */
static part_t cone_partials(double x, double y, double th, double v, double cx, double cy, double tho,
                            double vo, double a, double beta) {
    part_t o;
    double s_vx = v * cos(th), s_vy = v * sin(th);
    double o_vx = vo * cos(tho + beta), o_vy = vo * sin(tho + beta);
    double prx = x - cx, pry = y - cy;
    double vrx = s_vx - o_vx, vry = s_vy - o_vy;
    double dist = sqrt(prx * prx + pry * pry);
    double vrn = sqrt(vrx * vrx + vry * vry);
    double cb, cos_phi;
    if (fabs(dist) > fabs(a)) cb = sqrt(dist * dist - a * a) + ZERO_TOL; else cb = ZERO_TOL;
    if (dist > ZERO_TOL) cos_phi = cb / dist; else cos_phi = 0.0;
    o.h = (prx * vrx + pry * vry) + (dist * vrn * cos_phi);
    o.hx = (s_vx - o_vx) + vrn * (x - cx) / (cb + ZERO_TOL);
    o.hy = (s_vy - o_vy) + vrn * (y - cy) / (cb + ZERO_TOL);
    double cbt = cos(th + beta), sbt = sin(th + beta);
    o.hv = ((x - cx) * cbt + (y - cy) * sbt) + ((s_vx - o_vx) * cbt + (s_vy - o_vy) * sbt) * cb / (vrn + ZERO_TOL);
    o.hth = (-(x - cx) * s_vy + (y - cy) * s_vx) + (-(s_vx - o_vx) * s_vy + (s_vy - o_vy) * s_vx) * cb / (vrn + ZERO_TOL);
    o.ht = (-(s_vx - o_vx) * o_vx - (s_vy - o_vy) * o_vy) + (-vrn * ((x - cx) * o_vx + (y - cy) * o_vy) / (cb + ZERO_TOL));
    return o;
}

/* Horner value / derivatives -- numpy Polynomial order, cbf/obstacles.py:589-592 */
static void poly3(const double* c, double x, double* g, double* dg, double* ddg) {
    double v = 0.0, d1 = 0.0, d2 = 0.0;
    for (int i = 5; i >= 0; --i) v = c[i] + v * x;
    for (int i = 5; i >= 1; --i) d1 = (i * c[i]) + d1 * x;
    for (int i = 5; i >= 2; --i) d2 = ((i - 1) * (i * c[i])) + d2 * x;
    *g = v; *dg = d1; *ddg = d2;
}

/* PolyLane.get_shortest_distance_x -- cbf/obstacles.py:641-679 (see oracle.py:lane_closest_x) */
static double lane_closest_x(const double* c, double px, double py) {
    double x = px;
    for (int it = 0; it < LANE_MAX_IT; ++it) {
        double g, dg, ddg;
        poly3(c, x, &g, &dg, &ddg);
        double ex = x - px, ey = g - py;
        double grad = ex + ey * dg;
        double hess = (1 + dg * dg) + ey * ddg;
        double step = (hess > 0) ? -grad / hess : -grad;
        double D0 = ex * ex + ey * ey;
        double Dacc = D0 + LANE_DTOL * D0;
        double t = 1.0, xn = x;
        int ok = 0;
        for (int ls = 0; ls < LANE_LS_MAX; ++ls) {
            double gn, u1, u2;
            xn = x + t * step;
            poly3(c, xn, &gn, &u1, &u2);
            double Dn = (xn - px) * (xn - px) + (gn - py) * (gn - py);
            if (Dn <= Dacc) { ok = 1; break; }
            t = t * 0.5;
        }
        if (!ok) break;
        double dxn = fabs(xn - x);
        double lim = LANE_XTOL * (1 + fabs(x));
        x = xn;
        if (dxn <= lim) break;
    }
    return x;
}

/* PolyLane.update/evaluate/dx/dy -- cbf/obstacles.py:620-636,607-612,681-689 */
static part_t lane_partials(double x, double y, const double* c, double buffer, int sqrt_form) {
    part_t o;
    double cx = lane_closest_x(c, x, y), g, dg, ddg;
    poly3(c, cx, &g, &dg, &ddg);
    double eta = 1 + dg * ddg + dg * dg - y * ddg;
    if (fabs(eta) < ZERO_TOL) eta = ZERO_TOL;
    o.h = (cx - x) * (cx - x) + (g - y) * (g - y) - buffer;
    o.hx = (2 / eta) * ((x - cx) * (eta - 1) - (y - g) * dg);
    o.hy = (2 / eta) * (-(x - cx) * dg + (y - g) * (eta - dg * dg));
    o.hth = o.hv = o.ht = 0.0;
    if (sqrt_form) {   /* CBF_lane_sqrt -- test_scripts/stanley_controller_ellipse.py:489-492 */
        o.h = sqrt((cx - x) * (cx - x) + (g - y) * (g - y)) - buffer;
        o.hx = o.hx / (2 * (o.h + buffer));
        o.hy = o.hy / (2 * (o.h + buffer));
    }
    return o;
}

/* ELLIPSE_PREP (include/sccav_cbf.h): an Ellipse2D whose vehicle-independent terms were evaluated at
 * ingest (prepare_obstacles below) -- same h, grad h, h_t as ellipse_partials up to a few ulp */
static part_t ellipse_prep_partials(double x, double y, double cx, double cy, double m00, double m01, double m10,
                                    double m11, double wx, double wy) {
    part_t o;
    double dx = x - cx, dy = y - cy;
    double pa = m00 * dx + m01 * dy;
    double qb = m10 * dx + m11 * dy;
    o.h = (pa * pa + qb * qb) - 1;
    o.hx = 2 * (m00 * pa + m10 * qb);
    o.hy = 2 * (m01 * pa + m11 * qb);
    o.hth = 0.0;
    o.hv = 0.0;
    o.ht = -2 * (dx * wx + dy * wy);
    return o;
}

static part_t slot_partials(int desc, const double* f, int64_t fs, double x, double y, double th, double v) {
    const int type = desc & SCCAV_SLOT_TYPE_MASK;
    const int is_static = (desc & SCCAV_SLOT_STATIC) != 0;      /* velocity fields are not read */
    switch (type) {
        case SCCAV_SLOT_ELLIPSE: return ellipse_partials(x, y, f[0], f[fs], f[2 * fs], f[3 * fs], f[4 * fs],
                                                         is_static ? 0.0 : f[5 * fs], is_static ? 0.0 : f[6 * fs]);
        case SCCAV_SLOT_ELLIPSE_PREP: return ellipse_prep_partials(x, y, f[0], f[fs], f[2 * fs], f[3 * fs], f[4 * fs], f[5 * fs],
                                                                   is_static ? 0.0 : f[6 * fs], is_static ? 0.0 : f[7 * fs]);
        case SCCAV_SLOT_CONE: return cone_partials(x, y, th, v, f[0], f[fs], f[2 * fs], f[3 * fs], f[4 * fs], f[5 * fs]);
        case SCCAV_SLOT_LANE:
        case SCCAV_SLOT_LANE_SQRT: {
            double c[6] = {f[fs], f[2 * fs], f[3 * fs], f[4 * fs], f[5 * fs], f[6 * fs]};
            return lane_partials(x, y, c, f[0], type == SCCAV_SLOT_LANE_SQRT);
        }
        case SCCAV_SLOT_RADIAL: return radial_partials(x, y, v, f[0], f[fs], f[2 * fs], f[3 * fs], f[4 * fs], f[5 * fs], f[6 * fs]);
        default: return distance_partials(x, y, f[0], f[fs], f[2 * fs]);
    }
}

/* DBM_CBF_2DS gc/fc + F (cbf/cbf.py:159-164,200-207); KBM_VC_CBF2D F (cbf/cbf.py:94-101) */
static void make_row(int model, const part_t* p, double th, double v, double alpha, double lr,
                     double* A0, double* A1, double* b) {
    double c = cos(th), s = sin(th);
    if (model == SCCAV_MODEL_KBM) {
        *A0 = p->hx * c + p->hy * s;
        *A1 = p->hth;
        *b = -(alpha * p->h);
    } else if (model == SCCAV_MODEL_DUM) {                                         /* cbf/cbf.py:237-245,277-286 */
        *A0 = p->hv;
        *A1 = p->hth;
        double Lf = p->hx * (v * c) + p->hy * (v * s);
        *b = -((Lf + alpha * p->h) + p->ht);
    } else {
        *A0 = p->hv;
        *A1 = (p->hx * (-v * s) + p->hy * (v * c)) + p->hth * (v / lr);
        double Lf = p->hx * (v * c) + p->hy * (v * s);
        *b = -((Lf + alpha * p->h) + p->ht);
    }
}

/* ---- exact 2-variable QP (what cvxopt.solvers.cp approximates at cbf/cbf.py:213) ------------ */
static int qp_check(int m, const double* A0, const double* A1, const double* b, double u0, double u1,
                    int skip_a, int skip_b, double* worst) {
    int feas = 1;
    double w = -INFINITY;
    for (int k = 0; k < m; ++k) {
        double rk = (A0[k] * u0 + A1[k] * u1) - b[k];
        if (-rk > w) w = -rk;
        if (k == skip_a || k == skip_b) continue;
        double tol = QP_FEAS_EPS * (fabs(A0[k] * u0) + fabs(A1[k] * u1) + fabs(b[k]));
        if (!(rk >= -tol)) feas = 0;
    }
    *worst = w;
    return feas;
}

static int qp2_exact(int m, const double* A0, const double* A1, const double* b, double r0, double r1,
                     const double* R, double* u0o, double* u1o, uint32_t* masko) {
    double worst;
    if (qp_check(m, A0, A1, b, r0, r1, -1, -1, &worst)) { *u0o = r0; *u1o = r1; *masko = 0; return SCCAV_STATUS_INACTIVE; }
    double fbw = worst, fb0 = r0, fb1 = r1;
    uint32_t fbm = 0;
    double det = R[0] * R[3] - R[1] * R[2];
    double Ri00 = R[3] / det, Ri01 = -R[1] / det, Ri10 = -R[2] / det, Ri11 = R[0] / det;
    for (int k = 0; k < m; ++k) {
        double rk = (A0[k] * r0 + A1[k] * r1) - b[k];
        if (!(rk < 0)) continue;
        double g0 = Ri00 * A0[k] + Ri01 * A1[k];
        double g1 = Ri10 * A0[k] + Ri11 * A1[k];
        double den = A0[k] * g0 + A1[k] * g1;
        if (!(den > 0)) continue;
        double t = (-rk) / den;
        double u0 = r0 + g0 * t, u1 = r1 + g1 * t;
        if (qp_check(m, A0, A1, b, u0, u1, k, -1, &worst)) { *u0o = u0; *u1o = u1; *masko = 1u << k; return SCCAV_STATUS_ACTIVE; }
        if (worst < fbw - QP_TIE_EPS * (fabs(worst) + fabs(fbw))) { fbw = worst; fb0 = u0; fb1 = u1; fbm = 1u << k; }
    }
    for (int j = 0; j < m; ++j)
        for (int k = j + 1; k < m; ++k) {
            double t1 = A0[j] * A1[k], t2 = A1[j] * A0[k];
            double det2 = t1 - t2;
            if (!(fabs(det2) > QP_PAR_EPS * (fabs(t1) + fabs(t2)))) continue;
            double u0 = (b[j] * A1[k] - A1[j] * b[k]) / det2;
            double u1 = (A0[j] * b[k] - b[j] * A0[k]) / det2;
            double e0 = u0 - r0, e1 = u1 - r1;
            double w0 = 2 * (R[0] * e0 + R[1] * e1);
            double w1 = 2 * (R[2] * e0 + R[3] * e1);
            double lj = (w0 * A1[k] - A0[k] * w1) / det2;
            double lk = (A0[j] * w1 - w0 * A1[j]) / det2;
            int feas = qp_check(m, A0, A1, b, u0, u1, j, k, &worst);
            if (feas && lj >= 0 && lk >= 0) { *u0o = u0; *u1o = u1; *masko = (1u << j) | (1u << k); return SCCAV_STATUS_ACTIVE; }
            if (worst < fbw - QP_TIE_EPS * (fabs(worst) + fabs(fbw))) { fbw = worst; fb0 = u0; fb1 = u1; fbm = (1u << j) | (1u << k); }
        }
    *u0o = fb0; *u1o = fb1; *masko = fbm;
    return SCCAV_STATUS_INFEASIBLE;
}

/* one solve_cbf for vehicle n (cbf/cbf.py:166-220 / :67-110) */
static int filter_vehicle(const sccav_params* p, const uint8_t* sd, int M, int64_t N, int64_t n, const double* obst,
                          double x, double y, double th, double v, double alpha, const double* R,
                          double ur0, double ur1, double* u0, double* u1, uint32_t* mask, double* hmin,
                          double* A0, double* A1, double* b) {
    *hmin = INFINITY;
    for (int m = 0; m < M; ++m) {
        int64_t nn = (sd[m] & SCCAV_SLOT_SHARED) ? 0 : n;
        part_t pt = slot_partials(sd[m], obst + (int64_t)m * SCCAV_NFIELD * N + nn, N, x, y, th, v);
        make_row(p->model, &pt, th, v, alpha, p->lr, &A0[m], &A1[m], &b[m]);
        if (pt.h < *hmin) *hmin = pt.h;
    }
    double r0 = ur0, r1;
    if (p->model == SCCAV_MODEL_DUM) r1 = ur1;                                 /* cbf.py:253 */
    else if (p->model == SCCAV_MODEL_KBM) r1 = ur0 * tan(ur1) / p->L;           /* cbf.py:75 */
    else r1 = atan2(p->lr * tan(ur1), p->lf + p->lr);                          /* cbf.py:175 */
    double q0, q1;
    int st = qp2_exact(M, A0, A1, b, r0, r1, R, &q0, &q1, mask);
    *u0 = q0;
    if (p->model == SCCAV_MODEL_DUM) {
        *u1 = q1;                                                              /* cbf.py:293 */
    } else if (p->model == SCCAV_MODEL_KBM) {
        if (p->kbm_driver_delta) *u1 = atan(q1 * p->L / q0);                   /* sce.py:652 */
        else *u1 = atan2(q1 * p->L, r0);                                       /* cbf.py:109 */
    } else {
        *u1 = atan2((p->lf + p->lr) * tan(q1), p->lr);                         /* cbf.py:216 */
    }
    return st;
}

/* obstacles of vehicle n: its first count[n] slots (sccav_pervehicle.count), all M without it */
static int slot_count(const sccav_pervehicle* pv, int M, int64_t n) {
    if (!pv || !pv->count) return M;
    int c = pv->count[n];
    return c < 0 ? 0 : (c > M ? M : c);
}

static void weights(const sccav_params* p, const sccav_pervehicle* pv, int64_t N, int64_t n, double* alpha, double* R) {
    *alpha = (pv && pv->alpha) ? ((const double*)pv->alpha)[n] : p->alpha;
    if (pv && pv->R) { const double* r = (const double*)pv->R; R[0] = r[n]; R[1] = r[N + n]; R[2] = r[2 * N + n]; R[3] = r[3 * N + n]; }
    else { R[0] = p->R[0]; R[1] = p->R[1]; R[2] = p->R[2]; R[3] = p->R[3]; }
}

/* ---- parallel-for over independent vehicles (pthreads; libgomp is not in the image) ---------- */
typedef void (*body_fn)(void* ctx, int64_t n);
typedef struct { body_fn fn; void* ctx; int64_t N, chunk; int64_t next; } pf_t;

static void* pf_worker(void* arg) {
    pf_t* pf = (pf_t*)arg;
    for (;;) {
        int64_t lo = __atomic_fetch_add(&pf->next, pf->chunk, __ATOMIC_RELAXED);
        if (lo >= pf->N) break;
        int64_t hi = lo + pf->chunk < pf->N ? lo + pf->chunk : pf->N;
        for (int64_t n = lo; n < hi; ++n) pf->fn(pf->ctx, n);
    }
    return NULL;
}

int orc_num_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

static void parallel_for(int64_t N, int nthreads, int64_t chunk, body_fn fn, void* ctx) {
    if (nthreads <= 0) nthreads = orc_num_threads();
    if (nthreads > 256) nthreads = 256;
    if ((int64_t)nthreads > N) nthreads = (int)(N > 0 ? N : 1);
    pf_t pf = {fn, ctx, N, chunk, 0};
    if (nthreads == 1) { pf_worker(&pf); return; }
    pthread_t th[256];
    int started = 0;
    for (int i = 0; i < nthreads - 1; ++i)
        if (pthread_create(&th[started], NULL, pf_worker, &pf) == 0) ++started;
    pf_worker(&pf);
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
}

/* Batched solve_cbf; optional row outputs A_out [2][M][N], b_out [M][N]. */
typedef struct {
    const sccav_params* p; const uint8_t* sd; int32_t M; int64_t N; const double* state; const double* obst;
    const double* u_ref; const sccav_pervehicle* pv; double* u_out; uint32_t* mask_out; uint8_t* status_out;
    double* hmin_out; double* A_out; double* b_out;
} fctx_t;

static void filter_body(void* vctx, int64_t n) {
    fctx_t* c = (fctx_t*)vctx;
    const int64_t N = c->N;
    const int M = c->M;
    double A0[SCCAV_MAX_ROWS], A1[SCCAV_MAX_ROWS], b[SCCAV_MAX_ROWS], alpha, R[4], u0, u1, hmin;
    uint32_t mask;
    weights(c->p, c->pv, N, n, &alpha, R);
    const int Mv = slot_count(c->pv, M, n);
    int st = SCCAV_STATUS_INACTIVE;
    for (int m = Mv; m < M; ++m) { A0[m] = 0.0; A1[m] = 0.0; b[m] = -INFINITY; }      /* empty slots: vacuous rows */
    if (Mv > 0)
        st = filter_vehicle(c->p, c->sd, Mv, N, n, c->obst, c->state[n], c->state[N + n], c->state[2 * N + n],
                            c->state[3 * N + n], alpha, R, c->u_ref[n], c->u_ref[N + n], &u0, &u1, &mask, &hmin, A0, A1, b);
    else { u0 = c->u_ref[n]; u1 = c->u_ref[N + n]; mask = 0; hmin = INFINITY; }      /* carla_ml.py:935-936 */
    c->u_out[n] = u0; c->u_out[N + n] = u1;
    if (c->mask_out) c->mask_out[n] = mask;
    if (c->status_out) c->status_out[n] = (uint8_t)st;
    if (c->hmin_out) c->hmin_out[n] = hmin;
    for (int m = 0; m < M; ++m) {
        if (c->A_out) { c->A_out[(int64_t)m * N + n] = A0[m]; c->A_out[((int64_t)M + m) * N + n] = A1[m]; }
        if (c->b_out) c->b_out[(int64_t)m * N + n] = b[m];
    }
}

int orc_filter_step(const sccav_params* p, const uint8_t* sd, int32_t M, int64_t N, const double* state,
                    const double* obst, const double* u_ref, const sccav_pervehicle* pv, double* u_out,
                    uint32_t* mask_out, uint8_t* status_out, double* hmin_out, double* A_out, double* b_out,
                    int nthreads) {
    if (M < 1 || M > SCCAV_MAX_ROWS) return SCCAV_EINVAL;
    fctx_t c = {p, sd, M, N, state, obst, u_ref, pv, u_out, mask_out, status_out, hmin_out, A_out, b_out};
    parallel_for(N, nthreads, 256, filter_body, &c);
    return SCCAV_OK;
}

/* calc_target_index -- stanley_controller_ellipse.py:188-212: np.hypot + first-minimum argmin */
static int calc_target_index(double x, double y, double yaw, const double* cx, const double* cy, int P, double L, double* e) {
    double fx = x + L * cos(yaw), fy = y + L * sin(yaw);
    double best = INFINITY;
    int idx = 0;
    for (int i = 0; i < P; ++i) {
        double d = hypot(fx - cx[i], fy - cy[i]);
        if (d < best) { best = d; idx = i; }
    }
    double f0 = -cos(yaw + PI_ / 2), f1 = -sin(yaw + PI_ / 2);
    *e = (fx - cx[idx]) * f0 + (fy - cy[idx]) * f1;
    return idx;
}

/* RadialObstacleSpawner.update_seekers -- radial_dynamic_obstacles.py:193-239 */
static void seeker_update(double* f, int64_t fs, double ex, double ey, double dt, double k, double vmin) {
    double cx = f[0], cy = f[fs];
    double yaw = atan2(ey - cy, ex - cx);
    double vmag = k * hypot(ex - cx, ey - cy);
    if (vmag < vmin) vmag = vmin;
    double vx = vmag * cos(yaw), vy = vmag * sin(yaw);
    f[5 * fs] = vx; f[6 * fs] = vy;
    f[0] = cx + vx * dt; f[fs] = cy + vy * dt;
}

/* Closed loop for all vehicles (stanley_controller_ellipse.py:630-830, radial_dynamic_obstacles.py:427-507).
 * Same argument meaning as sccav_rollout_host_f64; all pointers are host memory. */
typedef struct {
    const sccav_params* p; const uint8_t* sd; int32_t M; int64_t N; int32_t T; const double* state; double* obst;
    const double* cx; const double* cy; const double* cyaw; int32_t P; const sccav_pervehicle* pv;
    const sccav_rollout_out* out;
} rctx_t;

static void rollout_body(void* vctx, int64_t n) {
    rctx_t* c_ = (rctx_t*)vctx;
    const sccav_params* p = c_->p; const uint8_t* sd = c_->sd; const int32_t M = c_->M; const int64_t N = c_->N;
    const int32_t T = c_->T; const double* state = c_->state; double* obst = c_->obst;
    const double* cx = c_->cx; const double* cy = c_->cy; const double* cyaw = c_->cyaw; const int32_t P = c_->P;
    const sccav_pervehicle* pv = c_->pv; const sccav_rollout_out* out = c_->out;
    const int stan = p->nominal == SCCAV_NOMINAL_STANLEY;
    {
        double x = state[n], y = state[N + n], yaw = state[2 * N + n], v = state[3 * N + n];
        double alpha, R[4];
        weights(p, pv, N, n, &alpha, R);
        double tspeed = (pv && pv->target_speed) ? ((const double*)pv->target_speed)[n] : p->target_speed;
        const int Mv = slot_count(pv, M, n);
        int last_idx = P - 1, target_idx = 0, steps = 0, nact = 0, ninf = 0;
        double time = 0.0, e;
        double hmin_all = INFINITY, bmin = INFINITY, bmax = -INFINITY, bint = 0.0;
        if (stan) target_idx = calc_target_index(x, y, yaw, cx, cy, P, p->L, &e);            /* sce.py:605 */
        while (steps < T) {
            if (p->terminate && !(p->t_max >= time && last_idx > target_idx)) break;           /* sce.py:630 */
            double ur0, ur1;
            if (stan) {
                double a_ref = p->Kp * (tspeed - v);                                            /* sce.py:135-143 */
                int idx = calc_target_index(x, y, yaw, cx, cy, P, p->L, &e);                    /* sce.py:146-169 */
                if (target_idx >= idx) idx = target_idx;
                double theta_e = normalize_angle(cyaw[idx] - yaw);
                double theta_d = atan2(p->k_stanley * e, v + p->ks_stanley);
                target_idx = idx;
                ur0 = (p->model == SCCAV_MODEL_KBM) ? tspeed : a_ref;                           /* sce.py:646-648 */
                ur1 = theta_e + theta_d;
            } else { ur0 = p->uref0; ur1 = p->uref1; }
            double u0 = ur0, u1 = ur1, hmin = INFINITY;
            uint32_t mask = 0;
            int status = SCCAV_STATUS_INACTIVE;
            if (Mv > 0 && p->model != SCCAV_MODEL_NONE) {
                double A0[SCCAV_MAX_ROWS], A1[SCCAV_MAX_ROWS], b[SCCAV_MAX_ROWS];
                status = filter_vehicle(p, sd, Mv, N, n, obst, x, y, yaw, v, alpha, R, ur0, ur1, &u0, &u1, &mask, &hmin, A0, A1, b);
            }
            double px = x, py = y, pyaw = yaw, pv_ = v, beta = 0.0;
            double delta = u1;
            if (delta < -p->max_steer) delta = -p->max_steer;
            if (delta > p->max_steer) delta = p->max_steer;
            if (p->model == SCCAV_MODEL_DBM) {                                                  /* update_com, sce.py:122-131 */
                beta = atan2(p->lr * tan(delta), p->lf + p->lr);
                double c = cos(yaw), s = sin(yaw);
                x += (v * c - v * s * beta) * p->dt;
                y += (v * s + v * c * beta) * p->dt;
                yaw += (v * beta / p->lr) * p->dt;
                v += u0 * p->dt;
            } else {                                                                            /* update / update_by_vel, sce.py:86-120 */
                x += v * cos(yaw) * p->dt;
                y += v * sin(pyaw) * p->dt;
                yaw += v / p->L * tan(delta) * p->dt;
                yaw = normalize_angle(yaw);
                if (p->model == SCCAV_MODEL_KBM) v = u0; else v += u0 * p->dt;
            }
            if (p->seeker)
                for (int m = 0; m < Mv; ++m)
                    if ((sd[m] & SCCAV_SLOT_TYPE_MASK) == SCCAV_SLOT_RADIAL && !(sd[m] & SCCAV_SLOT_SHARED))
                        seeker_update(obst + (int64_t)m * SCCAV_NFIELD * N + n, N, x, y, p->dt, p->seeker_k, p->seeker_vmin);
            if (p->record_stride > 0 && (steps % p->record_stride) == 0) {
                int64_t rec = steps / p->record_stride;
                if (out->traj) {
                    double* tr = (double*)out->traj + rec * SCCAV_TRAJ_FIELDS * N + n;
                    tr[0] = px; tr[N] = py; tr[2 * N] = pyaw; tr[3 * N] = pv_; tr[4 * N] = u0; tr[5 * N] = u1; tr[6 * N] = beta;
                }
                if (out->traj_idx) out->traj_idx[rec * N + n] = target_idx;
                if (out->traj_mask) out->traj_mask[rec * N + n] = mask;
            }
            time += p->dt;                                                                      /* sce.py:830 */
            ++steps;
            nact += (mask != 0);
            ninf += (status == SCCAV_STATUS_INFEASIBLE);
            if (hmin < hmin_all) hmin_all = hmin;
            if (beta < bmin) bmin = beta;
            if (beta > bmax) bmax = beta;
            bint += beta * p->dt;
        }
        double* os = (double*)out->state;
        os[n] = x; os[N + n] = y; os[2 * N + n] = yaw; os[3 * N + n] = v;
        if (out->steps) out->steps[n] = steps;
        if (out->target_idx) out->target_idx[n] = target_idx;
        if (out->n_active) out->n_active[n] = nact;
        if (out->n_infeasible) out->n_infeasible[n] = ninf;
        if (out->h_min) ((double*)out->h_min)[n] = hmin_all;
        if (out->beta_min) ((double*)out->beta_min)[n] = bmin;
        if (out->beta_max) ((double*)out->beta_max)[n] = bmax;
        if (out->beta_int) ((double*)out->beta_int)[n] = bint;
    }
}

int orc_rollout(const sccav_params* p, const uint8_t* sd, int32_t M, int64_t N, int32_t T, const double* state,
                double* obst, const double* cx, const double* cy, const double* cyaw, int32_t P,
                const sccav_pervehicle* pv, const sccav_rollout_out* out, int nthreads) {
    if (M < 0 || M > SCCAV_MAX_ROWS) return SCCAV_EINVAL;
    rctx_t c = {p, sd, M, N, T, state, obst, cx, cy, cyaw, P, pv, out};
    parallel_for(N, nthreads, 4, rollout_body, &c);
    return SCCAV_OK;
}
