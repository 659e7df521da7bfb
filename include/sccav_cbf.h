/*
 * sccav_cbf.h -- C-ABI of the B200-native batched CBF-QP safety filter (libsccav_cbf.so).
 *
 * This is the drop-in boundary for ONE hot path of Safety-Critical-Control-WIRIN/sccav_cbf:
 *     barrier evaluation -> 2-variable CBF-QP -> closed-loop Stanley / bicycle rollout.
 * The reference has no FFI of its own (pure Python); its boundary is the class API
 *     DBM_CBF_2DS.solve_cbf            cbf/cbf.py:166-220
 *     KBM_VC_CBF2D.solve_cbf           cbf/cbf.py:67-110
 *     DUM_CBF_2DS / SADBM_CBF_2DS.solve_cbf   cbf/cbf.py:247-298, 348-437
 *     ObstacleList2D.f/dx/dy/dtheta/dv/dt   cbf/obstacles.py:879-925
 *     the per-tick loop of             test_scripts/stanley_controller_ellipse.py:630-830
 *                                      test_scripts/radial_dynamic_obstacles.py:427-507
 * and, either side of the path: ObstacleList2D.update_by_bounding_box (obstacles.py:833-858), calc_spline_course
 * (cubic_spline_planner.py:178-190), PolyLane.fit_polynomial_curve (obstacles.py:715-773), the actuator block of
 * carla_scripts/multi_obstacle_CBF_local_with_lanes.py:955-980.
 * Each entry point below names the reference interface it replaces.  INTEGRATION.md shows the
 * ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++ / torch types; no exceptions cross the ABI.
 *  - every function returns 0 on success, a negative SCCAV_E* code otherwise;
 *    sccav_last_error() returns a thread-local message for the last failure.
 *  - "dev" pointers are CUDA device pointers on the CURRENT device, BORROWED for the duration of
 *    the call (stream-ordered: the call only enqueues work on `stream`, a cudaStream_t passed as
 *    void*; NULL = the legacy default stream).  "host" pointers are ordinary host memory.
 *  - no internal threads, no global mutable state besides a per-device attribute cache;
 *    re-entrant per stream.
 *  - batch layout is structure-of-arrays with the vehicle index n fastest:
 *        state  [4][N]        x, y, theta(yaw), v
 *        obst   [M][8][N]     slot m, field f, vehicle n  (field meaning depends on slot type)
 *        u      [2][N]        DBM: (a, delta)   KBM: (v, delta)
 *        A      [2][M][N], b [M][N]   constraint rows  A0*u0 + A1*u1 >= b  (u1 = beta / omega)
 *  - `_f64` entry points take double arrays, `_f32` float arrays; integer outputs are the same.
 */
#ifndef SCCAV_CBF_H
#define SCCAV_CBF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCCAV_VERSION 100          /* 0.1.0 */
#define SCCAV_NFIELD 8             /* fields per obstacle slot */
#define SCCAV_MAX_ROWS 32          /* M <= 32: the active set is a uint32 bit mask */
#define SCCAV_TRAJ_FIELDS 7        /* x, y, yaw, v, u0, u1(delta), beta */

/* ---- obstacle slot types (low 6 bits of slot_desc[m]) ------------------------------------ */
/* ELLIPSE  Ellipse2D, cbf/obstacles.py:139-331.  fields: cx, cy, a, b, theta, vx, vy, -
 *          (a, b already include the buffer, obstacles.py:159-160)                            */
#define SCCAV_SLOT_ELLIPSE 0
/* CONE     CollisionCone2D, cbf/obstacles.py:333-543.  fields: cx, cy, theta_o, v_o, a, beta, -, -
 *          (a already includes the buffer, obstacles.py:357)                                  */
#define SCCAV_SLOT_CONE 1
/* LANE     PolyLane, cbf/obstacles.py:545-689.  fields: buffer, c0, c1, c2, c3, c4, c5, -     */
#define SCCAV_SLOT_LANE 2
/* RADIAL   single_obstacle_CBF1, test_scripts/radial_dynamic_obstacles.py:366-425.
 *          fields: cx, cy, a, b, kv, vx, vy, -                                                */
#define SCCAV_SLOT_RADIAL 3
/* DISTANCE D_CBF, test_scripts/stanley_controller_ellipse.py:240-275.  fields: cx, cy, Ds     */
#define SCCAV_SLOT_DISTANCE 4
/* ELLIPSE_PREP  an ELLIPSE slot after sccav_prepare_obstacles_*: the terms that do not depend on
 *          the vehicle are evaluated once at ingest instead of at every solve --
 *          fields: cx, cy, m00 = cos(theta)/a, m01 = sin(theta)/a, m10 = -sin(theta)/b, m11 = cos(theta)/b,
 *          wx = vx/a^2, wy = vy/b^2.   With d = (x - cx, y - cy), (pa, qb) = M d:
 *          h = pa^2 + qb^2 - 1, grad h = 2 M^T (pa, qb), h_t = -2 (dx wx + dy wy)
 *          (the same functions as cbf/obstacles.py:193,218,229,316, a few ulp apart)              */
#define SCCAV_SLOT_ELLIPSE_PREP 5
/* LANE_SQRT  the distance form of the lane barrier, CBF_lane_sqrt / CBF_lane_cf_sqrt,
 *          test_scripts/stanley_controller_ellipse.py:465-512,546-579: h = sqrt(d^2) - buffer, and the gradient of
 *          the LANE barrier divided by 2 (h + buffer).  fields as LANE: buffer, c0, c1, c2, c3, c4, c5, -  */
#define SCCAV_SLOT_LANE_SQRT 6
#define SCCAV_SLOT_TYPE_MASK 0x3f
/* flag: the obstacle does not move -- its velocity fields are not read and h_t = 0            */
#define SCCAV_SLOT_STATIC 0x40
/* flag: the slot's 8 fields are shared by all vehicles (read from n = 0), e.g. global lanes   */
#define SCCAV_SLOT_SHARED 0x80

#define SCCAV_MODEL_DBM 0          /* DBM_CBF_2DS,  u = (a, beta),  cbf/cbf.py:112-220 */
#define SCCAV_MODEL_KBM 1          /* KBM_VC_CBF2D, u = (v, omega), cbf/cbf.py:33-110  */
#define SCCAV_MODEL_NONE 2         /* rollout only: USE_CBF = False, plant State.update (sce.py:86-101,828) */
#define SCCAV_MODEL_SADBM 4        /* SADBM_CBF_2DS, state-augmented steer-rate model, state (x, y, theta, v, beta),
                                      u = (a, d(beta)/dt) inside, (a, delta) in and out, cbf/cbf.py:300-437.  Filter entry
                                      points only; stateful: sccav_pervehicle.aug carries (beta, last beta_ref) from call to
                                      call, sccav_params.sadbm_dt is the FIXED step of the class (its default dt = 0.001;
                                      the reference's wall-clock mode dt = None is measured by the caller and passed here).
                                      Every CONE slot uses the vehicle's beta (cbf.py:424-426), not its own field 5.   */
#define SCCAV_MODEL_DUM 3          /* DUM_CBF_2DS, dynamic unicycle, u = (a, omega) in and out, cbf/cbf.py:222-298
                                      (g_c columns [0,0,0,1], [0,0,1,0]; its fc is taken as the 4-vector it lists,
                                      the reference declares it 5x1 and raises).  Filter entry points only.  */

#define SCCAV_NOMINAL_STANLEY 0    /* Stanley + P speed, stanley_controller_ellipse.py:135-212 */
#define SCCAV_NOMINAL_CONST 1      /* constant u_ref, radial_dynamic_obstacles.py:444          */

/* sccav_params.flags */
/* rollout: ELLIPSE slots are ingested once per launch (sccav_prepare_obstacles_*) and evaluated in
 * their prepared form at every step -- same results to a few ulp, no division or sincos per row   */
#define SCCAV_FLAG_PREPARED_ROWS 1
/* QP: always run the full active-set enumeration ({} -> singles -> pairs, first KKT point).  By default a
 * one-scan shortcut answers the problems whose optimum has one active row (the most violated row in the
 * metric of R) and hands every other problem -- and every near-tie -- to the enumeration; results are
 * bit-identical, the flag exists so that tests can prove it.                                      */
#define SCCAV_FLAG_QP_ENUMERATE 2
/* rollout, model DBM: the plant takes beta = clamp(beta*, +-beta(max_steer)) straight from the QP's beta* instead of
 * beta* -> delta = atan2(L tan beta*, lr) -> clip(delta, +-max_steer) -> beta = atan2(lr tan delta, L)
 * (cbf.py:216, sce.py:122-125).  tan and atan are monotone, so it is the same function -- bit-identical whenever the
 * steering clips, a few ulp apart otherwise (two tan + two atan2 per step are not evaluated); |beta*| >= 1.5 takes the
 * literal sequence.  delta is still produced for the steps that are recorded.  The reference order stays the default. */
#define SCCAV_FLAG_FUSED_STEER 4
/* filter entry points (sccav_filter_step_*), model DBM: the steering component of u_ref and of u is the slip angle
 * beta -- the QP's own coordinate -- instead of the steering angle delta: neither beta_ref = atan2(lr tan delta_ref, L)
 * (cbf.py:175) nor delta = atan2(L tan beta*, lr) (cbf.py:216) is evaluated.  For callers that already hold beta (a plant
 * integrated in beta, like State.update_com, sce.py:122-131): the same QP on the same rows, two tan + two atan2 per
 * solve cheaper.  Rejected (SCCAV_EINVAL) by every other model and by the rollout / drive-tick entry points.          */
#define SCCAV_FLAG_BETA_IO 8
/* rollout with seekers (sccav_params.seeker): the seeker's velocity direction is the normalised offset (dx, dy) / hypot(dx, dy)
 * instead of sincos(atan2(dy, dx)) (radial_dynamic_obstacles.py:193-239): the same unit vector to a few ulp, one atan2 and
 * one sincos per seeker and step cheaper.  With SCCAV_FLAG_PREPARED_ROWS the rollout also evaluates RADIAL rows with the
 * reciprocals 1 / a, 1 / b formed once per launch and v / (1 + v), 1 / (1 + v)^2 once per step (eight divisions per row in
 * the reference's order, rdo.py:391-405).  Opt-in like the other fast forms; the reference order stays the default.       */
#define SCCAV_FLAG_SEEKER_DIRECT 16

#define SCCAV_STATUS_INACTIVE 0    /* u == u_ref                                               */
#define SCCAV_STATUS_ACTIVE 1      /* KKT optimum with 1 or 2 active rows                      */
#define SCCAV_STATUS_INFEASIBLE 2  /* no KKT point; least-violation candidate returned         */

#define SCCAV_OK 0
#define SCCAV_EINVAL (-1)          /* bad argument (M = 0 maps to the reference's ValueError, cbf.py:177) */
#define SCCAV_ECUDA (-2)           /* a CUDA runtime call failed                               */
#define SCCAV_ENOMEM (-3)

/* Parameters of the filter and of the closed loop.  Doubles for both precisions (the _f32 entry
 * points round them once).  Defaults = test_scripts/stanley_controller_ellipse.py:52-58,590.  */
typedef struct sccav_params {
    int32_t model;             /* SCCAV_MODEL_*                                                */
    int32_t nominal;           /* SCCAV_NOMINAL_*                                              */
    int32_t terminate;         /* 1: stop a vehicle when !(t_max >= time && last_idx > target_idx), sce.py:630 */
    int32_t seeker;            /* 1: RADIAL slots chase the ego after every step, rdo.py:193-239 */
    int32_t kbm_driver_delta;  /* KBM only. 0: delta = atan2(w L, v_ref) (cbf.py:109); 1: atan(w L / v) (sce.py:652) */
    int32_t record_stride;     /* rollout: 0 = no trajectory, k = record every k-th step       */
    int32_t flags;             /* SCCAV_FLAG_* (0 = none)                                     */
    int32_t reserved1;
    double alpha;              /* class-K gain (gamma), cbf.py:128                             */
    double lr, lf, L;          /* cbf.py:150-152, cbf.py:61                                    */
    double max_steer;          /* sce.py:58                                                    */
    double dt;                 /* sce.py:54                                                    */
    double k_stanley;          /* sce.py:52                                                    */
    double ks_stanley;         /* softening of LateralStanley, controllers.py:144 (0 = function form) */
    double Kp;                 /* sce.py:53                                                    */
    double target_speed;       /* sce.py:590                                                   */
    double t_max;              /* sce.py:592                                                   */
    double R[4];               /* QP weight, row-major 2x2 SPD, cbf.py:154                     */
    double seeker_k, seeker_vmin;  /* rdo.py:193                                               */
    double uref0, uref1;       /* NOMINAL_CONST reference                                      */
    double sadbm_dt;           /* SADBM only: the class's dt, cbf.py:323,367,419 (0 = its default 0.001) */
} sccav_params;

/* Optional per-vehicle overrides (Monte-Carlo sweeps); any pointer may be NULL.  dev pointers. */
typedef struct sccav_pervehicle {
    const void* alpha;         /* [N]    overrides params.alpha                                */
    const void* R;             /* [4][N] overrides params.R                                    */
    const void* target_speed;  /* [N]    overrides params.target_speed                         */
    const int32_t* count;      /* [N]    obstacles of vehicle n = its first count[n] slots (0..M); the rest
                                  are empty.  The batched form of a per-vehicle ObstacleList2D whose length
                                  varies (cbf/obstacles.py:798-858).  count[n] = 0: u = u_ref, the guard
                                  callers put around solve_cbf (carla_ml.py:935-936)           */
    void* aug;                 /* [2][N]  SADBM only, read AND written: beta (the augmented state, cbf.py:336,419)
                                  and the converted reference of the previous call (beta_ref_last, cbf.py:335,369) */
} sccav_pervehicle;

/* Outputs of a rollout; any pointer except `state` may be NULL.  dev pointers.                */
typedef struct sccav_rollout_out {
    void* state;               /* [4][N] final state (may alias the input state)               */
    int32_t* steps;            /* [N] steps executed                                           */
    int32_t* target_idx;       /* [N] last Stanley target index                                */
    int32_t* n_active;         /* [N] steps with a non-empty active set                        */
    int32_t* n_infeasible;     /* [N] steps with status INFEASIBLE                             */
    void* h_min;               /* [N] min over steps and slots of h                            */
    void* beta_min;            /* [N]                                                          */
    void* beta_max;            /* [N]                                                          */
    void* beta_int;            /* [N] sum beta*dt                                              */
    void* traj;                /* [T_rec][7][N] pre-step x,y,yaw,v + u0,u1,beta of the step; T_rec = ceil(T/stride) */
    int32_t* traj_idx;         /* [T_rec][N] target index used by the step                     */
    uint32_t* traj_mask;       /* [T_rec][N] active set of the step                            */
    int32_t* n_evals;          /* [N] way-point distance evaluations + bounding-circle tests of the
                                  Stanley nearest-index search (roofline accounting)            */
} sccav_rollout_out;

int sccav_version(void);
const char* sccav_last_error(void);
/* 1 if a usable CUDA device is present (no compute is launched). */
int sccav_device_ok(void);

void sccav_default_params(sccav_params* p);

/* K1 -- barrier evaluation + row assembly.
 * Replaces ObstacleList2D.f/dx/dy/dtheta/dv/dt (obstacles.py:879-925) and the row assembly inside
 * DBM_CBF_2DS.solve_cbf F() (cbf.py:194-207) / KBM_VC_CBF2D.solve_cbf F() (cbf.py:94-101).
 * slot_desc: host, M bytes.  h_out (dev [M][N]) may be NULL. */
int sccav_barrier_rows_f64(const sccav_params* p, const uint8_t* slot_desc, int32_t M, int64_t N,
                           const double* state, const double* obst, const sccav_pervehicle* pv,
                           double* A_out, double* b_out, double* h_out, void* stream);
int sccav_barrier_rows_f32(const sccav_params* p, const uint8_t* slot_desc, int32_t M, int64_t N,
                           const float* state, const float* obst, const sccav_pervehicle* pv,
                           float* A_out, float* b_out, float* h_out, void* stream);

/* K0 -- barrier values and partials only: out dev [M][6][N] = h, h_x, h_y, h_theta, h_v, h_t.
 * Replaces the per-obstacle getters f/evaluate, dx, dy, dtheta, dv, dt (obstacles.py:183-236,304-317,
 * 401-458,607-612,681-689) and their stacked ObstacleList2D forms (obstacles.py:879-925). */
int sccav_barrier_partials_f64(const uint8_t* slot_desc, int32_t M, int64_t N, const double* state,
                               const double* obst, double* out, void* stream);
int sccav_barrier_partials_f32(const uint8_t* slot_desc, int32_t M, int64_t N, const float* state,
                               const float* obst, float* out, void* stream);

/* KS -- one Stanley steering call for N vehicles.  Replaces LateralStanley.control
 * (controllers.py:104-151) / stanley_control (stanley_controller_ellipse.py:146-169): exact global
 * nearest way-point (first minimum over all P points), front-axle error, monotone index clamp,
 * delta = normalize(cyaw[idx] - yaw) + atan2(k e, v + ks).  Uses params L (front-axle offset: L in the
 * function form, lf in the class form), k_stanley, ks_stanley.  front: dev [2][N] externally supplied
 * front-axle coordinates (controllers.py:105-110) or NULL.  target_idx: dev [N], in = last target
 * index, out = new one.  err_out (dev [N]) may be NULL. */
int sccav_stanley_control_f64(const sccav_params* p, int64_t N, const double* state, const double* front,
                              const double* course_x, const double* course_y, const double* course_yaw, int32_t P,
                              int32_t* target_idx, double* delta_out, double* err_out, void* stream);
int sccav_stanley_control_f32(const sccav_params* p, int64_t N, const float* state, const float* front,
                              const float* course_x, const float* course_y, const float* course_yaw, int32_t P,
                              int32_t* target_idx, float* delta_out, float* err_out, void* stream);

/* K2 -- the 2-variable QP  min (u-r)^T R (u-r)  s.t.  A u >= b.
 * Replaces cvxopt.solvers.cp(F) at cbf.py:107,213 (exact KKT optimum instead of an IPM iterate).
 * r: dev [2][N] reference already in QP coordinates (beta / omega).
 * warp_per_problem != 0 selects the warp-cooperative kernel (one problem per warp). */
int sccav_qp2_solve_f64(const sccav_params* p, int32_t M, int64_t N, const double* A, const double* b,
                        const double* r, const sccav_pervehicle* pv, double* u_out,
                        uint32_t* active_out, uint8_t* status_out, int32_t warp_per_problem, void* stream);
int sccav_qp2_solve_f32(const sccav_params* p, int32_t M, int64_t N, const float* A, const float* b,
                        const float* r, const sccav_pervehicle* pv, float* u_out,
                        uint32_t* active_out, uint8_t* status_out, int32_t warp_per_problem, void* stream);

/* K1+K2 fused -- one batched solve_cbf(u_ref) call including delta<->beta (cbf.py:166-220) or
 * the KBM conversions (cbf.py:75,109).  u_ref/u_out: dev [2][N] in (a|v, delta).
 * h_min_out (dev [N]) may be NULL. */
int sccav_filter_step_f64(const sccav_params* p, const uint8_t* slot_desc, int32_t M, int64_t N,
                          const double* state, const double* obst, const double* u_ref,
                          const sccav_pervehicle* pv, double* u_out, uint32_t* active_out,
                          uint8_t* status_out, double* h_min_out, void* stream);
int sccav_filter_step_f32(const sccav_params* p, const uint8_t* slot_desc, int32_t M, int64_t N,
                          const float* state, const float* obst, const float* u_ref,
                          const sccav_pervehicle* pv, float* u_out, uint32_t* active_out,
                          uint8_t* status_out, float* h_min_out, void* stream);

/* K3 -- persistent closed-loop rollout: T steps of
 *   nominal (Stanley + P speed | const) -> filter -> plant (update_com | update_by_vel) [-> seekers]
 * Replaces the while-loop of stanley_controller_ellipse.py:630-830 and animate() of
 * radial_dynamic_obstacles.py:427-507.  state: dev [4][N] initial state (read).  obst: dev
 * [M][8][N], READ-WRITE when params.seeker (moving centres are written back).  course_*: dev [P].
 * M may be 0 (no filter: u = u_ref). */
int sccav_rollout_f64(const sccav_params* p, const uint8_t* slot_desc, int32_t M, int64_t N, int32_t T,
                      const double* state, double* obst, const double* course_x, const double* course_y,
                      const double* course_yaw, int32_t P, const sccav_pervehicle* pv,
                      const sccav_rollout_out* out, void* stream);
int sccav_rollout_f32(const sccav_params* p, const uint8_t* slot_desc, int32_t M, int64_t N, int32_t T,
                      const float* state, float* obst, const float* course_x, const float* course_y,
                      const float* course_yaw, int32_t P, const sccav_pervehicle* pv,
                      const sccav_rollout_out* out, void* stream);

/* K3 over SEVERAL roads in one launch (Monte-Carlo over roads: the courses of sccav_spline_course_* go straight in).
 * course_x / course_y / course_yaw: dev [C][P_max] (the layout sccav_spline_course_* writes), course_np: dev [C] int32 =
 * points course c has (clamped to [1, P_max]).  The N vehicles are grouped by road: vehicles [c N/C, (c+1) N/C) drive
 * road c (N must be a multiple of C).  Every road is staged, indexed and driven exactly as a launch of sccav_rollout_*
 * with that road alone would: the results are bit-identical, the C launches and C course builds are not paid.
 * Replaces a loop over calc_spline_course (cubic_spline_planner.py:178-190) + the per-road while-loop of
 * stanley_controller_ellipse.py:630-830.  Stanley nominal control only. */
int sccav_rollout_roads_f64(const sccav_params* p, const uint8_t* slot_desc, int32_t M, int64_t N, int32_t T,
                            const double* state, double* obst, const double* course_x, const double* course_y,
                            const double* course_yaw, int32_t P_max, int32_t C, const int32_t* course_np,
                            const sccav_pervehicle* pv, const sccav_rollout_out* out, void* stream);
int sccav_rollout_roads_f32(const sccav_params* p, const uint8_t* slot_desc, int32_t M, int64_t N, int32_t T,
                            const float* state, float* obst, const float* course_x, const float* course_y,
                            const float* course_yaw, int32_t P_max, int32_t C, const int32_t* course_np,
                            const sccav_pervehicle* pv, const sccav_rollout_out* out, void* stream);

/* Host-buffer variants (what a host-language caller of the reference binds): all array pointers
 * (including those inside pv / out) are HOST memory; the call copies inputs to the current
 * device, runs the kernel, copies results back and synchronises `stream` before returning. */
int sccav_filter_step_host_f64(const sccav_params* p, const uint8_t* slot_desc, int32_t M, int64_t N,
                               const double* state, const double* obst, const double* u_ref,
                               const sccav_pervehicle* pv, double* u_out, uint32_t* active_out,
                               uint8_t* status_out, double* h_min_out, void* stream);
int sccav_filter_step_host_f32(const sccav_params* p, const uint8_t* slot_desc, int32_t M, int64_t N,
                               const float* state, const float* obst, const float* u_ref,
                               const sccav_pervehicle* pv, float* u_out, uint32_t* active_out,
                               uint8_t* status_out, float* h_min_out, void* stream);
int sccav_rollout_host_f64(const sccav_params* p, const uint8_t* slot_desc, int32_t M, int64_t N, int32_t T,
                           const double* state, double* obst, const double* course_x, const double* course_y,
                           const double* course_yaw, int32_t P, const sccav_pervehicle* pv,
                           const sccav_rollout_out* out, void* stream);
int sccav_rollout_host_f32(const sccav_params* p, const uint8_t* slot_desc, int32_t M, int64_t N, int32_t T,
                           const float* state, float* obst, const float* course_x, const float* course_y,
                           const float* course_yaw, int32_t P, const sccav_pervehicle* pv,
                           const sccav_rollout_out* out, void* stream);

/* Pipelined host API for repeated rollouts of one shape (Monte-Carlo batches streamed from the host): a handle owns
 * the device buffers, three streams and the events; the course -- and, optionally, static obstacles -- are uploaded
 * ONCE at creation and stay resident; every submission copies its inputs (HOST pointers, ideally pinned) on a copy
 * stream, runs the rollout on a compute stream and copies the per-vehicle summaries back on a third stream, so that the
 * upload of submission i + 1 and the download of submission i - 1 overlap the kernel of submission i.
 *   create : `depth` (1..8) submissions may be in flight; obst_resident (host, may be NULL) = obstacles shared by every
 *            submission (not with params.seeker); record_stride must be 0 (summaries only).
 *   submit : returns at once with a ticket; obst = NULL uses the resident obstacles; the host buffers of a submission
 *            (inputs AND outputs) must stay untouched until its ticket has been waited for.  With params.seeker the
 *            moved obstacles are written back into `obst`.
 *   wait   : blocks until the outputs of `ticket` (one of the last `depth` submissions) are in host memory.
 * Same results as sccav_rollout_host_* on the same inputs (same kernels).  Not re-entrant per handle. */
int sccav_pipeline_create_f64(const sccav_params* p, const uint8_t* slot_desc, int32_t M, int64_t N, int32_t T,
                              const double* course_x, const double* course_y, const double* course_yaw, int32_t P,
                              const double* obst_resident, int32_t depth, void** handle_out);
int sccav_pipeline_create_f32(const sccav_params* p, const uint8_t* slot_desc, int32_t M, int64_t N, int32_t T,
                              const float* course_x, const float* course_y, const float* course_yaw, int32_t P,
                              const float* obst_resident, int32_t depth, void** handle_out);
int sccav_pipeline_submit_f64(void* handle, const double* state, const double* obst, const sccav_pervehicle* pv,
                              const sccav_rollout_out* out, int64_t* ticket_out);
int sccav_pipeline_submit_f32(void* handle, const float* state, const float* obst, const sccav_pervehicle* pv,
                              const sccav_rollout_out* out, int64_t* ticket_out);
int sccav_pipeline_wait_f64(void* handle, int64_t ticket);
int sccav_pipeline_wait_f32(void* handle, int64_t ticket);
int sccav_pipeline_destroy_f64(void* handle);
int sccav_pipeline_destroy_f32(void* handle);

/* Obstacle ingest for repeated solves (the batched counterpart of constructing Ellipse2D objects,
 * cbf/obstacles.py:146-165, once and then calling solve_cbf every tick): every ELLIPSE slot of
 * obst_in [M][8][N] is rewritten as an ELLIPSE_PREP slot into obst_out [M][8][N] (may alias obst_in),
 * all other slots are copied; slot_desc_out[M] receives the new descriptors (flags preserved).
 * DEVICE pointers, asynchronous on `stream`. */
int sccav_prepare_obstacles_f64(const uint8_t* slot_desc, int32_t M, int64_t N, const double* obst_in, double* obst_out,
                                uint8_t* slot_desc_out, void* stream);
int sccav_prepare_obstacles_f32(const uint8_t* slot_desc, int32_t M, int64_t N, const float* obst_in, float* obst_out,
                                uint8_t* slot_desc_out, void* stream);

/* KB -- batched ObstacleList2D.update_by_bounding_box (cbf/obstacles.py:833-858) for N vehicles that
 * each keep their own obstacle list of up to M entries, fed every tick with up to K bounding boxes
 * (the on-wire obstacle format of the CARLA drivers: actor id, extent, location, yaw, speed --
 * multi_obstacle_CBF_local_with_lanes.py:906-932, obstacle_map.py:144-156).
 *   box_id  [K][N] int32   actor id per box, < 0 = no box (padding); ids are unique per vehicle
 *   box     [K][6][N]      SCCAV_BOX_*: extent.x, extent.y, location.x, location.y, rotation.yaw, velocity
 *   slot_id [M][N] int32   in/out: id held by slot m (the dict key), -1 = empty
 *   obst    [M][8][N]      in/out: the slots, ELLIPSE or CONE layout (obs_type)
 *   count   [N]   int32    in/out: entries of vehicle n = its first count[n] slots (-> sccav_pervehicle.count)
 *   dropped [N]   int32    out (may be NULL): new ids that found no free slot
 * SCCAV_INGEST_UPDATE: an id already held is updated in place by <obstacle>.update_by_bounding_box
 *   (Ellipse2D obstacles.py:294-302: a, b, centre, theta <- extent, location, yaw, buffer NOT re-applied;
 *   CollisionCone2D :512-530: a <- hypot(extent), s_obs <- [x, y, 0, velocity]); ids not in the boxes are
 *   removed; new ids are appended in box order by <obstacle>.from_bounding_box (a + buffer,
 *   :319-331, :532-543).  The resulting order -- surviving entries in their old order, then the new
 *   ones -- is the dict order of the reference, i.e. the constraint index of every row.
 * SCCAV_INGEST_REBUILD: the list is rebuilt from the boxes alone, every entry made by from_bounding_box
 *   (what the CARLA driver does each tick, multi_obstacle_CBF_local_with_lanes.py:918-928).
 * DEVICE pointers, asynchronous on `stream`; M <= 32, K <= 32. */
#define SCCAV_BOX_FIELDS 6
#define SCCAV_INGEST_UPDATE 0
#define SCCAV_INGEST_REBUILD 1
int sccav_ingest_boxes_f64(int32_t obs_type, int32_t mode, double buffer, int32_t M, int32_t K, int64_t N,
                           const int32_t* box_id, const double* box, int32_t* slot_id, double* obst, int32_t* count,
                           int32_t* dropped, void* stream);
int sccav_ingest_boxes_f32(int32_t obs_type, int32_t mode, double buffer, int32_t M, int32_t K, int64_t N,
                           const int32_t* box_id, const float* box, int32_t* slot_id, float* obst, int32_t* count,
                           int32_t* dropped, void* stream);

/* KA -- actuator shaping after the filter, the step that follows solve_cbf in the CARLA drivers
 * (multi_obstacle_CBF_local_with_lanes.py:955-980): u = (a, delta) -> (throttle, brake, steer).
 *   a > 0 : throttle = clamp(tanh(a), 0, 1), raised by at most `rate` per tick above throttle_prev;
 *           brake keeps its previous value (the driver does not reset it, :958-968) unless
 *           SCCAV_ACT_RESET_BRAKE is set, which zeroes it
 *   a <= 0: throttle = 0, brake = clamp(-tanh(a), 0, 1), raised by at most `rate` per tick above brake_prev
 *   steer = delta clamped to [0, max_steer] (delta > 0) or [-max_steer, 0]           (:973-976)
 * throttle_prev / brake_prev [N] are read and then overwritten with the new values (:970-971).
 * DEVICE pointers, asynchronous on `stream`. */
#define SCCAV_ACT_RESET_BRAKE 1
int sccav_actuator_shaping_f64(int64_t N, const double* u, double max_steer, double rate, int32_t flags,
                               double* throttle_prev, double* brake_prev, double* throttle_out, double* brake_out,
                               double* steer_out, void* stream);
int sccav_actuator_shaping_f32(int64_t N, const float* u, double max_steer, double rate, int32_t flags,
                               float* throttle_prev, float* brake_prev, float* throttle_out, float* brake_out,
                               float* steer_out, void* stream);

/* KC -- course generation for C roads at once, the step before the Stanley controller:
 * calc_spline_course (test_scripts/PathPlanning/CubicSpline/cubic_spline_planner.py:178-190) = two natural
 * cubic splines x(s), y(s) over the cumulative chord length (Spline2D :118-175, Spline :12-115; the
 * tridiagonal system of :95-115 solved by the Thomas recurrence instead of np.linalg.solve), sampled at
 * t_j = j ds, j < np = ceil(s_end / ds) (np.arange, :180), yaw = atan2(y', x') (:167-173), curvature (:156-165).
 *   wx, wy  [C][K]      way-points (K >= 2 per course, K <= 64)
 *   cx, cy, cyaw, ck [C][P_max]   samples; ck may be NULL; only the first min(np, P_max) of a course are written
 *   np_out  [C] int32   number of samples course c HAS (compare with P_max)
 * DEVICE pointers, asynchronous on `stream`.  Per-scenario roads: generate C courses, then hand all of them to ONE
 * sccav_rollout_roads_* launch. */
#define SCCAV_MAX_KNOTS 64
int sccav_spline_course_f64(int32_t C, int32_t K, const double* wx, const double* wy, double ds, int32_t P_max,
                            double* cx, double* cy, double* cyaw, double* ck, int32_t* np_out, void* stream);
int sccav_spline_course_f32(int32_t C, int32_t K, const float* wx, const float* wy, double ds, int32_t P_max,
                            float* cx, float* cy, float* cyaw, float* ck, int32_t* np_out, void* stream);

/* KL -- weighted polynomial lane fit for C lanes at once, the step before the lane barrier:
 * PolyLane.fit_polynomial_curve (cbf/obstacles.py:715-773: scipy curve_fit of a polynomial with per-point sigma;
 * fixed points enter as ordinary points with the small sigma `alpha`, :749-756) == PolynomialLaneCurve.lsq_curve
 * (test_scripts/lane_cbf_test.py:108-138) for equal weights.  The model is linear in its coefficients, so the
 * least-squares optimum curve_fit iterates to is THE weighted least-squares solution; it is computed in the
 * centred, scaled abscissa t = (x - mid) / half_range with polynomials orthogonal over the weighted points
 * (Forsythe's three-term recurrence: no normal equations), then expanded back to powers of x.
 *   x, y    [K][C]     points of lane c (the caller appends its fixed points, as the reference does)
 *   sigma   [K][C]     per-point sigma (weight 1 / sigma^2), or NULL = 10 everywhere (obstacles.py:738-739)
 *   count   [C] int32  lane c uses its first count[c] points (NULL = all K); needs count > degree
 *   degree  1..5
 *   coeffs  [6][C]     c0..c5 of y = sum c_i x^i -- the coefficient fields of a LANE slot; unused ones are 0
 *   status  [C] int32  0 = ok, 1 = too few points / singular system (coeffs = NaN); may be NULL
 * DEVICE pointers, asynchronous on `stream`. */
int sccav_fit_lanes_f64(int32_t C, int32_t K, const double* x, const double* y, const double* sigma, const int32_t* count,
                        int32_t degree, double* coeffs, int32_t* status, void* stream);
int sccav_fit_lanes_f32(int32_t C, int32_t K, const float* x, const float* y, const float* sigma, const int32_t* count,
                        int32_t degree, float* coeffs, int32_t* status, void* stream);

/* KD -- the per-tick loop of the CARLA driver for N ego vehicles, T ticks in ONE launch
 * (carla_scripts/multi_obstacle_CBF_local_with_lanes.py:861-983), everything a tick carries over kept in registers:
 *   delta, idx = lateral_stanley.control()      cbf/controllers.py:104-151 -- class form: front axle at params.L (= lf),
 *                                               atan2(k e, v + ks) with params.k_stanley / ks_stanley, its own last target index
 *   delta *= rad_to_steer                       :876-877
 *   acc_pid.set_dt(dt_t); u_a = acc_pid.control(v, trajectory[idx][3])    cbf/controllers.py:153-180 (kp, ki, kd)
 *   obstacle list = the n_fixed leading slots of `obst` as the caller set them (the two PolyLane boundaries, :913-916)
 *                   + one fresh CollisionCone2D(hypot(extent), s, [x, y, yaw, |v|]) per box of the tick (:918-928)
 *   u = solve_cbf([u_a, delta])                 DBM_CBF_2DS, cbf/cbf.py:166-220; an empty list passes u_ref through (:935-936)
 *   throttle / brake / steer                    :955-980 (sccav_actuator_shaping_* semantics)
 * Inputs per tick (time-major, DEVICE pointers): ego [T][4][N] -- the simulator owns the plant; ego = NULL closes the
 * loop with a stand-in plant instead (State.update_com on the filtered (a, delta) with the tick's dt, from state0 [4][N]);
 * box_id [T][K][N] (< 0 = none) / box [T][K][6][N] as for sccav_ingest_boxes_* (K = 0: the slots of `obst` are used as
 * they are, pv->count gives each vehicle's number); dt [T] (NULL: params.dt every tick).  slot_desc[M]: slots
 * [n_fixed, M) must be per-vehicle CONE slots when K > 0.  trajectory = (x, y, yaw, v)[P].
 * Carried between calls (read and written): target_idx [N] (LateralStanley's last target index), carry [4][N] =
 * PID e_prev, PID integral, throttle_prev, brake_prev.  Outputs: act_out [T][3][N] = throttle, brake, steer;
 * optional u_out [T][2][N] filtered (a, delta), active_out [T][N], target_idx_out [T][N], state_out [4][N]. */
typedef struct sccav_drive_params {
    double kp, ki, kd;          /* PID1 gains (the driver: 1.0, 0.01, 0.01, :660)                       */
    double rad_to_steer;        /* convert_rad_to_steer (:876)                                          */
    double max_steer_cmd;       /* clamp of the steer command (the driver: 1, :875)                      */
    double rate;                /* largest rise of throttle / brake per tick (the driver: 0.1, :961,967) */
    double cone_buffer;         /* CollisionCone2D default buffer 1.5 (cbf/obstacles.py:341)             */
    int32_t act_flags;          /* SCCAV_ACT_*                                                          */
    int32_t reserved;
} sccav_drive_params;
int sccav_drive_ticks_f64(const sccav_params* p, const sccav_drive_params* dp, const uint8_t* slot_desc, int32_t M, int32_t n_fixed,
                          int64_t N, int32_t T, const double* state0, const double* ego, int32_t K, const int32_t* box_id,
                          const double* box, const double* dt, double* obst, const double* traj_x, const double* traj_y,
                          const double* traj_yaw, const double* traj_v, int32_t P, const sccav_pervehicle* pv, int32_t* target_idx,
                          double* carry, double* act_out, double* u_out, uint32_t* active_out, int32_t* target_idx_out,
                          double* state_out, void* stream);
int sccav_drive_ticks_f32(const sccav_params* p, const sccav_drive_params* dp, const uint8_t* slot_desc, int32_t M, int32_t n_fixed,
                          int64_t N, int32_t T, const float* state0, const float* ego, int32_t K, const int32_t* box_id,
                          const float* box, const float* dt, float* obst, const float* traj_x, const float* traj_y,
                          const float* traj_yaw, const float* traj_v, int32_t P, const sccav_pervehicle* pv, int32_t* target_idx,
                          float* carry, float* act_out, float* u_out, uint32_t* active_out, int32_t* target_idx_out,
                          float* state_out, void* stream);

/* Measurement helpers used by bench.py (not part of the reference-facing path). */
/* Launches an unrolled FMA chain kernel and returns the achieved TFLOP/s (FMA = 2 flop) on the
 * current device; `dtype` 64 or 32.  Used as the measured FP64/FP32 CUDA-core peak. */
int sccav_measure_fma_peak(int32_t dtype, double* tflops_out);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t sccav_launch_count(void);
/* The library allocates its scratch and the device side of the host-buffer entry points from a stream-ordered memory
 * pool of its own (one per device; freed blocks are kept for the next call up to 1 GiB, SCCAV_POOL_KEEP_MB overrides).
 * The device's default pool is not touched.  This call returns everything the pool holds to the driver. */
int sccav_trim_pool(void);

/* Launch shape the rollout entry point would use for (slot_desc[M], N, P) on the current device:
 * info[8] = grid, block, dynamic smem bytes, registers/thread, max threads/block of the kernel,
 * resident CTAs/SM at that shape, course-in-shared-memory flag, SM count.  Launches nothing. */
int sccav_rollout_launch_info_f64(const uint8_t* slot_desc, int32_t M, int64_t N, int32_t P, int32_t* info);
int sccav_rollout_launch_info_f32(const uint8_t* slot_desc, int32_t M, int64_t N, int32_t P, int32_t* info);

/* TEST HOOK, host-only, launches nothing: runs the rollout kernel's pruned nearest-way-point search
 * (csrc/course_index.cuh, the same __host__ __device__ code) on the CPU for nq query points, next to
 * the exhaustive scan of calc_target_index (stanley_controller_ellipse.py:188-212), so that the
 * CPU test-suite can prove index equality (ties, far points, NaN) without a GPU.  All pointers are
 * HOST memory; hint / idx_full_out / evals_out may be NULL; dtype 64 or 32. */
int sccav_debug_course_index_host(const double* cx, const double* cy, int32_t P, const double* fx, const double* fy,
                                  const int32_t* hint, int64_t nq, int32_t dtype, int32_t* idx_out,
                                  int32_t* idx_full_out, int64_t* evals_out);

/* TEST HOOK, host-only, launches nothing: the cover table of the search for a course of P points (csrc/course_index.cuh,
 * cover_row).  For every window leaf w in [0, nleaf): up to kc cover nodes as (level, index) pairs -- node (h, j) = leaves
 * [j 2^h, (j + 1) 2^h) -- nearest first; count_out[w] = how many.  *nleaf_out, *nlev_out, *kc_out describe the tree; a first
 * call with level_out = NULL returns only those.  level_out / index_out hold nleaf * kc entries each. */
int sccav_debug_cover_host(int32_t P, int32_t* nleaf_out, int32_t* nlev_out, int32_t* kc_out, int32_t* count_out,
                           int32_t* level_out, int32_t* index_out);

#ifdef __cplusplus
}
#endif
#endif /* SCCAV_CBF_H */
