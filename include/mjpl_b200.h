/*
 * mjpl_b200.h -- C ABI of the B200 configuration-validity engine.
 *
 * This is the drop-in boundary for mjpl's validity hot path.  The reference has no FFI of its
 * own (it is pure Python calling the `mujoco` wheel), so each entry point below cites the
 * reference interface it replaces (paths relative to the reference checkout) and
 * INTEGRATION.md shows the ctypes stub a maintainer would add to mjpl.
 *
 * Conventions: plain C, every call returns 0 on success or a non-zero status with a message in
 * mjb_last_error() (thread local).  Device pointers are caller-owned; launches are
 * stream-ordered on `stream` (a cudaStream_t passed as void*) with no hidden synchronisation
 * unless stated.  A handle is bound to the CUDA device that was current at mjb_model_create
 * and, like the reference's CollisionConstraint (which owns one mutable MjData,
 * src/mjpl/constraint/collision_constraint.py:23), it is NOT re-entrant: one call at a time
 * per handle (its scratch buffers are per handle).  There is no CPU fallback: every compute
 * entry point fails with MJB_ERR_CUDA when no device is usable.
 */
#ifndef MJPL_B200_H
#define MJPL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MJB_OK 0
#define MJB_ERR_ARG 1     /* bad argument (ValueError on the Python side) */
#define MJB_ERR_MODEL 2   /* model uses something outside the supported subset */
#define MJB_ERR_CUDA 3    /* CUDA runtime failure / no device */

/* flags of the check entry points */
#define MJB_CHECK_LIMITS 1u     /* JointLimitConstraint.valid_config  (joint_limit_constraint.py:19-20) */
#define MJB_CHECK_COLLISION 2u  /* CollisionConstraint.valid_config   (collision_constraint.py:26-30)  */
#define MJB_NO_OBB_CULL 4u      /* debugging: skip the OBB mid-phase (results must not change) */
#define MJB_NO_FP64_RECHECK 8u  /* debugging: leave uncertain rows marked 2 instead of re-evaluating */
#define MJB_LIMITS_OUTWARD 16u  /* with MJB_CHECK_LIMITS: compare the fp32 row against the limits rounded OUTWARD to
                                   fp32.  For callers whose rows were fp64 before the cast: such a caller decides the
                                   limits itself on the fp64 values (joint_limit_constraint.py:19-20 is an fp64
                                   compare) and ANDs its mask with the result; a row exactly on a limit that fp32
                                   cannot represent is then not lost to the cast, and rows beyond the limits still
                                   skip the collision work */

/*
 * Constant tables taken from the MuJoCo model, under MjModel's own field names so that a real
 * mujoco.MjModel can be passed through field by field.  Everything is copied at
 * mjb_model_create; the caller may free its arrays afterwards.
 * Replaces: the MjModel argument of CollisionConstraint.__init__ / JointLimitConstraint.__init__
 * (collision_constraint.py:10-24, joint_limit_constraint.py:10-17) and the body-name ->
 * sorted-id allow-list built by CollisionRuleset.__init__ (collision_constraint.py:42-64).
 */
typedef struct mjb_model_desc {
  int32_t nq, nbody, njnt, ngeom, nmesh, nmeshvert, nexclude, nallowed;
  int32_t disable_contact, disable_filterparent;   /* opt.disableflags bits */
  const int32_t *body_parentid, *body_weldid, *body_jntadr, *body_jntnum;
  const double *body_pos, *body_quat;               /* (nbody,3) (nbody,4 wxyz) */
  const int32_t *jnt_type, *jnt_qposadr, *jnt_bodyid, *jnt_limited;
  const double *jnt_pos, *jnt_axis, *jnt_range, *qpos0;
  const int32_t *geom_type, *geom_bodyid, *geom_contype, *geom_conaffinity, *geom_dataid;
  const double *geom_size, *geom_pos, *geom_quat, *geom_margin, *geom_gap;
  const int32_t *mesh_vertadr, *mesh_vertnum;       /* convex-hull vertices only */
  const double *mesh_vert;                          /* (nmeshvert,3) */
  const int64_t *exclude_signature;                 /* (b1<<16)+b2 per <exclude> */
  const int32_t *allowed_body_pairs;                /* (nallowed,2) body ids */
  /* optional hull graphs (MjModel.mesh_graphadr / mesh_graph, local ids over the hull vertices
   * above): neighbour lists for hill-climbing support queries; NULL / -1 = scan all vertices */
  const int32_t *mesh_graphadr, *mesh_graph;
  int32_t nmeshgraph;
} mjb_model_desc;

typedef struct mjb_model mjb_model;

typedef struct mjb_stats {
  int64_t rows;            /* configurations submitted since the last reset */
  int64_t narrow_items;    /* (row, pair) items that reached the narrow phase */
  int64_t uncertain_rows;  /* rows re-evaluated by the fp64 kernel */
  int64_t queue_overflow;  /* extend chains clipped because a tree was at capacity (mjb_rrt_extend) */
  int64_t launches;        /* kernels launched by this handle */
} mjb_stats;

const char *mjb_last_error(void);
int mjb_device_count(void);

/* Build device tables on the current CUDA device. */
int mjb_model_create(const mjb_model_desc *desc, mjb_model **out);
void mjb_model_destroy(mjb_model *m);

/* Static geom-pair list after MuJoCo's filters minus allowed body pairs (introspection). */
int32_t mjb_model_npair(const mjb_model *m);
int mjb_model_pairs(const mjb_model *m, int32_t *geom1, int32_t *geom2);

/*
 * valid[i] = 1 iff row i passes the selected checks, else 0.
 * Replaces, for a whole block of rows: obeys_constraints(q, [JointLimitConstraint,
 * CollisionConstraint]) (src/mjpl/constraint/utils.py:6-19), i.e. per row
 * JointLimitConstraint.valid_config and data.qpos=q; mj_kinematics; mj_collision;
 * CollisionRuleset.obeys_ruleset(data.contact.geom) (collision_constraint.py:27-30).
 * d_q: (n, nq) fp32, row stride ldq elements; d_valid: (n,) bytes.
 */
int mjb_check_configs(mjb_model *m, const float *d_q, int64_t n, int32_t ldq, uint8_t *d_valid,
                      uint32_t flags, void *stream);

/* Same, host buffers: H2D copy, kernels, D2H copy, stream synchronised before returning.
 * This is what a Python caller holding numpy arrays uses (the end-to-end path).  Batches of more
 * than 65536 rows are copied in chunks on a second stream while the (single) validity launch is
 * already consuming them; pinned host memory makes those copies asynchronous. */
int mjb_check_configs_host(mjb_model *m, const float *h_q, int64_t n, uint8_t *h_valid,
                           uint32_t flags);

/* mj_kinematics for a block of rows (call site collision_constraint.py:28):
 * d_xpos (n,nbody,3), d_xquat (n,nbody,4 wxyz), fp32.  Parity/debug entry point. */
int mjb_fk(mjb_model *m, const float *d_q, int64_t n, int32_t ldq, float *d_xpos, float *d_xquat,
           void *stream);

/*
 * Signed distance to contact of every row, for band accounting (the north star counts the rows whose
 * reference signed distance lies within 1e-5 of the margin): d_dist[i] = min over the static pair list
 * of (signed distance of the pair - its margin), i.e. the row is in contact iff d_dist[i] <= 0 (SURVEY
 * A.3); d_pair[i] (optional) = index of the arg-min pair in mjb_model_pairs order.  fp64 throughout
 * (rows are read as fp32, the precision the validity kernels see): FK, closed forms for plane and
 * capsule pairs, GJK run to convergence for convex pairs and an expanding-polytope depth when the
 * cores intersect.  Capped: values above far_cap are reported as far_cap with pair -1 (far_cap <= 0:
 * MJB_DIST_FAR_DEFAULT), penetration deeper than MJB_DEPTH_CAP as -MJB_DEPTH_CAP.
 */
#define MJB_DIST_FAR_DEFAULT 0.01
#define MJB_DEPTH_CAP 1e-3
int mjb_min_distance(mjb_model *m, const float *d_q, int64_t n, int32_t ldq, double far_cap, double *d_dist,
                     int32_t *d_pair, void *stream);

/*
 * Edge validation: _valid_collision_interval(start, end, step, constraint)
 * (src/mjpl/planning/utils.py:188-216) for ne edges at once.  Interior waypoints
 * q0 + k*step*(q1-q0)/|q1-q0|, k = 1..K, K = ceil(|q1-q0|/step)-1, are generated on the device;
 * d_valid[e] = 1 iff all of them pass `flags` (the reference checks collisions only);
 * d_first_bad[e] (optional) = smallest failing k-1, or -1.
 */
int mjb_check_edges(mjb_model *m, const float *d_q0, const float *d_q1, int64_t ne, int32_t ldq,
                    float step, uint8_t *d_valid, int32_t *d_first_bad, uint32_t flags,
                    void *stream);

/*
 * Validity sweep with rows generated on the device: row r (global index row0+i), joint j is
 * lo_j + u*(hi_j-lo_j) with u a counter-based hash of (seed, r, j) -- no host->device traffic.
 * mjb_sweep_rows writes the same rows out (n,nq) so a host checker can see them.
 */
int mjb_check_sweep(mjb_model *m, uint64_t seed, int64_t row0, int64_t n, uint8_t *d_valid,
                    uint32_t flags, void *stream);
int mjb_sweep_rows(mjb_model *m, uint64_t seed, int64_t row0, int64_t n, float *d_q, void *stream);

/*
 * Tree.nearest_neighbor (src/mjpl/planning/tree.py:57-66) for n queries at once.  Trees are
 * rows of a padded (ntrees, cap, nq) fp64 array with per-tree node counts; d_rows (optional)
 * selects the tree of each query (NULL: query i uses tree i).  d_out[i] = index of the nearest
 * node (squared Euclidean distance, lowest index wins ties; non-finite nodes never win).
 */
int mjb_nearest_batch(const double *d_nodes, int64_t cap, int32_t nq, const int64_t *d_count,
                      const int64_t *d_rows, const double *d_targets, int64_t n, int64_t *d_out,
                      void *stream);

/*
 * Tree.get_path (src/mjpl/planning/tree.py:68-81) for n trees at once: from node d_first[i] of tree d_rows[i]
 * (NULL: tree i) follow the parent links (d_parent: (ntrees, cap) int64, -1 at the root) to the root.
 * d_steps: (n, max_depth) int64, row i = the node indices first .. root, padded with -1; d_len[i] = their
 * number (a chain longer than max_depth is cut there and reported with d_len[i] = -1).
 */
int mjb_tree_paths(const int64_t *d_parent, int64_t cap, const int64_t *d_rows, const int64_t *d_first, int64_t n,
                   int64_t max_depth, int64_t *d_steps, int64_t *d_len, void *stream);

/*
 * One _constrained_extend (src/mjpl/planning/utils.py:105-164) for n trees at once, for
 * non-projecting constraints, entirely on the device: nearest node, the chain
 * near + k*eps*(target-near)/|target-near| (k = 1..min(ceil(dist/eps), kcap); the step that covers
 * the remaining distance lands on the target itself), validity of every step with `flags`, the
 * reference's stop rules, and the append of the valid prefix to the tree.
 * Trees: d_nodes (ntrees, cap, nq) fp64, d_parent (ntrees, cap), d_count (ntrees); d_slots
 * (optional) = tree of each query.  Out: d_reached (n,nq) = configuration reached (the nearest
 * node itself if no step was valid), d_last (n,) = its node index.  The caller keeps
 * count + kcap <= cap (chains are clipped at the capacity and counted in stats.queue_overflow).
 */
int mjb_rrt_extend(mjb_model *m, double *d_nodes, int64_t *d_parent, int64_t *d_count, int64_t cap,
                   const int64_t *d_slots, const double *d_targets, int64_t n, double eps,
                   int32_t kcap, uint32_t flags, double *d_reached, int64_t *d_last, void *stream);

/* How many chain rows the following mjb_rrt_extend / mjb_rrt_extend_masked calls on this handle are expected to
 * check per call (0: the default, 24 per query).  A planner whose queries have mostly finished tells the library
 * so, and the calls take the small-launch path (results do not depend on the hint, only launch choices do). */
int mjb_set_chain_hint(mjb_model *m, int64_t expected_rows);

/* Same, with a per-query mask: queries with d_active[i] == 0 (or with NaN targets) build no chain and
 * append nothing (their d_reached is the nearest node).  d_active may be NULL (all active). */
int mjb_rrt_extend_masked(mjb_model *m, double *d_nodes, int64_t *d_parent, int64_t *d_count, int64_t cap,
                          const int64_t *d_slots, const double *d_targets, const uint8_t *d_active, int64_t n,
                          double eps, int32_t kcap, uint32_t flags, double *d_reached, int64_t *d_last, void *stream);

/*
 * The rest of one iteration of RRT.plan_to_configs (src/mjpl/planning/rrt.py:195-235) for S queries held
 * in S slots, with all planner state on the device so that whole iterations can be enqueued -- or
 * captured into a CUDA graph and replayed -- without a host round trip:
 *   mjb_rrt_sample  the sampling step (:206-215): with probability goal_bias the other tree's root
 *                   (q_goal, or q_init when the trees are swapped), else q_init with the planning
 *                   joints (d_plan_mask) drawn uniformly in [lo, hi]; counter-based random stream keyed
 *                   by (seed, slot, iteration).  Slots with d_active == 0 get NaN targets.
 *   mjb_rrt_meet    the connection test (:217-229): slots whose two extends reached the same
 *                   configuration record their connecting nodes (d_res_start / d_res_goal: node index in
 *                   the start / goal tree) and are retired; slots older than max_age iterations are
 *                   retired unsolved; the iteration counter advances (the trees swap roles, :231-235).
 * d_counters (int64[8], zero-initialised by the caller): [0] iteration, [1] solved, [2] gave up,
 * [3] active slots after the last iteration, [4] scratch.  The parity of [0] is `swapped`: on odd
 * iterations the caller extends the goal tree first and passes (qa, ia) of the goal tree.
 */
int mjb_rrt_sample(uint64_t seed, const int64_t *d_counters, int64_t nslots, int32_t nq, const double *d_q_init,
                   const double *d_q_goal, const uint8_t *d_plan_mask, const double *d_lo, const double *d_hi,
                   double goal_bias, const uint8_t *d_active, double *d_targets, void *stream);
int mjb_rrt_meet(int64_t nslots, int32_t nq, const double *d_qa, const double *d_qb, const int64_t *d_ia,
                 const int64_t *d_ib, int64_t max_age, uint8_t *d_active, int64_t *d_age, int64_t *d_res_start,
                 int64_t *d_res_goal, int64_t *d_counters, void *stream);

/*
 * PoseConstraint (src/mjpl/constraint/pose_constraint.py:11-171): a site must stay inside a box
 * of translations and roll/pitch/yaw expressed in a constraint frame.  fp64, rows (n,nq) of
 * doubles.  mjb_pose_valid = valid_config (:72-76: joint limits, then |displacement| <=
 * tolerance); mjb_pose_project = apply (:78-91): q -= J^T pinv(J J^T) dx until the displacement
 * is within tolerance (ok=1, row written to d_q_out) or the row leaves the joint limits / moves
 * more than 2*q_step from q_old / exceeds max_iters (ok=0: the reference returns None).
 */
typedef struct mjb_pose_spec {
  int32_t site_bodyid;                /* model.site_bodyid[site]  (pose_constraint.py:70) */
  double site_pos[3], site_quat[4];   /* model.site_pos / site_quat */
  double ref_pos[3], ref_quat[4];     /* reference_frame (world_T_C), quaternion wxyz */
  double lower[6], upper[6];          /* x, y, z, roll, pitch, yaw limits (may be +-inf) */
  double tolerance, q_step;
} mjb_pose_spec;

/* site_pose (src/mjpl/utils.py:60-75) for a block of fp64 rows: world position (n,3) and
 * quaternion (n,4 wxyz) of a site given by its body id and local pose. */
int mjb_site_pose(mjb_model *m, int32_t site_bodyid, const double *site_pos, const double *site_quat,
                  const double *d_q, int64_t n, double *d_pos, double *d_quat, void *stream);

int mjb_pose_valid(mjb_model *m, const mjb_pose_spec *spec, const double *d_q, int64_t n,
                   uint8_t *d_valid, void *stream);
int mjb_pose_project(mjb_model *m, const mjb_pose_spec *spec, const double *d_q_old,
                     const double *d_q, int64_t n, int32_t max_iters, double *d_q_out,
                     uint8_t *d_ok, int32_t *d_iters, void *stream);

/*
 * CBiRRT with a projecting constraint (BASELINE configs[3]): RRT.plan_to_configs
 * (src/mjpl/planning/rrt.py:195-235) over the step-by-step _constrained_extend
 * (src/mjpl/planning/utils.py:139-164) for S queries held in S slots, all state on the device.  The
 * slots advance asynchronously: one call = one TICK = every slot takes one projected step of the
 * extend it is in (propose `_step`, joint limits of the proposal, PoseConstraint.apply, collision
 * check of the projected row, the reference's stop rules, append), or sets up its next extend (sample
 * / nearest node), or runs its connection test.  Nothing returns to the host; the caller looks at
 * d_counters every few ticks.  Arrays are device pointers owned by the caller:
 *   nodes[k] (S,cap,nq) fp64, parent[k] (S,cap), count[k] (S): k = 0 start trees (root q_init, count 1),
 *   k = 1 goal trees (root q_goal); phase (S) int32 zero-initialised; swapped (S); age (S);
 *   target, tip, qa, cand, proj (S,nq) fp64 scratch; cand32 (S,nq) fp32; last, ia (S);
 *   proj_ok, valid, stepping (S) bytes; res_start / res_goal (S) = -1 until the slot's query is solved
 *   (then: connecting node in the start / goal tree); counters int64[8] zero-initialised:
 *   [0] ticks, [1] solved, [2] gave up (max_age iterations), [3] slots not yet retired after the last
 *   tick, [4] scratch, [5] appends refused because a tree was at capacity.
 */
typedef struct mjb_cbirrt_state {
  int64_t nslots, cap;
  int32_t nq, check_limits_before;   /* a JointLimitConstraint precedes the PoseConstraint in the list */
  double eps, goal_bias;
  uint64_t seed;
  int64_t max_age;
  const double *q_init, *q_goal;
  const uint8_t *plan_mask;
  const double *lo, *hi;
  double *nodes[2];
  int64_t *parent[2], *count[2];
  int32_t *phase;
  uint8_t *swapped;
  int64_t *age;
  double *target, *tip, *qa;
  int64_t *last, *ia;
  double *cand;
  float *cand32;
  double *proj;
  uint8_t *proj_ok, *valid, *stepping;
  int64_t *res_start, *res_goal, *counters;
} mjb_cbirrt_state;

int mjb_cbirrt_tick(mjb_model *m, const mjb_cbirrt_state *state, const mjb_pose_spec *pose, int32_t pose_max_iters,
                    uint32_t flags, void *stream);

/*
 * IKSolver.solve_ik (src/mjpl/inverse_kinematics/ik_solver_interface.py:11-28) for n (target,
 * initial guess) rows at once; replaces the per-attempt iteration loop of the stock solver
 * (mink_ik_solver.py:93-108: iterate until |position error| <= pos_tolerance and |orientation
 * error| <= ori_tolerance, at most `iterations` times).  The update is Levenberg-Marquardt damped
 * least squares on the geometric site Jacobian, fp64, with every iterate clamped to the joint
 * limits (the role of mink.ConfigurationLimit, :87); joints outside movable_mask are held fixed
 * (the DampingTask of :64-70).  Retries from random configurations and the constraint check on
 * the result (:99-116) are the caller's (mjpl_b200.inverse_kinematics).
 * d_target_pos (n,3), d_target_quat (n,4 wxyz), d_q_init / d_q_out (n,nq) fp64; d_ok[i] = 1 iff
 * row i converged; d_iters (optional); d_err (optional, (n,2)): final position / rotation error.
 */
typedef struct mjb_ik_spec {
  int32_t site_bodyid;
  double site_pos[3], site_quat[4];
  uint32_t movable_mask;             /* bit j: joint id j may move */
  double pos_tolerance, ori_tolerance;
  double lm_damping;                 /* error-proportional damping (mink FrameTask lm_damping, :80); <0 = default 0.1 */
  double damping;                    /* constant damping; <=0 = default 1e-9 */
  double max_step;                   /* cap on |dq|_inf per iteration; <=0 = default 0.5 */
  int32_t iterations;
} mjb_ik_spec;

int mjb_ik_solve(mjb_model *m, const mjb_ik_spec *spec, const double *d_target_pos,
                 const double *d_target_quat, const double *d_q_init, int64_t n, double *d_q_out,
                 uint8_t *d_ok, int32_t *d_iters, double *d_err, void *stream);

/*
 * Per-kernel timing of the validity launches (measurement aid for bench.py; off by default).
 * enable != 0 switches CUDA-event recording on for later launches (up to 2048 launches between
 * reads).  If ms4 != NULL the device is synchronised and ms4 receives the summed durations since
 * the last read: [0] validity_kernel, or fk_cull_kernel when the batch ran as the multi-kernel
 * pipeline, [1] mid_kernel, [2] narrow_kernel (both 0 for the single kernel), [3] the fp64 item
 * pass; *launches = number of validity launches summed.
 */
int mjb_kernel_timing(mjb_model *m, int enable, double *ms4, int64_t *launches);

/*
 * Measured FP32 FMA throughput of the current device in TFLOP/s (2 flops per FMA): the roofline
 * denominator for this path, which is bound by the CUDA cores and not by HBM or tensor cores
 * (BASELINE.md section 2 asks for a measured figure).  Runs ~25 ms; *ms_out (optional) = duration
 * of the best pass.
 */
int mjb_fma_peak(double *tflops, double *ms_out);

int mjb_get_stats(mjb_model *m, mjb_stats *out);   /* synchronises the handle's last stream */
int mjb_reset_stats(mjb_model *m);

#ifdef __cplusplus
}
#endif
#endif
