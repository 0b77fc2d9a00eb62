/*
 * oracle.h -- fp64 CPU restatement of mjpl's configuration-validity path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under mjpl_b200/ may include, link or call this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker / the timed CPU baseline.
 *
 * PARITY STATUS: "parity unpinned" against a real MuJoCo.  The arithmetic the reference
 * runs (reference: src/mjpl/constraint/collision_constraint.py:26-30) lives in the
 * third-party `mujoco` wheel (pyproject.toml:12, "mujoco >= 3", unpinned, not vendored, not
 * installable here: no network, no wheel).  This file restates MuJoCo 3.x's published
 * semantics and is pinned only by the reference's own known-answer
 * tests (tests/test_oracle.py lists them with reference file:line) and by
 * closed-form analytic cases.  tools/make_golden.py regenerates real-MuJoCo vectors whenever
 * a MuJoCo install is reachable.
 */
#ifndef MJPL_ORACLE_H
#define MJPL_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Raw model tables under MuJoCo's own MjModel field names (same layout as
 * include/mjpl_b200.h:mjb_model_desc so one marshaller serves both). */
typedef struct orc_model_desc {
  int32_t nq, nbody, njnt, ngeom, nmesh, nmeshvert, nexclude, nallowed;
  int32_t disable_contact, disable_filterparent;
  const int32_t *body_parentid, *body_weldid, *body_jntadr, *body_jntnum;
  const double *body_pos, *body_quat;
  const int32_t *jnt_type, *jnt_qposadr, *jnt_bodyid, *jnt_limited;
  const double *jnt_pos, *jnt_axis, *jnt_range, *qpos0;
  const int32_t *geom_type, *geom_bodyid, *geom_contype, *geom_conaffinity, *geom_dataid;
  const double *geom_size, *geom_pos, *geom_quat, *geom_margin, *geom_gap;
  const int32_t *mesh_vertadr, *mesh_vertnum;
  const double *mesh_vert;
  const int64_t *exclude_signature;
  const int32_t *allowed_body_pairs; /* nallowed x 2 body ids (CollisionRuleset) */
} orc_model_desc;

typedef struct orc_model orc_model;

#define ORC_CHECK_LIMITS 1u
#define ORC_CHECK_COLLISION 2u

int orc_model_create(const orc_model_desc *desc, orc_model **out);
void orc_model_destroy(orc_model *m);
const char *orc_last_error(void);

/* static pair list after MuJoCo's body/type filters minus the allowed body pairs */
int32_t orc_npair(const orc_model *m);
void orc_pairs(const orc_model *m, int32_t *geom1, int32_t *geom2);

/* mj_kinematics restated: xpos (n,nbody,3), xquat (n,nbody,4) */
int orc_fk(const orc_model *m, const double *q, int64_t n, double *xpos, double *xquat);
/* geom poses: geom_xpos (n,ngeom,3), geom_xmat (n,ngeom,9) */
int orc_geom_poses(const orc_model *m, const double *q, int64_t n, double *gxpos, double *gxmat);

/* valid[i] = (limits ok if asked) && (no tested pair at signed distance <= margin if asked).
 * min_dist[i] (optional) = min over tested pairs of (signed distance - margin), +1e30 if no
 * pair passed the conservative bounding-sphere cull; min_pair[i] (optional) = its pair index
 * or -1.  Penetration deeper than ORC_DEPTH_CAP is reported as -ORC_DEPTH_CAP. */
#define ORC_DEPTH_CAP 1e-3
int orc_check(const orc_model *m, const double *q, int64_t n, uint32_t flags, uint8_t *valid,
              double *min_dist, int32_t *min_pair);

/* worker threads used by orc_check (default 1; one pinned-free pthread per unit) */
void orc_set_threads(int n);
int orc_get_threads(void);

/* signed distance of one static pair at one configuration (no culling) */
int orc_pair_distance(const orc_model *m, const double *q, int32_t pair, double *dist);

#ifdef __cplusplus
}
#endif
#endif
