/*
 * oracle.c -- fp64 CPU restatement of mjpl's configuration-validity path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  "parity unpinned" vs a real MuJoCo: the
 * arithmetic the reference executes is inside the third-party `mujoco` wheel
 * (reference: pyproject.toml:12; call sites src/mjpl/constraint/collision_constraint.py:27-30),
 * which cannot be installed here.  Each function below names the reference line or the
 * MuJoCo 3.x routine whose published behaviour it restates.
 *
 * Deliberately written in a different style from the CUDA product path (closest-point
 * region GJK + full EPA here; signed-volume GJK with certified bounds there) so that the
 * two implementations fail independently.
 */
#define _GNU_SOURCE
#include "oracle.h"

#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum { G_PLANE = 0, G_HFIELD = 1, G_SPHERE = 2, G_CAPSULE = 3, G_ELLIPSOID = 4, G_CYLINDER = 5,
       G_BOX = 6, G_MESH = 7 };
enum { J_FREE = 0, J_BALL = 1, J_SLIDE = 2, J_HINGE = 3 };

struct orc_model {
  int nq, nbody, njnt, ngeom, nmesh, nmeshvert;
  int32_t *body_parentid, *body_weldid, *body_jntadr, *body_jntnum;
  double *body_pos, *body_quat;
  int32_t *jnt_type, *jnt_qposadr, *jnt_bodyid;
  double *jnt_pos, *jnt_axis, *jnt_range, *qpos0;
  int32_t *geom_type, *geom_bodyid, *geom_dataid;
  double *geom_size, *geom_pos, *geom_quat, *geom_margin;
  int32_t *mesh_vertadr, *mesh_vertnum;
  double *mesh_vert;
  double *geom_bcen; /* bounding-sphere centre, geom frame */
  double *geom_brad; /* bounding-sphere radius (plane: <0) */
  int npair;
  int32_t *pair_g1, *pair_g2;
};

static __thread char g_err[256];
const char *orc_last_error(void) { return g_err; }

/* ------------------------------------------------------------------ small math */
static inline double dot3(const double *a, const double *b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
static inline void cross3(const double *a, const double *b, double *c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
static inline void sub3(const double *a, const double *b, double *c) {
  c[0] = a[0] - b[0]; c[1] = a[1] - b[1]; c[2] = a[2] - b[2];
}
static inline double norm3(const double *a) { return sqrt(dot3(a, a)); }

/* Hamilton product, (w,x,y,z) -- mju_mulQuat */
static void quat_mul(const double *a, const double *b, double *r) {
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static void quat_normalize(double *q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < 1e-15) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
/* mju_quat2Mat: row-major 3x3 */
static void quat2mat(const double *q, double *m) {
  double q00 = q[0] * q[0], q11 = q[1] * q[1], q22 = q[2] * q[2], q33 = q[3] * q[3];
  double q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
  double q12 = q[1] * q[2], q13 = q[1] * q[3], q23 = q[2] * q[3];
  m[0] = q00 + q11 - q22 - q33; m[4] = q00 - q11 + q22 - q33; m[8] = q00 - q11 - q22 + q33;
  m[1] = 2 * (q12 - q03); m[2] = 2 * (q13 + q02);
  m[3] = 2 * (q12 + q03); m[5] = 2 * (q23 - q01);
  m[6] = 2 * (q13 - q02); m[7] = 2 * (q23 + q01);
}
static inline void mat_vec(const double *m, const double *v, double *r) {
  r[0] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  r[1] = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  r[2] = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
}
static inline void matT_vec(const double *m, const double *v, double *r) {
  r[0] = m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
  r[1] = m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
  r[2] = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
}
/* rotate vector by quaternion -- mju_rotVecQuat */
static void rot_vec_quat(const double *v, const double *q, double *r) {
  double m[9];
  quat2mat(q, m);
  mat_vec(m, v, r);
}

/* ------------------------------------------------------------------ model */
static void *dup_mem(const void *src, size_t bytes) {
  void *p = malloc(bytes ? bytes : 1);
  if (bytes && src) memcpy(p, src, bytes);
  return p;
}
#define DUP(field, count, type) m->field = (type *)dup_mem(d->field, (size_t)(count) * sizeof(type))

static int body_pair_allowed(const orc_model_desc *d, int b1, int b2) {
  /* reference: collision_constraint.py:93-95 -- sorted TRUE body ids (geom_bodyid) must
   * match a sorted allowed pair */
  int lo = b1 < b2 ? b1 : b2, hi = b1 < b2 ? b2 : b1;
  for (int k = 0; k < d->nallowed; k++) {
    int a = d->allowed_body_pairs[2 * k], b = d->allowed_body_pairs[2 * k + 1];
    int alo = a < b ? a : b, ahi = a < b ? b : a;
    if (alo == lo && ahi == hi) return 1;
  }
  return 0;
}

static void build_bounds(orc_model *m) {
  m->geom_bcen = (double *)calloc((size_t)m->ngeom * 3, sizeof(double));
  m->geom_brad = (double *)calloc((size_t)m->ngeom, sizeof(double));
  for (int g = 0; g < m->ngeom; g++) {
    const double *s = m->geom_size + 3 * g;
    double r = 0;
    switch (m->geom_type[g]) {
      case G_PLANE: r = -1; break;
      case G_SPHERE: r = s[0]; break;
      case G_CAPSULE: r = s[0] + s[1]; break;
      case G_CYLINDER: r = sqrt(s[0] * s[0] + s[1] * s[1]); break;
      case G_BOX: r = sqrt(dot3(s, s)); break;
      case G_ELLIPSOID: r = fmax(s[0], fmax(s[1], s[2])); break;
      case G_MESH: {
        int id = m->geom_dataid[g];
        if (id < 0 || m->mesh_vertnum[id] == 0) { r = 0; break; }
        const double *v = m->mesh_vert + 3 * m->mesh_vertadr[id];
        int n = m->mesh_vertnum[id];
        double lo[3] = {1e30, 1e30, 1e30}, hi[3] = {-1e30, -1e30, -1e30};
        for (int i = 0; i < n; i++)
          for (int k = 0; k < 3; k++) {
            lo[k] = fmin(lo[k], v[3 * i + k]);
            hi[k] = fmax(hi[k], v[3 * i + k]);
          }
        double *c = m->geom_bcen + 3 * g;
        for (int k = 0; k < 3; k++) c[k] = 0.5 * (lo[k] + hi[k]);
        for (int i = 0; i < n; i++) {
          double dd[3];
          sub3(v + 3 * i, c, dd);
          r = fmax(r, norm3(dd));
        }
        break;
      }
      default: r = 1e30;
    }
    m->geom_brad[g] = r;
  }
}

int orc_model_create(const orc_model_desc *d, orc_model **out) {
  *out = NULL;
  for (int j = 0; j < d->njnt; j++)
    if (d->jnt_type[j] != J_HINGE && d->jnt_type[j] != J_SLIDE) {
      /* mjpl itself excludes ball/free joints (reference README.md:19-20) */
      snprintf(g_err, sizeof g_err, "joint %d: only hinge/slide joints are supported", j);
      return 1;
    }
  orc_model *m = (orc_model *)calloc(1, sizeof *m);
  m->nq = d->nq; m->nbody = d->nbody; m->njnt = d->njnt; m->ngeom = d->ngeom;
  m->nmesh = d->nmesh; m->nmeshvert = d->nmeshvert;
  DUP(body_parentid, d->nbody, int32_t); DUP(body_weldid, d->nbody, int32_t);
  DUP(body_jntadr, d->nbody, int32_t); DUP(body_jntnum, d->nbody, int32_t);
  DUP(body_pos, d->nbody * 3, double); DUP(body_quat, d->nbody * 4, double);
  DUP(jnt_type, d->njnt, int32_t); DUP(jnt_qposadr, d->njnt, int32_t);
  DUP(jnt_bodyid, d->njnt, int32_t);
  DUP(jnt_pos, d->njnt * 3, double); DUP(jnt_axis, d->njnt * 3, double);
  DUP(jnt_range, d->njnt * 2, double); DUP(qpos0, d->nq, double);
  DUP(geom_type, d->ngeom, int32_t); DUP(geom_bodyid, d->ngeom, int32_t);
  DUP(geom_dataid, d->ngeom, int32_t);
  DUP(geom_size, d->ngeom * 3, double); DUP(geom_pos, d->ngeom * 3, double);
  DUP(geom_quat, d->ngeom * 4, double); DUP(geom_margin, d->ngeom, double);
  DUP(mesh_vertadr, d->nmesh, int32_t); DUP(mesh_vertnum, d->nmesh, int32_t);
  DUP(mesh_vert, d->nmeshvert * 3, double);
  build_bounds(m);

  /* Static pair list.  MuJoCo mj_collision (engine_collision_driver.c): a geom pair is
   * tested iff contacts are enabled, the two bodies pass filterBodyPair (different weld
   * body; not weld-parent/child unless either weld body is the world; parent filter can be
   * disabled), the body pair is not <exclude>d, and
   * (contype1 & conaffinity2) || (contype2 & conaffinity1)  [mj_contactFilter].
   * mjpl then ignores contacts between allowed body pairs (collision_constraint.py:83-95),
   * which is the same as deleting those pairs up front. */
  int cap = 64;
  m->pair_g1 = (int32_t *)malloc(cap * sizeof(int32_t));
  m->pair_g2 = (int32_t *)malloc(cap * sizeof(int32_t));
  if (!d->disable_contact)
    for (int g1 = 0; g1 < d->ngeom; g1++)
      for (int g2 = g1 + 1; g2 < d->ngeom; g2++) {
        int b1 = d->geom_bodyid[g1], b2 = d->geom_bodyid[g2];
        int w1 = d->body_weldid[b1], w2 = d->body_weldid[b2];
        if (w1 == w2) continue;
        int wp1 = d->body_weldid[d->body_parentid[w1]];
        int wp2 = d->body_weldid[d->body_parentid[w2]];
        if (!d->disable_filterparent && w1 != 0 && w2 != 0 && (w1 == wp2 || w2 == wp1)) continue;
        int lo = b1 < b2 ? b1 : b2, hi = b1 < b2 ? b2 : b1, excluded = 0;
        for (int k = 0; k < d->nexclude; k++)
          if (d->exclude_signature[k] == (((int64_t)lo << 16) + hi)) excluded = 1;
        if (excluded) continue;
        if (!((d->geom_contype[g1] & d->geom_conaffinity[g2]) ||
              (d->geom_contype[g2] & d->geom_conaffinity[g1])))
          continue;
        if (body_pair_allowed(d, b1, b2)) continue;
        int t1 = d->geom_type[g1], t2 = d->geom_type[g2];
        if (t1 == G_HFIELD || t2 == G_HFIELD) {
          snprintf(g_err, sizeof g_err, "height fields are not supported");
          orc_model_destroy(m);
          return 1;
        }
        if (t1 == G_PLANE && t2 == G_PLANE) continue; /* MuJoCo has no plane-plane collider */
        if (m->npair == cap) {
          cap *= 2;
          m->pair_g1 = (int32_t *)realloc(m->pair_g1, cap * sizeof(int32_t));
          m->pair_g2 = (int32_t *)realloc(m->pair_g2, cap * sizeof(int32_t));
        }
        m->pair_g1[m->npair] = g1;
        m->pair_g2[m->npair] = g2;
        m->npair++;
      }
  *out = m;
  return 0;
}

void orc_model_destroy(orc_model *m) {
  if (!m) return;
  free(m->body_parentid); free(m->body_weldid); free(m->body_jntadr); free(m->body_jntnum);
  free(m->body_pos); free(m->body_quat); free(m->jnt_type); free(m->jnt_qposadr);
  free(m->jnt_bodyid); free(m->jnt_pos); free(m->jnt_axis); free(m->jnt_range); free(m->qpos0);
  free(m->geom_type); free(m->geom_bodyid); free(m->geom_dataid); free(m->geom_size);
  free(m->geom_pos); free(m->geom_quat); free(m->geom_margin); free(m->mesh_vertadr);
  free(m->mesh_vertnum); free(m->mesh_vert); free(m->geom_bcen); free(m->geom_brad);
  free(m->pair_g1); free(m->pair_g2);
  free(m);
}

int32_t orc_npair(const orc_model *m) { return m->npair; }
void orc_pairs(const orc_model *m, int32_t *g1, int32_t *g2) {
  memcpy(g1, m->pair_g1, m->npair * sizeof(int32_t));
  memcpy(g2, m->pair_g2, m->npair * sizeof(int32_t));
}

/* ------------------------------------------------------------------ forward kinematics
 * MuJoCo mj_kinematics (engine_core_smooth.c), hinge/slide/fixed bodies, SURVEY.md A.1. */
static void fk_one(const orc_model *m, const double *q, double *xpos, double *xquat) {
  xpos[0] = xpos[1] = xpos[2] = 0;
  xquat[0] = 1; xquat[1] = xquat[2] = xquat[3] = 0;
  for (int i = 1; i < m->nbody; i++) {
    int p = m->body_parentid[i];
    double *pos = xpos + 3 * i, *quat = xquat + 4 * i, t[3];
    rot_vec_quat(m->body_pos + 3 * i, xquat + 4 * p, t);
    pos[0] = xpos[3 * p] + t[0]; pos[1] = xpos[3 * p + 1] + t[1]; pos[2] = xpos[3 * p + 2] + t[2];
    quat_mul(xquat + 4 * p, m->body_quat + 4 * i, quat);
    for (int k = 0; k < m->body_jntnum[i]; k++) {
      int j = m->body_jntadr[i] + k, a = m->jnt_qposadr[j];
      double xanchor[3], xaxis[3], tmp[3];
      rot_vec_quat(m->jnt_pos + 3 * j, quat, tmp);
      xanchor[0] = pos[0] + tmp[0]; xanchor[1] = pos[1] + tmp[1]; xanchor[2] = pos[2] + tmp[2];
      rot_vec_quat(m->jnt_axis + 3 * j, quat, xaxis);
      double dq = q[a] - m->qpos0[a];
      if (m->jnt_type[j] == J_SLIDE) {
        pos[0] += xaxis[0] * dq; pos[1] += xaxis[1] * dq; pos[2] += xaxis[2] * dq;
      } else { /* hinge: local rotation about jnt_axis, then off-centre correction */
        double s = sin(0.5 * dq), c = cos(0.5 * dq);
        double qloc[4] = {c, m->jnt_axis[3 * j] * s, m->jnt_axis[3 * j + 1] * s,
                          m->jnt_axis[3 * j + 2] * s};
        double qn[4];
        quat_mul(quat, qloc, qn);
        memcpy(quat, qn, sizeof qn);
        rot_vec_quat(m->jnt_pos + 3 * j, quat, tmp);
        pos[0] = xanchor[0] - tmp[0]; pos[1] = xanchor[1] - tmp[1]; pos[2] = xanchor[2] - tmp[2];
      }
    }
    quat_normalize(quat);
  }
}

static void geom_pose(const orc_model *m, const double *xpos, const double *xquat, int g,
                      double *gpos, double *gmat) {
  int b = m->geom_bodyid[g];
  double t[3], gq[4];
  rot_vec_quat(m->geom_pos + 3 * g, xquat + 4 * b, t);
  gpos[0] = xpos[3 * b] + t[0]; gpos[1] = xpos[3 * b + 1] + t[1]; gpos[2] = xpos[3 * b + 2] + t[2];
  quat_mul(xquat + 4 * b, m->geom_quat + 4 * g, gq);
  quat_normalize(gq);
  quat2mat(gq, gmat);
}

int orc_fk(const orc_model *m, const double *q, int64_t n, double *xpos, double *xquat) {
  for (int64_t i = 0; i < n; i++)
    fk_one(m, q + i * m->nq, xpos + i * m->nbody * 3, xquat + i * m->nbody * 4);
  return 0;
}

int orc_geom_poses(const orc_model *m, const double *q, int64_t n, double *gxpos, double *gxmat) {
  double *xpos = (double *)malloc(m->nbody * 3 * sizeof(double));
  double *xquat = (double *)malloc(m->nbody * 4 * sizeof(double));
  for (int64_t i = 0; i < n; i++) {
    fk_one(m, q + i * m->nq, xpos, xquat);
    for (int g = 0; g < m->ngeom; g++)
      geom_pose(m, xpos, xquat, g, gxpos + (i * m->ngeom + g) * 3, gxmat + (i * m->ngeom + g) * 9);
  }
  free(xpos); free(xquat);
  return 0;
}

/* ------------------------------------------------------------------ convex shapes */
typedef struct {
  int type;
  const double *size;
  double pos[3], mat[9];
  const double *verts;
  int nvert;
} shape_t;

static double swept_radius(const shape_t *s) {
  return (s->type == G_SPHERE || s->type == G_CAPSULE) ? s->size[0] : 0.0;
}

/* support point of the CORE shape (sphere -> centre, capsule -> segment) in world frame */
static void support(const shape_t *s, const double *dir, double *out) {
  double dl[3], pl[3] = {0, 0, 0};
  matT_vec(s->mat, dir, dl);
  switch (s->type) {
    case G_SPHERE: break;
    case G_CAPSULE: pl[2] = dl[2] >= 0 ? s->size[1] : -s->size[1]; break;
    case G_CYLINDER: {
      double r = sqrt(dl[0] * dl[0] + dl[1] * dl[1]);
      if (r > 1e-300) { pl[0] = s->size[0] * dl[0] / r; pl[1] = s->size[0] * dl[1] / r; }
      pl[2] = dl[2] >= 0 ? s->size[1] : -s->size[1];
      break;
    }
    case G_BOX:
      for (int k = 0; k < 3; k++) pl[k] = dl[k] >= 0 ? s->size[k] : -s->size[k];
      break;
    case G_ELLIPSOID: {
      double t[3] = {s->size[0] * dl[0], s->size[1] * dl[1], s->size[2] * dl[2]};
      double n = norm3(t);
      if (n > 1e-300) for (int k = 0; k < 3; k++) pl[k] = s->size[k] * t[k] / n;
      break;
    }
    case G_MESH: {
      double best = -1e300;
      int bi = 0;
      for (int i = 0; i < s->nvert; i++) {
        double v = dot3(s->verts + 3 * i, dl);
        if (v > best) { best = v; bi = i; }
      }
      pl[0] = s->verts[3 * bi]; pl[1] = s->verts[3 * bi + 1]; pl[2] = s->verts[3 * bi + 2];
      break;
    }
    default: break;
  }
  mat_vec(s->mat, pl, out);
  out[0] += s->pos[0]; out[1] += s->pos[1]; out[2] += s->pos[2];
}

static void mink_support(const shape_t *a, const shape_t *b, const double *dir, double *w) {
  double pa[3], pb[3], nd[3] = {-dir[0], -dir[1], -dir[2]};
  support(a, dir, pa);
  support(b, nd, pb);
  sub3(pa, pb, w);
}

/* ---- closest point of a simplex to the origin (Ericson, Real-Time Collision Detection
 * 5.1.2 / 5.1.5 / 5.1.6).  Returns barycentric weights; zero weight = vertex dropped. */
static void closest_segment(const double *a, const double *b, double *lam) {
  double ab[3];
  sub3(b, a, ab);
  double den = dot3(ab, ab);
  double t = den > 0 ? -dot3(a, ab) / den : 0;
  if (t <= 0) { lam[0] = 1; lam[1] = 0; }
  else if (t >= 1) { lam[0] = 0; lam[1] = 1; }
  else { lam[0] = 1 - t; lam[1] = t; }
}

static void closest_triangle(const double *a, const double *b, const double *c, double *lam) {
  double ab[3], ac[3];
  sub3(b, a, ab); sub3(c, a, ac);
  double d1 = -dot3(ab, a), d2 = -dot3(ac, a);
  lam[0] = lam[1] = lam[2] = 0;
  if (d1 <= 0 && d2 <= 0) { lam[0] = 1; return; }
  double d3 = -dot3(ab, b), d4 = -dot3(ac, b);
  if (d3 >= 0 && d4 <= d3) { lam[1] = 1; return; }
  double vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) { double v = d1 / (d1 - d3); lam[0] = 1 - v; lam[1] = v; return; }
  double d5 = -dot3(ab, c), d6 = -dot3(ac, c);
  if (d6 >= 0 && d5 <= d6) { lam[2] = 1; return; }
  double vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) { double w = d2 / (d2 - d6); lam[0] = 1 - w; lam[2] = w; return; }
  double va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
    double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
    lam[1] = 1 - w; lam[2] = w; return;
  }
  double den = 1.0 / (va + vb + vc);
  lam[1] = vb * den; lam[2] = vc * den; lam[0] = 1 - lam[1] - lam[2];
}

static void lincomb(double W[][3], const double *lam, int n, double *v) {
  v[0] = v[1] = v[2] = 0;
  for (int i = 0; i < n; i++)
    for (int k = 0; k < 3; k++) v[k] += lam[i] * W[i][k];
}

/* origin on the outside of plane (a,b,c) relative to d?  >0 outside, sign-robust */
static double outside_face(const double *a, const double *b, const double *c, const double *d) {
  double ab[3], ac[3], n[3], ad[3];
  sub3(b, a, ab); sub3(c, a, ac); cross3(ab, ac, n);
  sub3(d, a, ad);
  double sd = dot3(n, ad);   /* side of d */
  double so = -dot3(n, a);   /* side of origin */
  return -so * sd;           /* >0: opposite sides; 0: degenerate / on the plane */
}

/* returns 1 if the origin is inside the tetrahedron */
static int closest_tetra(double W[4][3], double *lam) {
  static const int F[4][3] = {{0, 1, 2}, {0, 1, 3}, {0, 2, 3}, {1, 2, 3}};
  static const int O[4] = {3, 2, 1, 0};
  /* A flat tetrahedron (GJK landing on a face of A-B and adding a fourth, coplanar support
   * point) has no inside: its face-side tests are rounding noise.  Detect it by volume and
   * fall back to the closest of all four faces. */
  double e1[3], e2[3], e3[3], cr[3];
  sub3(W[1], W[0], e1); sub3(W[2], W[0], e2); sub3(W[3], W[0], e3);
  cross3(e1, e2, cr);
  double vol6 = fabs(dot3(cr, e3));
  double len = fmax(norm3(e1), fmax(norm3(e2), norm3(e3)));
  int flat = !(vol6 > 1e-10 * len * len * len);
  double best = 1e300;
  int any = 0;
  for (int f = 0; f < 4; f++) {
    const double *a = W[F[f][0]], *b = W[F[f][1]], *c = W[F[f][2]];
    if (flat || outside_face(a, b, c, W[O[f]]) >= 0) {
      double l3[3], v[3];
      closest_triangle(a, b, c, l3);
      for (int k = 0; k < 3; k++) v[k] = l3[0] * a[k] + l3[1] * b[k] + l3[2] * c[k];
      double d = dot3(v, v);
      if (d < best) {
        best = d; any = 1;
        lam[0] = lam[1] = lam[2] = lam[3] = 0;
        lam[F[f][0]] = l3[0]; lam[F[f][1]] = l3[1]; lam[F[f][2]] = l3[2];
      }
    }
  }
  return any ? 0 : 1;
}

/* GJK distance between core shapes.  Returns distance (>0) or 0 when intersecting/touching;
 * in the latter case the final simplex (n points of A-B) is left in W for EPA. */
#define GJK_MAXIT 200
static double gjk_distance(const shape_t *A, const shape_t *B, double W[4][3], int *nW) {
  double v[3], w[3];
  sub3(A->pos, B->pos, v);
  if (dot3(v, v) < 1e-24) { v[0] = 1; v[1] = v[2] = 0; }
  int n = 0;
  double scale2 = 0;
  for (int it = 0; it < GJK_MAXIT; it++) {
    double nv[3] = {-v[0], -v[1], -v[2]};
    mink_support(A, B, nv, w);
    double vv = dot3(v, v), vw = dot3(v, w);
    if (n > 0) {
      double gap = vv - vw; /* >= 0 up to rounding; |v| - gap/|v| is a lower bound */
      if (gap <= 1e-11 * vv || gap <= 1e-14 * sqrt(vv)) break;
      int dup = 0;
      for (int i = 0; i < n; i++)
        if (W[i][0] == w[0] && W[i][1] == w[1] && W[i][2] == w[2]) dup = 1;
      if (dup) break;
    }
    memcpy(W[n++], w, sizeof w);
    scale2 = fmax(scale2, dot3(w, w));
    double lam[4] = {1, 0, 0, 0};
    int inside = 0;
    if (n == 2) closest_segment(W[0], W[1], lam);
    else if (n == 3) closest_triangle(W[0], W[1], W[2], lam);
    else if (n == 4) inside = closest_tetra(W, lam);
    if (inside) { *nW = 4; return 0.0; }
    lincomb(W, lam, n, v);
    /* drop vertices with zero weight */
    int k = 0;
    for (int i = 0; i < n; i++)
      if (lam[i] > 0) { if (k != i) memcpy(W[k], W[i], sizeof w); k++; }
    n = k;
    if (dot3(v, v) <= 1e-30 * fmax(scale2, 1e-300)) { *nW = n; return 0.0; }
  }
  *nW = n;
  return norm3(v);
}

/* ---- EPA: penetration depth of the origin inside A-B.  Stops early once the depth is
 * certified >= ORC_DEPTH_CAP (the band accounting only needs depths near zero). */
#define EPA_MAXV 160
#define EPA_MAXF 320
typedef struct { int v[3]; double n[3]; double d; int alive; } epa_face;

static int epa_make_face(double P[][3], const double *inner, int a, int b, int c, epa_face *f) {
  double ab[3], ac[3], tmp[3];
  sub3(P[b], P[a], ab); sub3(P[c], P[a], ac);
  cross3(ab, ac, f->n);
  double len = norm3(f->n);
  if (len < 1e-300) return 0;
  f->n[0] /= len; f->n[1] /= len; f->n[2] /= len;
  f->v[0] = a; f->v[1] = b; f->v[2] = c;
  sub3(P[a], inner, tmp);
  if (dot3(f->n, tmp) < 0) { /* orient away from an interior point */
    f->n[0] = -f->n[0]; f->n[1] = -f->n[1]; f->n[2] = -f->n[2];
    f->v[1] = c; f->v[2] = b;
  }
  f->d = dot3(f->n, P[a]);
  f->alive = 1;
  return 1;
}

static double epa_depth(const shape_t *A, const shape_t *B, double W[4][3], int nW) {
  static const double AX[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
  double P[EPA_MAXV][3];
  int np = nW;
  memcpy(P, W, sizeof(double) * 3 * nW);
  /* grow a degenerate GJK simplex into a tetrahedron that contains the origin */
  const double tiny = 1e-10;
  if (np == 1) {
    for (int k = 0; k < 6 && np == 1; k++) {
      double w[3], dd[3];
      mink_support(A, B, AX[k], w);
      sub3(w, P[0], dd);
      if (norm3(dd) > tiny) memcpy(P[np++], w, sizeof w);
    }
    if (np == 1) return 0.0;
  }
  if (np == 2) {
    double e[3];
    sub3(P[1], P[0], e);
    double best = 1e300;
    int bk = 0;
    for (int k = 0; k < 3; k++) if (fabs(e[k]) < best) { best = fabs(e[k]); bk = k; }
    double ax[3] = {0, 0, 0}, d1[3];
    ax[bk] = 1;
    cross3(e, ax, d1);
    for (int s = 0; s < 6 && np == 2; s++) {
      /* rotate the search direction around the edge in 60-degree steps */
      double ang = s * (M_PI / 3.0), el = norm3(e), u[3] = {e[0] / el, e[1] / el, e[2] / el}, d2[3], dir[3];
      cross3(u, d1, d2);
      for (int k = 0; k < 3; k++) dir[k] = cos(ang) * d1[k] + sin(ang) * d2[k];
      double w[3], t[3], cr[3];
      mink_support(A, B, dir, w);
      sub3(w, P[0], t);
      cross3(e, t, cr);
      if (norm3(cr) > tiny * el) memcpy(P[np++], w, sizeof w);
    }
    if (np == 2) return 0.0;
  }
  if (np == 3) {
    double ab[3], ac[3], n[3];
    sub3(P[1], P[0], ab); sub3(P[2], P[0], ac); cross3(ab, ac, n);
    double len = norm3(n);
    if (len < 1e-300) return 0.0;
    n[0] /= len; n[1] /= len; n[2] /= len;
    double w1[3], w2[3], nn[3] = {-n[0], -n[1], -n[2]}, t[3];
    mink_support(A, B, n, w1);
    mink_support(A, B, nn, w2);
    sub3(w1, P[0], t);
    double h1 = fabs(dot3(t, n));
    sub3(w2, P[0], t);
    double h2 = fabs(dot3(t, n));
    if (h1 < tiny && h2 < tiny) return 0.0;
    memcpy(P[np++], h1 >= h2 ? w1 : w2, sizeof w1);
  }
  double inner[3];
  for (int k = 0; k < 3; k++) inner[k] = 0.25 * (P[0][k] + P[1][k] + P[2][k] + P[3][k]);
  epa_face F[EPA_MAXF];
  int nf = 0;
  static const int T[4][3] = {{0, 1, 2}, {0, 1, 3}, {0, 2, 3}, {1, 2, 3}};
  for (int f = 0; f < 4; f++)
    if (epa_make_face(P, inner, T[f][0], T[f][1], T[f][2], &F[nf])) nf++;
  if (nf < 4) return 0.0;
  /* the origin may sit marginally outside the start tetra (touching case) */
  double lower = 0;
  for (int it = 0; it < 128; it++) {
    int bf = -1;
    double bd = 1e300;
    for (int f = 0; f < nf; f++)
      if (F[f].alive && F[f].d < bd) { bd = F[f].d; bf = f; }
    if (bf < 0) break;
    lower = bd > 0 ? bd : 0;
    if (lower >= ORC_DEPTH_CAP) return lower;
    double w[3];
    mink_support(A, B, F[bf].n, w);
    double h = dot3(w, F[bf].n);
    if (h - bd < 1e-10 || np >= EPA_MAXV) return h > 0 ? h : 0;
    memcpy(P[np], w, sizeof w);
    /* remove faces visible from w, collect the horizon */
    int E[EPA_MAXF * 3][2], ne = 0;
    for (int f = 0; f < nf; f++) {
      if (!F[f].alive) continue;
      double t[3];
      sub3(w, P[F[f].v[0]], t);
      if (dot3(F[f].n, t) > 1e-14) {
        F[f].alive = 0;
        for (int e = 0; e < 3; e++) {
          int a = F[f].v[e], b = F[f].v[(e + 1) % 3], found = -1;
          for (int k = 0; k < ne; k++)
            if (E[k][0] == b && E[k][1] == a) { found = k; break; }
          if (found >= 0) { E[found][0] = E[ne - 1][0]; E[found][1] = E[ne - 1][1]; ne--; }
          else { E[ne][0] = a; E[ne][1] = b; ne++; }
        }
      }
    }
    if (ne == 0) return h > 0 ? h : 0;
    for (int k = 0; k < ne; k++) {
      int slot = -1;
      for (int f = 0; f < nf; f++) if (!F[f].alive) { slot = f; break; }
      if (slot < 0) { if (nf >= EPA_MAXF) return lower; slot = nf++; }
      if (!epa_make_face(P, inner, E[k][0], E[k][1], np, &F[slot])) F[slot].alive = 0;
    }
    np++;
  }
  return lower;
}

static double convex_distance(const shape_t *A, const shape_t *B) {
  double W[4][3];
  int nW = 0;
  double r = swept_radius(A) + swept_radius(B);
  double d = gjk_distance(A, B, W, &nW);
  if (d > 0) return d - r;
  if (r >= ORC_DEPTH_CAP) return -r; /* cores touch: deeper than the cap already */
  return -epa_depth(A, B, W, nW) - r;
}

/* ------------------------------------------------------------------ analytic primitives */
static void capsule_ends(const shape_t *c, double *p, double *q) {
  for (int k = 0; k < 3; k++) {
    double ax = c->mat[3 * k + 2] * c->size[1];
    p[k] = c->pos[k] - ax;
    q[k] = c->pos[k] + ax;
  }
}

static double point_segment(const double *x, const double *p, const double *q) {
  double d[3], xp[3];
  sub3(q, p, d); sub3(x, p, xp);
  double den = dot3(d, d), t = den > 0 ? dot3(xp, d) / den : 0;
  t = t < 0 ? 0 : (t > 1 ? 1 : t);
  double c[3] = {p[0] + t * d[0] - x[0], p[1] + t * d[1] - x[1], p[2] + t * d[2] - x[2]};
  return norm3(c);
}

/* Ericson 5.1.9 ClosestPtSegmentSegment */
static double segment_segment(const double *p1, const double *q1, const double *p2, const double *q2) {
  double d1[3], d2[3], r[3];
  sub3(q1, p1, d1); sub3(q2, p2, d2); sub3(p1, p2, r);
  double a = dot3(d1, d1), e = dot3(d2, d2), f = dot3(d2, r), s, t;
  const double EPS = 1e-300;
  if (a <= EPS && e <= EPS) return norm3(r);
  if (a <= EPS) { s = 0; t = f / e; t = t < 0 ? 0 : (t > 1 ? 1 : t); }
  else {
    double c = dot3(d1, r);
    if (e <= EPS) { t = 0; s = -c / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
    else {
      double b = dot3(d1, d2), den = a * e - b * b;
      s = den > 1e-30 * a * e ? (b * f - c * e) / den : 0;
      s = s < 0 ? 0 : (s > 1 ? 1 : s);
      t = (b * s + f) / e;
      if (t < 0) { t = 0; s = -c / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
      else if (t > 1) { t = 1; s = (b - c) / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
    }
  }
  double c1[3], c2[3], dd[3];
  for (int k = 0; k < 3; k++) { c1[k] = p1[k] + s * d1[k]; c2[k] = p2[k] + t * d2[k]; }
  sub3(c1, c2, dd);
  return norm3(dd);
}

static double sphere_box(const shape_t *s, const shape_t *b) {
  double rel[3], l[3];
  sub3(s->pos, b->pos, rel);
  matT_vec(b->mat, rel, l);
  double out2 = 0, inside = 1e300;
  for (int k = 0; k < 3; k++) {
    double e = fabs(l[k]) - b->size[k];
    if (e > 0) out2 += e * e;
    if (-e < inside) inside = -e;
  }
  if (out2 > 0) return sqrt(out2) - s->size[0];
  return -inside - s->size[0];
}

/* signed distance plane (A) vs anything (B); plane normal = 3rd column of its frame */
static double plane_distance(const shape_t *P, const shape_t *B) {
  double n[3] = {P->mat[2], P->mat[5], P->mat[8]}, rel[3];
  sub3(B->pos, P->pos, rel);
  double dc = dot3(n, rel);
  switch (B->type) {
    case G_SPHERE: return dc - B->size[0];
    case G_CAPSULE: { /* mjc_PlaneCapsule: the two end spheres */
      double ax = n[0] * B->mat[2] + n[1] * B->mat[5] + n[2] * B->mat[8];
      return dc - fabs(ax) * B->size[1] - B->size[0];
    }
    case G_CYLINDER: { /* mjc_PlaneCylinder */
      double ax = n[0] * B->mat[2] + n[1] * B->mat[5] + n[2] * B->mat[8];
      double rad = 1 - ax * ax;
      return dc - fabs(ax) * B->size[1] - B->size[0] * sqrt(rad > 0 ? rad : 0);
    }
    case G_BOX: { /* mjc_PlaneBox: deepest corner */
      double nl[3];
      matT_vec(B->mat, n, nl);
      return dc - fabs(nl[0]) * B->size[0] - fabs(nl[1]) * B->size[1] - fabs(nl[2]) * B->size[2];
    }
    default: { /* mjc_PlaneConvex: support point along -n (mesh hull vertex, ellipsoid) */
      double nn[3] = {-n[0], -n[1], -n[2]}, s[3], t[3];
      support(B, nn, s);
      sub3(s, P->pos, t);
      return dot3(n, t) - swept_radius(B);
    }
  }
}

static double pair_signed_distance(const shape_t *A, const shape_t *B) {
  /* order by type like MuJoCo's collision table (type1 <= type2) */
  if (A->type > B->type) { const shape_t *t = A; A = B; B = t; }
  if (A->type == G_PLANE) return plane_distance(A, B);
  if (A->type == G_SPHERE && B->type == G_SPHERE) { /* mjc_SphereSphere */
    double d[3];
    sub3(A->pos, B->pos, d);
    return norm3(d) - A->size[0] - B->size[0];
  }
  if (A->type == G_SPHERE && B->type == G_CAPSULE) { /* mjc_SphereCapsule */
    double p[3], q[3];
    capsule_ends(B, p, q);
    return point_segment(A->pos, p, q) - A->size[0] - B->size[0];
  }
  if (A->type == G_CAPSULE && B->type == G_CAPSULE) { /* mjc_CapsuleCapsule */
    double p1[3], q1[3], p2[3], q2[3];
    capsule_ends(A, p1, q1);
    capsule_ends(B, p2, q2);
    return segment_segment(p1, q1, p2, q2) - A->size[0] - B->size[0];
  }
  if (A->type == G_SPHERE && B->type == G_BOX) return sphere_box(A, B); /* mjc_SphereBox */
  /* mjc_CapsuleBox, mjc_BoxBox, mjc_SphereCylinder and mjc_Convex pairs: with margin 0 each
   * reduces to "closed convex sets at distance <= 0"; evaluated with exact GJK/EPA here. */
  return convex_distance(A, B);
}

static void make_shape(const orc_model *m, const double *xpos, const double *xquat, int g, shape_t *s) {
  s->type = m->geom_type[g];
  s->size = m->geom_size + 3 * g;
  geom_pose(m, xpos, xquat, g, s->pos, s->mat);
  s->verts = NULL; s->nvert = 0;
  if (s->type == G_MESH) {
    int id = m->geom_dataid[g];
    s->verts = m->mesh_vert + 3 * m->mesh_vertadr[id];
    s->nvert = m->mesh_vertnum[id];
  }
}

/* ------------------------------------------------------------------ validity */
static int limits_ok(const orc_model *m, const double *q) {
  /* reference: joint_limit_constraint.py:16-20 -- np.all((q >= lower) & (q <= upper)) with
   * lower/upper = model.jnt_range columns (needs nq == njnt; closed interval) */
  for (int j = 0; j < m->njnt; j++)
    if (!(q[j] >= m->jnt_range[2 * j] && q[j] <= m->jnt_range[2 * j + 1])) return 0;
  return 1;
}

static void check_one(const orc_model *m, const double *q, uint32_t flags, double *xpos,
                      double *xquat, shape_t *shapes, double *bcen, uint8_t *valid,
                      double *min_dist, int32_t *min_pair, int want_dist) {
  int ok = 1;
  double best = 1e30;
  int bestp = -1;
  if ((flags & ORC_CHECK_LIMITS) && !limits_ok(m, q)) ok = 0;
  if ((flags & ORC_CHECK_COLLISION) && (ok || want_dist)) {
    fk_one(m, q, xpos, xquat);
    for (int g = 0; g < m->ngeom; g++) shapes[g].type = -1;
    for (int p = 0; p < m->npair; p++) {
      int g1 = m->pair_g1[p], g2 = m->pair_g2[p];
      for (int s = 0; s < 2; s++) {
        int g = s ? g2 : g1;
        if (shapes[g].type < 0) {
          make_shape(m, xpos, xquat, g, &shapes[g]);
          double c[3];
          mat_vec(shapes[g].mat, m->geom_bcen + 3 * g, c);
          for (int k = 0; k < 3; k++) bcen[3 * g + k] = shapes[g].pos[k] + c[k];
        }
      }
      double margin = fmax(m->geom_margin[g1], m->geom_margin[g2]);
      /* conservative bounding-sphere cull (mj_filterSphere analogue), 1 mm slack */
      double r1 = m->geom_brad[g1], r2 = m->geom_brad[g2];
      if (r1 >= 0 && r2 >= 0) {
        double d[3];
        sub3(bcen + 3 * g1, bcen + 3 * g2, d);
        double rs = r1 + r2 + margin + 1e-3;
        if (dot3(d, d) > rs * rs) continue;
      } else if (r1 < 0 && r2 >= 0) {
        double n[3] = {shapes[g1].mat[2], shapes[g1].mat[5], shapes[g1].mat[8]}, d[3];
        sub3(bcen + 3 * g2, shapes[g1].pos, d);
        if (dot3(n, d) > r2 + margin + 1e-3) continue;
      }
      double dist = pair_signed_distance(&shapes[g1], &shapes[g2]) - margin;
      if (dist < -ORC_DEPTH_CAP) dist = -ORC_DEPTH_CAP;
      if (dist < best) { best = dist; bestp = p; }
      if (dist <= 0) { ok = 0; if (!want_dist) break; }
    }
  }
  *valid = (uint8_t)ok;
  if (min_dist) *min_dist = best;
  if (min_pair) *min_pair = bestp;
}

typedef struct {
  const orc_model *m;
  const double *q;
  int64_t n;
  uint32_t flags;
  uint8_t *valid;
  double *min_dist;
  int32_t *min_pair;
  int64_t *next; /* shared work counter (chunks of CHUNK rows) */
} check_job;

#define CHUNK 256
static void *check_worker(void *arg) {
  check_job *j = (check_job *)arg;
  const orc_model *m = j->m;
  int want = (j->min_dist != NULL) || (j->min_pair != NULL);
  double *xpos = (double *)malloc(m->nbody * 3 * sizeof(double));
  double *xquat = (double *)malloc(m->nbody * 4 * sizeof(double));
  shape_t *shapes = (shape_t *)malloc((m->ngeom + 1) * sizeof(shape_t));
  double *bcen = (double *)malloc((m->ngeom + 1) * 3 * sizeof(double));
  for (;;) {
    int64_t lo = __atomic_fetch_add(j->next, CHUNK, __ATOMIC_RELAXED);
    if (lo >= j->n) break;
    int64_t hi = lo + CHUNK < j->n ? lo + CHUNK : j->n;
    for (int64_t i = lo; i < hi; i++)
      check_one(m, j->q + i * m->nq, j->flags, xpos, xquat, shapes, bcen, j->valid + i,
                j->min_dist ? j->min_dist + i : NULL, j->min_pair ? j->min_pair + i : NULL, want);
  }
  free(xpos); free(xquat); free(shapes); free(bcen);
  return NULL;
}

static int g_threads = 1;
void orc_set_threads(int n) { g_threads = n < 1 ? 1 : (n > 256 ? 256 : n); }
int orc_get_threads(void) { return g_threads; }

int orc_check(const orc_model *m, const double *q, int64_t n, uint32_t flags, uint8_t *valid,
              double *min_dist, int32_t *min_pair) {
  if ((flags & ORC_CHECK_LIMITS) && m->nq != m->njnt) {
    snprintf(g_err, sizeof g_err, "joint limits need nq == njnt");
    return 1;
  }
  int64_t next = 0;
  check_job job = {m, q, n, flags, valid, min_dist, min_pair, &next};
  int nt = g_threads;
  if (n < 2 * CHUNK) nt = 1;
  if (nt == 1) { check_worker(&job); return 0; }
  /* one worker per core, PINNED (north star: "reference path timed on the box's own host cores ...
   * threads pinned"): worker t runs on the t-th core of the calling thread's affinity mask */
  pthread_t th[256];
  cpu_set_t allowed;
  int cores[256], ncores = 0;
  if (sched_getaffinity(0, sizeof allowed, &allowed) == 0)
    for (int c = 0; c < CPU_SETSIZE && ncores < 256; c++)
      if (CPU_ISSET(c, &allowed)) cores[ncores++] = c;
  for (int t = 0; t < nt; t++) {
    pthread_attr_t at;
    pthread_attr_init(&at);
    if (ncores > 0) {
      cpu_set_t one;
      CPU_ZERO(&one);
      CPU_SET(cores[t % ncores], &one);
      pthread_attr_setaffinity_np(&at, sizeof one, &one);
    }
    if (pthread_create(&th[t], &at, check_worker, &job) != 0) pthread_create(&th[t], NULL, check_worker, &job);
    pthread_attr_destroy(&at);
  }
  for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
  return 0;
}

int orc_pair_distance(const orc_model *m, const double *q, int32_t pair, double *dist) {
  if (pair < 0 || pair >= m->npair) { snprintf(g_err, sizeof g_err, "bad pair index"); return 1; }
  double *xpos = (double *)malloc(m->nbody * 3 * sizeof(double));
  double *xquat = (double *)malloc(m->nbody * 4 * sizeof(double));
  shape_t a, b;
  fk_one(m, q, xpos, xquat);
  make_shape(m, xpos, xquat, m->pair_g1[pair], &a);
  make_shape(m, xpos, xquat, m->pair_g2[pair], &b);
  double margin = fmax(m->geom_margin[m->pair_g1[pair]], m->geom_margin[m->pair_g2[pair]]);
  *dist = pair_signed_distance(&a, &b) - margin;
  free(xpos); free(xquat);
  return 0;
}
