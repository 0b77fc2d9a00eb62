// vk_build.h -- host-side (C++) construction of the device tables from MjModel-named arrays.
//
// Restates MuJoCo's static collision filtering (engine_collision_driver.c: filterBodyPair,
// mj_contactFilter) and mjpl's allow-list rule (reference:
// src/mjpl/constraint/collision_constraint.py:42-64, 83-95 -- deleting allowed body pairs up
// front is equivalent to ignoring their contacts afterwards), then lowers every colliding geom
// to a sphere-swept vertex set / cylinder / plane in its body frame.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mjpl_b200.h"
#include "vk_core.cuh"

namespace vkb {

#ifndef VK_HILL
#define VK_HILL_HOST 0
#else
#define VK_HILL_HOST VK_HILL
#endif

using namespace vk;

enum { G_PLANE = 0, G_HFIELD = 1, G_SPHERE = 2, G_CAPSULE = 3, G_ELLIPSOID = 4, G_CYLINDER = 5, G_BOX = 6, G_MESH = 7 };
enum { J_FREE = 0, J_BALL = 1, J_SLIDE = 2, J_HINGE = 3 };

// direction grid of a hull for the inscribed-radius bound below
struct HullDirs { std::vector<V3<double>> c0, u0; std::vector<double> delta; };

struct HostModel {
  int nq = 0, nbody = 0, njnt = 0, ngeom = 0, nslot = 0;
  FkTables<double> fk;                  // indexed by SLOT (moving bodies only)
  std::vector<int> slot_body;           // slot -> body id
  std::vector<int> body_slot;           // body -> slot or -1
  std::vector<Pose<double>> static_pose;  // world pose of world-fixed bodies (by body id)
  std::vector<double> jnt_lo, jnt_hi;   // fp64 limits (reference compares in fp64)
  std::vector<Shape<double>> shapes;    // moving shapes first (sorted by slot), then static
  int nmoving_shapes = 0;
  std::vector<int> slot_shape_adr, slot_shape_num;
  std::vector<Vtx<double>> verts;
  std::vector<uint16_t> adj_start;      // per vertex (+1 sentinel per table end): offsets into adj; all zero w/o graphs
  std::vector<uint8_t> adj;             // local neighbour ids
  std::vector<Pair> pairs;              // processing order
  std::vector<int> pair_g1, pair_g2;    // MuJoCo geom ids, same order as `pairs`
  std::vector<double> pair_rsum64, pair_bsum64;  // fp64 copies of Pair::rsum / bsum
  // support maps (vk_core.cuh): per hull 6 x R x R cells (offset << 8 | count) and the candidate vertex ids
  std::vector<uint32_t> smap_cells;
  std::vector<uint8_t> smap_ids;
  std::vector<struct HullDirs> hull_dirs;   // per shape, build time only (inner shapes); cleared by build_groups
  double group_rin_any[MAX_GROUP] = {0};    // census only: largest inner ball of a member shape around the group centre
  // cull groups (vk_pipe.cuh): moving bodies that carry shapes, and every world-fixed shape by itself
  int ngroup_moving = 0;
  int slot_group_adr[MAX_BODY], slot_group_num[MAX_BODY];   // pose slot -> its moving groups (usually one; none if it carries no shape)
  double group_c[MAX_GROUP][3];          // bounding-sphere centre of a moving group, body frame
  double group_r[MAX_GROUP] = {0};       // its radius (covers the swept radii of the member shapes)
  std::vector<StaticGroup> static_groups;
  std::vector<GroupPair> group_pairs;    // sorted by kind, then by moving group
  std::vector<uint16_t> gp_member;       // pair indices (into `pairs`), grouped by group pair
  int gp_kind_end[3] = {0, 0, 0};        // group_pairs[0 .. end[0]) spheres, [end[0] .. end[1]) capsules, then planes
  double l0_sq_err = 0;                  // what level 0's squared limits carry for fp32 rounding (build_groups)
  double calib_l0_per_row = 0, calib_sub_per_row = 0;   // calibrated level-0 survivors / expanded shape pairs per row
  double bin_expect[NBIN] = {0};   // calibrated narrow-phase items per row that land in each bin (vk_split.cuh)
  int nrounds = 0;
  int round_start[MAX_ROUNDS + 1] = {0};  // pair index range of each round
  int round_gjk[MAX_ROUNDS] = {0};        // 1: the round holds GJK pairs, 0: plane / segment pairs
  double calib_sphere_per_row = 0, calib_items_per_row = 0, calib_pen_rows = 0;
  std::string err;
};

inline Q4<double> qd(const double *q) { Q4<double> r; r.w = q[0]; r.x = q[1]; r.y = q[2]; r.z = q[3]; return r; }
inline V3<double> vd(const double *v) { return mk<double>(v[0], v[1], v[2]); }

// symmetric 3x3 eigen decomposition (cyclic Jacobi); columns of V are eigenvectors
inline void jacobi3(double A[3][3], double V[3][3]) {
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) V[i][j] = i == j;
  for (int sweep = 0; sweep < 32; sweep++) {
    double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    if (off < 1e-18) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        if (fabs(A[p][q]) < 1e-300) continue;
        double th = (A[q][q] - A[p][p]) / (2 * A[p][q]);
        double t = (th >= 0 ? 1 : -1) / (fabs(th) + sqrt(th * th + 1));
        double c = 1 / sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < 3; k++) {
          double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; k++) {
          double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; k++) {
          double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
        }
      }
  }
}

inline void orthobasis_from_z(V3<double> z, double R[9]) {
  V3<double> a = fabs(z.x) < 0.9 ? mk<double>(1, 0, 0) : mk<double>(0, 1, 0);
  V3<double> x = cross(a, z);
  double n = sqrt(dot(x, x));
  x = x * (1.0 / n);
  V3<double> y = cross(z, x);
  R[0] = x.x; R[1] = y.x; R[2] = z.x;
  R[3] = x.y; R[4] = y.y; R[5] = z.y;
  R[6] = x.z; R[7] = y.z; R[8] = z.z;
}

// bounding sphere + OBB of a sphere-swept vertex set
inline void fit_bounds(Shape<double> &s, const std::vector<Vtx<double>> &verts) {
  const int n = s.nvert;
  const Vtx<double> *v = verts.data() + s.vadr;
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int i = 0; i < n; i++) {
    const double p[3] = {v[i].x, v[i].y, v[i].z};
    for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); }
  }
  // bounding sphere: start at the AABB centre, then a few Ritter-style shrink steps
  double c[3] = {0.5 * (lo[0] + hi[0]), 0.5 * (lo[1] + hi[1]), 0.5 * (lo[2] + hi[2])};
  auto radius_at = [&](const double *cc) {
    double r = 0;
    for (int i = 0; i < n; i++) {
      double dx = v[i].x - cc[0], dy = v[i].y - cc[1], dz = v[i].z - cc[2];
      r = std::max(r, sqrt(dx * dx + dy * dy + dz * dz));
    }
    return r;
  };
  double r = radius_at(c);
  for (int it = 0; it < 200 && n > 1; it++) {  // move towards the farthest point while it helps
    int far = 0; double fr = -1;
    for (int i = 0; i < n; i++) {
      double dx = v[i].x - c[0], dy = v[i].y - c[1], dz = v[i].z - c[2];
      double d = dx * dx + dy * dy + dz * dz;
      if (d > fr) { fr = d; far = i; }
    }
    double step = 0.05 / (1 + it * 0.1);
    double c2[3] = {c[0] + step * (v[far].x - c[0]), c[1] + step * (v[far].y - c[1]), c[2] + step * (v[far].z - c[2])};
    double r2 = radius_at(c2);
    if (r2 < r) { r = r2; memcpy(c, c2, sizeof c); }
  }
  s.bc[0] = c[0]; s.bc[1] = c[1]; s.bc[2] = c[2];
  s.brad = r + s.radius;
  // OBB
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (n == 2) {
    V3<double> d = mk<double>(v[1].x - v[0].x, v[1].y - v[0].y, v[1].z - v[0].z);
    double len = sqrt(dot(d, d));
    if (len > 1e-12) orthobasis_from_z(d * (1.0 / len), R);
  } else if (n >= 4) {
    double m[3] = {0, 0, 0};
    for (int i = 0; i < n; i++) { m[0] += v[i].x; m[1] += v[i].y; m[2] += v[i].z; }
    for (int k = 0; k < 3; k++) m[k] /= n;
    double C[3][3] = {{0}}, V[3][3];
    for (int i = 0; i < n; i++) {
      double d[3] = {v[i].x - m[0], v[i].y - m[1], v[i].z - m[2]};
      for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) C[a][b] += d[a] * d[b];
    }
    jacobi3(C, V);
    // make it a proper rotation
    V3<double> x = mk<double>(V[0][0], V[1][0], V[2][0]), y = mk<double>(V[0][1], V[1][1], V[2][1]);
    double nx = sqrt(dot(x, x)); x = x * (1.0 / nx);
    y = y - x * dot(x, y);
    double ny = sqrt(dot(y, y));
    if (ny > 1e-9) {
      y = y * (1.0 / ny);
      V3<double> z = cross(x, y);
      R[0] = x.x; R[1] = y.x; R[2] = z.x; R[3] = x.y; R[4] = y.y; R[5] = z.y; R[6] = x.z; R[7] = y.z; R[8] = z.z;
    }
  }
  // try both the PCA frame and the body-axis frame, keep the smaller volume
  double bestvol = 1e300;
  for (int cand = 0; cand < 2; cand++) {
    double Rc[9];
    if (cand == 0) memcpy(Rc, R, sizeof Rc);
    else { double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}; memcpy(Rc, I, sizeof Rc); }
    double l[3] = {1e300, 1e300, 1e300}, h[3] = {-1e300, -1e300, -1e300};
    for (int i = 0; i < n; i++)
      for (int k = 0; k < 3; k++) {
        double pr = Rc[k] * v[i].x + Rc[3 + k] * v[i].y + Rc[6 + k] * v[i].z;  // column k
        l[k] = std::min(l[k], pr); h[k] = std::max(h[k], pr);
      }
    double half[3], cen[3];
    for (int k = 0; k < 3; k++) { half[k] = 0.5 * (h[k] - l[k]) + s.radius; cen[k] = 0.5 * (h[k] + l[k]); }
    double vol = (half[0] + 1e-4) * (half[1] + 1e-4) * (half[2] + 1e-4);
    if (vol < bestvol) {
      bestvol = vol;
      for (int k = 0; k < 9; k++) s.orot[k] = Rc[k];
      for (int k = 0; k < 3; k++) s.ohalf[k] = half[k];
      for (int k = 0; k < 3; k++) s.oc[k] = Rc[3 * k] * cen[0] + Rc[3 * k + 1] * cen[1] + Rc[3 * k + 2] * cen[2];
    }
  }
}

// distance from point p to the segment a..b
inline double point_segment_dist(const double *p, const double *a, const double *b) {
  double ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, ap[3] = {p[0] - a[0], p[1] - a[1], p[2] - a[2]};
  double l2 = ab[0] * ab[0] + ab[1] * ab[1] + ab[2] * ab[2];
  double t = l2 > 0 ? (ap[0] * ab[0] + ap[1] * ab[1] + ap[2] * ab[2]) / l2 : 0.0;
  t = t < 0 ? 0 : (t > 1 ? 1 : t);
  double d[3] = {ap[0] - t * ab[0], ap[1] - t * ab[1], ap[2] - t * ab[2]};
  return sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
}

// Bounding capsule of a sphere-swept vertex set: axis = principal direction of the vertices (or a
// given direction), through the centroid or through the mid-point of the perpendicular extent; for a
// few radius factors the shortest segment that keeps every vertex within that radius; the smallest
// volume wins.  The radius is finally raised to the largest vertex distance actually found, so the
// capsule contains the set by construction, whatever the heuristics above did.
inline void fit_capsule(Shape<double> &s, const std::vector<Vtx<double>> &verts) {
  const int n = s.nvert;
  const Vtx<double> *v = verts.data() + s.vadr;
  if (n == 1) {
    s.ca[0] = s.cb[0] = v[0].x; s.ca[1] = s.cb[1] = v[0].y; s.ca[2] = s.cb[2] = v[0].z;
    s.crad = s.radius; s.caplen = 0;
    return;
  }
  double m[3] = {0, 0, 0};
  for (int i = 0; i < n; i++) { m[0] += v[i].x; m[1] += v[i].y; m[2] += v[i].z; }
  for (int k = 0; k < 3; k++) m[k] /= n;
  double C[3][3] = {{0}}, V[3][3];
  for (int i = 0; i < n; i++) {
    double d[3] = {v[i].x - m[0], v[i].y - m[1], v[i].z - m[2]};
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) C[a][b] += d[a] * d[b];
  }
  jacobi3(C, V);
  int kmax = 0;
  for (int k = 1; k < 3; k++) if (C[k][k] > C[kmax][kmax]) kmax = k;
  double u[3] = {V[0][kmax], V[1][kmax], V[2][kmax]};
  double un = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
  if (!(un > 1e-12)) { u[0] = 1; u[1] = u[2] = 0; un = 1; }
  for (int k = 0; k < 3; k++) u[k] /= un;
  // mid-point of the perpendicular extent (AABB of the vertices projected on the plane normal to u)
  double plo[3] = {1e300, 1e300, 1e300}, phi[3] = {-1e300, -1e300, -1e300};
  for (int i = 0; i < n; i++) {
    double d[3] = {v[i].x - m[0], v[i].y - m[1], v[i].z - m[2]};
    double t = d[0] * u[0] + d[1] * u[1] + d[2] * u[2];
    for (int k = 0; k < 3; k++) { double pk = d[k] - t * u[k]; plo[k] = std::min(plo[k], pk); phi[k] = std::max(phi[k], pk); }
  }
  double best = 1e300;
  for (int cand = 0; cand < 2; cand++) {
    double c0[3];
    for (int k = 0; k < 3; k++) c0[k] = m[k] + (cand ? 0.5 * (plo[k] + phi[k]) : 0.0);
    std::vector<double> t(n), ri(n);
    double rmax = 0;
    for (int i = 0; i < n; i++) {
      double d[3] = {v[i].x - c0[0], v[i].y - c0[1], v[i].z - c0[2]};
      t[i] = d[0] * u[0] + d[1] * u[1] + d[2] * u[2];
      double pr[3] = {d[0] - t[i] * u[0], d[1] - t[i] * u[1], d[2] - t[i] * u[2]};
      ri[i] = sqrt(pr[0] * pr[0] + pr[1] * pr[1] + pr[2] * pr[2]);
      rmax = std::max(rmax, ri[i]);
    }
    const double scales[6] = {1.0, 1.05, 1.1, 1.2, 1.35, 1.5};
    for (double sc : scales) {
      double r = std::max(rmax * sc, 1e-9);
      double lo = 1e300, hi = -1e300;
      for (int i = 0; i < n; i++) {
        double h = sqrt(std::max(r * r - ri[i] * ri[i], 0.0));
        lo = std::min(lo, t[i] + h); hi = std::max(hi, t[i] - h);
      }
      if (lo > hi) lo = hi = 0.5 * (lo + hi);
      double a[3], b[3];
      for (int k = 0; k < 3; k++) { a[k] = c0[k] + lo * u[k]; b[k] = c0[k] + hi * u[k]; }
      for (int i = 0; i < n; i++) { double p[3] = {v[i].x, v[i].y, v[i].z}; r = std::max(r, point_segment_dist(p, a, b)); }
      double len = hi - lo;
      double vol = M_PI * r * r * len + 4.0 / 3.0 * M_PI * r * r * r;
      if (vol < best) {
        best = vol;
        for (int k = 0; k < 3; k++) { s.ca[k] = a[k]; s.cb[k] = b[k]; }
        s.crad = r * (1 + 1e-9) + 1e-12 + s.radius; s.caplen = len;
      }
    }
  }
  if (s.caplen < 1e-6) {   // degenerate segment: a sphere; re-centre on the segment's mid-point
    for (int k = 0; k < 3; k++) s.ca[k] = s.cb[k] = 0.5 * (s.ca[k] + s.cb[k]);
    double r = 0;
    for (int i = 0; i < n; i++) { double p[3] = {v[i].x, v[i].y, v[i].z}; r = std::max(r, point_segment_dist(p, s.ca, s.cb)); }
    s.crad = r * (1 + 1e-9) + 1e-12 + s.radius; s.caplen = 0;
  }
}


// ---- inner shapes: the certain-contact shortcut ------------------------------------------------------
// A ball B(c, r) lies inside the hull of the vertices iff h(d) - c.d >= r for every unit direction d
// (h = support function).  The directions are covered by the cells of a cube map (centre c0, every
// direction of the cell within delta of c0, u0 = a support vertex for c0); for d in the cell
//   h(d) - c.d >= (u0 - c).d >= (u0 - c).c0 - |u0 - c| delta,
// so the minimum of the right-hand side over the cells is a rigorous lower bound of the inscribed radius
// around c.  It is concave in c (a minimum of concave functions), which the searches below rely on.
inline void hull_dirs_build(const Vtx<double> *v, int n, HullDirs &D, int R = 32) {
  D.c0.clear(); D.u0.clear(); D.delta.clear();
  for (int f = 0; f < 6; f++) {
    const int k = f / 2;
    const double sg = (f % 2) ? -1.0 : 1.0;
    for (int iu = 0; iu < R; iu++)
      for (int iv = 0; iv < R; iv++) {
        auto dir = [&](double u, double w) {
          double d[3];
          d[k] = sg; d[(k + 1) % 3] = u; d[(k + 2) % 3] = w;
          const double nn = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
          return mk<double>(d[0] / nn, d[1] / nn, d[2] / nn);
        };
        const double u0 = -1 + 2.0 * iu / R, u1 = -1 + 2.0 * (iu + 1) / R, w0 = -1 + 2.0 * iv / R, w1 = -1 + 2.0 * (iv + 1) / R;
        const V3<double> c = dir(0.5 * (u0 + u1), 0.5 * (w0 + w1));
        double delta = 0;
        for (int a = 0; a < 2; a++)
          for (int b = 0; b < 2; b++) {
            const V3<double> q = dir(a ? u1 : u0, b ? w1 : w0) - c;
            delta = std::max(delta, sqrt(dot(q, q)));
          }
        int best = 0; double hb = -1e300;
        for (int i = 0; i < n; i++) { const double h = v[i].x * c.x + v[i].y * c.y + v[i].z * c.z; if (h > hb) { hb = h; best = i; } }
        D.c0.push_back(c); D.u0.push_back(mk<double>(v[best].x, v[best].y, v[best].z)); D.delta.push_back(delta * 1.02 + 1e-6);
      }
  }
}
inline double inscribed_radius(const HullDirs &D, V3<double> c) {
  double r = 1e300;
  for (size_t i = 0; i < D.c0.size(); i++) {
    const V3<double> w = D.u0[i] - c;
    r = std::min(r, dot(w, D.c0[i]) - sqrt(dot(w, w)) * D.delta[i]);
  }
  return r;
}
// radius of the largest ball around c (shape frame) known to lie inside the shape; <= 0: none
inline double inner_ball(const Shape<double> &s, const std::vector<Vtx<double>> &verts, const HullDirs *D, V3<double> c) {
  if (s.kind == SK_CYL) {
    const V3<double> e = c - mk<double>(s.c[0], s.c[1], s.c[2]), ax = mk<double>(s.ax[0], s.ax[1], s.ax[2]);
    const double t = dot(e, ax);
    const V3<double> pr = e - ax * t;
    return std::min(s.radius - sqrt(dot(pr, pr)), s.halflen - fabs(t));
  }
  if (s.kind != SK_VERTS) return -1;
  const Vtx<double> *v = verts.data() + s.vadr;
  if (s.nvert <= 2) {
    const double p[3] = {c.x, c.y, c.z}, a[3] = {v[0].x, v[0].y, v[0].z};
    const double b[3] = {v[s.nvert - 1].x, v[s.nvert - 1].y, v[s.nvert - 1].z};
    return s.radius - point_segment_dist(p, a, b);
  }
  if (!D || D->c0.empty()) return -1;
  const double r = inscribed_radius(*D, c);
  return r >= 0 ? r + s.radius : -1;
}
// the longest sub-segment m +- t u (0 <= t <= tmax) on which inner_ball >= target (bisection; inner_ball is concave)
template <typename F> inline double inner_extent(F ball, V3<double> m, V3<double> u, double tmax, double target, bool symmetric, int sign) {
  auto ok = [&](double t) {
    if (symmetric) return ball(m + u * t) >= target && ball(m - u * t) >= target;
    return ball(m + u * (sign * t)) >= target;
  };
  if (ok(tmax)) return tmax;
  double lo = 0, hi = tmax;
  for (int it = 0; it < 14; it++) { const double mid = 0.5 * (lo + hi); if (ok(mid)) lo = mid; else hi = mid; }
  return lo;
}
// Inner capsule of a shape: centre of a large inscribed ball (pattern search from a few candidates), then for a
// few radius factors the longest segment along the bounding capsule's axis; the largest volume wins.
inline void fit_inner(Shape<double> &s, const std::vector<Vtx<double>> &verts, HullDirs &D) {
  for (int k = 0; k < 3; k++) s.ia[k] = s.ib[k] = 0;
  s.irad = 0;
  if (s.kind == SK_CYL) {
    const double r = std::min(s.radius, s.halflen), h = std::max(s.halflen - r, 0.0);
    for (int k = 0; k < 3; k++) { s.ia[k] = s.c[k] - h * s.ax[k]; s.ib[k] = s.c[k] + h * s.ax[k]; }
    s.irad = r;
    return;
  }
  if (s.kind != SK_VERTS) return;
  const Vtx<double> *v = verts.data() + s.vadr;
  if (s.nvert <= 2) {
    s.ia[0] = v[0].x; s.ia[1] = v[0].y; s.ia[2] = v[0].z;
    s.ib[0] = v[s.nvert - 1].x; s.ib[1] = v[s.nvert - 1].y; s.ib[2] = v[s.nvert - 1].z;
    s.irad = s.radius;
    return;
  }
  hull_dirs_build(v, s.nvert, D);
  auto ball = [&](V3<double> c) { return inner_ball(s, verts, &D, c); };
  V3<double> cen = mk<double>(0, 0, 0), lo = mk<double>(1e300, 1e300, 1e300), hi = mk<double>(-1e300, -1e300, -1e300);
  for (int i = 0; i < s.nvert; i++) {
    cen = cen + mk<double>(v[i].x, v[i].y, v[i].z) * (1.0 / s.nvert);
    lo.x = std::min(lo.x, v[i].x); lo.y = std::min(lo.y, v[i].y); lo.z = std::min(lo.z, v[i].z);
    hi.x = std::max(hi.x, v[i].x); hi.y = std::max(hi.y, v[i].y); hi.z = std::max(hi.z, v[i].z);
  }
  const V3<double> cand[4] = {cen, (lo + hi) * 0.5, mk<double>(s.bc[0], s.bc[1], s.bc[2]),
                              (mk<double>(s.ca[0], s.ca[1], s.ca[2]) + mk<double>(s.cb[0], s.cb[1], s.cb[2])) * 0.5};
  V3<double> c = cand[0];
  double r = -1e300;
  for (const V3<double> &cc : cand) { const double rr = ball(cc); if (rr > r) { r = rr; c = cc; } }
  double step = 0.25 * sqrt(dot(hi - lo, hi - lo));
  for (int it = 0; it < 16; it++, step *= 0.6) {
    for (int a = 0; a < 3; a++)
      for (int sg = -1; sg <= 1; sg += 2) {
        V3<double> c2 = c;
        (a == 0 ? c2.x : (a == 1 ? c2.y : c2.z)) += sg * step;
        const double r2 = ball(c2);
        if (r2 > r) { r = r2; c = c2; }
      }
  }
  if (!(r > 0)) return;
  V3<double> a0 = c, b0 = c;
  double best = 4.0 / 3.0 * M_PI * r * r * r, rbest = r;
  if (s.caplen > 1e-9) {
    const V3<double> u = (mk<double>(s.cb[0], s.cb[1], s.cb[2]) - mk<double>(s.ca[0], s.ca[1], s.ca[2])) * (1.0 / s.caplen);
    const double fac[4] = {0.95, 0.85, 0.7, 0.55};
    for (double f : fac) {
      const double rr = f * r;
      const double tp = inner_extent(ball, c, u, s.caplen, rr, false, +1), tm = inner_extent(ball, c, u, s.caplen, rr, false, -1);
      const double vol = M_PI * rr * rr * (tp + tm) + 4.0 / 3.0 * M_PI * rr * rr * rr;
      if (vol > best) { best = vol; rbest = rr; a0 = c - u * tm; b0 = c + u * tp; }
    }
  }
  s.ia[0] = a0.x; s.ia[1] = a0.y; s.ia[2] = a0.z;
  s.ib[0] = b0.x; s.ib[1] = b0.y; s.ib[2] = b0.z;
  s.irad = rbest;
}

template <typename T> inline FkTables<T> convert_fk(const FkTables<double> &s) {
  FkTables<T> o; memset(&o, 0, sizeof o);
  o.nq = s.nq; o.nbody = s.nbody; o.njnt = s.njnt;
  for (int i = 0; i < MAX_BODY; i++) {
    o.body_parent[i] = s.body_parent[i]; o.body_jntadr[i] = s.body_jntadr[i];
    o.body_jntnum[i] = s.body_jntnum[i]; o.body_slot[i] = s.body_slot[i];
    for (int k = 0; k < 3; k++) o.body_pos[i][k] = (T)s.body_pos[i][k];
    for (int k = 0; k < 4; k++) o.body_quat[i][k] = (T)s.body_quat[i][k];
  }
  for (int j = 0; j < MAX_JNT; j++) {
    o.jnt_type[j] = s.jnt_type[j]; o.jnt_qadr[j] = s.jnt_qadr[j];
    for (int k = 0; k < 3; k++) { o.jnt_pos[j][k] = (T)s.jnt_pos[j][k]; o.jnt_axis[j][k] = (T)s.jnt_axis[j][k]; }
    o.jnt_lo[j] = (T)s.jnt_lo[j]; o.jnt_hi[j] = (T)s.jnt_hi[j]; o.qpos0[j] = (T)s.qpos0[j];
  }
  return o;
}

template <typename T> inline Shape<T> convert_shape(const Shape<double> &s) {
  Shape<T> o; memset(&o, 0, sizeof o);
  o.kind = s.kind; o.slot = s.slot; o.vadr = s.vadr; o.nvert = s.nvert; o.geom = s.geom;
  o.graph = s.graph; memcpy(o.ext, s.ext, sizeof o.ext);
  o.radius = (T)s.radius; o.halflen = (T)s.halflen; o.brad = (T)s.brad;
  for (int k = 0; k < 3; k++) { o.c[k] = (T)s.c[k]; o.ax[k] = (T)s.ax[k]; o.bc[k] = (T)s.bc[k]; o.oc[k] = (T)s.oc[k]; o.ohalf[k] = (T)s.ohalf[k]; }
  for (int k = 0; k < 9; k++) o.orot[k] = (T)s.orot[k];
  o.group = s.group;
  for (int k = 0; k < 3; k++) { o.ca[k] = (T)s.ca[k]; o.cb[k] = (T)s.cb[k]; }
  // rounding the end points can move them by half an ulp: the radius absorbs it
  o.crad = (T)(s.crad * (1.0 + 2e-7) + (sizeof(T) == 4 ? 1e-6 : 0.0)); o.caplen = (T)s.caplen;
  o.map = s.map;
  // inner capsule: rounding moves the end points by half an ulp, the radius gives that up (and a little more)
  for (int k = 0; k < 3; k++) { o.ia[k] = (T)s.ia[k]; o.ib[k] = (T)s.ib[k]; }
  o.irad = s.irad > 0 ? (T)(s.irad * (1.0 - 2e-7) - (sizeof(T) == 4 ? 1e-6 : 0.0)) : (T)0;
  return o;
}

// ---- cull groups and group pairs (level 0 of the pipeline's broad phase, vk_pipe.cuh) ---------------
// A moving body with all its shapes is one group (bounding sphere in the body frame); a world-fixed
// shape is a group of its own (bounding capsule / plane in the world frame: the reference scenes'
// obstacles are long thin boxes and capsules, for which a sphere is a useless bound).  Every shape pair
// belongs to exactly one group pair; a group pair that passes the level-0 test is expanded into its
// shape pairs, each of which then faces the bounding-capsule test and the OBB test.
inline bool build_groups(HostModel &H) {
  const double slack = 1e-4;
  for (int s = 0; s < MAX_BODY; s++) { H.slot_group_adr[s] = 0; H.slot_group_num[s] = 0; }
  H.ngroup_moving = 0;
  // bounding sphere of the shapes [s0, s1) (all on one body): centre by a shrinking axis search, radius =
  // farthest vertex (+ swept radius) of any member
  auto fit_group = [&](int s0, int s1, double *c, double &r) {
    auto radius_at = [&](const double *cc) {
      double rr = 0;
      for (int k = s0; k < s1; k++) {
        const Shape<double> &sh = H.shapes[k];
        if (sh.kind == SK_VERTS) {
          for (int i = 0; i < sh.nvert; i++) {
            const Vtx<double> &v = H.verts[sh.vadr + i];
            rr = std::max(rr, sqrt((v.x - cc[0]) * (v.x - cc[0]) + (v.y - cc[1]) * (v.y - cc[1]) + (v.z - cc[2]) * (v.z - cc[2])) + sh.radius);
          }
        } else {
          rr = std::max(rr, sqrt((sh.bc[0] - cc[0]) * (sh.bc[0] - cc[0]) + (sh.bc[1] - cc[1]) * (sh.bc[1] - cc[1]) + (sh.bc[2] - cc[2]) * (sh.bc[2] - cc[2])) + sh.brad);
        }
      }
      return rr;
    };
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int k = s0; k < s1; k++)
      for (int a = 0; a < 3; a++) {
        lo[a] = std::min(lo[a], H.shapes[k].bc[a] - H.shapes[k].brad);
        hi[a] = std::max(hi[a], H.shapes[k].bc[a] + H.shapes[k].brad);
      }
    for (int a = 0; a < 3; a++) c[a] = 0.5 * (lo[a] + hi[a]);
    r = radius_at(c);
    for (int it = 0; it < 300; it++) {   // shrink: moves along the axes while they help
      const double step = 0.25 * r / (1 + it * 0.1);
      bool moved = false;
      for (int a = 0; a < 3 && !moved; a++)
        for (int sg = -1; sg <= 1 && !moved; sg += 2) {
          double c2[3] = {c[0], c[1], c[2]};
          c2[a] += sg * step;
          const double r2 = radius_at(c2);
          if (r2 < r) { r = r2; memcpy(c, c2, sizeof(double) * 3); moved = true; }
        }
    }
  };
  for (int sl = 0; sl < H.nslot; sl++) {
    if (H.slot_shape_num[sl] == 0) continue;
    const int s0 = H.slot_shape_adr[sl], s1 = s0 + H.slot_shape_num[sl];
    double c[3], r;
    fit_group(s0, s1, c, r);
    // A body whose shapes are spread out (Franka link5: three meshes along the link, union sphere 0.17 m
    // against 0.08-0.10 m for the members) makes a useless group: its pairs survive level 0 all the time
    // and each survival expands into every member pair.  Such a body gets one group per shape instead.
    double rmax = 0;
    for (int k = s0; k < s1; k++) rmax = std::max(rmax, H.shapes[k].brad);
    const bool split = (s1 - s0) > 1 && r > 1.4 * rmax && H.ngroup_moving + (s1 - s0) <= MAX_GROUP;
    H.slot_group_adr[sl] = H.ngroup_moving;
    for (int k0 = s0; k0 < s1; k0 = split ? k0 + 1 : s1) {
      const int k1 = split ? k0 + 1 : s1;
      if (split) fit_group(k0, k1, c, r);
      if (H.ngroup_moving >= MAX_GROUP) { H.err = "too many cull groups"; return false; }
      const int g = H.ngroup_moving++;
      H.slot_group_num[sl]++;
      for (int a = 0; a < 3; a++) H.group_c[g][a] = c[a];
      H.group_r[g] = r * (1 + 1e-6) + 1e-6;   // covers the fp32 rounding of centre and vertices
      for (int k = k0; k < k1; k++) H.shapes[k].group = g;
    }
  }
  H.static_groups.clear();
  std::vector<double> tube_r(H.shapes.size(), 0.0);
  for (size_t k = (size_t)H.nmoving_shapes; k < H.shapes.size(); k++) {
    Shape<double> &sh = H.shapes[k];
    StaticGroup sg; memset(&sg, 0, sizeof sg);
    if (sh.kind == SK_PLANE) {
      for (int a = 0; a < 3; a++) { sg.a[a] = (float)sh.c[a]; sg.u[a] = (float)sh.ax[a]; }
    } else {
      double l2 = 0;
      for (int a = 0; a < 3; a++) l2 += (sh.cb[a] - sh.ca[a]) * (sh.cb[a] - sh.ca[a]);
      const double len = sqrt(l2);
      for (int a = 0; a < 3; a++) { sg.a[a] = (float)sh.ca[a]; sg.u[a] = len > 1e-6 ? (float)((sh.cb[a] - sh.ca[a]) / len) : 0.f; }
      sg.len = len > 1e-6 ? (float)len : 0.f;
    }
    // inner tube: a piece of the bounding capsule's segment, symmetric about its mid-point, and a radius such
    // that the swept piece lies inside the shape
    sg.th = -1.f;
    tube_r[k] = 0;
    if (sh.kind != SK_PLANE && sh.irad > 0) {
      const HullDirs *D = k < H.hull_dirs.size() ? &H.hull_dirs[k] : nullptr;
      auto ball = [&](V3<double> c) { return inner_ball(sh, H.verts, D, c); };
      const V3<double> ca = mk<double>(sh.ca[0], sh.ca[1], sh.ca[2]), cb = mk<double>(sh.cb[0], sh.cb[1], sh.cb[2]);
      const V3<double> m = (ca + cb) * 0.5;
      const double rm = ball(m);
      if (rm > 0) {
        if (!(sh.caplen > 1e-6)) { sg.th = 0.f; tube_r[k] = rm; }
        else {
          const V3<double> u = (cb - ca) * (1.0 / sh.caplen);
          double best = -1;
          const double fac[5] = {0.98, 0.9, 0.75, 0.6, 0.45};
          for (double f : fac) {
            const double t = inner_extent(ball, m, u, 0.5 * sh.caplen, f * rm, true, 0);
            const double vol = M_PI * f * rm * f * rm * 2 * t + 4.0 / 3.0 * M_PI * f * rm * f * rm * f * rm;
            if (vol > best) { best = vol; tube_r[k] = f * rm; sg.th = (float)(t * (1 - 1e-6)); }
          }
        }
      }
    }
    sh.group = H.ngroup_moving + (int)H.static_groups.size();
    H.static_groups.push_back(sg);
  }
  // inner ball of every moving shape around ITS GROUP'S centre (the point level 0 knows in the world frame)
  std::vector<double> gin(H.shapes.size(), 0.0);
  for (int g = 0; g < MAX_GROUP; g++) H.group_rin_any[g] = 0;
  for (int k = 0; k < H.nmoving_shapes; k++) {
    const Shape<double> &sh = H.shapes[k];
    if (sh.group < 0 || !(sh.irad > 0)) continue;
    const int g = sh.group;
    gin[k] = inner_ball(sh, H.verts, (size_t)k < H.hull_dirs.size() ? &H.hull_dirs[k] : nullptr,
                        mk<double>(H.group_c[g][0], H.group_c[g][1], H.group_c[g][2]));
    H.group_rin_any[g] = std::max(H.group_rin_any[g], gin[k]);
  }
  // group pairs
  struct Key { int kind, ga, gb; };
  std::vector<Key> keys;
  std::vector<std::vector<uint16_t>> members;
  std::vector<double> lims, lims_in;
  const double inner_safety = 1e-4;   // fp32 forward kinematics, centres and the expanded squares of level 0: a few 1e-6
  for (size_t p = 0; p < H.pairs.size(); p++) {
    const Pair &pr = H.pairs[p];
    const Shape<double> &A = H.shapes[pr.sa], &B = H.shapes[pr.sb];
    if (A.slot < 0 && B.slot < 0) { H.err = "a pair of two world-fixed geoms survived the filters"; return false; }
    const double margin = H.pair_rsum64[p] - (A.kind == SK_VERTS ? A.radius : 0.0) - (B.kind == SK_VERTS ? B.radius : 0.0);
    Key k;
    double lim, lim_in = -1e30;   // lim_in: the pair's inner balls / tube around the centres level 0 uses
    if (A.slot >= 0 && B.slot >= 0) {
      k.kind = GK_SPHERE; k.ga = std::min(A.group, B.group); k.gb = std::max(A.group, B.group);
      lim = H.group_r[A.group] + H.group_r[B.group] + margin + slack;
      if (gin[pr.sa] > 0 && gin[pr.sb] > 0) lim_in = gin[pr.sa] + gin[pr.sb] - inner_safety;
    } else {
      const Shape<double> &M = A.slot >= 0 ? A : B, &S = A.slot >= 0 ? B : A;
      const int im = A.slot >= 0 ? pr.sa : pr.sb, is = A.slot >= 0 ? pr.sb : pr.sa;
      k.kind = S.kind == SK_PLANE ? GK_PLANE : GK_CAPSULE;
      k.ga = M.group; k.gb = S.group - H.ngroup_moving;
      lim = H.group_r[M.group] + (S.kind == SK_PLANE ? 0.0 : S.crad * (1 + 2e-7) + 1e-6) + margin + slack;
      if (S.kind == SK_PLANE) { if (gin[im] > inner_safety) lim_in = gin[im] - inner_safety; }
      else if (gin[im] > 0 && tube_r[is] > 0) lim_in = gin[im] + tube_r[is] - inner_safety;
    }
    size_t idx = 0;
    for (; idx < keys.size(); idx++) if (keys[idx].kind == k.kind && keys[idx].ga == k.ga && keys[idx].gb == k.gb) break;
    if (idx == keys.size()) { keys.push_back(k); members.emplace_back(); lims.push_back(lim); lims_in.push_back(lim_in); }
    members[idx].push_back((uint16_t)p);
    lims[idx] = std::max(lims[idx], lim);
    lims_in[idx] = std::max(lims_in[idx], lim_in);
  }
  // Rounding of level 0's squared distances: |e - s u|^2 is evaluated as |e|^2 - s (2 e.u - s) in fp32
  // (point_segment_d2), which costs a few ulps of |e|^2 -- measured 2.1 eps32 |e|^2 (hs_point_segment_check).  |e| is
  // bounded by the reach of the kinematic tree (link offsets, hinge anchors, slide ranges: rotations keep lengths)
  // plus the far end of the world-fixed segments; the squared limits carry 8 eps32 x that bound squared.
  double reach = 0;
  {
    std::vector<double> r(std::max(H.nslot, 1), 0.0);
    for (int sl = 0; sl < H.nslot; sl++) {
      const int ps = H.fk.body_parent[sl];
      double a = ps >= 0 ? r[ps] : 0.0;
      a += sqrt(H.fk.body_pos[sl][0] * H.fk.body_pos[sl][0] + H.fk.body_pos[sl][1] * H.fk.body_pos[sl][1] + H.fk.body_pos[sl][2] * H.fk.body_pos[sl][2]);
      for (int j = H.fk.body_jntadr[sl]; j < H.fk.body_jntadr[sl] + H.fk.body_jntnum[sl]; j++) {
        if (H.fk.jnt_type[j] == JK_SLIDE) a += std::max(fabs(H.fk.jnt_lo[j] - H.fk.qpos0[H.fk.jnt_qadr[j]]), fabs(H.fk.jnt_hi[j] - H.fk.qpos0[H.fk.jnt_qadr[j]]));   // (rows within the joint range)
        else a += 2.0 * sqrt(H.fk.jnt_pos[j][0] * H.fk.jnt_pos[j][0] + H.fk.jnt_pos[j][1] * H.fk.jnt_pos[j][1] + H.fk.jnt_pos[j][2] * H.fk.jnt_pos[j][2]);
      }
      r[sl] = a;
      for (int g = H.slot_group_adr[sl]; g < H.slot_group_adr[sl] + H.slot_group_num[sl]; g++)
        reach = std::max(reach, a + sqrt(H.group_c[g][0] * H.group_c[g][0] + H.group_c[g][1] * H.group_c[g][1] + H.group_c[g][2] * H.group_c[g][2]));
    }
  }
  double far_static = 0;
  for (const StaticGroup &sg : H.static_groups)
    far_static = std::max(far_static, sqrt((double)sg.a[0] * sg.a[0] + (double)sg.a[1] * sg.a[1] + (double)sg.a[2] * sg.a[2]) + (double)sg.len);
  const double e_max = reach + far_static;
  const double sq_err = 8.0 * 5.97e-8 * e_max * e_max + 1e-9;
  H.l0_sq_err = sq_err;
  std::vector<int> ord(keys.size());
  for (size_t i = 0; i < ord.size(); i++) ord[i] = (int)i;
  std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) {
    if (keys[x].kind != keys[y].kind) return keys[x].kind < keys[y].kind;
    if (keys[x].ga != keys[y].ga) return keys[x].ga < keys[y].ga;
    return keys[x].gb < keys[y].gb;
  });
  H.group_pairs.clear(); H.gp_member.clear();
  H.gp_kind_end[0] = H.gp_kind_end[1] = H.gp_kind_end[2] = 0;
  for (int i : ord) {
    GroupPair g; memset(&g, 0, sizeof g);
    g.ga = (uint16_t)keys[i].ga; g.gb = (uint16_t)keys[i].gb; g.kind = (uint8_t)keys[i].kind;
    if (H.gp_member.size() + members[i].size() > 65535) { H.err = "too many geom pairs"; return false; }
    if (members[i].size() > 255) { H.err = "too many geom pairs between two bodies"; return false; }
    g.first = (uint16_t)H.gp_member.size(); g.n = (uint8_t)members[i].size();
    const bool squared = keys[i].kind != GK_PLANE;
    g.lim = (float)((squared ? lims[i] * lims[i] + sq_err : lims[i]) * (1 + 4e-7));
    // sphere / capsule kinds compare squares: "none" is 0 there; the plane kind compares the signed height
    // (the inner limit gives the same amount up; a sum that leaves nothing means: no certain-contact test)
    const bool none = !(lims_in[i] > 0) || (squared && !(lims_in[i] * lims_in[i] > 2.0 * sq_err)) || getenv("MJB_NO_INNER0");
    g.lim_in = none ? (squared ? 0.f : -1e30f) : (float)((squared ? lims_in[i] * lims_in[i] - sq_err : lims_in[i]) * (1 - 4e-7));
    for (uint16_t p : members[i]) H.gp_member.push_back(p);
    H.group_pairs.push_back(g);
    for (int k = keys[i].kind; k < 3; k++) H.gp_kind_end[k]++;
  }
  if (H.group_pairs.size() > 2047) { H.err = "too many body pairs"; return false; }   // 11 bits in a level-0 queue entry
  // calibration of the level-0 output (buffer sizing only): the same fp32 test on seeded rows
  {
    FkTables<float> fk32 = convert_fk<float>(H.fk);
    Pose<float> ident; ident.p = mk<float>(0, 0, 0); ident.q.w = 1; ident.q.x = ident.q.y = ident.q.z = 0;
    const int NCAL = 1024;
    double n0 = 0, nsub = 0;
    for (int r = 0; r < NCAL; r++) {
      float q[MAX_JNT];
      for (int j = 0; j < H.nq; j++) {
        float lo = (float)H.jnt_lo[j], hi = (float)H.jnt_hi[j];
        if (!(hi > lo)) { lo = -3.14159f; hi = 3.14159f; }
        q[j] = sweep_value(0x5eedull, (uint64_t)r, (uint32_t)j, lo, hi);
      }
      Pose<float> P[MAX_BODY];
      V3<float> cen[MAX_GROUP];
      for (int k = 0; k < H.nslot; k++) {
        int ps = fk32.body_parent[k];
        P[k] = fk_body(fk32, k, ps < 0 ? ident : P[ps], q);
        for (int g = H.slot_group_adr[k]; g < H.slot_group_adr[k] + H.slot_group_num[k]; g++)
          cen[g] = P[k].p + qrot(P[k].q, mk<float>((float)H.group_c[g][0], (float)H.group_c[g][1], (float)H.group_c[g][2]));
      }
      for (const GroupPair &g : H.group_pairs)
        if (group_pair_near(g, cen[g.ga], g.kind == GK_SPHERE ? cen[g.gb] : cen[g.ga],
                            g.kind == GK_SPHERE ? nullptr : &H.static_groups[g.gb])) { n0 += 1; nsub += g.n; }
    }
    H.calib_l0_per_row = n0 / NCAL; H.calib_sub_per_row = nsub / NCAL;
  }
  H.hull_dirs.clear(); H.hull_dirs.shrink_to_fit();
  return true;
}

inline bool build_host_model(const mjb_model_desc *d, HostModel &H) {
  H.nq = d->nq; H.nbody = d->nbody; H.njnt = d->njnt; H.ngeom = d->ngeom;
  if (d->nbody < 1 || d->nbody > 4096 || d->njnt < 0 || d->ngeom < 0) { H.err = "bad model sizes"; return false; }
  if (d->njnt > MAX_JNT) { H.err = "too many joints (max 32)"; return false; }
  int nq_expected = 0;
  for (int j = 0; j < d->njnt; j++) {
    if (d->jnt_type[j] != J_HINGE && d->jnt_type[j] != J_SLIDE) {
      // mjpl itself only plans for hinge/slide joints (reference README.md:19-20)
      H.err = "only hinge and slide joints are supported (joint " + std::to_string(j) + ")";
      return false;
    }
    nq_expected++;
  }
  if (nq_expected != d->nq) { H.err = "nq does not match the joint list"; return false; }

  // ---- moving bodies get pose slots; world-fixed bodies get constant world poses ------------
  H.body_slot.assign(d->nbody, -1);
  H.static_pose.resize(d->nbody);
  Pose<double> ident; ident.p = mk<double>(0, 0, 0); ident.q.w = 1; ident.q.x = ident.q.y = ident.q.z = 0;
  H.static_pose[0] = ident;
  memset(&H.fk, 0, sizeof H.fk);
  H.fk.nq = d->nq; H.fk.njnt = d->njnt;
  for (int b = 1; b < d->nbody; b++) {
    int p = d->body_parentid[b];
    if (p < 0 || p >= b) { H.err = "bodies must be ordered parent-first"; return false; }
    bool moving = d->body_weldid[b] != 0;
    Pose<double> local; local.p = vd(d->body_pos + 3 * b); local.q = qnormalize(qd(d->body_quat + 4 * b));
    if (!moving) {
      if (d->body_jntnum[b] != 0) { H.err = "inconsistent body_weldid"; return false; }
      const Pose<double> &P = H.static_pose[p];
      H.static_pose[b].p = P.p + qrot(P.q, local.p);
      H.static_pose[b].q = qnormalize(qmul(P.q, local.q));
      continue;
    }
    int s = H.nslot++;
    if (s >= MAX_BODY) { H.err = "too many moving bodies (max 32)"; return false; }
    H.body_slot[b] = s;
    H.slot_body.push_back(b);
    int ps = H.body_slot[p];
    H.fk.body_parent[s] = ps;
    if (ps < 0) {  // fold the fixed parent's world pose into this body's local offset
      const Pose<double> &P = H.static_pose[p];
      Pose<double> f; f.p = P.p + qrot(P.q, local.p); f.q = qnormalize(qmul(P.q, local.q));
      local = f;
    }
    H.fk.body_pos[s][0] = local.p.x; H.fk.body_pos[s][1] = local.p.y; H.fk.body_pos[s][2] = local.p.z;
    H.fk.body_quat[s][0] = local.q.w; H.fk.body_quat[s][1] = local.q.x; H.fk.body_quat[s][2] = local.q.y; H.fk.body_quat[s][3] = local.q.z;
    H.fk.body_jntadr[s] = d->body_jntadr[b] < 0 ? 0 : d->body_jntadr[b];
    H.fk.body_jntnum[s] = d->body_jntnum[b];
    H.fk.body_slot[s] = b;
  }
  H.fk.nbody = H.nslot;
  H.jnt_lo.resize(d->njnt); H.jnt_hi.resize(d->njnt);
  for (int j = 0; j < d->njnt; j++) {
    H.fk.jnt_type[j] = d->jnt_type[j];
    H.fk.jnt_qadr[j] = d->jnt_qposadr[j];
    for (int k = 0; k < 3; k++) { H.fk.jnt_pos[j][k] = d->jnt_pos[3 * j + k]; H.fk.jnt_axis[j][k] = d->jnt_axis[3 * j + k]; }
    H.fk.jnt_lo[j] = H.jnt_lo[j] = d->jnt_range[2 * j];
    H.fk.jnt_hi[j] = H.jnt_hi[j] = d->jnt_range[2 * j + 1];
  }
  for (int a = 0; a < d->nq; a++) H.fk.qpos0[a] = d->qpos0[a];

  // ---- static pair list (MuJoCo filters, then mjpl's allow-list) -------------------------------
  struct RawPair { int g1, g2; };
  std::vector<RawPair> raw;
  auto allowed = [&](int b1, int b2) {
    int lo = std::min(b1, b2), hi = std::max(b1, b2);
    for (int k = 0; k < d->nallowed; k++) {
      int a = d->allowed_body_pairs[2 * k], b = d->allowed_body_pairs[2 * k + 1];
      if (std::min(a, b) == lo && std::max(a, b) == hi) return true;
    }
    return false;
  };
  if (!d->disable_contact)
    for (int g1 = 0; g1 < d->ngeom; g1++)
      for (int g2 = g1 + 1; g2 < d->ngeom; g2++) {
        int b1 = d->geom_bodyid[g1], b2 = d->geom_bodyid[g2];
        int w1 = d->body_weldid[b1], w2 = d->body_weldid[b2];
        if (w1 == w2) continue;                       // same weld body
        int wp1 = d->body_weldid[d->body_parentid[w1]], wp2 = d->body_weldid[d->body_parentid[w2]];
        if (!d->disable_filterparent && w1 != 0 && w2 != 0 && (w1 == wp2 || w2 == wp1)) continue;
        int64_t sig = ((int64_t)std::min(b1, b2) << 16) + std::max(b1, b2);
        bool ex = false;
        for (int k = 0; k < d->nexclude; k++) ex |= d->exclude_signature[k] == sig;
        if (ex) continue;
        if (!((d->geom_contype[g1] & d->geom_conaffinity[g2]) || (d->geom_contype[g2] & d->geom_conaffinity[g1]))) continue;
        if (allowed(b1, b2)) continue;
        int t1 = d->geom_type[g1], t2 = d->geom_type[g2];
        if (t1 == G_PLANE && t2 == G_PLANE) continue;  // MuJoCo has no plane-plane collider
        if (t1 == G_HFIELD || t2 == G_HFIELD) { H.err = "height fields are not supported"; return false; }
        if (t1 == G_ELLIPSOID || t2 == G_ELLIPSOID) { H.err = "ellipsoid geoms are not supported"; return false; }
        raw.push_back({g1, g2});
      }

  // ---- shapes ----------------------------------------------------------------------------------
  std::vector<int> geom_shape(d->ngeom, -1);
  std::vector<Shape<double>> tmp;
  std::vector<int> tmp_slot;
  std::vector<HullDirs> tmp_dirs;
  auto make_shape = [&](int g) -> bool {
    if (geom_shape[g] >= 0) return true;
    Shape<double> s; memset(&s, 0, sizeof s);
    s.geom = g;
    int b = d->geom_bodyid[g];
    s.slot = H.body_slot[b];
    Pose<double> T; T.p = vd(d->geom_pos + 3 * g); T.q = qnormalize(qd(d->geom_quat + 4 * g));
    if (s.slot < 0) {  // world-fixed: express in the world frame once and for all
      const Pose<double> &P = H.static_pose[b];
      Pose<double> W; W.p = P.p + qrot(P.q, T.p); W.q = qnormalize(qmul(P.q, T.q));
      T = W;
    }
    M3<double> R = q2mat(T.q);
    const double *sz = d->geom_size + 3 * g;
    auto push = [&](double x, double y, double z) {
      V3<double> w = T.p + mul(R, mk<double>(x, y, z));
      Vtx<double> v; v.x = w.x; v.y = w.y; v.z = w.z; v.w = 0;
      H.verts.push_back(v);
    };
    s.vadr = (int)H.verts.size();
    switch (d->geom_type[g]) {
      case G_PLANE:
        s.kind = SK_PLANE;
        if (s.slot >= 0) { H.err = "plane geoms must be attached to a world-fixed body"; return false; }
        s.c[0] = T.p.x; s.c[1] = T.p.y; s.c[2] = T.p.z;
        s.ax[0] = R.m[2]; s.ax[1] = R.m[5]; s.ax[2] = R.m[8];
        break;
      case G_SPHERE: s.kind = SK_VERTS; s.radius = sz[0]; push(0, 0, 0); break;
      case G_CAPSULE: s.kind = SK_VERTS; s.radius = sz[0]; push(0, 0, -sz[1]); push(0, 0, sz[1]); break;
      case G_BOX:
        s.kind = SK_VERTS;
        for (int i = 0; i < 8; i++) push((i & 1 ? 1 : -1) * sz[0], (i & 2 ? 1 : -1) * sz[1], (i & 4 ? 1 : -1) * sz[2]);
        break;
      case G_CYLINDER: {
        s.kind = SK_CYL; s.radius = sz[0]; s.halflen = sz[1];
        s.c[0] = T.p.x; s.c[1] = T.p.y; s.c[2] = T.p.z;
        s.ax[0] = R.m[2]; s.ax[1] = R.m[5]; s.ax[2] = R.m[8];
        s.bc[0] = s.c[0]; s.bc[1] = s.c[1]; s.bc[2] = s.c[2];
        s.brad = sqrt(sz[0] * sz[0] + sz[1] * sz[1]);
        s.oc[0] = s.c[0]; s.oc[1] = s.c[1]; s.oc[2] = s.c[2];
        orthobasis_from_z(mk<double>(s.ax[0], s.ax[1], s.ax[2]), s.orot);
        s.ohalf[0] = s.ohalf[1] = sz[0]; s.ohalf[2] = sz[1];
        break;
      }
      case G_MESH: {
        s.kind = SK_VERTS;
        int id = d->geom_dataid[g];
        if (id < 0 || id >= d->nmesh || d->mesh_vertnum[id] < 1) { H.err = "mesh geom without hull vertices"; return false; }
        const double *mv = d->mesh_vert + 3 * d->mesh_vertadr[id];
        for (int i = 0; i < d->mesh_vertnum[id]; i++) push(mv[3 * i], mv[3 * i + 1], mv[3 * i + 2]);
        break;
      }
      default: H.err = "unsupported geom type"; return false;
    }
    s.nvert = (int)H.verts.size() - s.vadr;
    if (s.kind == SK_VERTS) { fit_bounds(s, H.verts); fit_capsule(s, H.verts); }
    if (s.kind == SK_CYL) {   // the axis segment swept by the cylinder's radius contains the cylinder
      for (int k = 0; k < 3; k++) { s.ca[k] = s.c[k] - s.halflen * s.ax[k]; s.cb[k] = s.c[k] + s.halflen * s.ax[k]; }
      s.crad = s.radius * (1 + 1e-9); s.caplen = 2 * s.halflen;
    }
    tmp_dirs.emplace_back();
    if (!getenv("MJB_NO_INNER")) fit_inner(s, H.verts, tmp_dirs.back());
    geom_shape[g] = (int)tmp.size();
    tmp.push_back(s);
    tmp_slot.push_back(s.slot);
    return true;
  };
  for (auto &rp : raw) { if (!make_shape(rp.g1) || !make_shape(rp.g2)) return false; }

  // order shapes: moving ones by slot, then static ones
  std::vector<int> order(tmp.size());
  for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    int sa = tmp_slot[a] < 0 ? 1 << 20 : tmp_slot[a], sb = tmp_slot[b] < 0 ? 1 << 20 : tmp_slot[b];
    return sa < sb;
  });
  std::vector<int> newidx(tmp.size());
  H.slot_shape_adr.assign(std::max(H.nslot, 1), 0);
  H.slot_shape_num.assign(std::max(H.nslot, 1), 0);
  for (size_t i = 0; i < order.size(); i++) {
    newidx[order[i]] = (int)i;
    const Shape<double> &s = tmp[order[i]];
    H.shapes.push_back(s);
    H.hull_dirs.push_back(std::move(tmp_dirs[order[i]]));
    if (s.slot >= 0) {
      if (H.slot_shape_num[s.slot] == 0) H.slot_shape_adr[s.slot] = (int)i;
      H.slot_shape_num[s.slot]++;
      H.nmoving_shapes++;
    }
  }
  if (H.shapes.size() > 4096) { H.err = "too many collision geoms"; return false; }

  // ---- hull graphs -> per-vertex neighbour lists (hill-climbing support queries) ----------------
  // One offset per vertex of the global vertex table (+1 sentinel); a shape's list for local
  // vertex i is adj[adj_start[vadr+i] .. adj_start[vadr+i+1]).  Graphs are validated (ids in range,
  // symmetric, every vertex has a neighbour); anything else falls back to scanning all vertices.
  H.adj_start.assign(H.verts.size() + 1, 0);
  // Measured on B200 (1M Franka rows, hulls of 41-152 vertices): hill-climbing needs 3.8x fewer
  // dot products than scanning, but its dependent load chain is latency bound at 16 warps/SM and the
  // step got SLOWER (5.2 ms vs 2.9 ms).  It is therefore only switched on for large hulls.
  const char *hm = getenv("MJB_HILL_MIN");
  const int hill_min = hm ? atoi(hm) : 200;
  {
    std::vector<std::vector<uint8_t>> lists(H.verts.size());
    for (auto &sh : H.shapes) {
      sh.graph = 0;
      if (!VK_HILL_HOST) continue;
      if (sh.kind != SK_VERTS || d->geom_type[sh.geom] != G_MESH || !d->mesh_graph || !d->mesh_graphadr) continue;
      const int id = d->geom_dataid[sh.geom];
      const int ga = d->mesh_graphadr[id];
      const int nv = sh.nvert;
      if (ga < 0 || nv < hill_min || nv > 256 || ga + 2 > d->nmeshgraph) continue;
      const int32_t *g = d->mesh_graph + ga;
      const int gnv = g[0], gnf = g[1];
      if (gnv != nv || ga + 2 + 3 * gnv + 6 * gnf > d->nmeshgraph) continue;
      const int32_t *edgeadr = g + 2, *globalid = g + 2 + gnv, *edges = g + 2 + 2 * gnv;
      const int nedge = gnv + 3 * gnf;
      bool ok = true;
      std::vector<std::vector<uint8_t>> loc(nv);
      for (int i = 0; i < nv && ok; i++) {
        if (globalid[i] != i) ok = false;  // mesh_vert must already be in hull-local order
        for (int e = edgeadr[i]; ok; e++) {
          if (e < 0 || e >= nedge) { ok = false; break; }
          int j = edges[e];
          if (j < 0) break;
          if (j >= nv || j == i) { ok = false; break; }
          loc[i].push_back((uint8_t)j);
        }
        if (loc[i].empty()) ok = false;
      }
      for (int i = 0; i < nv && ok; i++)
        for (uint8_t j : loc[i])
          if (std::find(loc[j].begin(), loc[j].end(), (uint8_t)i) == loc[j].end()) ok = false;
      if (!ok) continue;
      for (int i = 0; i < nv; i++) lists[sh.vadr + i] = loc[i];
      sh.graph = 1;
      for (int k = 0; k < 3; k++)
        for (int sgn = 0; sgn < 2; sgn++) {
          int bi = 0; double bv = -1e300;
          for (int i = 0; i < nv; i++) {
            const Vtx<double> &p = H.verts[sh.vadr + i];
            double c = k == 0 ? p.x : (k == 1 ? p.y : p.z);
            if (sgn) c = -c;
            if (c > bv) { bv = c; bi = i; }
          }
          sh.ext[2 * k + sgn] = (uint8_t)bi;
        }
    }
    size_t total = 0;
    for (auto &l : lists) total += l.size();
    if (total > 65000) {  // offsets are 16 bit: give up on graphs for huge models
      for (auto &sh : H.shapes) sh.graph = 0;
      for (auto &l : lists) l.clear();
    }
    for (size_t i = 0; i < lists.size(); i++) {
      H.adj_start[i] = (uint16_t)H.adj.size();
      for (uint8_t j : lists[i]) H.adj.push_back(j);
    }
    H.adj_start[lists.size()] = (uint16_t)H.adj.size();
  }
  for (auto &sh : H.shapes) { sh.group = -1; sh.map = -1; }
  // ---- support maps for the larger hulls (see vk_core.cuh for the superset argument) --------------
  {
    const char *sm = getenv("MJB_SMAP_MIN");
    const int smap_min = sm ? atoi(sm) : 24;   // hulls with fewer vertices are scanned
    for (auto &sh : H.shapes) {
      if (sh.kind != SK_VERTS || sh.nvert < smap_min || sh.nvert > 255) continue;
      if (H.smap_ids.size() > (1u << 23)) break;
      const Vtx<double> *v = H.verts.data() + sh.vadr;
      const int R = SMAP_R;
      sh.map = (int)H.smap_cells.size();
      for (int f = 0; f < 6; f++) {
        const int k = f / 2;
        const double sg = (f % 2) ? -1.0 : 1.0;
        for (int iu = 0; iu < R; iu++)
          for (int iv = 0; iv < R; iv++) {
            auto dir = [&](double u, double w) {
              double d[3];
              d[k] = sg; d[(k + 1) % 3] = u; d[(k + 2) % 3] = w;
              const double n = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
              return mk<double>(d[0] / n, d[1] / n, d[2] / n);
            };
            const double u0 = -1 + 2.0 * iu / R, u1 = -1 + 2.0 * (iu + 1) / R, w0 = -1 + 2.0 * iv / R, w1 = -1 + 2.0 * (iv + 1) / R;
            const V3<double> c = dir(0.5 * (u0 + u1), 0.5 * (w0 + w1));
            double delta = 0;
            for (int a = 0; a < 2; a++)
              for (int b = 0; b < 2; b++) {
                const V3<double> q = dir(a ? u1 : u0, b ? w1 : w0) - c;
                delta = std::max(delta, sqrt(dot(q, q)));
              }
            delta = delta * 1.02 + 1e-6;
            int best = 0; double hb = -1e300;
            for (int i = 0; i < sh.nvert; i++) { const double h = v[i].x * c.x + v[i].y * c.y + v[i].z * c.z; if (h > hb) { hb = h; best = i; } }
            std::vector<uint8_t> cand;
            for (int i = 0; i < sh.nvert; i++) {
              const double h = v[i].x * c.x + v[i].y * c.y + v[i].z * c.z;
              const double dx = v[i].x - v[best].x, dy = v[i].y - v[best].y, dz = v[i].z - v[best].z;
              if (hb - h <= sqrt(dx * dx + dy * dy + dz * dz) * delta + 1e-9) cand.push_back((uint8_t)i);
            }
            const uint32_t word = (uint32_t)(H.smap_ids.size() << 8) | (uint32_t)cand.size();   // <= 255 vertices per mapped hull
            H.smap_ids.insert(H.smap_ids.end(), cand.begin(), cand.end());
            H.smap_cells.push_back(word);
          }
      }
    }
  }

  // ---- pairs in processing order --------------------------------------------------------------------
  struct Tmp { Pair p; int g1, g2; long key; double rsum, bsum; };
  std::vector<Tmp> tp;
  for (auto &rp : raw) {
    int a = newidx[geom_shape[rp.g1]], b = newidx[geom_shape[rp.g2]];
    const Shape<double> *A = &H.shapes[a], *B = &H.shapes[b];
    double margin = std::max(d->geom_margin[rp.g1], d->geom_margin[rp.g2]);
    // plane first; otherwise the shape with more vertices is A (its frame hosts the GJK)
    if (B->kind == SK_PLANE || (A->kind != SK_PLANE && B->nvert > A->nvert)) { std::swap(a, b); std::swap(A, B); }
    Tmp t; memset(&t, 0, sizeof t);
    t.g1 = rp.g1; t.g2 = rp.g2;
    t.p.sa = (uint16_t)a; t.p.sb = (uint16_t)b;
    if (A->kind == SK_PLANE) {
      t.p.kind = PK_PLANE;
      t.rsum = (B->kind == SK_VERTS ? B->radius : 0.0) + margin;
      t.bsum = B->brad + margin;
      t.key = 0;
      if (B->kind == SK_VERTS && B->nvert >= 8) t.p.flags = PF_OBB;  // OBB-above-plane test before the vertex scan
    } else {
      bool seg = A->kind == SK_VERTS && B->kind == SK_VERTS && A->nvert <= 2 && B->nvert <= 2;
      t.p.kind = seg ? PK_SEGSEG : PK_GJK;
      double ra = A->kind == SK_VERTS ? A->radius : 0.0, rb = B->kind == SK_VERTS ? B->radius : 0.0;
      t.rsum = ra + rb + margin;
      t.bsum = A->brad + B->brad + margin;
      int cost = (A->kind == SK_CYL ? 16 : A->nvert) + (B->kind == SK_CYL ? 16 : B->nvert);
      t.p.flags = (!seg && cost >= 12) ? PF_OBB : 0;
      t.key = seg ? 1 : 1000 + cost;
    }
    t.p.rsum = (float)t.rsum;
    t.p.bsum = (float)t.bsum;
    tp.push_back(t);
  }
  for (auto &t : tp) {
    const Shape<double> &A = H.shapes[t.p.sa], &B = H.shapes[t.p.sb];
    if (A.slot < 0) t.p.flags |= PF_A_STATIC;
    if (B.slot < 0) t.p.flags |= PF_B_STATIC;
  }
  // ---- calibration: run the same fp32 core on seeded rows (uniform in the joint ranges) to
  // learn, per pair, how often it survives the sphere cull / the OBB cull / ends in contact.
  // This only ORDERS the work (likely contacts first => early exit; bounded work-queue fill
  // per round); it never decides a result.
  const int NCAL = 2048;
  std::vector<double> n_sph(tp.size(), 0), n_obb(tp.size(), 0), n_pen(tp.size(), 0);
  {
    FkTables<float> fk32 = convert_fk<float>(H.fk);
    std::vector<Shape<float>> s32; std::vector<Vtx<float>> v32;
    for (auto &sh : H.shapes) s32.push_back(convert_shape<float>(sh));
    for (auto &v : H.verts) { Vtx<float> f; f.x = (float)v.x; f.y = (float)v.y; f.z = (float)v.z; f.w = 0; v32.push_back(f); }
    Pose<float> ident; ident.p = mk<float>(0, 0, 0); ident.q.w = 1; ident.q.x = ident.q.y = ident.q.z = 0;
    for (int r = 0; r < NCAL; r++) {
      float q[MAX_JNT];
      for (int j = 0; j < d->nq; j++) {
        float lo = (float)H.jnt_lo[j], hi = (float)H.jnt_hi[j];
        if (!(hi > lo)) { lo = -3.14159f; hi = 3.14159f; }
        q[j] = sweep_value(0x5eedull, (uint64_t)r, (uint32_t)j, lo, hi);
      }
      Pose<float> P[MAX_BODY];
      for (int k = 0; k < H.nslot; k++) { int ps = fk32.body_parent[k]; P[k] = fk_body(fk32, k, ps < 0 ? ident : P[ps], q); }
      bool row_pen = false;
      for (size_t p = 0; p < tp.size(); p++) {
        const Pair &pr = tp[p].p;
        const Shape<float> &A = s32[pr.sa], &B = s32[pr.sb];
        const Pose<float> &PA = A.slot < 0 ? ident : P[A.slot];
        const Pose<float> &PB = B.slot < 0 ? ident : P[B.slot];
        V3<float> cB = PB.p + qrot(PB.q, mk<float>(B.bc[0], B.bc[1], B.bc[2]));
        const float slack = 1e-4f;
        if (pr.kind == PK_PLANE) {
          float dd = A.ax[0] * (cB.x - A.c[0]) + A.ax[1] * (cB.y - A.c[1]) + A.ax[2] * (cB.z - A.c[2]);
          if (dd > pr.bsum + slack) continue;
        } else {
          V3<float> cA = PA.p + qrot(PA.q, mk<float>(A.bc[0], A.bc[1], A.bc[2]));
          V3<float> dd = cA - cB;
          if (dot(dd, dd) > (pr.bsum + slack) * (pr.bsum + slack)) continue;
        }
        n_sph[p]++;
        if (midphase_cull(pr, A, B, PA, PB, pr.rsum - swept_radius(A) - swept_radius(B), slack)) continue;
        n_obb[p]++;
        int v = narrow_item<float>(pr.kind, A, B, v32.data(), PA, PB, pr.rsum);
        if (v == V_PEN) { n_pen[p]++; row_pen = true; }
      }
      H.calib_pen_rows += row_pen;
    }
    for (size_t p = 0; p < tp.size(); p++) {
      H.calib_sphere_per_row += n_sph[p] / NCAL; H.calib_items_per_row += n_obb[p] / NCAL;
      const Shape<double> &A = H.shapes[tp[p].p.sa], &B = H.shapes[tp[p].p.sb];
      (void)B;
      H.bin_expect[item_bin(tp[p].p, A, B)] += n_obb[p] / NCAL;   // every surviving item goes to a bin (closed forms: bin 0)
    }
    H.calib_pen_rows /= NCAL;
  }
  // ---- order: cheap analytic kinds first, then GJK pairs by contact likelihood (descending),
  // ties by fewer narrow-phase items and fewer vertices
  std::vector<int> ord(tp.size());
  for (size_t i = 0; i < ord.size(); i++) ord[i] = (int)i;
  // MJB_ORDER=ratio (experiment): contacts found per unit of expected work instead of raw likelihood
  const char *order_env = getenv("MJB_ORDER");
  const bool by_ratio = order_env && !strcmp(order_env, "ratio");
  std::vector<double> gain(tp.size(), 0.0);
  for (size_t i = 0; i < tp.size(); i++) {
    const Shape<double> &A = H.shapes[tp[i].p.sa], &B = H.shapes[tp[i].p.sb];
    const double cost = 1.0 + 12.0 * n_sph[i] / NCAL + (n_obb[i] / NCAL) * 3.0 * (double)(A.nvert + B.nvert + 16);
    gain[i] = by_ratio ? (n_pen[i] / NCAL) / cost : n_pen[i];
  }
  std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) {
    bool gx = tp[x].p.kind == PK_GJK, gy = tp[y].p.kind == PK_GJK;
    if (gx != gy) return !gx;
    if (gain[x] != gain[y]) return gain[x] > gain[y];
    if (n_obb[x] != n_obb[y]) return n_obb[x] < n_obb[y];
    if (tp[x].key != tp[y].key) return tp[x].key < tp[y].key;
    if (tp[x].p.sa != tp[y].p.sa) return tp[x].p.sa < tp[y].p.sa;
    return tp[x].p.sb < tp[y].p.sb;
  });
  // ---- rounds: bounded expected queue fill per row (sphere survivors <= 5, items <= 2.5; measured on
  // B200: 3/1.5 -> 3.15 ms per 1M Franka rows, 5/2.5 -> 3.03, 8/4 -> 3.04)
  H.nrounds = 0;
  const char *rs_env = getenv("MJB_ROUND_SPH"), *ri_env = getenv("MJB_ROUND_ITEMS");
  const double lim_s = rs_env ? atof(rs_env) : 5.0, lim_o = ri_env ? atof(ri_env) : 2.5;
  double acc_s = 0, acc_o = 0;
  int cur_gjk = -1;
  std::vector<int> round_of(ord.size(), 0);
  for (size_t k = 0; k < ord.size(); k++) {
    int i = ord[k];
    int g = tp[i].p.kind == PK_GJK ? 1 : 0;
    double es = n_sph[i] / NCAL, eo = n_obb[i] / NCAL;
    bool fresh = (k == 0) || (g != cur_gjk) || (acc_s + es > lim_s) || (acc_o + eo > lim_o);
    if (fresh && H.nrounds < MAX_ROUNDS) {
      H.round_start[H.nrounds] = (int)k;
      H.round_gjk[H.nrounds] = g;
      H.nrounds++;
      acc_s = acc_o = 0;
      cur_gjk = g;
    }
    acc_s += es; acc_o += eo;
    round_of[k] = H.nrounds - 1;
  }
  // inside a round the order is free: sort by shape A so the sphere stage fetches A's centre
  // once per run of pairs
  for (int r = 0; r < H.nrounds; r++) {
    int b = H.round_start[r], e = (r + 1 < H.nrounds) ? H.round_start[r + 1] : (int)ord.size();
    const char *so = getenv("MJB_ROUND_SORT");   // experiment: "size" = similar hull sizes adjacent
    const bool by_size = so && !strcmp(so, "size");
    std::stable_sort(ord.begin() + b, ord.begin() + e, [&](int x, int y) {
      if (by_size) {
        const int ax = H.shapes[tp[x].p.sa].nvert, ay = H.shapes[tp[y].p.sa].nvert;
        if (ax != ay) return ax > ay;
        if (tp[x].p.sa != tp[y].p.sa) return tp[x].p.sa < tp[y].p.sa;
        const int bx = H.shapes[tp[x].p.sb].nvert, by = H.shapes[tp[y].p.sb].nvert;
        if (bx != by) return bx > by;
        return tp[x].p.sb < tp[y].p.sb;
      }
      if (tp[x].p.sa != tp[y].p.sa) return tp[x].p.sa < tp[y].p.sa;
      return tp[x].p.sb < tp[y].p.sb;
    });
  }
  for (size_t k = 0; k < ord.size(); k++) {
    Tmp &t = tp[ord[k]];
    t.p.round = (uint16_t)round_of[k];
    H.pairs.push_back(t.p); H.pair_g1.push_back(t.g1); H.pair_g2.push_back(t.g2);
    H.pair_rsum64.push_back(t.rsum); H.pair_bsum64.push_back(t.bsum);
  }
  H.round_start[H.nrounds] = (int)ord.size();
  if (H.pairs.size() > 60000) { H.err = "too many geom pairs"; return false; }
  return build_groups(H);
}

// mjb_pose_spec (caller's view: site in its body frame, reference frame world_T_C) -> PoseSpec
// (kernel's view: pose slot, C_T_world, joints that move the site)
inline bool make_pose_spec(const HostModel &H, const mjb_pose_spec *in, PoseSpec &sp, std::string &err) {
  if (!in) { err = "null pose spec"; return false; }
  // reference: ValueError texts of PoseConstraint.__init__ (pose_constraint.py:51-54)
  if (in->tolerance < 0.0) { err = "`tolerance` must be >= 0."; return false; }
  if (!(in->q_step > 0.0)) { err = "`q_step` must be > 0."; return false; }
  if (in->site_bodyid < 0 || in->site_bodyid >= H.nbody) { err = "bad site body id"; return false; }
  memset(&sp, 0, sizeof sp);
  const int b = in->site_bodyid;
  sp.site_slot = H.body_slot[b];
  Pose<double> S; S.p = mk<double>(in->site_pos[0], in->site_pos[1], in->site_pos[2]);
  S.q = qnormalize(qd(in->site_quat));
  if (sp.site_slot < 0) {  // world-fixed body: fold its pose in
    const Pose<double> &Pb = H.static_pose[b];
    Pose<double> W; W.p = Pb.p + qrot(Pb.q, S.p); W.q = qnormalize(qmul(Pb.q, S.q));
    S = W;
  }
  sp.site_pos[0] = S.p.x; sp.site_pos[1] = S.p.y; sp.site_pos[2] = S.p.z;
  sp.site_quat[0] = S.q.w; sp.site_quat[1] = S.q.x; sp.site_quat[2] = S.q.y; sp.site_quat[3] = S.q.z;
  // C_T_world = inverse(reference frame)
  Q4<double> iq = qconj(qnormalize(qd(in->ref_quat)));
  V3<double> ip = -qrot(iq, mk<double>(in->ref_pos[0], in->ref_pos[1], in->ref_pos[2]));
  sp.cw_pos[0] = ip.x; sp.cw_pos[1] = ip.y; sp.cw_pos[2] = ip.z;
  sp.cw_quat[0] = iq.w; sp.cw_quat[1] = iq.x; sp.cw_quat[2] = iq.y; sp.cw_quat[3] = iq.z;
  for (int i = 0; i < 6; i++) { sp.lo[i] = in->lower[i]; sp.hi[i] = in->upper[i]; }
  sp.tolerance = in->tolerance; sp.q_step = in->q_step;
  // joints that move the site: joints of the site's body and of all its ancestors
  for (int s = sp.site_slot; s >= 0; s = H.fk.body_parent[s])
    for (int k = 0; k < H.fk.body_jntnum[s]; k++) sp.jnt_mask |= 1u << (H.fk.body_jntadr[s] + k);
  return true;
}

// mjb_ik_spec -> IkSpec (argument checks follow MinkIKSolver.__init__, mink_ik_solver.py:47-52)
inline bool make_ik_spec(const HostModel &H, const mjb_ik_spec *in, IkSpec &out, std::string &err) {
  if (!in) { err = "null ik spec"; return false; }
  if (in->movable_mask == 0) { err = "`joints` cannot be empty."; return false; }
  if (in->iterations < 1) { err = "`iterations` must be > 0."; return false; }
  if (!(in->pos_tolerance >= 0.0) || !(in->ori_tolerance >= 0.0)) { err = "tolerances must be >= 0."; return false; }
  mjb_pose_spec ps;
  memset(&ps, 0, sizeof ps);
  ps.site_bodyid = in->site_bodyid;
  memcpy(ps.site_pos, in->site_pos, sizeof ps.site_pos);
  memcpy(ps.site_quat, in->site_quat, sizeof ps.site_quat);
  ps.ref_quat[0] = 1.0; ps.q_step = 1.0;
  PoseSpec sp;
  if (!make_pose_spec(H, &ps, sp, err)) return false;
  memset(&out, 0, sizeof out);
  out.site_slot = sp.site_slot;
  out.jnt_mask = sp.jnt_mask & in->movable_mask;
  memcpy(out.site_pos, sp.site_pos, sizeof sp.site_pos);
  memcpy(out.site_quat, sp.site_quat, sizeof sp.site_quat);
  out.pos_tol = in->pos_tolerance; out.ori_tol = in->ori_tolerance;
  out.lm_damping = in->lm_damping >= 0.0 ? in->lm_damping : 0.1;
  out.damping = in->damping > 0.0 ? in->damping : 1e-9;
  out.max_step = in->max_step > 0.0 ? in->max_step : 0.5;
  out.iterations = in->iterations;
  return true;
}

}  // namespace vkb
