// vk_core.cuh -- geometric core of the validity kernels (host+device, templated on scalar).
//
// float  instance: the sm_100a fast path (vk_kernels.cu).
// double instance: the on-device re-evaluation of rows the fast path could not certify.
//
// What it restates (MuJoCo 3.x semantics; reference call sites
// src/mjpl/constraint/collision_constraint.py:27-30):
//   * mj_kinematics for hinge / slide / fixed bodies                      -> fk_body()
//   * the boolean "closed convex sets at signed distance <= margin" that every MuJoCo
//     narrow-phase routine reduces to when it decides whether to emit a contact
//     (mjc_Convex / mjc_BoxBox / mjc_CapsuleBox / ... / mjc_PlaneConvex)   -> gjk_classify(),
//     segseg_classify(), plane tests in the kernels.
//
// Design: every convex collision geom is a "sphere-swept vertex set" in its BODY frame
// (mesh = hull vertices, box = 8 corners, capsule = 2 end points + radius, sphere = 1 point
// + radius), or a cylinder.  One GJK loop serves them all; it returns a three-way verdict
// with certified bounds so that the fp32 path never guesses:
//   SEP  : a separating direction proves   distance >  R + tol
//   PEN  : a witness point / enclosed origin proves distance <  R - tol
//   UNC  : neither (|distance - R| <~ tol, or no convergence) -> re-evaluated in fp64
// The simplex solver is the signed-volume formulation (barycentric cofactors) rather than
// Voronoi-region tests; the CPU oracle deliberately uses the other one.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define VK_HD __host__ __device__ __forceinline__
#define VK_HD_NOINLINE __host__ __device__ __noinline__
#ifndef VK_SOLVE_INLINE
#define VK_SOLVE_HD VK_HD_NOINLINE
#else
#define VK_SOLVE_HD VK_HD
#endif
#else
#define VK_HD inline
#define VK_HD_NOINLINE
#define VK_SOLVE_HD inline
#endif

namespace vk {

enum Verdict : int { V_SEP = 0, V_PEN = 1, V_UNC = 2 };
enum ShapeKind : int { SK_VERTS = 0, SK_CYL = 1, SK_PLANE = 2 };
enum PairKind : int { PK_PLANE = 0, PK_SEGSEG = 1, PK_GJK = 2, PK_BOXBOX = 3 };
enum JointKind : int { JK_SLIDE = 2, JK_HINGE = 3 };

constexpr int MAX_BODY = 32;
constexpr int MAX_JNT = 32;

template <typename T> struct Num;
template <> struct Num<float> {
  static constexpr float tol = 4e-6f;        // certification slack (FK + GJK rounding)
  static constexpr float conv_rel = 1e-6f;   // GJK relative convergence on |v|^2 - v.w
  static constexpr float tiny = 1e-30f;
#ifndef VK_MAXIT32
#define VK_MAXIT32 12   // B200, 1M Franka rows: 24 -> narrow 0.330 ms + fp64 0.048, 12 -> 0.302 + 0.054, 9 -> 0.301 + 0.055 (the kernel ends with its slowest item)
#endif
  static constexpr int maxit = VK_MAXIT32;   // an item that has not been certified by then is handed to the fp64 pass
};
template <> struct Num<double> {
  static constexpr double tol = 0.0;
  static constexpr double conv_rel = 1e-12;
  static constexpr double tiny = 1e-280;
  static constexpr int maxit = 64;
};

template <typename T> struct V3 { T x, y, z; };
template <typename T> struct Q4 { T w, x, y, z; };
template <typename T> struct M3 { T m[9]; };  // row-major
template <typename T> struct Pose { V3<T> p; Q4<T> q; };

template <typename T> VK_HD V3<T> mk(T x, T y, T z) { V3<T> r; r.x = x; r.y = y; r.z = z; return r; }
template <typename T> VK_HD V3<T> operator+(V3<T> a, V3<T> b) { return mk<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> VK_HD V3<T> operator-(V3<T> a, V3<T> b) { return mk<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> VK_HD V3<T> operator-(V3<T> a) { return mk<T>(-a.x, -a.y, -a.z); }
template <typename T> VK_HD V3<T> operator*(V3<T> a, T s) { return mk<T>(a.x * s, a.y * s, a.z * s); }
template <typename T> VK_HD T dot(V3<T> a, V3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> VK_HD V3<T> cross(V3<T> a, V3<T> b) {
  return mk<T>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
template <typename T> VK_HD T det3(V3<T> a, V3<T> b, V3<T> c) { return dot(a, cross(b, c)); }

VK_HD float vk_sqrt(float x) { return sqrtf(x); }
VK_HD double vk_sqrt(double x) { return sqrt(x); }
VK_HD float vk_abs(float x) { return fabsf(x); }
VK_HD double vk_abs(double x) { return fabs(x); }
VK_HD float vk_min(float a, float b) { return fminf(a, b); }
VK_HD double vk_min(double a, double b) { return fmin(a, b); }
VK_HD float vk_max(float a, float b) { return fmaxf(a, b); }
VK_HD double vk_max(double a, double b) { return fmax(a, b); }
VK_HD void vk_sincos(float x, float *s, float *c) {
#if defined(__CUDA_ARCH__)
  sincosf(x, s, c);
#else
  *s = sinf(x); *c = cosf(x);
#endif
}
VK_HD void vk_sincos(double x, double *s, double *c) {
#if defined(__CUDA_ARCH__)
  sincos(x, s, c);
#else
  *s = sin(x); *c = cos(x);
#endif
}

// Hamilton product (w,x,y,z) -- mju_mulQuat
template <typename T> VK_HD Q4<T> qmul(Q4<T> a, Q4<T> b) {
  Q4<T> r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
  r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
  return r;
}
template <typename T> VK_HD Q4<T> qconj(Q4<T> a) { Q4<T> r; r.w = a.w; r.x = -a.x; r.y = -a.y; r.z = -a.z; return r; }
// rotate v by unit quaternion q:  v + 2w(u x v) + 2 u x (u x v)
template <typename T> VK_HD V3<T> qrot(Q4<T> q, V3<T> v) {
  V3<T> u = mk<T>(q.x, q.y, q.z);
  V3<T> t = cross(u, v) * T(2);
  return v + t * q.w + cross(u, t);
}
template <typename T> VK_HD Q4<T> qnormalize(Q4<T> q) {
  T n = vk_sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  T inv = T(1) / n;
  Q4<T> r; r.w = q.w * inv; r.x = q.x * inv; r.y = q.y * inv; r.z = q.z * inv;
  return r;
}
template <typename T> VK_HD M3<T> q2mat(Q4<T> q) {
  M3<T> r;
  T q00 = q.w * q.w, q11 = q.x * q.x, q22 = q.y * q.y, q33 = q.z * q.z;
  T q01 = q.w * q.x, q02 = q.w * q.y, q03 = q.w * q.z, q12 = q.x * q.y, q13 = q.x * q.z, q23 = q.y * q.z;
  r.m[0] = q00 + q11 - q22 - q33; r.m[4] = q00 - q11 + q22 - q33; r.m[8] = q00 - q11 - q22 + q33;
  r.m[1] = 2 * (q12 - q03); r.m[2] = 2 * (q13 + q02);
  r.m[3] = 2 * (q12 + q03); r.m[5] = 2 * (q23 - q01);
  r.m[6] = 2 * (q13 - q02); r.m[7] = 2 * (q23 + q01);
  return r;
}
template <typename T> VK_HD V3<T> mul(const M3<T> &m, V3<T> v) {
  return mk<T>(m.m[0] * v.x + m.m[1] * v.y + m.m[2] * v.z, m.m[3] * v.x + m.m[4] * v.y + m.m[5] * v.z,
               m.m[6] * v.x + m.m[7] * v.y + m.m[8] * v.z);
}
template <typename T> VK_HD V3<T> mulT(const M3<T> &m, V3<T> v) {
  return mk<T>(m.m[0] * v.x + m.m[3] * v.y + m.m[6] * v.z, m.m[1] * v.x + m.m[4] * v.y + m.m[7] * v.z,
               m.m[2] * v.x + m.m[5] * v.y + m.m[8] * v.z);
}

// ------------------------------------------------------------------------------ model tables
// Kinematic tree (passed to the fp32 kernel as a __grid_constant__ parameter: constant bank).
template <typename T> struct FkTables {
  int nq, nbody, njnt, pad;
  int body_parent[MAX_BODY];
  int body_jntadr[MAX_BODY];
  int body_jntnum[MAX_BODY];
  int body_slot[MAX_BODY];     // pose slot of a moving body, -1 for world-fixed bodies
  T body_pos[MAX_BODY][3];
  T body_quat[MAX_BODY][4];
  int jnt_type[MAX_JNT];
  int jnt_qadr[MAX_JNT];
  T jnt_pos[MAX_JNT][3];
  T jnt_axis[MAX_JNT][3];
  T jnt_lo[MAX_JNT], jnt_hi[MAX_JNT], qpos0[MAX_JNT];
};

// One collision geom, expressed in the frame of its body (world-fixed geoms: body = -1 and the
// data is already in the world frame).
template <typename T> struct Shape {
  int kind;        // ShapeKind
  int slot;        // pose slot of the carrying body, -1 = world frame (static)
  int vadr, nvert; // SK_VERTS: range in the vertex table
  T radius;        // swept radius (sphere / capsule) or cylinder radius
  T halflen;       // cylinder half length
  T c[3];          // SK_CYL: centre; SK_PLANE: a point of the plane
  T ax[3];         // SK_CYL: axis (unit); SK_PLANE: normal (unit)
  T bc[3];         // bounding-sphere centre
  T brad;          // bounding-sphere radius (includes swept radius)
  T oc[3];         // OBB centre
  T orot[9];       // OBB axes = columns, row-major 3x3
  T ohalf[3];      // OBB half extents (include swept radius)
  int geom;        // MuJoCo geom id
  int graph;       // 1: hull adjacency available (hill-climbing support), 0: scan all vertices
  uint8_t ext[8];  // graph: local ids of the extreme vertices along +x,-x,+y,-y,+z,-z (start points)
  int group;       // cull group of the shape (vk_pipe.cuh): its moving body, or the shape itself if world fixed
  // bounding capsule (line-swept sphere) in the same frame as the vertices: segment ca..cb, radius
  // crad (includes the swept radius); caplen == 0 marks a point-like capsule (a sphere)
  T ca[3], cb[3], crad, caplen;
  int map;         // support map of the hull (first cell in the cell table), -1: scan all vertices
  // inner capsule: segment ia..ib swept by irad lies INSIDE the shape (same frame as the vertices; irad <= 0: none).
  // Two shapes whose inner capsules overlap are certainly in contact -- no narrow phase needed.
  T ia[3], ib[3], irad;
};
static_assert(sizeof(Shape<float>) == 208, "Shape<float> must stay a multiple of 16 bytes (cp.async.bulk granularity)");

struct Pair {
  uint16_t sa, sb;  // shape indices (sa: plane if any; else the one with more vertices first)
  uint8_t kind;     // PairKind
  uint8_t flags;    // PF_* bits
  uint16_t round;   // processing round (pairs are sorted by round)
  float rsum;       // swept radii + margin: contact iff core distance <= rsum
  float bsum;       // bounding radii sum + margin
};
static_assert(sizeof(Pair) == 16, "Pair is loaded as one 16-byte word");
constexpr uint8_t PF_OBB = 1, PF_A_STATIC = 2, PF_B_STATIC = 4;
constexpr int MAX_ROUNDS = 32;

// ------------------------------------------------------------------------------ forward kinematics
// One body of mj_kinematics (engine_core_smooth.c), SURVEY.md A.1: parent pose -> body pose.
template <typename T, typename TQ>
VK_HD Pose<T> fk_body(const FkTables<T> &tb, int i, const Pose<T> &parent, const TQ *q) {
  Pose<T> o;
  V3<T> bp = mk<T>(tb.body_pos[i][0], tb.body_pos[i][1], tb.body_pos[i][2]);
  Q4<T> bq; bq.w = tb.body_quat[i][0]; bq.x = tb.body_quat[i][1]; bq.y = tb.body_quat[i][2]; bq.z = tb.body_quat[i][3];
  o.p = parent.p + qrot(parent.q, bp);
  o.q = qmul(parent.q, bq);
  const int ja = tb.body_jntadr[i], jn = tb.body_jntnum[i];
  for (int k = 0; k < jn; k++) {
    const int j = ja + k;
    const int a = tb.jnt_qadr[j];
    V3<T> jp = mk<T>(tb.jnt_pos[j][0], tb.jnt_pos[j][1], tb.jnt_pos[j][2]);
    V3<T> jax = mk<T>(tb.jnt_axis[j][0], tb.jnt_axis[j][1], tb.jnt_axis[j][2]);
    T dq = T(q[a]) - tb.qpos0[a];
    if (tb.jnt_type[j] == JK_SLIDE) {
      o.p = o.p + qrot(o.q, jax) * dq;
    } else {
      V3<T> anchor = o.p + qrot(o.q, jp);
      T s, c;
      vk_sincos(dq * T(0.5), &s, &c);
      Q4<T> ql; ql.w = c; ql.x = jax.x * s; ql.y = jax.y * s; ql.z = jax.z * s;
      o.q = qmul(o.q, ql);
      o.p = anchor - qrot(o.q, jp);
    }
  }
  o.q = qnormalize(o.q);
  return o;
}

// ------------------------------------------------------------------------------ simplex solver
// Closest point of conv{P0..P(n-1)} to the origin as barycentric weights (signed volumes).
template <typename T> struct Simplex {
  V3<T> p0, p1, p2, p3;
  int n;
};

template <typename T> VK_HD void solve1(V3<T> a, V3<T> b, T &la, T &lb) {
  V3<T> t = b - a;
  T tt = dot(t, t);
  T s = tt > T(0) ? -dot(a, t) / tt : T(0);
  s = s < T(0) ? T(0) : (s > T(1) ? T(1) : s);
  la = T(1) - s;
  lb = s;
}

template <typename T> VK_HD T comb2(V3<T> a, V3<T> b, T la, T lb) {
  V3<T> v = a * la + b * lb;
  return dot(v, v);
}

// solve2/solve3 are big and sit on rarely-taken paths: kept out of line so the kernel's hot loops
// stay small (instruction-fetch stalls were the second largest stall reason in the ncu capture)
template <typename T> VK_SOLVE_HD void solve2(V3<T> a, V3<T> b, V3<T> c, T &la, T &lb, T &lc) {
  V3<T> n = cross(b - a, c - a);
  T nn = dot(n, n);
  bool inside = false;
  T ca = T(0), cb = T(0), cc = T(0);
  if (nn > Num<T>::tiny) {
    V3<T> p = n * (dot(a, n) / nn);  // origin projected on the plane
    ca = dot(n, cross(b - p, c - p));
    cb = dot(n, cross(c - p, a - p));
    cc = dot(n, cross(a - p, b - p));
    inside = (ca >= T(0)) && (cb >= T(0)) && (cc >= T(0));
  }
  if (inside) {
    T s = T(1) / (ca + cb + cc);
    la = ca * s; lb = cb * s; lc = cc * s;
    return;
  }
  // best of the edges whose opposite signed area is negative (all three if degenerate)
  T best = T(-1);
  la = T(1); lb = T(0); lc = T(0);
  const bool deg = !(nn > Num<T>::tiny);
  if (deg || ca < T(0)) {  // edge bc
    T x, y; solve1(b, c, x, y);
    T d = comb2(b, c, x, y);
    best = d; la = T(0); lb = x; lc = y;
  }
  if (deg || cb < T(0)) {  // edge ca
    T x, y; solve1(c, a, x, y);
    T d = comb2(c, a, x, y);
    if (best < T(0) || d < best) { best = d; la = y; lb = T(0); lc = x; }
  }
  if (deg || cc < T(0)) {  // edge ab
    T x, y; solve1(a, b, x, y);
    T d = comb2(a, b, x, y);
    if (best < T(0) || d < best) { best = d; la = x; lb = y; lc = T(0); }
  }
}

template <typename T> VK_HD T comb3(V3<T> a, V3<T> b, V3<T> c, T la, T lb, T lc) {
  V3<T> v = a * la + b * lb + c * lc;
  return dot(v, v);
}

// returns true when the origin is inside the tetrahedron (weights then all > 0)
template <typename T>
VK_SOLVE_HD bool solve3(V3<T> a, V3<T> b, V3<T> c, V3<T> d, T &la, T &lb, T &lc, T &ld) {
  T Ca = -det3(b, c, d), Cb = det3(a, c, d), Cc = -det3(a, b, d), Cd = det3(a, b, c);
  T dm = Ca + Cb + Cc + Cd;
  // scale for the degeneracy test: product of edge lengths ~ volume scale
  T sc = vk_abs(Ca) + vk_abs(Cb) + vk_abs(Cc) + vk_abs(Cd);
  const bool deg = !(vk_abs(dm) > T(1e-6) * sc) || !(sc > Num<T>::tiny);
  if (dm < T(0)) { Ca = -Ca; Cb = -Cb; Cc = -Cc; Cd = -Cd; dm = -dm; }
  if (!deg && Ca > T(0) && Cb > T(0) && Cc > T(0) && Cd > T(0)) {
    T s = T(1) / dm;
    la = Ca * s; lb = Cb * s; lc = Cc * s; ld = Cd * s;
    return true;
  }
  T best = T(-1);
  la = T(1); lb = lc = ld = T(0);
  if (deg || Ca <= T(0)) {  // face bcd
    T x, y, z; solve2(b, c, d, x, y, z);
    best = comb3(b, c, d, x, y, z); la = T(0); lb = x; lc = y; ld = z;
  }
  if (deg || Cb <= T(0)) {  // face acd
    T x, y, z; solve2(a, c, d, x, y, z);
    T e = comb3(a, c, d, x, y, z);
    if (best < T(0) || e < best) { best = e; la = x; lb = T(0); lc = y; ld = z; }
  }
  if (deg || Cc <= T(0)) {  // face abd
    T x, y, z; solve2(a, b, d, x, y, z);
    T e = comb3(a, b, d, x, y, z);
    if (best < T(0) || e < best) { best = e; la = x; lb = y; lc = T(0); ld = z; }
  }
  if (deg || Cd <= T(0)) {  // face abc
    T x, y, z; solve2(a, b, c, x, y, z);
    T e = comb3(a, b, c, x, y, z);
    if (best < T(0) || e < best) { best = e; la = x; lb = y; lc = z; ld = T(0); }
  }
  return false;
}

// distance from the origin to the plane through (a,b,c); 0 for a degenerate triangle
template <typename T> VK_HD T plane_dist(V3<T> a, V3<T> b, V3<T> c) {
  V3<T> n = cross(b - a, c - a);
  T nn = dot(n, n);
  return nn > Num<T>::tiny ? vk_abs(dot(a, n)) / vk_sqrt(nn) : T(0);
}

template <typename T> VK_HD void cswap(bool p, V3<T> &a, V3<T> &b) {
  V3<T> ta = a, tb = b;
  a.x = p ? tb.x : ta.x; a.y = p ? tb.y : ta.y; a.z = p ? tb.z : ta.z;
  b.x = p ? ta.x : tb.x; b.y = p ? ta.y : tb.y; b.z = p ? ta.z : tb.z;
}
template <typename T> VK_HD void cswap(bool p, T &a, T &b) { T ta = a, tb = b; a = p ? tb : ta; b = p ? ta : tb; }

// ------------------------------------------------------------------------------ support functions
// Vertex tables are stored as 4 scalars per vertex (x,y,z,0).
template <typename T> struct Vtx { T x, y, z, w; };

template <typename T>
VK_HD V3<T> support_verts(const Vtx<T> *__restrict__ v, int n, V3<T> d) {
  T best = v[0].x * d.x + v[0].y * d.y + v[0].z * d.z;
  V3<T> bp = mk<T>(v[0].x, v[0].y, v[0].z);
  for (int i = 1; i < n; i++) {
    Vtx<T> p = v[i];
    T s = p.x * d.x + p.y * d.y + p.z * d.z;
    bool g = s > best;
    best = g ? s : best;
    bp.x = g ? p.x : bp.x; bp.y = g ? p.y : bp.y; bp.z = g ? p.z : bp.z;
  }
  return bp;
}

// ------------------------------------------------------------------------------ support maps
// A hull's support function max_v v.d is piecewise constant in the direction d.  The unit sphere of
// directions is cut into 6 x SMAP_R x SMAP_R cube-map cells; for every cell the host lists the vertices
// that can be the support for SOME direction of the cell, so a query reads one cell and scans its
// few candidates (6 on average for the Franka links at SMAP_R = 8) instead of all 41-152 vertices.
// The list is a rigorous superset: with c the cell's centre direction, u0 = support(c) and
// delta = max |d - c| over the unit directions d of the cell, a vertex v that is the support for such
// a d satisfies v.d >= u0.d, hence (u0 - v).c <= (u0 - v).(c - d) <= |u0 - v| delta.  Every v that
// passes this test (with delta inflated by 2 %: a direction rounded into the neighbouring cell is
// still covered) is listed.  The value returned is therefore the same as a full scan's, up to the
// rounding of the dot products that the certification tolerance already covers.
constexpr int SMAP_R = 8;
constexpr int SMAP_CELLS = 6 * SMAP_R * SMAP_R;
template <typename T> VK_HD int smap_cell(V3<T> d) {
  const T ax = vk_abs(d.x), ay = vk_abs(d.y), az = vk_abs(d.z);
  int k; T m, u, v;
  if (ax >= ay && ax >= az) { k = 0; m = d.x; u = d.y; v = d.z; }
  else if (ay >= az) { k = 1; m = d.y; u = d.z; v = d.x; }
  else { k = 2; m = d.z; u = d.x; v = d.y; }
  const T am = vk_abs(m);
  if (!(am > T(0))) return -1;
  const T inv = T(1) / am;
  int iu = (int)((u * inv + T(1)) * T(SMAP_R / 2)), iv = (int)((v * inv + T(1)) * T(SMAP_R / 2));
  iu = iu < 0 ? 0 : (iu > SMAP_R - 1 ? SMAP_R - 1 : iu);
  iv = iv < 0 ? 0 : (iv > SMAP_R - 1 ? SMAP_R - 1 : iv);
  return ((2 * k + (m < T(0) ? 1 : 0)) * SMAP_R + iu) * SMAP_R + iv;
}
// support vertex through the map (cells: offset << 8 | count into ids; ids: local vertex numbers)
template <typename T>
VK_HD int support_mapped(const Vtx<T> *__restrict__ v, int nvert, const uint32_t *__restrict__ cells, const uint8_t *__restrict__ ids, V3<T> d) {
  const int c = smap_cell(d);
  int bi = 0;
  if (c < 0) {
    T best = v[0].x * d.x + v[0].y * d.y + v[0].z * d.z;
    for (int i = 1; i < nvert; i++) { const T t = v[i].x * d.x + v[i].y * d.y + v[i].z * d.z; if (t > best) { best = t; bi = i; } }
    return bi;
  }
  const uint32_t e = cells[c];
  const uint8_t *id = ids + (e >> 8);
  const int n = (int)(e & 255u);
  bi = id[0];
  T best = v[bi].x * d.x + v[bi].y * d.y + v[bi].z * d.z;
  for (int k = 1; k < n; k++) {
    const int i = id[k];
    const T t = v[i].x * d.x + v[i].y * d.y + v[i].z * d.z;
    if (t > best) { best = t; bi = i; }
  }
  return bi;
}

// Hill-climbing support on the hull's vertex graph: from `start`, move to the best strictly
// better neighbour until none is better.  On a convex polytope with its full edge graph a vertex
// without a better neighbour is a global maximiser of the linear function, so this returns a
// true support vertex (up to rounding of the dot products).  adj_start is indexed by local
// vertex id (nvert+1 entries for this shape), adj holds local neighbour ids.
template <typename T>
VK_HD int support_hill(const Vtx<T> *__restrict__ v, const uint16_t *__restrict__ adj_start,
                       const uint8_t *__restrict__ adj, V3<T> d, int start) {
  int cur = start;
  T best = v[cur].x * d.x + v[cur].y * d.y + v[cur].z * d.z;
  for (;;) {
    int cj = cur;
    T cb = best;
    for (int e = adj_start[cur]; e < adj_start[cur + 1]; e++) {
      const int j = adj[e];
      const T t = v[j].x * d.x + v[j].y * d.y + v[j].z * d.z;
      if (t > cb) { cb = t; cj = j; }
    }
    if (cj == cur) return cur;
    cur = cj;
    best = cb;
  }
}

// start vertex for a cold query: the precomputed extreme vertex along the dominant axis of d
template <typename T> VK_HD int hill_start(const Shape<T> &s, V3<T> d) {
  const T ax = vk_abs(d.x), ay = vk_abs(d.y), az = vk_abs(d.z);
  int k = (ax >= ay && ax >= az) ? 0 : (ay >= az ? 1 : 2);
  const T c = k == 0 ? d.x : (k == 1 ? d.y : d.z);
  return s.ext[2 * k + (c < T(0) ? 1 : 0)];
}

template <typename T> VK_HD V3<T> support_cyl(const Shape<T> &s, V3<T> d) {
  V3<T> ax = mk<T>(s.ax[0], s.ax[1], s.ax[2]);
  V3<T> c = mk<T>(s.c[0], s.c[1], s.c[2]);
  T da = dot(d, ax);
  V3<T> rd = d - ax * da;  // radial part
  T rn = vk_sqrt(dot(rd, rd));
  V3<T> p = c + ax * (da >= T(0) ? s.halflen : -s.halflen);
  if (rn > Num<T>::tiny) p = p + rd * (s.radius / rn);
  return p;
}

template <typename T>
VK_HD V3<T> support_shape(const Shape<T> &s, const Vtx<T> *__restrict__ verts, V3<T> d) {
  if (s.kind == SK_CYL) return support_cyl(s, d);
  return support_verts(verts + s.vadr, s.nvert, d);
}

// swept radius of a shape (0 for cylinders, whose `radius` is part of the core)
template <typename T> VK_HD T swept_radius(const Shape<T> &s) { return s.kind == SK_VERTS ? s.radius : T(0); }

// relative pose of B in A's frame
template <typename T> struct Rel { M3<T> R; V3<T> t; };
template <typename T> VK_HD Rel<T> relative_pose(const Pose<T> &A, const Pose<T> &B) {
  Rel<T> r;
  Q4<T> ai = qconj(A.q);
  r.R = q2mat(qmul(ai, B.q));
  r.t = qrot(ai, B.p - A.p);
  return r;
}

// ------------------------------------------------------------------------------ GJK three-way classifier
// A in its own frame, B through `rel`.  R = swept radii + margin.  The loop is exposed as
// init + step so the kernel can run ONE iteration per trip of a persistent-lane loop (lanes
// that finish fetch the next work item instead of idling until the slowest lane is done).
template <typename T> struct GjkState {
  V3<T> v, p0, p1, p2, p3;
  int n, it;
};

template <typename T>
VK_HD void gjk_init(GjkState<T> &s, const Shape<T> &A, const Shape<T> &B, const Rel<T> &rel) {
  V3<T> cA = mk<T>(A.bc[0], A.bc[1], A.bc[2]);
  V3<T> cB = mul(rel.R, mk<T>(B.bc[0], B.bc[1], B.bc[2])) + rel.t;
  s.v = cA - cB;
  if (!(dot(s.v, s.v) > Num<T>::tiny)) s.v = mk<T>(T(1), T(0), T(0));
  s.p0 = s.p1 = s.p2 = s.p3 = s.v;
  s.n = 0;
  s.it = 0;
}

// one iteration; returns -1 to continue or the verdict
// supA(d): support point of A along d (A frame); supB(d): support point of B along d (B frame).
template <typename T, typename SupA, typename SupB>
VK_HD int gjk_step_impl(GjkState<T> &s, const Rel<T> &rel, T R, SupA supA, SupB supB) {
  const T tol = Num<T>::tol;
  if (s.it >= Num<T>::maxit) return V_UNC;
  s.it++;
  // support of A-B along -v
  V3<T> v = s.v;
  V3<T> sa = supA(-v);
  V3<T> sb = mul(rel.R, supB(mulT(rel.R, v))) + rel.t;
  V3<T> w = sa - sb;
  T vv = dot(v, v), vw = dot(v, w);
  // every x in A-B has x.v >= v.w  =>  distance >= v.w/|v|
  if (vw > T(0)) {
    T lim = R + tol;
    if (vw * vw > lim * lim * vv) return V_SEP;
  }
  // converged: |v| is the core distance up to rounding and it sits inside the tol band
  // (otherwise SEP above or PEN below would have fired) -> uncertain
  if (s.n > 0 && (vv - vw) <= Num<T>::conv_rel * vv) return V_UNC;
  // append w and solve for the closest point of the simplex
  T l0 = T(0), l1 = T(0), l2 = T(0), l3 = T(0);
  bool inside = false;
  if (s.n == 0) { s.p0 = w; l0 = T(1); }
  else if (s.n == 1) { s.p1 = w; solve1(s.p0, s.p1, l0, l1); }
  else if (s.n == 2) { s.p2 = w; solve2(s.p0, s.p1, s.p2, l0, l1, l2); }
  else { s.p3 = w; inside = solve3(s.p0, s.p1, s.p2, s.p3, l0, l1, l2, l3); }
  if (inside) {
    // origin enclosed by the core tetrahedron: depth >= min face distance
    T dep = vk_min(vk_min(plane_dist(s.p1, s.p2, s.p3), plane_dist(s.p0, s.p2, s.p3)),
                   vk_min(plane_dist(s.p0, s.p1, s.p3), plane_dist(s.p0, s.p1, s.p2)));
    return (dep + R > tol) ? V_PEN : V_UNC;
  }
  v = s.p0 * l0 + s.p1 * l1 + s.p2 * l2 + s.p3 * l3;
  s.v = v;
  // v is a point of A-B: core distance <= |v|
  T nvv = dot(v, v);
  if (R > tol && nvv < (R - tol) * (R - tol)) return V_PEN;
  if (!(nvv > tol * tol)) return V_UNC;  // cores (nearly) touching: cannot be certified either way
  // compact the simplex: keep vertices with positive weight at the front
  bool k0 = l0 > T(0), k1 = l1 > T(0), k2 = l2 > T(0), k3 = l3 > T(0);
  { bool c = !k0 && k1; cswap(c, s.p0, s.p1); cswap(c, k0, k1); }
  { bool c = !k1 && k2; cswap(c, s.p1, s.p2); cswap(c, k1, k2); }
  { bool c = !k2 && k3; cswap(c, s.p2, s.p3); cswap(c, k2, k3); }
  { bool c = !k0 && k1; cswap(c, s.p0, s.p1); cswap(c, k0, k1); }
  { bool c = !k1 && k2; cswap(c, s.p1, s.p2); cswap(c, k1, k2); }
  { bool c = !k0 && k1; cswap(c, s.p0, s.p1); cswap(c, k0, k1); }
  s.n = int(k0) + int(k1) + int(k2) + int(k3);
  if (s.n == 4) return V_UNC;  // cannot happen unless solve3 misreported
  return -1;
}

template <typename T>
VK_HD int gjk_step(GjkState<T> &s, const Shape<T> &A, const Shape<T> &B, const Vtx<T> *__restrict__ verts,
                   const Rel<T> &rel, T R) {
  return gjk_step_impl(
      s, rel, R, [&](V3<T> d) { return support_shape(A, verts, d); }, [&](V3<T> d) { return support_shape(B, verts, d); });
}

template <typename T>
VK_HD int gjk_classify(const Shape<T> &A, const Shape<T> &B, const Vtx<T> *__restrict__ verts,
                       const Rel<T> &rel, T R, int *iters) {
  GjkState<T> s;
  gjk_init(s, A, B, rel);
  int verdict;
  do { verdict = gjk_step(s, A, B, verts, rel, R); } while (verdict < 0);
  if (iters) *iters = s.it;
  return verdict;
}

// ------------------------------------------------------------------------------ segment-segment (sphere/capsule pairs)
// mjc_SphereSphere / mjc_SphereCapsule / mjc_CapsuleCapsule: core = point or segment.
template <typename T>
VK_HD int segseg_classify(V3<T> p1, V3<T> q1, V3<T> p2, V3<T> q2, T R) {
  V3<T> d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
  T a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r), s, t;
  if (!(a > Num<T>::tiny) && !(e > Num<T>::tiny)) { s = t = T(0); }
  else if (!(a > Num<T>::tiny)) { s = T(0); t = f / e; t = t < T(0) ? T(0) : (t > T(1) ? T(1) : t); }
  else {
    T c = dot(d1, r);
    if (!(e > Num<T>::tiny)) { t = T(0); s = -c / a; s = s < T(0) ? T(0) : (s > T(1) ? T(1) : s); }
    else {
      T b = dot(d1, d2), den = a * e - b * b;
      s = den > T(1e-7) * a * e ? (b * f - c * e) / den : T(0);
      s = s < T(0) ? T(0) : (s > T(1) ? T(1) : s);
      t = (b * s + f) / e;
      if (t < T(0)) { t = T(0); s = -c / a; s = s < T(0) ? T(0) : (s > T(1) ? T(1) : s); }
      else if (t > T(1)) { t = T(1); s = (b - c) / a; s = s < T(0) ? T(0) : (s > T(1) ? T(1) : s); }
    }
  }
  V3<T> dd = (p1 + d1 * s) - (p2 + d2 * t);
  T dist = vk_sqrt(dot(dd, dd));
  const T tol = Num<T>::tol;
  // near-parallel segments: the clamped solution may be a non-optimal (larger) distance,
  // which is only safe for the PEN verdict; send borderline cases to fp64
  if (dist > R + tol) {
    T b = dot(d1, d2), den = a * e - b * b;
    if (a > Num<T>::tiny && e > Num<T>::tiny && !(den > T(1e-4) * a * e) && dist < R + T(100) * tol + T(1e-3))
      return V_UNC;
    return V_SEP;
  }
  if (dist < R - tol) return V_PEN;
  return V_UNC;
}

// ------------------------------------------------------------------------------ OBB-OBB separating-axis cull
// true  => the two boxes are certainly disjoint by more than `slack` (conservative cull).
// Boxes: centre/axes/half extents of A in A's frame, B brought over with `rel`.
template <typename T>
VK_HD bool obb_disjoint(const Shape<T> &A, const Shape<T> &B, const Rel<T> &rel, T slack) {
  // express everything in A's OBB frame
  M3<T> Ra; for (int i = 0; i < 9; i++) Ra.m[i] = A.orot[i];
  M3<T> Rb; for (int i = 0; i < 9; i++) Rb.m[i] = B.orot[i];
  // C = Ra^T * rel.R * Rb
  M3<T> RB;  // rel.R * Rb
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      RB.m[3 * i + j] = rel.R.m[3 * i] * Rb.m[j] + rel.R.m[3 * i + 1] * Rb.m[3 + j] + rel.R.m[3 * i + 2] * Rb.m[6 + j];
  T Cm[9], Ab[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      Cm[3 * i + j] = Ra.m[i] * RB.m[j] + Ra.m[3 + i] * RB.m[3 + j] + Ra.m[6 + i] * RB.m[6 + j];
      Ab[3 * i + j] = vk_abs(Cm[3 * i + j]) + T(1e-6);
    }
  V3<T> cb = mul(rel.R, mk<T>(B.oc[0], B.oc[1], B.oc[2])) + rel.t;
  V3<T> dw = cb - mk<T>(A.oc[0], A.oc[1], A.oc[2]);
  T tt[3];
  { V3<T> t3 = mulT(Ra, dw); tt[0] = t3.x; tt[1] = t3.y; tt[2] = t3.z; }
  const T *a = A.ohalf, *b = B.ohalf;
  // A's face axes
  for (int i = 0; i < 3; i++) {
    T rb = b[0] * Ab[3 * i] + b[1] * Ab[3 * i + 1] + b[2] * Ab[3 * i + 2];
    if (vk_abs(tt[i]) > a[i] + rb + slack) return true;
  }
  // B's face axes
  for (int j = 0; j < 3; j++) {
    T ra = a[0] * Ab[j] + a[1] * Ab[3 + j] + a[2] * Ab[6 + j];
    T tp = tt[0] * Cm[j] + tt[1] * Cm[3 + j] + tt[2] * Cm[6 + j];
    if (vk_abs(tp) > ra + b[j] + slack) return true;
  }
  return false;
}

// ------------------------------------------------------------------------------ narrow phase of one item
// plane (A, world fixed) against B.  mjc_PlaneConvex / PlaneSphere / PlaneCapsule / PlaneBox:
// deepest point of B along -n;  mjc_PlaneCylinder analytic.  supB as in gjk_step_impl.
template <typename T, typename SupB>
VK_HD int plane_classify(const Shape<T> &A, const Shape<T> &B, const Pose<T> &PB, T R, SupB supB) {
  const T tol = Num<T>::tol;
  V3<T> n = mk<T>(A.ax[0], A.ax[1], A.ax[2]);
  V3<T> c = mk<T>(A.c[0], A.c[1], A.c[2]);
  T dist;
  if (B.kind == SK_CYL) {
    V3<T> ax = qrot(PB.q, mk<T>(B.ax[0], B.ax[1], B.ax[2]));
    V3<T> cb = PB.p + qrot(PB.q, mk<T>(B.c[0], B.c[1], B.c[2]));
    T na = dot(n, ax);
    T rad = T(1) - na * na;
    dist = dot(n, cb - c) - vk_abs(na) * B.halflen - B.radius * vk_sqrt(rad > T(0) ? rad : T(0));
  } else {
    V3<T> nl = qrot(qconj(PB.q), n);  // plane normal in B's frame
    V3<T> sp = supB(-nl);
    dist = dot(n, PB.p + qrot(PB.q, sp) - c);
  }
  if (dist > R + tol) return V_SEP;
  if (dist < R - tol) return V_PEN;
  return V_UNC;
}

// conservative cull: true => B's OBB is certainly above the plane by more than `slack`
template <typename T>
VK_HD bool obb_above_plane(const Shape<T> &A, const Shape<T> &B, const Pose<T> &PB, T slack) {
  V3<T> n = mk<T>(A.ax[0], A.ax[1], A.ax[2]);
  V3<T> nl = qrot(qconj(PB.q), n);
  T ext = vk_abs(nl.x * B.orot[0] + nl.y * B.orot[3] + nl.z * B.orot[6]) * B.ohalf[0] +
          vk_abs(nl.x * B.orot[1] + nl.y * B.orot[4] + nl.z * B.orot[7]) * B.ohalf[1] +
          vk_abs(nl.x * B.orot[2] + nl.y * B.orot[5] + nl.z * B.orot[8]) * B.ohalf[2];
  T cen = dot(n, PB.p - mk<T>(A.c[0], A.c[1], A.c[2])) + dot(nl, mk<T>(B.oc[0], B.oc[1], B.oc[2]));
  return cen - ext > slack + T(1e-6);
}

// segment-segment item (sphere / capsule cores), world frame
template <typename T>
VK_HD int segseg_item(const Shape<T> &A, const Shape<T> &B, const Vtx<T> *verts, const Pose<T> &PA, const Pose<T> &PB, T R) {
  Vtx<T> a0 = verts[A.vadr], a1 = verts[A.vadr + A.nvert - 1];
  Vtx<T> b0 = verts[B.vadr], b1 = verts[B.vadr + B.nvert - 1];
  V3<T> p1 = PA.p + qrot(PA.q, mk<T>(a0.x, a0.y, a0.z)), q1 = PA.p + qrot(PA.q, mk<T>(a1.x, a1.y, a1.z));
  V3<T> p2 = PB.p + qrot(PB.q, mk<T>(b0.x, b0.y, b0.z)), q2 = PB.p + qrot(PB.q, mk<T>(b1.x, b1.y, b1.z));
  return segseg_classify(p1, q1, p2, q2, R);
}

// ------------------------------------------------------------------------------ bounding-capsule cull
// squared distance between the segments p1..q1 and p2..q2 (either may be a point): closed form with
// clamping (exact in exact arithmetic: the clamped coordinate steps below can only decrease the
// distance of the pair found so far).  The scalar part runs in fp64 whatever T is: for nearly
// parallel segments `a e - b b` cancels, and in fp32 the pair that comes out can be off by
// angle x length (4e-4 m measured on unit segments) -- too much for a cull that has to be
// conservative.  B200 issues fp64 at half the fp32 rate, and this is ~35 flops per shape pair.
// inv_a / inv_e: 1 / |q1-p1|^2 and 1 / |q2-p2|^2 (0 for a point).  The three quotients are products
// with fp32-accurate reciprocals: the parameters s, t only need to be feasible and near the optimum
// (a relative error of 1e-7 in them moves the distance by 1e-7 x length), it is the cancellation in
// numerator and denominator that needs the wide format.
VK_HD float vk_rcp(float x) {
#if defined(__CUDA_ARCH__)
  return __frcp_rn(x);
#else
  return 1.0f / x;
#endif
}
template <typename T>
VK_HD T segseg_dist2(V3<T> p1, V3<T> q1, T inv_a, V3<T> p2, V3<T> q2, T inv_e) {
  const V3<double> P1 = mk<double>(p1.x, p1.y, p1.z), P2 = mk<double>(p2.x, p2.y, p2.z);
  const V3<double> d1 = mk<double>(q1.x, q1.y, q1.z) - P1, d2 = mk<double>(q2.x, q2.y, q2.z) - P2, r = P1 - P2;
  const double a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r), c = dot(d1, r), b = dot(d1, d2);
  const double den = a * e - b * b;
  double s = den > 1e-14 * a * e ? (b * f - c * e) * (double)vk_rcp((float)den) : 0.0;
  s = s < 0.0 ? 0.0 : (s > 1.0 ? 1.0 : s);
  double t = (b * s + f) * (double)inv_e;
  t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
  s = (b * t - c) * (double)inv_a;
  s = s < 0.0 ? 0.0 : (s > 1.0 ? 1.0 : s);
  const V3<double> dd = r + d1 * s - d2 * t;
  return (T)dot(dd, dd);
}
template <typename T> VK_HD T segseg_dist2(V3<T> p1, V3<T> q1, V3<T> p2, V3<T> q2) {
  const V3<T> d1 = q1 - p1, d2 = q2 - p2;
  const T a = dot(d1, d1), e = dot(d2, d2);
  return segseg_dist2(p1, q1, a > T(1e-30) ? T(1) / a : T(0), p2, q2, e > T(1e-30) ? T(1) / e : T(0));
}

// true => the bounding capsules of the two shapes are certainly further apart than rsum_core +
// slack, i.e. the pair cannot be in contact.  `lim` = margin + slack (the capsule radii already
// include the swept radii).  World-fixed shapes (slot < 0) carry world-frame capsules: no transform.
template <typename T>
VK_HD bool capsule_cull(const Pair &pr, const Shape<T> &A, const Shape<T> &B, const Pose<T> &PA, const Pose<T> &PB, T lim) {
  const bool bseg = B.caplen > T(0), bmov = B.slot >= 0;
  V3<T> b0 = mk<T>(B.ca[0], B.ca[1], B.ca[2]), b1 = mk<T>(B.cb[0], B.cb[1], B.cb[2]);
  if (bmov) { b0 = PB.p + qrot(PB.q, b0); b1 = bseg ? PB.p + qrot(PB.q, b1) : b0; }
  if (pr.kind == PK_PLANE) {
    const V3<T> n = mk<T>(A.ax[0], A.ax[1], A.ax[2]), c = mk<T>(A.c[0], A.c[1], A.c[2]);
    const T d = vk_min(dot(n, b0 - c), dot(n, b1 - c));
    return d > B.crad + lim;
  }
  const bool aseg = A.caplen > T(0), amov = A.slot >= 0;
  V3<T> a0 = mk<T>(A.ca[0], A.ca[1], A.ca[2]), a1 = mk<T>(A.cb[0], A.cb[1], A.cb[2]);
  if (amov) { a0 = PA.p + qrot(PA.q, a0); a1 = aseg ? PA.p + qrot(PA.q, a1) : a0; }
  const T r = A.crad + B.crad + lim;
  T d2;
  if (!aseg && !bseg) { const V3<T> d = a0 - b0; d2 = dot(d, d); }
  else d2 = segseg_dist2(a0, a1, aseg ? (T)vk_rcp((float)(A.caplen * A.caplen)) : T(0), b0, b1,
                         bseg ? (T)vk_rcp((float)(B.caplen * B.caplen)) : T(0));
  return d2 > r * r;
}

// the mid-phase cull of one (row, pair) that survived the bounding spheres: true => certainly no
// contact.  Bounding capsules first (cheap, and tight for the long thin shapes spheres are useless
// for), then the OBB separating axes for the pairs whose narrow phase is expensive enough (PF_OBB).
template <typename T>
VK_HD bool midphase_cull(const Pair &pr, const Shape<T> &A, const Shape<T> &B, const Pose<T> &PA, const Pose<T> &PB,
                         T margin, T slack) {
  if (pr.kind != PK_SEGSEG && capsule_cull(pr, A, B, PA, PB, margin + slack)) return true;
  if (!(pr.flags & PF_OBB)) return false;
  if (pr.kind == PK_PLANE) return obb_above_plane(A, B, PB, margin + slack);
  Rel<T> rel = relative_pose(PA, PB);
  return obb_disjoint(A, B, rel, margin + slack);
}

// ------------------------------------------------------------------------------ cull groups (vk_pipe.cuh)
// Level 0 of the broad phase works on GROUPS: a moving body with all its collision shapes (one
// bounding sphere in the body frame), or one world-fixed shape (its bounding capsule / plane in the
// world frame).  A group pair stands for all the shape pairs between its two groups.
enum GroupPairKind : int { GK_SPHERE = 0, GK_CAPSULE = 1, GK_PLANE = 2 };
struct GroupPair {
  uint16_t ga;      // moving group (index into the moving-group table = row of the centre array)
  uint16_t gb;      // GK_SPHERE: moving group; GK_CAPSULE / GK_PLANE: index into the static table
  uint16_t first;   // its shape pairs: member[first .. first + n)
  uint8_t n, kind;
  float lim;        // GK_SPHERE / GK_CAPSULE: (radii + largest margin of the member pairs + slack)^2, compared with a
                    // squared distance; GK_PLANE: radius + margin + slack, compared with the signed height
  float lim_in;     // the same for the two groups' INNER radii around the same centres (a ball around a moving
                    // group's centre that lies inside one of its shapes; the tube inside a world-fixed shape around
                    // its segment).  Centres closer than this: the two shapes intersect, the row is certainly in
                    // contact.  No inner shapes: 0 (squared kinds) / -1e30 (plane) -- the test never fires.
};
static_assert(sizeof(GroupPair) == 16, "GroupPair is loaded as one 16-byte word");
struct StaticGroup {   // world frame
  float a[3];          // GK_CAPSULE: segment start;   GK_PLANE: a point of the plane
  float len;           // GK_CAPSULE: segment length (0 for a point)
  float u[3];          // GK_CAPSULE: unit vector along the segment (0 for a point);  GK_PLANE: unit normal
  float th;            // GK_CAPSULE: the shape's inner tube runs along |s - len / 2| <= th of the segment (< 0: none)
};
static_assert(sizeof(StaticGroup) == 32, "StaticGroup is loaded as two 16-byte words");
constexpr int MAX_GROUP = MAX_BODY;

// level-0 test of one group pair.  Returns 0: no shape pair between the two groups can be in contact;
// 1: near (expand the group pair); 2: CERTAIN contact (the inner ball / tube of the two groups overlap).
// cA: world centre of moving group ga; cB: world centre of moving group gb (GK_SPHERE only);
// S: the static group (GK_CAPSULE / GK_PLANE).
// squared distance from e (relative to the segment's start) to the segment of unit direction u and length len;
// *s_out = parameter of the closest point.  |e - s u|^2 expanded: 14 flops; the cancellation costs a few ulps
// of |e|^2 (< 1e-6 m^2 at arm's length), far inside the slack / safety of the two limits it is compared with.
VK_HD float point_segment_d2(V3<float> e, V3<float> u, float len, float *s_out) {
  const float ee = dot(e, e), s = dot(e, u);
  const float sc = s < 0.f ? 0.f : (s > len ? len : s);
  *s_out = sc;
  return ee - sc * (2.f * s - sc);
}
VK_HD int group_pair_test(const GroupPair &g, V3<float> cA, V3<float> cB, const StaticGroup *S) {
  float d2;
  if (g.kind == GK_SPHERE) {
    const V3<float> d = cA - cB;
    d2 = dot(d, d);
  } else {
    const V3<float> e = cA - mk<float>(S->a[0], S->a[1], S->a[2]);
    const V3<float> u = mk<float>(S->u[0], S->u[1], S->u[2]);
    if (g.kind == GK_PLANE) {
      const float h = dot(e, u);
      return h < g.lim_in ? 2 : (h <= g.lim ? 1 : 0);
    }
    float sc;
    d2 = point_segment_d2(e, u, S->len, &sc);
    if (!(fabsf(sc - 0.5f * S->len) <= S->th)) return d2 <= g.lim ? 1 : 0;
  }
  return d2 < g.lim_in ? 2 : (d2 <= g.lim ? 1 : 0);
}
VK_HD bool group_pair_near(const GroupPair &g, V3<float> cA, V3<float> cB, const StaticGroup *S) { return group_pair_test(g, cA, cB, S) != 0; }

// Distance^2 between two points of the segments near their closest pair, all in the caller's precision.  The
// parameters are feasible (clamped to [0, 1]), so the value is the distance of two actual points of the
// segments: never BELOW the true minimum by more than rounding -- the safe side for a certain-contact test
// (the bounding-capsule cull needs the opposite guarantee and uses the fp64 routine above).
template <typename T>
VK_HD T segseg_upper2(V3<T> p1, V3<T> q1, T inv_a, V3<T> p2, V3<T> q2, T inv_e) {
  const V3<T> d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
  const T a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r), c = dot(d1, r), b = dot(d1, d2);
  const T den = a * e - b * b;
  T s = den > T(1e-6) * a * e ? (b * f - c * e) / den : T(0);
  s = s < T(0) ? T(0) : (s > T(1) ? T(1) : s);
  T t = (b * s + f) * inv_e;
  t = t < T(0) ? T(0) : (t > T(1) ? T(1) : t);
  s = (b * t - c) * inv_a;
  s = s < T(0) ? T(0) : (s > T(1) ? T(1) : s);
  const V3<T> dd = r + d1 * s - d2 * t;
  return dot(dd, dd);
}

// certain contact of one shape pair from the inner capsules: true => the shapes intersect (distance < 0 <= margin)
template <typename T>
VK_HD bool inner_contact(const Pair &pr, const Shape<T> &A, const Shape<T> &B, const Pose<T> &PA, const Pose<T> &PB) {
  if (!(B.irad > T(0))) return false;
  const bool bmov = B.slot >= 0;
  V3<T> b0 = mk<T>(B.ia[0], B.ia[1], B.ia[2]), b1 = mk<T>(B.ib[0], B.ib[1], B.ib[2]);
  const bool bseg = !(b0.x == b1.x && b0.y == b1.y && b0.z == b1.z);
  if (bmov) { b0 = PB.p + qrot(PB.q, b0); b1 = bseg ? PB.p + qrot(PB.q, b1) : b0; }
  const T safety = sizeof(T) == 4 ? T(1e-5) : T(0);   // fp32 forward kinematics and end points: a few 1e-6 at arm's length
  if (pr.kind == PK_PLANE) {
    const V3<T> n = mk<T>(A.ax[0], A.ax[1], A.ax[2]), c = mk<T>(A.c[0], A.c[1], A.c[2]);
    return vk_min(dot(n, b0 - c), dot(n, b1 - c)) < B.irad - safety;
  }
  if (!(A.irad > T(0))) return false;
  const bool amov = A.slot >= 0;
  V3<T> a0 = mk<T>(A.ia[0], A.ia[1], A.ia[2]), a1 = mk<T>(A.ib[0], A.ib[1], A.ib[2]);
  const bool aseg = !(a0.x == a1.x && a0.y == a1.y && a0.z == a1.z);
  if (amov) { a0 = PA.p + qrot(PA.q, a0); a1 = aseg ? PA.p + qrot(PA.q, a1) : a0; }
  const T r = A.irad + B.irad - safety;
  if (!(r > T(0))) return false;
  T d2;
  if (!aseg && !bseg) { const V3<T> d = a0 - b0; d2 = dot(d, d); }
  else {
    const V3<T> da = a1 - a0, db = b1 - b0;
    const T la = dot(da, da), lb = dot(db, db);
#ifdef VK_INNER_FP64
    d2 = segseg_dist2(a0, a1, aseg ? T(1) / la : T(0), b0, b1, bseg ? T(1) / lb : T(0));
#else
    d2 = segseg_upper2(a0, a1, aseg ? T(1) / la : T(0), b0, b1, bseg ? T(1) / lb : T(0));
#endif
  }
  return d2 < r * r;
}

// ------------------------------------------------------------------------------ narrow phase of one item (scalar)
template <typename T>
VK_HD int narrow_item(int kind, const Shape<T> &A, const Shape<T> &B, const Vtx<T> *verts,
                      const Pose<T> &PA, const Pose<T> &PB, T R) {
  if (kind == PK_PLANE)
    return plane_classify(A, B, PB, R, [&](V3<T> d) { return support_verts(verts + B.vadr, B.nvert, d); });
  if (kind == PK_SEGSEG) return segseg_item(A, B, verts, PA, PB, R);
  Rel<T> rel = relative_pose(PA, PB);
  return gjk_classify(A, B, verts, rel, R, (int *)nullptr);
}

// ------------------------------------------------------------------------------ pose constraint (CBiRRT projection)
// PoseConstraint (reference: src/mjpl/constraint/pose_constraint.py): a site must stay inside
// a box of translations / roll-pitch-yaw expressed in a constraint frame C.
//   _displacement_from_constraint :93-123, _get_jacobian :125-147, _e_rpy :150-171,
//   valid_config :72-76, apply :78-91 (Newton-like projection q -= J^T pinv(J J^T) dx).
// fp64 throughout, one lane per row (the iteration is a short dense-algebra loop per row).
struct PoseSpec {
  int site_slot;            // pose slot of the site's body, -1 = world fixed
  unsigned jnt_mask;        // bit j: joint j moves the site (ancestor joints of its body)
  double site_pos[3], site_quat[4];   // site in its body frame (world frame if site_slot < 0)
  double cw_pos[3], cw_quat[4];       // C_T_world = inverse of the reference frame
  double lo[6], hi[6];                // allowed x, y, z, roll, pitch, yaw in C
  double tolerance, q_step;
};

VK_HD void quat2rpy(const Q4<double> &q, double &roll, double &pitch, double &yaw) {
  // mink / jaxlie SO3.as_rpy_radians (quaternion wxyz)
  roll = atan2(2.0 * (q.w * q.x + q.y * q.z), 1.0 - 2.0 * (q.x * q.x + q.y * q.y));
  double sp = 2.0 * (q.w * q.y - q.z * q.x);
  sp = sp > 1.0 ? 1.0 : (sp < -1.0 ? -1.0 : sp);
  pitch = asin(sp);
  yaw = atan2(2.0 * (q.w * q.z + q.x * q.y), 1.0 - 2.0 * (q.y * q.y + q.z * q.z));
}

// FK of all slots + world anchors/axes of every joint (mj_kinematics' xanchor / xaxis)
VK_HD void fk_with_joints(const FkTables<double> &fk, int nslot, const double *q, Pose<double> *P,
                                               V3<double> *anchor, V3<double> *axis) {
  Pose<double> ident; ident.p = mk<double>(0, 0, 0); ident.q.w = 1; ident.q.x = ident.q.y = ident.q.z = 0;
  for (int s = 0; s < nslot; s++) {
    const int ps = fk.body_parent[s];
    const Pose<double> &par = ps < 0 ? ident : P[ps];
    Pose<double> o;
    o.p = par.p + qrot(par.q, mk<double>(fk.body_pos[s][0], fk.body_pos[s][1], fk.body_pos[s][2]));
    Q4<double> bq; bq.w = fk.body_quat[s][0]; bq.x = fk.body_quat[s][1]; bq.y = fk.body_quat[s][2]; bq.z = fk.body_quat[s][3];
    o.q = qmul(par.q, bq);
    for (int k = 0; k < fk.body_jntnum[s]; k++) {
      const int j = fk.body_jntadr[s] + k, a = fk.jnt_qadr[j];
      V3<double> jp = mk<double>(fk.jnt_pos[j][0], fk.jnt_pos[j][1], fk.jnt_pos[j][2]);
      V3<double> jax = mk<double>(fk.jnt_axis[j][0], fk.jnt_axis[j][1], fk.jnt_axis[j][2]);
      anchor[j] = o.p + qrot(o.q, jp);
      axis[j] = qrot(o.q, jax);
      const double dq = q[a] - fk.qpos0[a];
      if (fk.jnt_type[j] == JK_SLIDE) o.p = o.p + axis[j] * dq;
      else {
        double sn, cs;
        vk_sincos(0.5 * dq, &sn, &cs);
        Q4<double> ql; ql.w = cs; ql.x = jax.x * sn; ql.y = jax.y * sn; ql.z = jax.z * sn;
        o.q = qmul(o.q, ql);
        o.p = anchor[j] - qrot(o.q, jp);
      }
    }
    o.q = qnormalize(o.q);
    P[s] = o;
  }
}

// displacement of the site from the constraint box, and the site's world pose
VK_HD double pose_displacement(const PoseSpec &sp, const Pose<double> *P, double *dx, Pose<double> &site) {
  Pose<double> ident; ident.p = mk<double>(0, 0, 0); ident.q.w = 1; ident.q.x = ident.q.y = ident.q.z = 0;
  const Pose<double> &B = sp.site_slot < 0 ? ident : P[sp.site_slot];
  Q4<double> sq; sq.w = sp.site_quat[0]; sq.x = sp.site_quat[1]; sq.y = sp.site_quat[2]; sq.z = sp.site_quat[3];
  site.p = B.p + qrot(B.q, mk<double>(sp.site_pos[0], sp.site_pos[1], sp.site_pos[2]));
  site.q = qnormalize(qmul(B.q, sq));
  Q4<double> cq; cq.w = sp.cw_quat[0]; cq.x = sp.cw_quat[1]; cq.y = sp.cw_quat[2]; cq.z = sp.cw_quat[3];
  V3<double> t = mk<double>(sp.cw_pos[0], sp.cw_pos[1], sp.cw_pos[2]) + qrot(cq, site.p);
  Q4<double> r = qmul(cq, site.q);
  double d[6] = {t.x, t.y, t.z, 0, 0, 0};
  quat2rpy(r, d[3], d[4], d[5]);
  double n2 = 0;
  for (int i = 0; i < 6; i++) {
    dx[i] = d[i] > sp.hi[i] ? d[i] - sp.hi[i] : (d[i] < sp.lo[i] ? d[i] - sp.lo[i] : 0.0);
    n2 += dx[i] * dx[i];
  }
  return sqrt(n2);
}

// y = pinv(A) x for a symmetric 6x6 A (cyclic Jacobi eigen-decomposition; eigenvalues below
// 6*eps*max are dropped, which is numpy.linalg.pinv's default cut-off).
// The loops are deliberately NOT unrolled (runtime indices into local arrays): the fully
// unrolled form of this routine was miscompiled for sm_100a by nvcc 12.9 at every optimisation
// level except -G (host and device disagreed on well-conditioned inputs; tools/dbg/pinv_dbg.cu).
VK_HD void sym6_pinv_apply(double Ain[6][6], const double *x, double *y) {
  double A[36], V[36];
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int i = 0; i < 36; i++) { A[i] = Ain[i / 6][i % 6]; V[i] = (i / 6 == i % 6) ? 1.0 : 0.0; }
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int sweep = 0; sweep < 12; sweep++) {
    double off = 0, dg = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int i = 0; i < 36; i++) {
      const double a2 = A[i] * A[i];
      if (i / 6 == i % 6) dg += a2; else off += a2;
    }
    if (off <= 2e-30 * dg || off == 0.0) break;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int pq = 0; pq < 36; pq++) {
      const int p = pq / 6, q = pq % 6;
      if (q <= p) continue;
      const double apq = A[p * 6 + q];
      if (fabs(apq) < 1e-300) continue;
      const double th = (A[q * 6 + q] - A[p * 6 + p]) / (2.0 * apq);
      const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
      const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
      for (int k = 0; k < 6; k++) {
        const double akp = A[k * 6 + p], akq = A[k * 6 + q];
        A[k * 6 + p] = c * akp - s * akq;
        A[k * 6 + q] = s * akp + c * akq;
        const double vkp = V[k * 6 + p], vkq = V[k * 6 + q];
        V[k * 6 + p] = c * vkp - s * vkq;
        V[k * 6 + q] = s * vkp + c * vkq;
      }
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
      for (int k = 0; k < 6; k++) {
        const double apk = A[p * 6 + k], aqk = A[q * 6 + k];
        A[p * 6 + k] = c * apk - s * aqk;
        A[q * 6 + k] = s * apk + c * aqk;
      }
    }
  }
  double lmax = 0;
  for (int i = 0; i < 6; i++) lmax = fmax(lmax, fabs(A[i * 6 + i]));
  const double cut = 6.0 * 2.220446049250313e-16 * lmax;
  for (int i = 0; i < 6; i++) y[i] = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int k = 0; k < 6; k++) {
    const double lam = A[k * 6 + k];
    if (!(fabs(lam) > cut)) continue;
    double cc = 0;
    for (int i = 0; i < 6; i++) cc += V[i * 6 + k] * x[i];
    cc /= lam;
    for (int i = 0; i < 6; i++) y[i] += V[i * 6 + k] * cc;
  }
}

// valid_config of one row (:72-76)
VK_HD bool pose_valid_row(const FkTables<double> &fk, int nslot, const PoseSpec &spec, const double *qin) {
  for (int j = 0; j < fk.njnt; j++)
    if (!(qin[j] >= fk.jnt_lo[j] && qin[j] <= fk.jnt_hi[j])) return false;
  Pose<double> P[MAX_BODY];
  V3<double> anchor[MAX_JNT], axis[MAX_JNT];
  fk_with_joints(fk, nslot, qin, P, anchor, axis);
  double dx[6];
  Pose<double> site;
  return pose_displacement(spec, P, dx, site) <= spec.tolerance;
}

// apply of one row (:78-91); q is updated in place; returns true when the projection succeeded
VK_HD bool pose_project_row(const FkTables<double> &fk, int nslot, const PoseSpec &spec, const double *q_old, double *q,
                            int max_iters, int *iters_out) {
  const int nq = fk.nq;
  Pose<double> P[MAX_BODY];
  V3<double> anchor[MAX_JNT], axis[MAX_JNT];
  int it = 0;
  bool ok = false;
  for (; it < max_iters; it++) {
    fk_with_joints(fk, nslot, q, P, anchor, axis);
    double dx[6];
    Pose<double> site;
    if (pose_displacement(spec, P, dx, site) <= spec.tolerance) { ok = true; break; }
    // RPY Jacobian of the site: E_rpy(world rpy of the site) @ [jacp; jacr]  (:125-171).
    // NB the reference's E_rpy has cos(pitch) at [4,4] where the exact inverse has cos(yaw)
    // (pose_constraint.py:165); restated as is.
    double roll, pitch, yaw;
    quat2rpy(site.q, roll, pitch, yaw);
    const double cp = cos(pitch), sp_ = sin(pitch), cy = cos(yaw), sy = sin(yaw);
    const double E[3][3] = {{cy / cp, sy / cp, 0.0}, {-sy, cp, 0.0}, {cy * (sp_ / cp), sy * (sp_ / cp), 1.0}};
    double J[6][MAX_JNT];
    for (int j = 0; j < fk.njnt; j++) {
      V3<double> jp = mk<double>(0, 0, 0), jr = mk<double>(0, 0, 0);
      if ((spec.jnt_mask >> j) & 1u) {
        if (fk.jnt_type[j] == JK_SLIDE) jp = axis[j];
        else { jr = axis[j]; jp = cross(axis[j], site.p - anchor[j]); }
      }
      const int c = fk.jnt_qadr[j];  // dof index == qpos index for hinge / slide
      J[0][c] = jp.x; J[1][c] = jp.y; J[2][c] = jp.z;
      J[3][c] = E[0][0] * jr.x + E[0][1] * jr.y + E[0][2] * jr.z;
      J[4][c] = E[1][0] * jr.x + E[1][1] * jr.y + E[1][2] * jr.z;
      J[5][c] = E[2][0] * jr.x + E[2][1] * jr.y + E[2][2] * jr.z;
    }
    double A[6][6];
    for (int i = 0; i < 6; i++)
      for (int k = i; k < 6; k++) {
        double acc = 0;
        for (int c = 0; c < nq; c++) acc += J[i][c] * J[k][c];
        A[i][k] = A[k][i] = acc;
      }
    double y[6];
    sym6_pinv_apply(A, dx, y);
    double far2 = 0;
    bool lim = true;
    for (int c = 0; c < nq; c++) {
      double e = 0;
      for (int i = 0; i < 6; i++) e += J[i][c] * y[i];
      q[c] -= e;
      const double d = q[c] - q_old[c];
      far2 += d * d;
    }
    for (int j = 0; j < fk.njnt; j++) lim = lim && q[j] >= fk.jnt_lo[j] && q[j] <= fk.jnt_hi[j];
    if (!lim || sqrt(far2) > 2.0 * spec.q_step) break;  // :88-91
  }
  if (iters_out) *iters_out = it;
  return ok;
}

// ------------------------------------------------------------------------------ inverse kinematics (move-to-pose goals)
// IKSolver.solve_ik (reference: src/mjpl/inverse_kinematics/ik_solver_interface.py:11-28; the
// stock implementation, mink_ik_solver.py:74-117, iterates a QP-based differential IK from an
// initial guess until the site's pose error is within pos/ori tolerance).  The contract, not the
// QP, is what planners consume: a configuration whose site pose is within tolerance of the
// target.  Here each row runs Levenberg-Marquardt damped least squares in fp64:
//   e = [p_t - p; rotvec(R_t R^T)]   (world frame),  J = geometric site Jacobian (movable joints),
//   dq = J^T (J J^T + (lm |e|^2 + damping) I)^-1 e,  |dq|_inf capped, q clamped to joint limits.
struct IkSpec {
  int site_slot;
  unsigned jnt_mask;                 // ancestor joints of the site that are allowed to move
  double site_pos[3], site_quat[4];
  double pos_tol, ori_tol, lm_damping, damping, max_step;
  int iterations;
};

// solve the SPD 6x6 system A y = x (Cholesky; runtime-indexed loops, see sym6_pinv_apply)
VK_HD bool spd6_solve(const double *Ain, const double *x, double *y) {
  double L[36];
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int i = 0; i < 36; i++) L[i] = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int i = 0; i < 6; i++) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int j = 0; j <= i; j++) {
      double s = Ain[i * 6 + j];
      for (int k = 0; k < j; k++) s -= L[i * 6 + k] * L[j * 6 + k];
      if (i == j) {
        if (!(s > 0.0)) return false;
        L[i * 6 + i] = sqrt(s);
      } else L[i * 6 + j] = s / L[j * 6 + j];
    }
  }
  double z[6];
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int i = 0; i < 6; i++) {
    double s = x[i];
    for (int k = 0; k < i; k++) s -= L[i * 6 + k] * z[k];
    z[i] = s / L[i * 6 + i];
  }
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int i = 5; i >= 0; i--) {
    double s = z[i];
    for (int k = i + 1; k < 6; k++) s -= L[k * 6 + i] * y[k];
    y[i] = s / L[i * 6 + i];
  }
  return true;
}

// pose error of the site against a target (world frame): e[0:3] = p_t - p, e[3:6] = rotation
// vector of R_t R^T; returns the two norms
VK_HD void ik_error(const IkSpec &sp, const Pose<double> *P, const double *tpos, const double *tquat, double *e,
                    V3<double> &site_p, double &perr, double &oerr) {
  Pose<double> ident; ident.p = mk<double>(0, 0, 0); ident.q.w = 1; ident.q.x = ident.q.y = ident.q.z = 0;
  const Pose<double> &B = sp.site_slot < 0 ? ident : P[sp.site_slot];
  Q4<double> sq; sq.w = sp.site_quat[0]; sq.x = sp.site_quat[1]; sq.y = sp.site_quat[2]; sq.z = sp.site_quat[3];
  site_p = B.p + qrot(B.q, mk<double>(sp.site_pos[0], sp.site_pos[1], sp.site_pos[2]));
  const Q4<double> cq = qnormalize(qmul(B.q, sq));
  Q4<double> tq; tq.w = tquat[0]; tq.x = tquat[1]; tq.y = tquat[2]; tq.z = tquat[3];
  tq = qnormalize(tq);
  Q4<double> d = qmul(tq, qconj(cq));
  if (d.w < 0) { d.w = -d.w; d.x = -d.x; d.y = -d.y; d.z = -d.z; }
  const double vn = sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
  const double ang = 2.0 * atan2(vn, d.w);
  const double k = vn > 1e-12 ? ang / vn : 2.0;
  e[0] = tpos[0] - site_p.x; e[1] = tpos[1] - site_p.y; e[2] = tpos[2] - site_p.z;
  e[3] = k * d.x; e[4] = k * d.y; e[5] = k * d.z;
  perr = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
  oerr = ang;
}

VK_HD bool ik_row(const FkTables<double> &fk, int nslot, const IkSpec &spec, const double *tpos, const double *tquat,
                  double *q, int *iters_out, double *perr_out, double *oerr_out) {
  const int nq = fk.nq;
  Pose<double> P[MAX_BODY];
  V3<double> anchor[MAX_JNT], axis[MAX_JNT];
  for (int j = 0; j < fk.njnt; j++) {  // start inside the limits
    const int c = fk.jnt_qadr[j];
    if ((spec.jnt_mask >> j) & 1u) q[c] = q[c] < fk.jnt_lo[j] ? fk.jnt_lo[j] : (q[c] > fk.jnt_hi[j] ? fk.jnt_hi[j] : q[c]);
  }
  int it = 0;
  bool ok = false;
  double perr = 0, oerr = 0;
  for (;; it++) {
    fk_with_joints(fk, nslot, q, P, anchor, axis);
    double e[6];
    V3<double> sp;
    ik_error(spec, P, tpos, tquat, e, sp, perr, oerr);
    if (perr <= spec.pos_tol && oerr <= spec.ori_tol) { ok = true; break; }
    if (it >= spec.iterations) break;
    // Active set over the joint limits (the role of mink.ConfigurationLimit in the reference's QP,
    // mink_ik_solver.py:87): a joint that sits on a limit and would be pushed further out is
    // frozen and the step re-solved, so that the other joints take up its share.  The frozen set
    // is a weight per column rather than a bit mask tested while rebuilding J inside the pass
    // loop: that form came back from nvcc 12.9 (sm_100a, -O3) with the first pass's step reused
    // on every later pass (tools/dbg/ik_dbg.py shows device and host iterates side by side).
    double J[6][MAX_JNT], w[MAX_JNT], dq[MAX_JNT], sc = 1.0;
    for (int c = 0; c < nq; c++) {
      w[c] = 0.0;
      for (int i = 0; i < 6; i++) J[i][c] = 0.0;
    }
    for (int j = 0; j < fk.njnt; j++) {
      if (!((spec.jnt_mask >> j) & 1u)) continue;
      const int c = fk.jnt_qadr[j];
      V3<double> jp, jr = mk<double>(0, 0, 0);
      if (fk.jnt_type[j] == JK_SLIDE) jp = axis[j];
      else { jr = axis[j]; jp = cross(axis[j], sp - anchor[j]); }
      J[0][c] = jp.x; J[1][c] = jp.y; J[2][c] = jp.z;
      J[3][c] = jr.x; J[4][c] = jr.y; J[5][c] = jr.z;
      w[c] = 1.0;
    }
    const double mu = spec.lm_damping * (perr * perr + oerr * oerr) + spec.damping;
    bool solved = false;
    for (int pass = 0; pass < 4; pass++) {
      double A[36];
      for (int i = 0; i < 6; i++)
        for (int k = i; k < 6; k++) {
          double acc = 0;
          for (int c = 0; c < nq; c++) acc += w[c] * J[i][c] * J[k][c];
          if (i == k) acc += mu;
          A[i * 6 + k] = A[k * 6 + i] = acc;
        }
      double y[6];
      solved = spd6_solve(A, e, y);
      if (!solved) break;
      double big = 0;
      for (int c = 0; c < nq; c++) {
        double s = 0;
        for (int i = 0; i < 6; i++) s += J[i][c] * y[i];
        dq[c] = w[c] * s;
        big = fmax(big, fabs(dq[c]));
      }
      sc = big > spec.max_step ? spec.max_step / big : 1.0;
      int nblocked = 0, nfree = 0;
      for (int j = 0; j < fk.njnt; j++) {
        const int c = fk.jnt_qadr[j];
        if (w[c] == 0.0) continue;
        if ((q[c] <= fk.jnt_lo[j] && dq[c] < 0.0) || (q[c] >= fk.jnt_hi[j] && dq[c] > 0.0)) { w[c] = 0.0; dq[c] = 0.0; nblocked++; }
        else nfree++;
      }
      if (!nblocked) break;
      if (!nfree) { solved = false; break; }
    }
    if (!solved) break;
    for (int j = 0; j < fk.njnt; j++) {
      const int c = fk.jnt_qadr[j];
      if (w[c] == 0.0) continue;
      double v = q[c] + sc * dq[c];
      v = v < fk.jnt_lo[j] ? fk.jnt_lo[j] : (v > fk.jnt_hi[j] ? fk.jnt_hi[j] : v);
      q[c] = v;
    }
  }
  if (iters_out) *iters_out = it;
  if (perr_out) *perr_out = perr;
  if (oerr_out) *oerr_out = oerr;
  return ok;
}

// ------------------------------------------------------------------------------ signed distance (band accounting)
// mjb_min_distance: the signed distance of a pair, not just its three-way verdict.  fp64 only (it is
// an accounting pass over rows near the contact band, not the hot path).
//   cores apart          : GJK run to convergence, distance = |v| - R
//   cores intersecting   : expanding-polytope depth of the origin inside A - B, distance = -depth - R,
//                          stopped as soon as the depth is certified >= depth_cap
// Supports are plain scans of the vertex tables (support_shape).
constexpr int EPA_MAXV = 40, EPA_MAXF = 96;

struct EpaFace { double n[3], d; uint8_t v[3], alive; };

VK_HD bool epa_make_face(const V3<double> *P, V3<double> inner, int a, int b, int c, EpaFace &f) {
  V3<double> n = cross(P[b] - P[a], P[c] - P[a]);
  const double len = vk_sqrt(dot(n, n));
  if (!(len > 1e-300)) return false;
  n = n * (1.0 / len);
  f.v[0] = (uint8_t)a; f.v[1] = (uint8_t)b; f.v[2] = (uint8_t)c;
  if (dot(n, P[a] - inner) < 0) { n = -n; f.v[1] = (uint8_t)c; f.v[2] = (uint8_t)b; }   // outward
  f.n[0] = n.x; f.n[1] = n.y; f.n[2] = n.z;
  f.d = dot(n, P[a]);
  f.alive = 1;
  return true;
}

// support point of A - B along d (A frame)
template <typename SupA, typename SupB>
VK_HD V3<double> mink_support(const Rel<double> &rel, V3<double> d, SupA supA, SupB supB) {
  return supA(d) - (mul(rel.R, supB(mulT(rel.R, -d))) + rel.t);
}

// depth of the origin inside A - B, given the (possibly degenerate) simplex GJK ended with
template <typename SupA, typename SupB>
VK_HD double epa_depth(const GjkState<double> &gs, const Rel<double> &rel, double depth_cap, SupA supA, SupB supB) {
  V3<double> P[EPA_MAXV];
  int np = gs.n < 1 ? 1 : gs.n;
  P[0] = gs.p0; P[1] = gs.p1; P[2] = gs.p2; P[3] = gs.p3;
  const double tiny = 1e-10;
  if (np == 1) {   // grow a point into a segment
    const V3<double> AX[6] = {mk<double>(1, 0, 0), mk<double>(-1, 0, 0), mk<double>(0, 1, 0), mk<double>(0, -1, 0), mk<double>(0, 0, 1), mk<double>(0, 0, -1)};
    for (int k = 0; k < 6 && np == 1; k++) {
      const V3<double> w = mink_support(rel, AX[k], supA, supB);
      const V3<double> dd = w - P[0];
      if (vk_sqrt(dot(dd, dd)) > tiny) P[np++] = w;
    }
    if (np == 1) return 0.0;
  }
  if (np == 2) {   // ... a segment into a triangle: search directions around the edge
    const V3<double> e = P[1] - P[0];
    const double el = vk_sqrt(dot(e, e));
    const V3<double> u = e * (1.0 / el);
    const V3<double> ax = vk_abs(u.x) <= vk_abs(u.y) && vk_abs(u.x) <= vk_abs(u.z) ? mk<double>(1, 0, 0)
                          : (vk_abs(u.y) <= vk_abs(u.z) ? mk<double>(0, 1, 0) : mk<double>(0, 0, 1));
    V3<double> d1 = cross(u, ax);
    d1 = d1 * (1.0 / vk_sqrt(dot(d1, d1)));
    const V3<double> d2 = cross(u, d1);
    for (int k = 0; k < 6 && np == 2; k++) {
      const double ang = k * 1.0471975511965976;
      const V3<double> dir = d1 * cos(ang) + d2 * sin(ang);
      const V3<double> w = mink_support(rel, dir, supA, supB);
      const V3<double> cr = cross(e, w - P[0]);
      if (vk_sqrt(dot(cr, cr)) > tiny * el) P[np++] = w;
    }
    if (np == 2) return 0.0;
  }
  if (np == 3) {   // ... a triangle into a tetrahedron
    V3<double> n = cross(P[1] - P[0], P[2] - P[0]);
    const double len = vk_sqrt(dot(n, n));
    if (!(len > 1e-300)) return 0.0;
    n = n * (1.0 / len);
    const V3<double> w1 = mink_support(rel, n, supA, supB), w2 = mink_support(rel, -n, supA, supB);
    const double h1 = vk_abs(dot(w1 - P[0], n)), h2 = vk_abs(dot(w2 - P[0], n));
    if (h1 < tiny && h2 < tiny) return 0.0;
    P[np++] = h1 >= h2 ? w1 : w2;
  }
  const V3<double> inner = (P[0] + P[1] + P[2] + P[3]) * 0.25;
  EpaFace F[EPA_MAXF];
  int nf = 0;
  const int T[4][3] = {{0, 1, 2}, {0, 1, 3}, {0, 2, 3}, {1, 2, 3}};
  for (int f = 0; f < 4; f++)
    if (epa_make_face(P, inner, T[f][0], T[f][1], T[f][2], F[nf])) nf++;
  if (nf < 4) return 0.0;
  double lower = 0;
  for (int it = 0; it < 64; it++) {
    int bf = -1;
    double bd = 1e300;
    for (int f = 0; f < nf; f++)
      if (F[f].alive && F[f].d < bd) { bd = F[f].d; bf = f; }
    if (bf < 0) break;
    lower = bd > 0 ? bd : 0;
    if (lower >= depth_cap) return lower;
    const V3<double> fn = mk<double>(F[bf].n[0], F[bf].n[1], F[bf].n[2]);
    const V3<double> w = mink_support(rel, fn, supA, supB);
    const double h = dot(w, fn);
    if (h - bd < 1e-10 || np >= EPA_MAXV) return h > 0 ? h : 0;
    P[np] = w;
    // remove the faces w can see, collect the horizon edges
    uint8_t E[EPA_MAXF][2];
    int ne = 0;
    for (int f = 0; f < nf; f++) {
      if (!F[f].alive) continue;
      const V3<double> nn = mk<double>(F[f].n[0], F[f].n[1], F[f].n[2]);
      if (dot(nn, w - P[F[f].v[0]]) > 1e-14) {
        F[f].alive = 0;
        for (int e = 0; e < 3; e++) {
          const uint8_t a = F[f].v[e], b = F[f].v[(e + 1) % 3];
          int found = -1;
          for (int k = 0; k < ne; k++)
            if (E[k][0] == b && E[k][1] == a) { found = k; break; }
          if (found >= 0) { E[found][0] = E[ne - 1][0]; E[found][1] = E[ne - 1][1]; ne--; }
          else if (ne < EPA_MAXF) { E[ne][0] = a; E[ne][1] = b; ne++; }
        }
      }
    }
    if (ne == 0) return h > 0 ? h : 0;
    for (int k = 0; k < ne; k++) {
      int slot = -1;
      for (int f = 0; f < nf; f++)
        if (!F[f].alive) { slot = f; break; }
      if (slot < 0) { if (nf >= EPA_MAXF) return lower; slot = nf++; }
      if (!epa_make_face(P, inner, E[k][0], E[k][1], np, F[slot])) F[slot].alive = 0;
    }
    np++;
  }
  return lower;
}

// signed distance of the cores of A and B (swept radii and margin NOT subtracted): > 0 apart, < 0 = -depth
template <typename SupA, typename SupB>
VK_HD double gjk_signed_distance(const Shape<double> &A, const Shape<double> &B, const Rel<double> &rel, double depth_cap,
                                 SupA supA, SupB supB) {
  GjkState<double> s;
  gjk_init(s, A, B, rel);
  for (int it = 0; it < 96; it++) {
    const V3<double> v = s.v;
    const V3<double> w = mink_support(rel, -v, supA, supB);
    const double vv = dot(v, v), vw = dot(v, w);
    if (s.n > 0 && (vv - vw) <= 1e-13 * vv) return vk_sqrt(vv);   // converged: |v| is the distance
    double l0 = 0, l1 = 0, l2 = 0, l3 = 0;
    bool inside = false;
    if (s.n == 0) { s.p0 = w; l0 = 1; }
    else if (s.n == 1) { s.p1 = w; solve1(s.p0, s.p1, l0, l1); }
    else if (s.n == 2) { s.p2 = w; solve2(s.p0, s.p1, s.p2, l0, l1, l2); }
    else { s.p3 = w; inside = solve3(s.p0, s.p1, s.p2, s.p3, l0, l1, l2, l3); }
    if (inside) { s.n = 4; return -epa_depth(s, rel, depth_cap, supA, supB); }
    const V3<double> nv = s.p0 * l0 + s.p1 * l1 + s.p2 * l2 + s.p3 * l3;
    s.v = nv;
    bool k0 = l0 > 0, k1 = l1 > 0, k2 = l2 > 0, k3 = l3 > 0;
    { bool c = !k0 && k1; cswap(c, s.p0, s.p1); cswap(c, k0, k1); }
    { bool c = !k1 && k2; cswap(c, s.p1, s.p2); cswap(c, k1, k2); }
    { bool c = !k2 && k3; cswap(c, s.p2, s.p3); cswap(c, k2, k3); }
    { bool c = !k0 && k1; cswap(c, s.p0, s.p1); cswap(c, k0, k1); }
    { bool c = !k1 && k2; cswap(c, s.p1, s.p2); cswap(c, k1, k2); }
    { bool c = !k0 && k1; cswap(c, s.p0, s.p1); cswap(c, k0, k1); }
    s.n = int(k0) + int(k1) + int(k2) + int(k3);
    if (!(dot(nv, nv) > 1e-26)) return -epa_depth(s, rel, depth_cap, supA, supB);   // cores touch: depth from the simplex
    if (s.n == 4) return -epa_depth(s, rel, depth_cap, supA, supB);
  }
  return vk_sqrt(dot(s.v, s.v));
}

// signed distance of one pair minus the margin (rsum = swept radii + margin), clamped below at -depth_cap
VK_HD double pair_signed_distance(const Pair &pr, const Shape<double> &A, const Shape<double> &B, const Vtx<double> *verts,
                                  const Pose<double> &PA, const Pose<double> &PB, double rsum, double depth_cap) {
  double d;
  if (pr.kind == PK_PLANE) {
    const V3<double> n = mk<double>(A.ax[0], A.ax[1], A.ax[2]), c = mk<double>(A.c[0], A.c[1], A.c[2]);
    if (B.kind == SK_CYL) {
      const V3<double> ax = qrot(PB.q, mk<double>(B.ax[0], B.ax[1], B.ax[2]));
      const V3<double> cb = PB.p + qrot(PB.q, mk<double>(B.c[0], B.c[1], B.c[2]));
      const double na = dot(n, ax), rad = 1.0 - na * na;
      d = dot(n, cb - c) - vk_abs(na) * B.halflen - B.radius * vk_sqrt(rad > 0 ? rad : 0.0);
    } else {
      const V3<double> sp = support_verts(verts + B.vadr, B.nvert, -qrot(qconj(PB.q), n));
      d = dot(n, PB.p + qrot(PB.q, sp) - c);
    }
    d -= rsum;
  } else if (pr.kind == PK_SEGSEG) {
    const Vtx<double> a0 = verts[A.vadr], a1 = verts[A.vadr + A.nvert - 1], b0 = verts[B.vadr], b1 = verts[B.vadr + B.nvert - 1];
    const V3<double> p1 = PA.p + qrot(PA.q, mk<double>(a0.x, a0.y, a0.z)), q1 = PA.p + qrot(PA.q, mk<double>(a1.x, a1.y, a1.z));
    const V3<double> p2 = PB.p + qrot(PB.q, mk<double>(b0.x, b0.y, b0.z)), q2 = PB.p + qrot(PB.q, mk<double>(b1.x, b1.y, b1.z));
    d = vk_sqrt(segseg_dist2(p1, q1, p2, q2)) - rsum;
  } else {
    if (rsum >= depth_cap) {   // sphere-swept pair: a core distance of 0 is already deeper than the cap
      // fall through to the general routine, the clamp below takes care of it
    }
    const Rel<double> rel = relative_pose(PA, PB);
    d = gjk_signed_distance(A, B, rel, depth_cap + rsum, [&](V3<double> dir) { return support_shape(A, verts, dir); },
                            [&](V3<double> dir) { return support_shape(B, verts, dir); }) - rsum;
  }
  return d < -depth_cap ? -depth_cap : d;
}

// min over the static pair list of (signed distance - margin) for one row, fp64.  Pairs whose bounding
// capsules are further apart than the best distance so far cannot improve it and are skipped; the
// result is capped at far_cap from above (pair = -1: nothing closer than that) and at -depth_cap from
// below.  Reference semantics of the quantity: SURVEY A.3 (contact iff signed distance <= margin).
VK_HD void row_min_distance(const FkTables<double> &fk, int nslot, const Shape<double> *shapes, const Vtx<double> *verts,
                            const Pair *pairs, const double *pair_rsum, int npair, const double *q, double far_cap,
                            double depth_cap, double &best, int &bestp) {
  Pose<double> P[MAX_BODY], ident;
  ident.p = mk<double>(0, 0, 0); ident.q.w = 1; ident.q.x = ident.q.y = ident.q.z = 0;
  for (int s = 0; s < nslot; s++) {
    const int ps = fk.body_parent[s];
    P[s] = fk_body(fk, s, ps < 0 ? ident : P[ps], q);
  }
  best = far_cap; bestp = -1;
  for (int p = 0; p < npair; p++) {
    const Pair pr = pairs[p];
    const Shape<double> &A = shapes[pr.sa], &B = shapes[pr.sb];
    const Pose<double> &PA = A.slot < 0 ? ident : P[A.slot];
    const Pose<double> &PB = B.slot < 0 ? ident : P[B.slot];
    const double rsum = pair_rsum[p];
    const double margin = rsum - swept_radius(A) - swept_radius(B);
    if (capsule_cull(pr, A, B, PA, PB, margin + best)) continue;   // lower bound of the signed distance > best
    const double d = pair_signed_distance(pr, A, B, verts, PA, PB, rsum, depth_cap);
    if (d < best) { best = d; bestp = p; }
  }
}

// ------------------------------------------------------------------------------ two-kernel pipeline: item bins
// (shared by the device kernels in vk_split.cuh and the host-side capacity estimate in vk_build.h)
constexpr int NBIN = 8;
// does the item need the hull-scanning narrow phase (narrow_kernel), or is it closed form?
template <typename T> VK_HD bool item_needs_scan(const Pair &pr, const Shape<T> &B) {
  return pr.kind == PK_GJK || (pr.kind == PK_PLANE && B.kind == SK_VERTS && B.nvert > 8);
}
// Bin of an item, so that the lanes of a warp scan hulls of similar size: plane-hull items and
// tiny pairs first, then small-core-vs-hull pairs by hull size, then hull-vs-hull by total size.
template <typename T> VK_HD int item_bin(const Pair &pr, const Shape<T> &A, const Shape<T> &B) {
  if (pr.kind != PK_GJK) return 0;
  const int lo = A.nvert < B.nvert ? A.nvert : B.nvert, hi = A.nvert < B.nvert ? B.nvert : A.nvert;
  if (hi <= 8) return 0;
  if (lo <= 8) return hi <= 48 ? 1 : (hi <= 64 ? 2 : (hi <= 102 ? 3 : 4));
  const int nv = lo + hi;
  return nv <= 130 ? 5 : (nv <= 210 ? 6 : 7);
}

// ------------------------------------------------------------------------------ counter-based row generator
// splitmix64 finaliser over (seed, row, joint) -> 24-bit uniform in [0,1)
VK_HD uint32_t sweep_bits(uint64_t seed, uint64_t row, uint32_t j) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (row * 64ull + j + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return uint32_t(z >> 40);
}
VK_HD float sweep_value(uint64_t seed, uint64_t row, uint32_t j, float lo, float hi) {
  float u = float(sweep_bits(seed, row, j)) * (1.0f / 16777216.0f);
  // plain mul + add, each rounded (no fma) so a numpy float32 host mirror is bit-identical
#if defined(__CUDA_ARCH__)
  return __fadd_rn(lo, __fmul_rn(u, __fsub_rn(hi, lo)));
#else
  volatile float d = hi - lo;
  volatile float m = u * d;
  return lo + m;
#endif
}

}  // namespace vk
