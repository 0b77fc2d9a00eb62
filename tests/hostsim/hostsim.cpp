// hostsim.cpp -- TEST TOOL: runs the templated core of the CUDA kernels (vk_core.cuh,
// vk_build.h) row by row on the CPU so the geometry can be compared with the fp64 oracle in the
// `-m "not gpu"` suite, where no device exists.  It is NOT a product path: nothing in
// mjpl_b200/ loads it, it is built only by tests/hostsim/__init__.py, and the shipped library
// has no CPU fallback.
//
// Mirrors validity_kernel's per-row logic (P1 limits+FK, P2 sphere + OBB culls, P3 narrow
// phase with certified verdicts, fp64 re-evaluation of uncertain rows).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../mjpl_b200/csrc/vk_build.h"

using namespace vk;

struct Sim {
  vkb::HostModel H;
  FkTables<float> fk32;
  std::vector<Shape<float>> s32;
  std::vector<Vtx<float>> v32;
};

extern "C" {

int hs_create(const mjb_model_desc *d, Sim **out, char *err, int errlen) {
  Sim *s = new Sim();
  if (!vkb::build_host_model(d, s->H)) {
    snprintf(err, errlen, "%s", s->H.err.c_str());
    delete s;
    return 2;
  }
  s->fk32 = vkb::convert_fk<float>(s->H.fk);
  for (auto &sh : s->H.shapes) s->s32.push_back(vkb::convert_shape<float>(sh));
  for (auto &v : s->H.verts) { Vtx<float> f; f.x = (float)v.x; f.y = (float)v.y; f.z = (float)v.z; f.w = 0; s->v32.push_back(f); }
  *out = s;
  return 0;
}
void hs_destroy(Sim *s) { delete s; }
int hs_npair(Sim *s) { return (int)s->H.pairs.size(); }
void hs_pairs(Sim *s, int32_t *g1, int32_t *g2) {
  for (size_t i = 0; i < s->H.pairs.size(); i++) { g1[i] = s->H.pair_g1[i]; g2[i] = s->H.pair_g2[i]; }
}

// xpos (n,nbody,3), xquat (n,nbody,4) in fp32 arithmetic
void hs_fk(Sim *s, const float *q, int64_t n, float *xpos, float *xquat) {
  const auto &H = s->H;
  for (int64_t r = 0; r < n; r++) {
    Pose<float> P[MAX_BODY], ident;
    ident.p = mk<float>(0, 0, 0); ident.q.w = 1; ident.q.x = ident.q.y = ident.q.z = 0;
    for (int k = 0; k < H.nslot; k++) {
      int ps = s->fk32.body_parent[k];
      P[k] = fk_body(s->fk32, k, ps < 0 ? ident : P[ps], q + r * H.nq);
    }
    for (int b = 0; b < H.nbody; b++) {
      float *xp = xpos + (r * H.nbody + b) * 3, *xq = xquat + (r * H.nbody + b) * 4;
      int k = H.body_slot[b];
      if (k >= 0) {
        xp[0] = P[k].p.x; xp[1] = P[k].p.y; xp[2] = P[k].p.z;
        xq[0] = P[k].q.w; xq[1] = P[k].q.x; xq[2] = P[k].q.y; xq[3] = P[k].q.z;
      } else {
        const auto &S = H.static_pose[b];
        xp[0] = (float)S.p.x; xp[1] = (float)S.p.y; xp[2] = (float)S.p.z;
        xq[0] = (float)S.q.w; xq[1] = (float)S.q.x; xq[2] = (float)S.q.y; xq[3] = (float)S.q.z;
      }
    }
  }
}

}  // extern "C"

template <typename T>
static int row_verdict(const vkb::HostModel &H, const FkTables<T> &fk, const Shape<T> *shapes, const Vtx<T> *verts,
                       const float *q, bool use_obb, int64_t *stats) {
  // returns 0 = no contact, 1 = certain contact, 2 = uncertain (no certain contact)
  Pose<T> P[MAX_BODY], ident;
  ident.p = mk<T>(0, 0, 0); ident.q.w = 1; ident.q.x = ident.q.y = ident.q.z = 0;
  for (int k = 0; k < H.nslot; k++) {
    int ps = fk.body_parent[k];
    P[k] = fk_body(fk, k, ps < 0 ? ident : P[ps], q);
  }
  const T slack = sizeof(T) == 4 ? T(1e-4) : T(1e-6);
  bool unc = false;
  for (size_t p = 0; p < H.pairs.size(); p++) {
    const Pair pr = H.pairs[p];
    const Shape<T> &A = shapes[pr.sa], &B = shapes[pr.sb];
    const Pose<T> &PA = A.slot < 0 ? ident : P[A.slot];
    const Pose<T> &PB = B.slot < 0 ? ident : P[B.slot];
    const T bsum = sizeof(T) == 4 ? (T)pr.bsum : (T)H.pair_bsum64[p];
    const T rsum = sizeof(T) == 4 ? (T)pr.rsum : (T)H.pair_rsum64[p];
    V3<T> cB = PB.p + qrot(PB.q, mk<T>(B.bc[0], B.bc[1], B.bc[2]));
    if (pr.kind == PK_PLANE) {
      T d = A.ax[0] * (cB.x - A.c[0]) + A.ax[1] * (cB.y - A.c[1]) + A.ax[2] * (cB.z - A.c[2]);
      if (d > bsum + slack) continue;
    } else {
      V3<T> cA = PA.p + qrot(PA.q, mk<T>(A.bc[0], A.bc[1], A.bc[2]));
      V3<T> dd = cA - cB;
      if (dot(dd, dd) > (bsum + slack) * (bsum + slack)) continue;
    }
    if (stats) stats[4]++;
    if (use_obb && midphase_cull(pr, A, B, PA, PB, rsum - swept_radius(A) - swept_radius(B), slack)) continue;
    if (stats) stats[0]++;
    int v;
    if (pr.kind == PK_GJK) {
      int iters = 0;
      Rel<T> rel = relative_pose(PA, PB);
      v = gjk_classify(A, B, verts, rel, rsum, &iters);
      if (stats) stats[47] += (long long)iters * (A.nvert + B.nvert);
      if (stats) { stats[1] += iters; stats[5]++; if (iters > stats[6]) stats[6] = iters; stats[8 + (iters < 31 ? iters : 31)]++; stats[40 + v]++; stats[44 + v] += iters; }
    } else {
      v = narrow_item<T>(pr.kind, A, B, verts, PA, PB, rsum);
    }
    if (v == V_PEN) return 1;
    if (v == V_UNC) { unc = true; if (stats) stats[7]++; }
  }
  return unc ? 2 : 0;
}

extern "C" {

// valid[i]: 1 valid, 0 invalid.  stats: [0] narrow items, [1] gjk iterations, [2] uncertain rows,
// [3] rows, [4] sphere survivors, [5] gjk calls, [6] max gjk iterations, [7] uncertain items
void hs_check(Sim *s, const float *q, int64_t n, uint32_t flags, int use_obb, int use_recheck, uint8_t *valid,
              int64_t *stats) {
  const auto &H = s->H;
  for (int64_t r = 0; r < n; r++) {
    const float *qr = q + r * H.nq;
    bool ok = true;
    if (flags & 1u)
      for (int j = 0; j < H.njnt; j++) ok = ok && ((double)qr[j] >= H.jnt_lo[j]) && ((double)qr[j] <= H.jnt_hi[j]);
    int out = ok ? 1 : 0;
    if (ok && (flags & 2u)) {
      int v = row_verdict<float>(H, s->fk32, s->s32.data(), s->v32.data(), qr, use_obb != 0, stats);
      if (v == 2) {
        if (stats) stats[2]++;
        if (use_recheck) v = row_verdict<double>(H, H.fk, H.shapes.data(), H.verts.data(), qr, false, nullptr) != 0 ? 1 : 0;
      }
      out = v == 0 ? 1 : (v == 1 ? 0 : 2);
    }
    valid[r] = (uint8_t)out;
    if (stats) stats[3]++;
  }
}


// verdict of one pair (by MuJoCo geom ids) at one row, in fp32 (prec=0) or fp64 (prec=1); -1 if no such pair
int hs_pair_verdict(Sim *s, const float *q, int g1, int g2, int prec, int *iters) {
  const auto &H = s->H;
  for (size_t p = 0; p < H.pairs.size(); p++) {
    if (!((H.pair_g1[p] == g1 && H.pair_g2[p] == g2) || (H.pair_g1[p] == g2 && H.pair_g2[p] == g1))) continue;
    const Pair pr = H.pairs[p];
    if (prec == 0) {
      Pose<float> P[MAX_BODY], ident;
      ident.p = mk<float>(0, 0, 0); ident.q.w = 1; ident.q.x = ident.q.y = ident.q.z = 0;
      for (int k = 0; k < H.nslot; k++) { int ps = s->fk32.body_parent[k]; P[k] = fk_body(s->fk32, k, ps < 0 ? ident : P[ps], q); }
      const Shape<float> &A = s->s32[pr.sa], &B = s->s32[pr.sb];
      const Pose<float> &PA = A.slot < 0 ? ident : P[A.slot];
      const Pose<float> &PB = B.slot < 0 ? ident : P[B.slot];
      if (pr.kind == PK_GJK) { Rel<float> rel = relative_pose(PA, PB); return gjk_classify(A, B, s->v32.data(), rel, pr.rsum, iters); }
      return narrow_item<float>(pr.kind, A, B, s->v32.data(), PA, PB, pr.rsum);
    } else {
      Pose<double> P[MAX_BODY], ident;
      ident.p = mk<double>(0, 0, 0); ident.q.w = 1; ident.q.x = ident.q.y = ident.q.z = 0;
      for (int k = 0; k < H.nslot; k++) { int ps = H.fk.body_parent[k]; P[k] = fk_body(H.fk, k, ps < 0 ? ident : P[ps], q); }
      const Shape<double> &A = H.shapes[pr.sa], &B = H.shapes[pr.sb];
      const Pose<double> &PA = A.slot < 0 ? ident : P[A.slot];
      const Pose<double> &PB = B.slot < 0 ? ident : P[B.slot];
      if (pr.kind == PK_GJK) { Rel<double> rel = relative_pose(PA, PB); return gjk_classify(A, B, H.verts.data(), rel, H.pair_rsum64[p], iters); }
      return narrow_item<double>(pr.kind, A, B, H.verts.data(), PA, PB, H.pair_rsum64[p]);
    }
  }
  return -1;
}


// hill-climbing support vs scanning all vertices, for every shape with a hull graph and `ndir`
// random directions (cold start and warm start from the previous answer): returns the number of
// shapes with a graph; max_gap = largest shortfall of the hill-climbed support value.
int hs_support_check(Sim *s, int ndir, uint64_t seed, double *max_gap, int64_t *evals_hill, int64_t *evals_scan) {
  const auto &H = s->H;
  int ngraph = 0;
  *max_gap = 0; *evals_hill = 0; *evals_scan = 0;
  for (size_t k = 0; k < H.shapes.size(); k++) {
    const Shape<float> &sh = s->s32[k];
    if (!sh.graph) continue;
    ngraph++;
    const Vtx<float> *v = s->v32.data() + sh.vadr;
    const uint16_t *as = H.adj_start.data() + sh.vadr;
    int warm = -1;
    for (int i = 0; i < ndir; i++) {
      V3<float> d = mk<float>(sweep_value(seed, i, 0, -1.f, 1.f), sweep_value(seed, i, 1, -1.f, 1.f), sweep_value(seed, i, 2, -1.f, 1.f));
      if (i % 3 == 2) { d.x *= 1e-3f; }              // near-axis directions
      V3<float> brute = support_verts(v, sh.nvert, d);
      int start = (i % 2 == 0 || warm < 0) ? hill_start(sh, d) : warm;
      // count evaluations along the climb
      int cur = start;
      for (;;) {
        int nxt = cur; float cb = v[cur].x * d.x + v[cur].y * d.y + v[cur].z * d.z;
        for (int e = as[cur]; e < as[cur + 1]; e++) { int j = H.adj[e]; float t = v[j].x * d.x + v[j].y * d.y + v[j].z * d.z; (*evals_hill)++; if (t > cb) { cb = t; nxt = j; } }
        if (nxt == cur) break;
        cur = nxt;
      }
      int got = support_hill(v, as, H.adj.data(), d, start);
      if (got != cur) return -1;
      warm = got;
      *evals_scan += sh.nvert;
      double gap = (double)dot(brute, d) - (double)(v[got].x * d.x + v[got].y * d.y + v[got].z * d.z);
      if (gap > *max_gap) *max_gap = gap;
    }
  }
  return ngraph;
}


// PoseConstraint rows on the CPU through the same core (pose_valid_row / pose_project_row)
int hs_pose(Sim *s, const mjb_pose_spec *spec, const double *q_old, const double *q, int64_t n, int project, int max_iters,
            double *q_out, uint8_t *ok, int32_t *iters, char *err, int errlen) {
  PoseSpec sp;
  std::string e;
  if (!vkb::make_pose_spec(s->H, spec, sp, e)) { snprintf(err, errlen, "%s", e.c_str()); return 1; }
  const int nq = s->H.nq;
  for (int64_t r = 0; r < n; r++) {
    double row[MAX_JNT];
    for (int j = 0; j < nq; j++) row[j] = q[r * nq + j];
    if (!project) { ok[r] = pose_valid_row(s->H.fk, s->H.nslot, sp, row); continue; }
    int it = 0;
    bool good = pose_project_row(s->H.fk, s->H.nslot, sp, q_old + r * nq, row, max_iters, &it);
    ok[r] = good;
    if (iters) iters[r] = it;
    for (int j = 0; j < nq; j++) q_out[r * nq + j] = good ? row[j] : q[r * nq + j];
  }
  return 0;
}

// IK rows on the CPU through the same core (ik_row)
int hs_ik(Sim *s, const mjb_ik_spec *spec, const double *tpos, const double *tquat, const double *q_init, int64_t n,
          double *q_out, uint8_t *ok, int32_t *iters, double *errs, char *err, int errlen) {
  IkSpec sp;
  std::string e;
  if (!vkb::make_ik_spec(s->H, spec, sp, e)) { snprintf(err, errlen, "%s", e.c_str()); return 1; }
  const int nq = s->H.nq;
  for (int64_t r = 0; r < n; r++) {
    double row[MAX_JNT];
    for (int j = 0; j < nq; j++) row[j] = q_init[r * nq + j];
    int it = 0;
    double pe = 0, oe = 0;
    ok[r] = ik_row(s->H.fk, s->H.nslot, sp, tpos + r * 3, tquat + r * 4, row, &it, &pe, &oe);
    if (iters) iters[r] = it;
    if (errs) { errs[r * 2] = pe; errs[r * 2 + 1] = oe; }
    for (int j = 0; j < nq; j++) q_out[r * nq + j] = row[j];
  }
  return 0;
}

// per-pair work census (analysis aid): out[p*4 + {0,1,2,3}] = sphere survivors, narrow items,
// vertex evaluations, contacts -- every pair evaluated on every row (no early exit)
void hs_pair_census(Sim *s, const float *q, int64_t n, int64_t *out) {
  const auto &H = s->H;
  Pose<float> P[MAX_BODY], ident;
  ident.p = mk<float>(0, 0, 0); ident.q.w = 1; ident.q.x = ident.q.y = ident.q.z = 0;
  for (int64_t r = 0; r < n; r++) {
    const float *qr = q + r * H.nq;
    for (int k = 0; k < H.nslot; k++) { int ps = s->fk32.body_parent[k]; P[k] = fk_body(s->fk32, k, ps < 0 ? ident : P[ps], qr); }
    for (size_t p = 0; p < H.pairs.size(); p++) {
      const Pair pr = H.pairs[p];
      const Shape<float> &A = s->s32[pr.sa], &B = s->s32[pr.sb];
      const Pose<float> &PA = A.slot < 0 ? ident : P[A.slot];
      const Pose<float> &PB = B.slot < 0 ? ident : P[B.slot];
      const float slack = 1e-4f;
      V3<float> cB = PB.p + qrot(PB.q, mk<float>(B.bc[0], B.bc[1], B.bc[2]));
      if (pr.kind == PK_PLANE) {
        float d = A.ax[0] * (cB.x - A.c[0]) + A.ax[1] * (cB.y - A.c[1]) + A.ax[2] * (cB.z - A.c[2]);
        if (d > pr.bsum + slack) continue;
      } else {
        V3<float> cA = PA.p + qrot(PA.q, mk<float>(A.bc[0], A.bc[1], A.bc[2]));
        V3<float> dd = cA - cB;
        if (dot(dd, dd) > (pr.bsum + slack) * (pr.bsum + slack)) continue;
      }
      out[p * 4 + 0]++;
      if (midphase_cull(pr, A, B, PA, PB, pr.rsum - swept_radius(A) - swept_radius(B), slack)) continue;
      out[p * 4 + 1]++;
      int v, iters = 0;
      if (pr.kind == PK_GJK) {
        Rel<float> rel = relative_pose(PA, PB);
        v = gjk_classify(A, B, s->v32.data(), rel, pr.rsum, &iters);
        out[p * 4 + 2] += (int64_t)iters * (A.nvert + B.nvert);
      } else v = narrow_item<float>(pr.kind, A, B, s->v32.data(), PA, PB, pr.rsum);
      if (v == V_PEN) out[p * 4 + 3]++;
    }
  }
}

// calibrated narrow-phase items per row for each bin of the multi-kernel pipeline, the total over all
// pairs, and every pair's bin
void hs_bins(Sim *s, double *bin_expect, double *items_per_row, int32_t *pair_bin) {
  const auto &H = s->H;
  for (int b = 0; b < NBIN; b++) bin_expect[b] = H.bin_expect[b];
  *items_per_row = H.calib_items_per_row;
  for (size_t p = 0; p < H.pairs.size(); p++) {
    const Pair &pr = H.pairs[p];
    const Shape<double> &A = H.shapes[pr.sa], &B = H.shapes[pr.sb];
    (void)B;
    pair_bin[p] = item_bin(pr, A, B);   // closed-form kinds share bin 0 with the smallest scans
  }
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// The pipeline's culling order (vk_pipe.cuh): level-0 group test -> expansion into shape pairs ->
// bounding-capsule cull -> OBB cull -> narrow phase.  fp32, no early exit; fp64 re-evaluation of
// uncertain rows as in hs_check.  stats: [0] level-0 survivors, [1] expanded shape pairs,
// [2] capsule survivors, [3] OBB survivors (narrow items), [4] contacts
extern "C" void hs_check_pipe(Sim *s, const float *q, int64_t n, uint32_t flags, uint8_t *valid, int64_t *stats) {
  const auto &H = s->H;
  Pose<float> P[MAX_BODY], ident;
  ident.p = mk<float>(0, 0, 0); ident.q.w = 1; ident.q.x = ident.q.y = ident.q.z = 0;
  const float slack = 1e-4f;
  for (int64_t r = 0; r < n; r++) {
    const float *qr = q + r * H.nq;
    bool ok = true;
    if (flags & 1u)
      for (int j = 0; j < H.njnt; j++) ok = ok && ((double)qr[j] >= H.jnt_lo[j]) && ((double)qr[j] <= H.jnt_hi[j]);
    int out = ok ? 1 : 0;
    if (ok && (flags & 2u)) {
      V3<float> cen[MAX_GROUP];
      for (int k = 0; k < H.nslot; k++) {
        int ps = s->fk32.body_parent[k];
        P[k] = fk_body(s->fk32, k, ps < 0 ? ident : P[ps], qr);
        for (int g = H.slot_group_adr[k]; g < H.slot_group_adr[k] + H.slot_group_num[k]; g++)
          cen[g] = P[k].p + qrot(P[k].q, mk<float>((float)H.group_c[g][0], (float)H.group_c[g][1], (float)H.group_c[g][2]));
      }
      bool pen = false, unc = false;
      bool cert0 = false, cert1 = false;   // certain contact from the inner balls (level 0) / inner capsules (shape pairs)
      int64_t row_l0 = 0, row_items = 0, row_exp = 0;
      for (const GroupPair &g : H.group_pairs) {
        const int t0 = group_pair_test(g, cen[g.ga], g.kind == GK_SPHERE ? cen[g.gb] : cen[g.ga],
                                       g.kind == GK_SPHERE ? nullptr : &H.static_groups[g.gb]);
        if (t0 == 0) continue;
        if (t0 == 2) cert0 = true;
        stats[0]++; row_l0++; row_exp += g.n;
        for (int i = 0; i < g.n; i++) {
          const int p = H.gp_member[g.first + i];
          const Pair pr = H.pairs[p];
          const Shape<float> &A = s->s32[pr.sa], &B = s->s32[pr.sb];
          const Pose<float> &PA = A.slot < 0 ? ident : P[A.slot];
          const Pose<float> &PB = B.slot < 0 ? ident : P[B.slot];
          stats[1]++;
          const float margin = pr.rsum - swept_radius(A) - swept_radius(B);
          if (capsule_cull(pr, A, B, PA, PB, margin + slack)) continue;
          stats[2]++;
          if (pr.kind != PK_SEGSEG && inner_contact(pr, A, B, PA, PB)) cert1 = true;
          if (midphase_cull(pr, A, B, PA, PB, margin, slack)) continue;
          stats[3]++; row_items++;
          int v;
          if (pr.kind == PK_GJK) {
            int iters = 0;
            v = gjk_classify(A, B, s->v32.data(), relative_pose(PA, PB), pr.rsum, &iters);
            stats[8 + (iters < 15 ? iters : 15)]++;          // [8..23] GJK iteration histogram
            stats[24 + v]++;                                 // [24..26] GJK verdicts SEP / PEN / UNC
            if (iters == 1) stats[27 + v]++;                 // [27..29] verdicts of the items that end in one iteration
            if (v == V_UNC) {                                // [40..47] fp64 iterations of the items fp32 could not certify: <=4, <=8, <=12, <=16, <=24, <=32, <=48, more
              Pose<double> P64[MAX_BODY], id64;
              id64.p = mk<double>(0, 0, 0); id64.q.w = 1; id64.q.x = id64.q.y = id64.q.z = 0;
              for (int k = 0; k < H.nslot; k++) { int ps = H.fk.body_parent[k]; P64[k] = fk_body(H.fk, k, ps < 0 ? id64 : P64[ps], qr); }
              const Shape<double> &A64 = H.shapes[pr.sa], &B64 = H.shapes[pr.sb];
              int it64 = 0;
              gjk_classify(A64, B64, H.verts.data(), relative_pose(A64.slot < 0 ? id64 : P64[A64.slot], B64.slot < 0 ? id64 : P64[B64.slot]), H.pair_rsum64[p], &it64);
              const int edges[7] = {4, 8, 12, 16, 24, 32, 48};
              int b = 0;
              while (b < 7 && it64 > edges[b]) b++;
              stats[40 + b]++;
            }
          } else {
            v = narrow_item<float>(pr.kind, A, B, s->v32.data(), PA, PB, pr.rsum);
            stats[5]++;                                      // closed-form / plane items
          }
          if (v == V_PEN) { pen = true; stats[4]++; }
          if (v == V_UNC) unc = true;
        }
      }
      int v = pen ? 1 : (unc ? 2 : 0);
      if (v == 2) v = row_verdict<double>(H, H.fk, H.shapes.data(), H.verts.data(), qr, false, nullptr) != 0 ? 1 : 0;
      out = v == 0 ? 1 : 0;
      // [32..] the certain-contact shortcut: rows it catches, what they carry, and its false positives (must be 0)
      if (out == 0) stats[32]++;
      if (cert0) { stats[33]++; stats[35] += row_l0; stats[36] += row_exp; stats[37] += row_items; }
      if (cert0 || cert1) { stats[34]++; stats[38] += row_items; if (out) stats[39]++; }
      if (cert0 || cert1) out = 0;   // as on the device: the shortcut's word is final
    }
    valid[r] = (uint8_t)out;
  }
}

// every vertex of every shape lies inside the shape's bounding capsule and inside its group's bounding
// sphere (fp32 tables, as the device sees them): returns the largest violation (<= 0 means contained)
extern "C" double hs_bounds_check(Sim *s) {
  const auto &H = s->H;
  double worst = -1e300;
  for (size_t k = 0; k < s->s32.size(); k++) {
    const Shape<float> &sh = s->s32[k];
    if (sh.kind != SK_VERTS) continue;
    for (int i = 0; i < sh.nvert; i++) {
      const Vtx<float> &v = s->v32[sh.vadr + i];
      double p[3] = {v.x, v.y, v.z}, a[3] = {sh.ca[0], sh.ca[1], sh.ca[2]}, b[3] = {sh.cb[0], sh.cb[1], sh.cb[2]};
      worst = std::max(worst, vkb::point_segment_dist(p, a, b) + (double)sh.radius - (double)sh.crad);
      if (sh.slot >= 0) {
        const int g = sh.group;
        double d = sqrt((p[0] - (float)H.group_c[g][0]) * (p[0] - (float)H.group_c[g][0]) + (p[1] - (float)H.group_c[g][1]) * (p[1] - (float)H.group_c[g][1]) +
                        (p[2] - (float)H.group_c[g][2]) * (p[2] - (float)H.group_c[g][2]));
        worst = std::max(worst, d + (double)sh.radius - H.group_r[g]);
      }
    }
  }
  return worst;
}

// segseg_dist2 in fp32 against a brute-force fp64 minimum over a fine grid: returns the largest
// overestimate of the DISTANCE (the cull needs estimate <= truth + slack)
extern "C" double hs_segseg_check(int ncase, uint64_t seed) {
  double worst = 0;
  for (int c = 0; c < ncase; c++) {
    float x[12];
    for (int k = 0; k < 12; k++) x[k] = sweep_value(seed, c, k, -1.f, 1.f);
    V3<float> p1 = mk<float>(x[0], x[1], x[2]), q1 = mk<float>(x[3], x[4], x[5]), p2 = mk<float>(x[6], x[7], x[8]), q2 = mk<float>(x[9], x[10], x[11]);
    const int kind = c % 6;
    if (kind == 1) q2 = p2 + (q1 - p1) * 0.7f;                                   // parallel
    if (kind == 2) { q2 = p2 + (q1 - p1) * 0.7f; q2.x += 1e-4f * x[9]; }          // nearly parallel
    if (kind == 3) q1 = p1;                                                      // point vs segment
    if (kind == 4) { q1 = p1 + (q1 - p1) * 0.05f; q2 = p2 + (q2 - p2) * 0.05f; }  // short segments
    if (kind == 5) { p2 = p1 + (p2 - p1) * 0.01f; q2 = p2 + (q1 - p1) * 0.5f; q2.y += 3e-3f * x[10]; }  // close, nearly parallel
    const double est = sqrt((double)segseg_dist2(p1, q1, p2, q2));
    // brute force: the distance is convex in (s, t); nested ternary search in fp64
    auto dist = [&](double sA, double tB) {
      double d[3] = {(p1.x + sA * ((double)q1.x - p1.x)) - (p2.x + tB * ((double)q2.x - p2.x)),
                     (p1.y + sA * ((double)q1.y - p1.y)) - (p2.y + tB * ((double)q2.y - p2.y)),
                     (p1.z + sA * ((double)q1.z - p1.z)) - (p2.z + tB * ((double)q2.z - p2.z))};
      return sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    };
    auto best_t = [&](double sA) {
      double lo = 0, hi = 1;
      for (int it = 0; it < 100; it++) { double m1 = lo + (hi - lo) / 3, m2 = hi - (hi - lo) / 3; if (dist(sA, m1) < dist(sA, m2)) hi = m2; else lo = m1; }
      return dist(sA, 0.5 * (lo + hi));
    };
    double lo = 0, hi = 1;
    for (int it = 0; it < 100; it++) { double m1 = lo + (hi - lo) / 3, m2 = hi - (hi - lo) / 3; if (best_t(m1) < best_t(m2)) hi = m2; else lo = m1; }
    const double truth = best_t(0.5 * (lo + hi));
    worst = std::max(worst, est - truth);
  }
  return worst;
}

// mjb_min_distance's per-row routine on the CPU (fp64 core): dist (n), pair (n)
extern "C" void hs_min_distance(Sim *s, const float *q, int64_t n, double far_cap, double *dist, int32_t *pair) {
  const auto &H = s->H;
  for (int64_t r = 0; r < n; r++) {
    double row[MAX_JNT];
    for (int j = 0; j < H.nq; j++) row[j] = (double)q[r * H.nq + j];
    double best; int bp;
    row_min_distance(H.fk, H.nslot, H.shapes.data(), H.verts.data(), H.pairs.data(), H.pair_rsum64.data(), (int)H.pairs.size(), row,
                     far_cap, 1e-3, best, bp);
    dist[r] = best; pair[r] = bp;
  }
}

// support maps: for every mapped hull and `ndir` directions (random, near-axis, near cell borders) the
// mapped support VALUE against the full scan's, in fp32 as the device computes it.  Returns the number
// of mapped shapes; *worst = largest shortfall of the mapped value (0 = always the same maximum);
// *avg_candidates = mean candidates scanned per query.
extern "C" int hs_smap_check(Sim *s, int ndir, uint64_t seed, double *worst, double *avg_candidates) {
  const auto &H = s->H;
  int nmapped = 0;
  long long cand = 0, queries = 0;
  *worst = 0;
  for (size_t k = 0; k < H.shapes.size(); k++) {
    const Shape<float> &sh = s->s32[k];
    if (sh.kind != SK_VERTS || sh.map < 0) continue;
    nmapped++;
    const Vtx<float> *v = s->v32.data() + sh.vadr;
    for (int i = 0; i < ndir; i++) {
      V3<float> d = mk<float>(sweep_value(seed, i, 0, -1.f, 1.f), sweep_value(seed, i, 1, -1.f, 1.f), sweep_value(seed, i, 2, -1.f, 1.f));
      if (i % 4 == 1) { d.x *= 1e-4f; }                                       // near an axis plane
      if (i % 4 == 2) { d.y = d.x * (1.f + 1e-6f * (float)(i % 7 - 3)); }       // near a face boundary of the cube map
      if (i % 4 == 3) { const float q = 0.25f * (float)(i % 9 - 4); d.x = 1.f; d.y = q + 1e-7f * (float)(i % 5 - 2); }  // on cell borders
      const int c = smap_cell(d);
      if (c >= 0) { cand += (long long)(H.smap_cells[sh.map + c] & 255u); queries++; }
      const int bi = support_mapped(v, sh.nvert, H.smap_cells.data() + sh.map, H.smap_ids.data(), d);
      const V3<float> full = support_verts(v, sh.nvert, d);
      const double gap = (double)dot(full, d) - (double)(v[bi].x * d.x + v[bi].y * d.y + v[bi].z * d.z);
      if (gap > *worst) *worst = gap;
    }
  }
  *avg_candidates = queries ? (double)cand / (double)queries : 0.0;
  return nmapped;
}

// per group pair (level 0 of the pipeline): how often it survives on the given rows, its kind and its number of
// member shape pairs, and how many of its members then survive the capsule and the OBB culls: out[g*5 + {0..4}] =
// survivals, members, capsule survivors, OBB survivors, kind
extern "C" int hs_group_census(Sim *s, const float *q, int64_t n, int64_t *out, int32_t *ga, int32_t *gb) {
  const auto &H = s->H;
  Pose<float> P[MAX_BODY], ident;
  ident.p = mk<float>(0, 0, 0); ident.q.w = 1; ident.q.x = ident.q.y = ident.q.z = 0;
  const float slack = 1e-4f;
  for (size_t g = 0; g < H.group_pairs.size(); g++) { out[g * 5 + 1] = H.group_pairs[g].n; out[g * 5 + 4] = H.group_pairs[g].kind; ga[g] = H.group_pairs[g].ga; gb[g] = H.group_pairs[g].gb; }
  for (int64_t r = 0; r < n; r++) {
    const float *qr = q + r * H.nq;
    V3<float> cen[MAX_GROUP];
    for (int k = 0; k < H.nslot; k++) {
      int ps = s->fk32.body_parent[k];
      P[k] = fk_body(s->fk32, k, ps < 0 ? ident : P[ps], qr);
        for (int g = H.slot_group_adr[k]; g < H.slot_group_adr[k] + H.slot_group_num[k]; g++)
          cen[g] = P[k].p + qrot(P[k].q, mk<float>((float)H.group_c[g][0], (float)H.group_c[g][1], (float)H.group_c[g][2]));
    }
    for (size_t gi = 0; gi < H.group_pairs.size(); gi++) {
      const GroupPair &g = H.group_pairs[gi];
      if (!group_pair_near(g, cen[g.ga], g.kind == GK_SPHERE ? cen[g.gb] : cen[g.ga], g.kind == GK_SPHERE ? nullptr : &H.static_groups[g.gb])) continue;
      out[gi * 5 + 0]++;
      for (int i = 0; i < g.n; i++) {
        const Pair pr = H.pairs[H.gp_member[g.first + i]];
        const Shape<float> &A = s->s32[pr.sa], &B = s->s32[pr.sb];
        const Pose<float> &PA = A.slot < 0 ? ident : P[A.slot];
        const Pose<float> &PB = B.slot < 0 ? ident : P[B.slot];
        const float margin = pr.rsum - swept_radius(A) - swept_radius(B);
        if (pr.kind != PK_SEGSEG && capsule_cull(pr, A, B, PA, PB, margin + slack)) continue;
        out[gi * 5 + 2]++;
        if ((pr.flags & PF_OBB) && midphase_cull(pr, A, B, PA, PB, margin, slack)) continue;
        out[gi * 5 + 3]++;
      }
    }
  }
  return (int)H.group_pairs.size();
}

// The inner capsules (fp32 tables, as the device sees them) against the shapes they must lie in: points on
// the capsule's surface are tested against the support function of the shape in many directions (a point
// p is inside hull (+) ball(r) iff d.p <= h(d) + r for every unit d).  Returns the largest violation found
// (<= 0: no sample sticks out); *nshape = shapes that have an inner capsule.
extern "C" double hs_inner_check(Sim *s, int nsample, uint64_t seed, int *nshape) {
  double worst = -1e300;
  int cnt = 0;
  for (size_t k = 0; k < s->s32.size(); k++) {
    const Shape<float> &sh = s->s32[k];
    if (!(sh.irad > 0.f)) continue;
    cnt++;
    const double a[3] = {sh.ia[0], sh.ia[1], sh.ia[2]}, b[3] = {sh.ib[0], sh.ib[1], sh.ib[2]};
    for (int i = 0; i < nsample; i++) {
      double n[3], nn = 0;
      for (int c = 0; c < 3; c++) { n[c] = sweep_value(seed, (uint64_t)(k * 1000003 + i), (uint32_t)c, -1.f, 1.f); nn += n[c] * n[c]; }
      if (nn < 1e-6) continue;
      nn = sqrt(nn);
      const double t = 0.5 + 0.5 * sweep_value(seed, (uint64_t)(k * 1000003 + i), 3u, -1.f, 1.f);
      double p[3];
      for (int c = 0; c < 3; c++) { n[c] /= nn; p[c] = a[c] + t * (b[c] - a[c]) + (double)sh.irad * n[c]; }
      if (sh.kind == SK_CYL) {
        double e[3] = {p[0] - sh.c[0], p[1] - sh.c[1], p[2] - sh.c[2]};
        const double h = e[0] * sh.ax[0] + e[1] * sh.ax[1] + e[2] * sh.ax[2];
        double pr2 = 0;
        for (int c = 0; c < 3; c++) { const double w = e[c] - h * sh.ax[c]; pr2 += w * w; }
        worst = std::max(worst, std::max(sqrt(pr2) - (double)sh.radius, fabs(h) - (double)sh.halflen));
        continue;
      }
      // directions: the capsule's normal at p, and perturbations of it
      for (int j = 0; j < 24; j++) {
        double d[3], dn = 0;
        for (int c = 0; c < 3; c++) {
          d[c] = n[c] + (j ? (j < 12 ? 0.3 : 1.5) * sweep_value(seed + 7, (uint64_t)(k * 1000003 + i), (uint32_t)(4 + 3 * j + c), -1.f, 1.f) : 0.0);
          dn += d[c] * d[c];
        }
        dn = sqrt(dn);
        if (dn < 1e-6) continue;
        double h = -1e300;
        for (int v = 0; v < sh.nvert; v++) {
          const Vtx<float> &V = s->v32[sh.vadr + v];
          h = std::max(h, (d[0] * V.x + d[1] * V.y + d[2] * V.z) / dn);
        }
        worst = std::max(worst, (d[0] * p[0] + d[1] * p[1] + d[2] * p[2]) / dn - h - (double)sh.radius);
      }
    }
  }
  if (nshape) *nshape = cnt;
  return worst;
}

// point_segment_d2 (the expanded square level 0 compares with its squared limits) against the exact value in
// fp64: returns the largest |error| / |e|_max^2 of the squared distance over random points up to |e|_max = 3 sqrt(3) m
// from the segment's start and segments up to 2 m long (build_groups allows 8 eps32 |e|_max^2 in the squared limits)
extern "C" double hs_point_segment_check(int ncase, uint64_t seed) {
  double worst = 0;
  for (int c = 0; c < ncase; c++) {
    float x[7];
    for (int k = 0; k < 7; k++) x[k] = sweep_value(seed, c, k, -1.f, 1.f);
    const V3<float> e = mk<float>(3.f * x[0], 3.f * x[1], 3.f * x[2]);
    double un = sqrt((double)x[3] * x[3] + (double)x[4] * x[4] + (double)x[5] * x[5]);
    if (un < 1e-3) continue;
    const V3<float> u = mk<float>((float)(x[3] / un), (float)(x[4] / un), (float)(x[5] / un));
    const float len = (c % 5 == 0) ? 0.f : 1.f + x[6];
    float sc;
    const double got = (double)point_segment_d2(e, u, len, &sc);
    // exact: the segment the fp32 tables describe (start 0, direction u as stored, length len)
    const double ul = sqrt((double)u.x * u.x + (double)u.y * u.y + (double)u.z * u.z);
    const double s = ((double)e.x * u.x + (double)e.y * u.y + (double)e.z * u.z) / ul;
    const double t = s < 0 ? 0 : (s > len ? len : s);
    const double dx = e.x - t * u.x / ul, dy = e.y - t * u.y / ul, dz = e.z - t * u.z / ul;
    worst = std::max(worst, fabs(got - (dx * dx + dy * dy + dz * dz)));
  }
  return worst / 27.0;
}

extern "C" double hs_l0_sq_err(Sim *s) { return s->H.l0_sq_err; }
