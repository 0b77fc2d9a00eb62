// vk_row.cuh -- the validity path for SMALL launches: one warp per row.
//
// The planners spend their time in launches of a few hundred to a few thousand rows (the tail of a
// batch of bi-RRT queries, every tick of the constrained planner, scalar valid_config calls).  The
// throughput kernels put one LANE on a row, so such a launch runs as long as the longest dependent
// chain of one lane: FK, then ~320 pair tests one after the other, then GJK items whose hull scans
// are serial -- 120-150 us however few rows there are.  Here the 32 lanes of a warp share ONE row:
//   FK                 lane 0 (a chain of ~10 quaternion products; everything else waits ~2 us)
//   level 0            lanes over the group pairs (vk_pipe.cuh's hierarchy: body spheres, world capsules)
//   expansion, culls   lanes over the shape pairs of the surviving groups: bounding capsules, OBBs
//   narrow phase       closed forms lane = item; GJK with 8 lanes on an item (hull scans split 8 ways,
//                      4 items in flight per warp), first certain contact ends the row
// Same arithmetic and the same certified verdicts as the other kernels (shared core functions):
// identical results, which tests/test_gpu_parity.py checks across all three paths.
#pragma once

#include "vk_kernels.cuh"
#include "vk_split.cuh"

namespace vk {

constexpr int ROWK_THREADS = 256;          // 8 rows in flight per CTA
constexpr int ROWK_G = 8;                  // lanes per GJK item
constexpr int ROWK_MAXITEMS = 512;         // per-row lists (group pairs, shape pairs, items) in shared memory

struct RowLayout { size_t verts, shapes, pairs, gpairs, sgroups, member, adjs, adj, per_warp, bars, total, warp_bytes; };
__host__ __device__ inline RowLayout row_layout(int nvert, int nshape, int npair, int ngpair, int nsgroup, int nmember, int nadj,
                                                int nslot, int ngroup_moving, int nq) {
  RowLayout L;
  size_t o = 0;
  L.verts = o; o = align_up(o + (size_t)nvert * sizeof(Vtx<float>), 128);
  L.shapes = o; o = align_up(o + (size_t)nshape * sizeof(Shape<float>), 128);
  L.pairs = o; o = align_up(o + (size_t)npair * sizeof(Pair), 128);
  L.gpairs = o; o = align_up(o + (size_t)ngpair * sizeof(GroupPair), 128);
  L.sgroups = o; o = align_up(o + (size_t)(nsgroup > 0 ? nsgroup : 1) * sizeof(StaticGroup), 128);
  L.member = o; o = align_up(o + (size_t)nmember * sizeof(uint16_t), 128);
  L.adjs = o; o = align_up(o + (size_t)(nvert + 1) * sizeof(uint16_t), 128);
  L.adj = o; o = align_up(o + (size_t)nadj, 128);
  // per warp: row (nq), poses (nslot x 8), group centres (ngm x 4), a flag word, two u16 lists
  L.warp_bytes = align_up((size_t)(nq + nslot * 8 + (ngroup_moving > 0 ? ngroup_moving : 1) * 4 + 1) * sizeof(float) + 2 * ROWK_MAXITEMS * sizeof(uint16_t), 128);
  L.per_warp = o; o += L.warp_bytes * (ROWK_THREADS / 32);
  L.bars = o; o = align_up(o + 64, 128);
  L.total = o;
  return L;
}

__global__ void __launch_bounds__(ROWK_THREADS, 2) row_kernel(const __grid_constant__ KArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int nq = a.fk.nq;
  const RowLayout L = row_layout(a.nvert, a.nshape, a.npair, a.ngpair, a.nsgroup, a.nmember, a.nadj, a.nslot, a.ngroup_moving, nq);
  Vtx<float> *s_verts = reinterpret_cast<Vtx<float> *>(smem + L.verts);
  Shape<float> *s_shapes = reinterpret_cast<Shape<float> *>(smem + L.shapes);
  Pair *s_pairs = reinterpret_cast<Pair *>(smem + L.pairs);
  GroupPair *s_gp = reinterpret_cast<GroupPair *>(smem + L.gpairs);
  StaticGroup *s_sg = reinterpret_cast<StaticGroup *>(smem + L.sgroups);
  uint16_t *s_member = reinterpret_cast<uint16_t *>(smem + L.member);
  uint16_t *s_adjs = reinterpret_cast<uint16_t *>(smem + L.adjs);
  uint8_t *s_adj = smem + L.adj;
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem + L.bars);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ngm = a.ngroup_moving > 0 ? a.ngroup_moving : 1;
  unsigned char *wbase = smem + L.per_warp + (size_t)warp * L.warp_bytes;
  float *wq = reinterpret_cast<float *>(wbase);
  float *wpose = wq + nq;                 // [slot][8]: px py pz qw qx qy qz 0
  float *wcen = wpose + a.nslot * 8;      // [group][4]
  uint32_t *wflag = reinterpret_cast<uint32_t *>(wcen + ngm * 4);   // the fp64 item list was full for an item of this row
  uint16_t *list0 = reinterpret_cast<uint16_t *>(wflag + 1);
  uint16_t *list1 = list0 + ROWK_MAXITEMS;

  // small launches in the modes whose row count only the device knows (edges, chains) are dispatched
  // here: both this kernel and the throughput kernel are enqueued, and the one that is not in its
  // regime returns at once
  long long nrows = a.n;
  if (a.mode == MODE_EDGES || a.mode == MODE_CHAINS) nrows = a.edge_prefix[a.nedge];
  if (a.rowk_max >= 0 && nrows > a.rowk_max) return;
  if ((long long)blockIdx.x * (ROWK_THREADS / 32) >= nrows) return;   // the warps of the CTAs before this one cover every row

  const uint32_t bytes_v = (uint32_t)(a.nvert * sizeof(Vtx<float>));
  const uint32_t bytes_s = (uint32_t)(a.nshape * sizeof(Shape<float>));
  const uint32_t bytes_p = (uint32_t)(a.npair * sizeof(Pair));
  const uint32_t bytes_g = (uint32_t)(a.ngpair * sizeof(GroupPair));
  const uint32_t bytes_sg = (uint32_t)(a.nsgroup * sizeof(StaticGroup));
  const uint32_t bytes_m = (uint32_t)align_up((size_t)a.nmember * sizeof(uint16_t), 16);
  const uint32_t bytes_as = (uint32_t)align_up((size_t)(a.nvert + 1) * sizeof(uint16_t), 16);
  const uint32_t bytes_a = (uint32_t)align_up((size_t)a.nadj, 16);
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&s_bar[0], bytes_v + bytes_s + bytes_p + bytes_g + bytes_sg + bytes_m + bytes_as + bytes_a);
    if (bytes_v) bulk_g2s(s_verts, a.verts, bytes_v, &s_bar[0]);
    if (bytes_s) bulk_g2s(s_shapes, a.shapes, bytes_s, &s_bar[0]);
    if (bytes_p) bulk_g2s(s_pairs, a.pairs, bytes_p, &s_bar[0]);
    if (bytes_g) bulk_g2s(s_gp, a.gpairs, bytes_g, &s_bar[0]);
    if (bytes_sg) bulk_g2s(s_sg, a.sgroups, bytes_sg, &s_bar[0]);
    if (bytes_m) bulk_g2s(s_member, a.gp_member, bytes_m, &s_bar[0]);
    bulk_g2s(s_adjs, a.adj_start, bytes_as, &s_bar[0]);
    if (bytes_a) bulk_g2s(s_adj, a.adj, bytes_a, &s_bar[0]);
  }
  mbar_wait(&s_bar[0], 0);
  __syncthreads();

  const bool use_obb = !(a.flags & F_NO_OBB);
  const float slack = 1e-4f;
  const unsigned below = (1u << lane) - 1u;
  const int gl = lane & (ROWK_G - 1), grp = lane / ROWK_G;
  const unsigned gmask = ((1u << ROWK_G) - 1u) << (lane & ~(ROWK_G - 1));
  long long items_total = 0, rows_total = 0;

  for (;;) {
    long long row = 0;
    if (lane == 0) row = (long long)atomicAdd(&a.counters[C_TICKET], 1ull);
    row = __shfl_sync(0xffffffffu, row, 0);
    if (row >= nrows) break;
    rows_total++;
    if (lane == 0) *wflag = 0;
    // ---- the row, its joint-limit answer ---------------------------------------------------------------
    long long e_idx = 0;
    int e_k = 0;
    int lim_ok = 1;
    if (lane == 0) {
      if (a.mode == MODE_DENSE) {
        for (int j = 0; j < nq; j++) wq[j] = a.q[row * a.ldq + j];
      } else if (a.mode == MODE_EDGES) {
        edge_lookup(a.edge_prefix, a.nedge, row, e_idx, e_k);
        edge_row<float>(a.q0, a.q1, a.ldq, nq, a.step, e_idx, e_k, wq);
      } else if (a.mode == MODE_CHAINS) {
        edge_lookup(a.edge_prefix, a.nedge, row, e_idx, e_k);
        const bool lim = a.flags & F_LIMITS;
        lim_ok = chain_point<float>(a.c0, a.c1, nq, a.ceps, e_idx, e_k, wq, lim ? a.jnt_lo : nullptr, lim ? a.jnt_hi : nullptr);
      } else {
        for (int j = 0; j < nq; j++)
          wq[j] = sweep_value(a.seed, (uint64_t)(a.row0 + row), (uint32_t)j, a.fk.jnt_lo[j], a.fk.jnt_hi[j]);
      }
      if ((a.flags & F_LIMITS) && a.mode != MODE_CHAINS) lim_ok = limits_ok(wq, a.fk.njnt, a.jnt_lo, a.jnt_hi, a.flags & F_LIMITS_OUTWARD);
      // ---- FK: poses and group centres of this row -> shared -----------------------------------------
      if (lim_ok && (a.flags & F_COLLISION)) {
        Pose<float> prev;
        prev.p = mk<float>(0, 0, 0); prev.q.w = 1; prev.q.x = prev.q.y = prev.q.z = 0;
        int prev_slot = -1;
#pragma unroll 1
        for (int s = 0; s < a.nslot; s++) {
          const int ps = a.fk.body_parent[s];
          Pose<float> P = prev;
          if (ps != prev_slot) {
            if (ps < 0) { P.p = mk<float>(0, 0, 0); P.q.w = 1; P.q.x = P.q.y = P.q.z = 0; }
            else { const float *b = wpose + ps * 8; P.p = mk<float>(b[0], b[1], b[2]); P.q.w = b[3]; P.q.x = b[4]; P.q.y = b[5]; P.q.z = b[6]; }
          }
          const Pose<float> B = fk_body(a.fk, s, P, wq);
          prev = B; prev_slot = s;
          float *b = wpose + s * 8;
          b[0] = B.p.x; b[1] = B.p.y; b[2] = B.p.z; b[3] = B.q.w; b[4] = B.q.x; b[5] = B.q.y; b[6] = B.q.z;
          for (int g = a.slot_group_adr[s]; g < a.slot_group_adr[s] + a.slot_group_num[s]; g++) {
            const V3<float> c = B.p + qrot(B.q, mk<float>(a.group_c[g][0], a.group_c[g][1], a.group_c[g][2]));
            wcen[g * 4] = c.x; wcen[g * 4 + 1] = c.y; wcen[g * 4 + 2] = c.z;
          }
        }
      }
    }
    lim_ok = __shfl_sync(0xffffffffu, lim_ok, 0);
    e_idx = __shfl_sync(0xffffffffu, e_idx, 0);
    e_k = __shfl_sync(0xffffffffu, e_k, 0);
    __syncwarp();
    bool hit = false;
    unsigned unc_any = 0;
    if (lim_ok && (a.flags & F_COLLISION)) {
      auto pose_of = [&](int slot) {
        Pose<float> P;
        if (slot < 0) { P.p = mk<float>(0, 0, 0); P.q.w = 1; P.q.x = P.q.y = P.q.z = 0; return P; }
        const float *b = wpose + slot * 8;
        P.p = mk<float>(b[0], b[1], b[2]); P.q.w = b[3]; P.q.x = b[4]; P.q.y = b[5]; P.q.z = b[6];
        return P;
      };
      // ---- level 0: lanes over the group pairs -> list0 (group pair ids) ---------------------------------
      // (a group pair whose inner balls / tube overlap is a certain contact: the row is settled, see GroupPair::lim_in)
      int n0 = 0;
      for (int p0 = 0; p0 < a.ngpair && !hit; p0 += 32) {
        const int p = p0 + lane;
        int near = 0;
        if (p < a.ngpair) {
          const GroupPair g = s_gp[p];
          const V3<float> cA = mk<float>(wcen[g.ga * 4], wcen[g.ga * 4 + 1], wcen[g.ga * 4 + 2]);
          const V3<float> cB = g.kind == GK_SPHERE ? mk<float>(wcen[g.gb * 4], wcen[g.gb * 4 + 1], wcen[g.gb * 4 + 2]) : cA;
          near = group_pair_test(g, cA, cB, g.kind == GK_SPHERE ? nullptr : &s_sg[g.gb]);
        }
        const unsigned m = __ballot_sync(0xffffffffu, near != 0);
        if (near) list0[n0 + __popc(m & below)] = (uint16_t)p;
        n0 += __popc(m);
        hit = __any_sync(0xffffffffu, near == 2);
      }
      if (hit) n0 = 0;
      __syncwarp();
      // ---- expansion: the shape pairs of the surviving group pairs -> list1 -------------------------------
      int T = 0;
      for (int i0 = 0; i0 < n0; i0 += 32) {
        const int i = i0 + lane;
        int first = 0, n = 0;
        if (i < n0) { const GroupPair g = s_gp[list0[i]]; first = g.first; n = g.n; }
        int off = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, off, o);
          if (lane >= o) off += v;
        }
        const int tot = __shfl_sync(0xffffffffu, off, 31);
        off += T - n;
        for (int k = 0; k < n; k++)
          if (off + k < ROWK_MAXITEMS) list1[off + k] = s_member[first + k];
        T += tot;
      }
      __syncwarp();
      bool overflow = T > ROWK_MAXITEMS;   // cannot happen while the model has <= 512 shape pairs per row's survivors
      if (T > ROWK_MAXITEMS) T = ROWK_MAXITEMS;
      // ---- culls: lanes over the shape pairs (bounding capsules, then OBBs) -> list0 (items) --------------
      int ni = 0;
      for (int i0 = 0; i0 < T; i0 += 32) {
        const int i = i0 + lane;
        bool keep = false, certain = false;
        int ip = 0;
        if (i < T) {
          ip = list1[i];
          const Pair pr = s_pairs[ip];
          keep = true;
          if (use_obb) {
            const Shape<float> &A = s_shapes[pr.sa];
            const Shape<float> &B = s_shapes[pr.sb];
            const Pose<float> PA = pose_of(A.slot), PB = pose_of(B.slot);
            keep = !midphase_cull(pr, A, B, PA, PB, pr.rsum - swept_radius(A) - swept_radius(B), slack);
            if (keep && pr.kind != PK_SEGSEG && inner_contact(pr, A, B, PA, PB)) certain = true;   // inner capsules overlap
          }
        }
        if (__any_sync(0xffffffffu, certain)) { hit = true; break; }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        __syncwarp();
        if (keep) list0[ni + __popc(m & below)] = (uint16_t)ip;   // ni <= i0: never ahead of the entries still to be read from list1
        ni += __popc(m);
      }
      __syncwarp();
      if (hit) ni = 0;
      items_total += ni;
      // ---- narrow phase ---------------------------------------------------------------------------------------
      // closed forms and plane items: lane = item
      for (int i0 = 0; i0 < ni && !hit; i0 += 32) {
        const int i = i0 + lane;
        int v = V_SEP, ip = 0;
        if (i < ni) {
          ip = list0[i];
          const Pair pr = s_pairs[ip];
          if (pr.kind != PK_GJK) {
            const Shape<float> &A = s_shapes[pr.sa];
            const Shape<float> &B = s_shapes[pr.sb];
            const Pose<float> PA = pose_of(A.slot), PB = pose_of(B.slot);
            if (pr.kind == PK_PLANE) v = plane_classify(A, B, PB, pr.rsum, [&](V3<float> d) { return support_shape(B, s_verts, d); });
            else v = segseg_item(A, B, s_verts, PA, PB, pr.rsum);
          }
        }
        if (v == V_UNC && !(a.flags & F_NO_RECHECK)) note_uncertain(a.counters, a.recheck_items, a.item_cap, row, ip, wflag);
        hit = __any_sync(0xffffffffu, v == V_PEN);
        unc_any |= __ballot_sync(0xffffffffu, v == V_UNC);
      }
      // GJK items: ROWK_G lanes per item, 32 / ROWK_G items in flight; a group that is done takes the next item
      if (!hit) {
        int next = 0;          // warp-uniform cursor into list0
        bool have = false;
        GjkState<float> gs;
        Rel<float> rel;
        const Shape<float> *SA = s_shapes, *SB = s_shapes;
        float R = 0.f;
        int pidx = 0, wa = -1, wb = -1;
#pragma unroll 1
        for (;;) {
          // hand out items to the groups that need one (lane gl == 0 of each group asks)
          const unsigned need = __ballot_sync(0xffffffffu, !have && gl == 0);
          int mine = -1;
          if (need) {
            // every lane walks the cursor identically: find the next GJK items, one per needing group in group order
            int cur = next;
            unsigned todo = need;
            while (todo && cur < ni) {
              if (s_pairs[list0[cur]].kind == PK_GJK) {
                const int g0 = (__ffs(todo) - 1) / ROWK_G;
                if (g0 == grp) mine = cur;
                todo &= todo - 1;
              }
              cur++;
            }
            next = cur;
          }
          if (!have && mine >= 0) {
            pidx = list0[mine];
            const Pair pr = s_pairs[pidx];
            SA = s_shapes + pr.sa; SB = s_shapes + pr.sb; R = pr.rsum;
            rel = relative_pose(pose_of(SA->slot), pose_of(SB->slot));
            gjk_init(gs, *SA, *SB, rel);
            wa = wb = -1;
            have = true;
          }
          if (__ballot_sync(0xffffffffu, have) == 0) break;   // no item left anywhere
          int v = -1;
          if (have) {
            const Shape<float> &As = *SA, &Bs = *SB;
            v = gjk_step_impl(
                gs, rel, R, [&](V3<float> d) { return group_support<ROWK_G>(As, s_verts, s_adjs, s_adj, d, gl, gmask, wa); },
                [&](V3<float> d) { return group_support<ROWK_G>(Bs, s_verts, s_adjs, s_adj, d, gl, gmask, wb); });
            if (v >= 0) {
              have = false;
              if (v == V_UNC && gl == 0 && !(a.flags & F_NO_RECHECK)) note_uncertain(a.counters, a.recheck_items, a.item_cap, row, pidx, wflag);
            }
          }
          unc_any |= __ballot_sync(0xffffffffu, v == V_UNC);
          if (__any_sync(0xffffffffu, v == V_PEN)) { hit = true; break; }
        }
      }
      __syncwarp();
      if ((overflow || *wflag) && !hit && lane == 0) {   // lists too small for this row / fp64 item list full: whole row in fp64
        const unsigned long long slot = atomicAdd(&a.counters[C_RECHECK], 1ull);
        a.recheck_rows[slot] = row;
      }
    }
    // ---- the row's answer ---------------------------------------------------------------------------------
    if (lane == 0) {
      const bool ok = lim_ok && !hit;
      const bool pending = ok && unc_any != 0;
      if (pending) atomicAdd(&a.counters[C_UNCERTAIN], 1ull);
      if (a.mode == MODE_EDGES || a.mode == MODE_CHAINS) {
        if (!ok || (pending && (a.flags & F_NO_RECHECK))) atomicMin(&a.first_bad[e_idx], e_k);
      } else {
        a.valid[row] = pending ? (uint8_t)((a.flags & F_NO_RECHECK) ? 2 : 1) : (uint8_t)(ok ? 1 : 0);
      }
    }
    __syncwarp();
  }
  if (lane == 0 && items_total) atomicAdd(&a.counters[C_ITEMS], (unsigned long long)items_total);
  if (lane == 0 && rows_total) atomicAdd(&a.counters[C_ROWS], (unsigned long long)rows_total);
}

}  // namespace vk
