// vk_pipe.cuh -- the validity path for large batches: a pipeline of small kernels.
//
//   fk_cull_kernel  lane = row.  Joint limits, forward kinematics (poses to an L2-resident
//                   [row][slot][8] array), then LEVEL 0 of the broad phase on cull GROUPS: a moving body
//                   with all its shapes is one bounding sphere, a world-fixed shape is its bounding
//                   capsule or plane.  ~120 group tests per row instead of ~320 shape-pair tests, and no
//                   per-shape state in shared memory: 6.4 KB per warp, <= 64 registers, 32 warps per SM.
//                   Surviving (row, group pair) entries are staged per warp and flushed to a global list
//                   with ONE atomic per flush.
//   mid_kernel      lane = entry / shape pair.  Expands every surviving group pair into its shape pairs,
//                   culls them with the shapes' bounding CAPSULES (the reference scenes' obstacles are
//                   long thin boxes and capsules: spheres around them are useless, capsules are tight),
//                   compacts, culls the survivors with the OBB separating-axis test, and appends what is
//                   left to the narrow phase's bins (closed-form kinds into the first bin).
//   narrow_kernel   (vk_split.cuh) persistent lanes over the bins, certified verdicts, fp64 list.
//
// Measured on the uniform Franka sweep (scene_with_obstacles, tests/hostsim census): 6.1 group pairs
// per row survive level 0, they expand to 20 shape pairs, 3.8 survive the capsules, 2.6 the OBBs --
// against 30 sphere survivors and 7.8 narrow-phase items per row with per-shape bounding spheres.
// Culls are conservative (slack 1e-4 m, far above fp32 rounding) and never decide a result.
#pragma once

#include "vk_kernels.cuh"
#include "vk_split.cuh"

namespace vk {

#ifndef VK_L0_QCAP
#define VK_L0_QCAP 256
#endif
constexpr int L0_QCAP = VK_L0_QCAP;   // per-warp staging queue of level-0 survivors (entries)
constexpr int PIPE_FK_THREADS = 256;  // fk_cull_kernel: 8 warps per CTA, 4 CTAs per SM
constexpr int MID_THREADS = 256;
#ifndef VK_MID_CTAS
#define VK_MID_CTAS 2
#endif
constexpr int MID_CTAS = VK_MID_CTAS;      // resident CTAs per SM the register budget is set for
constexpr int MID_Q1CAP = 1024;       // per-warp queue of expanded shape pairs
constexpr int MID_Q2CAP = 96;         // per-warp queue of capsule survivors waiting for the OBB test
#ifndef VK_MID_CHUNK
#define VK_MID_CHUNK 256
#endif
constexpr int MID_CHUNK = VK_MID_CHUNK;        // level-0 entries a warp claims per ticket
#ifndef VK_MID_AHEAD
#define VK_MID_AHEAD 1
#endif
constexpr int MID_AHEAD = VK_MID_AHEAD;        // batches before the end of its chunk at which a warp claims the next ticket
#ifndef VK_MID_STAGE
#define VK_MID_STAGE 32
#endif
static_assert(VK_MID_STAGE >= 32, "a drain batch can put 32 items into one bin");
constexpr int MID_STAGE = VK_MID_STAGE;         // per-warp, per-bin staging of narrow-phase items (entries)

struct PipeFkLayout { size_t gpairs, sgroups, cen, qtile, q0, bars, total; };
__host__ __device__ inline PipeFkLayout pipe_fk_layout(int ngpair, int nsgroup, int ngroup_moving, int nq) {
  PipeFkLayout L;
  const int W = PIPE_FK_THREADS / 32;
  size_t o = 0;
  L.gpairs = o; o = align_up(o + (size_t)ngpair * sizeof(GroupPair), 128);
  L.sgroups = o; o = align_up(o + (size_t)(nsgroup > 0 ? nsgroup : 1) * sizeof(StaticGroup), 128);
  L.cen = o; o = align_up(o + (size_t)W * (ngroup_moving > 0 ? ngroup_moving : 1) * 3 * 32 * sizeof(float), 128);
  L.qtile = o; o = align_up(o + (size_t)W * 32 * nq * sizeof(float), 128);
  L.q0 = o; o = align_up(o + (size_t)W * L0_QCAP * sizeof(uint32_t), 128);
  L.bars = o; o = align_up(o + 64 + 8 * W, 128);
  L.total = o;
  return L;
}

// rows whose level-0 entries did not fit the global list are re-evaluated whole in fp64 (the list is
// sized at several times the calibrated average; only degenerate batches get here)
__device__ __forceinline__ void pipe_row_overflow(const KArgs &a, long long row) {
  const uint32_t bit = 1u << ((unsigned)(row & 3) * 8u);
  if (!(atomicOr(&a.row_flags[row >> 2], bit) & bit)) {
    const unsigned long long s = atomicAdd(&a.counters[C_RECHECK], 1ull);
    a.recheck_rows[s] = row;
  }
}

// One warp's queued level-0 entries -> the global list (one atomic).  Entries of rows that were settled
// after they queued (`live` bit clear) stay behind.  Out of line: it runs once per ~200 entries and would
// otherwise be inlined into every push site of the level-0 loops.
__device__ __noinline__ void l0_flush(const KArgs &a, const uint32_t *q0, int n0, unsigned live, unsigned coll_lanes, long long row_base) {
  const int lane = threadIdx.x & 31;
  const unsigned below = (1u << lane) - 1u;
  int cnt = n0;
  if (live != coll_lanes) {   // some row of the tile has been settled: its entries stay behind
    cnt = 0;
    for (int i0 = 0; i0 < n0; i0 += 32) {
      const int i = i0 + lane;
      cnt += __popc(__ballot_sync(0xffffffffu, i < n0 && ((live >> (q0[i < n0 ? i : 0] & 31u)) & 1u)));
    }
  }
  unsigned long long base = 0;
  if (lane == 0 && cnt) base = atomicAdd(&a.counters[C_L0], (unsigned long long)cnt);
  base = __shfl_sync(0xffffffffu, base, 0);
  for (int i0 = 0; i0 < n0; i0 += 32) {
    const int i = i0 + lane;
    const uint32_t e = q0[i < n0 ? i : 0];
    const bool k = i < n0 && ((live >> (e & 31u)) & 1u);
    const unsigned m = __ballot_sync(0xffffffffu, k);
    if (k) {
      const unsigned long long pos = base + __popc(m & below);
      const unsigned long long r = (unsigned long long)(row_base + (e & 31u));
      if (pos < a.l0_cap) a.l0_items[pos] = r | ((unsigned long long)(e >> 5) << 40);
      else pipe_row_overflow(a, (long long)r);
    }
    base += __popc(m);
  }
  __syncwarp();
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

#ifndef VK_FK_CTAS
#define VK_FK_CTAS 4
#endif
#ifndef VK_L0_UNROLL
#define VK_L0_UNROLL 2   // B200, 1M Franka rows: 1 -> 0.381 ms, 2 -> 0.361, 3 -> 0.380, 4 -> 0.374, 8 -> 0.409
                         // (the tables as a kernel parameter, read with indexed LDC instead of LDS: 0.428)
#endif
__global__ void __launch_bounds__(PIPE_FK_THREADS, VK_FK_CTAS) fk_cull_kernel(const __grid_constant__ KArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int nq = a.fk.nq;
  const PipeFkLayout L = pipe_fk_layout(a.ngpair, a.nsgroup, a.ngroup_moving, nq);
  GroupPair *s_gp = reinterpret_cast<GroupPair *>(smem + L.gpairs);
  StaticGroup *s_sg = reinterpret_cast<StaticGroup *>(smem + L.sgroups);
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem + L.bars);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ngm = a.ngroup_moving > 0 ? a.ngroup_moving : 1;
  float *cen = reinterpret_cast<float *>(smem + L.cen) + (size_t)warp * ngm * 96;   // [group][xyz][lane]
  float *wq = reinterpret_cast<float *>(smem + L.qtile) + (size_t)warp * 32 * nq;   // this warp's 32 rows
  uint32_t *q0 = reinterpret_cast<uint32_t *>(smem + L.q0) + (size_t)warp * L0_QCAP;

  // ---- one-time: group tables -> shared memory through the bulk-copy engine ----------------------
  const uint32_t bytes_g = (uint32_t)(a.ngpair * sizeof(GroupPair));
  const uint32_t bytes_s = (uint32_t)(a.nsgroup * sizeof(StaticGroup));
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&s_bar[0], bytes_g + bytes_s);
    if (bytes_g) bulk_g2s(s_gp, a.gpairs, bytes_g, &s_bar[0]);
    if (bytes_s) bulk_g2s(s_sg, a.sgroups, bytes_s, &s_bar[0]);
  }
  mbar_wait(&s_bar[0], 0);
  __syncthreads();

  long long nrows = a.n;
  if (a.mode == MODE_EDGES || a.mode == MODE_CHAINS) nrows = a.edge_prefix[a.nedge];
  const long long ntiles = (nrows + 31) / 32;
  uint32_t row_parity = 0;
  const bool dense_bulk = (a.mode == MODE_DENSE) && (a.ldq == nq) && ((reinterpret_cast<uintptr_t>(a.q) & 15) == 0);
  long long rows_total = 0;
  uint64_t *wbar = s_bar + 8 + warp;
  if (lane == 0) mbar_init(wbar, 1);
  fence_barrier_init();
  __syncwarp();
  const unsigned below = (1u << lane) - 1u;

  for (;;) {
    long long tile = 0;
    if (lane == 0) tile = (long long)atomicAdd(&a.counters[C_TICKET], 1ull);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile >= ntiles) break;
    const long long row_base = tile * 32;
    const int rows_here = (int)((nrows - row_base) < 32 ? (nrows - row_base) : 32);
    const long long row = row_base + lane;
    const bool active = lane < rows_here;
    rows_total += rows_here;

    // ---- P0: the warp's rows -> shared (one TMA bulk copy per tile) ---------------------------------
    if (a.mode == MODE_DENSE) {
      if (a.rows_ready) {   // host -> device copy still in flight (mjb_check_configs_host)
        if (lane == 0) {
          const unsigned long long need = (unsigned long long)(a.row0 + row_base + rows_here);   // row0: this launch's first row within the host batch
          const long long t0 = clock64();
          while (ld_acquire_sys(a.rows_ready) < need) {
            __nanosleep(256);
            if (clock64() - t0 > 8000000000ll) __trap();
          }
        }
        __syncwarp();
      }
      if (dense_bulk && rows_here == 32) {
        if (lane == 0) {
          fence_proxy_async();
          mbar_expect_tx(wbar, (uint32_t)(32 * nq * sizeof(float)));
          bulk_g2s(wq, a.q + row_base * nq, (uint32_t)(32 * nq * sizeof(float)), wbar);
        }
        mbar_wait(wbar, row_parity);
        row_parity ^= 1;
      } else {
        for (int i = lane; i < rows_here * nq; i += 32) {
          int r = i / nq, j = i - r * nq;
          wq[i] = a.q[(row_base + r) * a.ldq + j];
        }
      }
      __syncwarp();
    }

    // ---- P1: limits + FK, lane = row; the row's preliminary answer ----------------------------------
    float *q = wq + lane * nq;
    long long e_idx = 0;
    int e_k = 0;
    bool lim_ok = true;
    if (active) {
      if (a.mode == MODE_EDGES) {
        edge_lookup(a.edge_prefix, a.nedge, row, e_idx, e_k);
        edge_row<float>(a.q0, a.q1, a.ldq, nq, a.step, e_idx, e_k, q);
      } else if (a.mode == MODE_CHAINS) {
        edge_lookup(a.edge_prefix, a.nedge, row, e_idx, e_k);
        const bool lim = a.flags & F_LIMITS;
        lim_ok = chain_point<float>(a.c0, a.c1, nq, a.ceps, e_idx, e_k, q, lim ? a.jnt_lo : nullptr, lim ? a.jnt_hi : nullptr);
      } else if (a.mode == MODE_SWEEP) {
#pragma unroll 1
        for (int j = 0; j < nq; j++)
          q[j] = sweep_value(a.seed, (uint64_t)(a.row0 + row), (uint32_t)j, a.fk.jnt_lo[j], a.fk.jnt_hi[j]);
      }
      if ((a.flags & F_LIMITS) && a.mode != MODE_CHAINS)
        lim_ok = limits_ok(q, a.fk.njnt, a.jnt_lo, a.jnt_hi, a.flags & F_LIMITS_OUTWARD);
      if ((a.mode == MODE_EDGES || a.mode == MODE_CHAINS) && !(a.flags & F_ROWMASK)) {
        if (!lim_ok) atomicMin(&a.first_bad[e_idx], e_k);
      } else {
        a.valid[row] = lim_ok ? 1 : 0;   // the later kernels can only turn it to 0
      }
    }
    const bool do_coll = active && lim_ok && (a.flags & F_COLLISION);
    const unsigned coll_lanes = __ballot_sync(0xffffffffu, do_coll);
    if (coll_lanes == 0) continue;
    if (do_coll) {
      Pose<float> prev;
      prev.p = mk<float>(0, 0, 0); prev.q.w = 1; prev.q.x = prev.q.y = prev.q.z = 0;
      int prev_slot = -1;
#pragma unroll 1
      for (int s = 0; s < a.nslot; s++) {
        const int ps = a.fk.body_parent[s];
        Pose<float> P = (ps == prev_slot) ? prev : load_pose8(a.pose8, a.nslot, row, ps);
        Pose<float> B = fk_body(a.fk, s, P, q);
        prev = B; prev_slot = s;
        float4 *b = reinterpret_cast<float4 *>(a.pose8 + ((size_t)row * a.nslot + s) * 8);
        b[0] = make_float4(B.p.x, B.p.y, B.p.z, B.q.w);
        b[1] = make_float4(B.q.x, B.q.y, B.q.z, 0.f);
        for (int g = a.slot_group_adr[s]; g < a.slot_group_adr[s] + a.slot_group_num[s]; g++) {
          const V3<float> c = B.p + qrot(B.q, mk<float>(a.group_c[g][0], a.group_c[g][1], a.group_c[g][2]));
          float *cc = cen + g * 96 + lane;
          cc[0] = c.x; cc[32] = c.y; cc[64] = c.z;
        }
      }
    }
    __syncwarp();

    // ---- level 0: group pairs, lane = row; survivors -> per-warp queue -> global list --------------
    int n0 = 0;   // warp-uniform queue fill
    // A group pair whose INNER balls / tube overlap settles the row on the spot (certain contact, see
    // GroupPair::lim_in): the row is marked, stops queueing, and what it queued so far is dropped at the
    // next flush.
    bool alive = do_coll;
    // (the certain-contact test is only looked at when some lane of the warp queues the pair: a certain pair is
    // a near pair, and most trips of the loops below end at the vote)
    auto push = [&](bool s, int p, auto certain) {
      const unsigned m = __ballot_sync(0xffffffffu, s);
      if (m) {
        if (s) {
          q0[n0 + __popc(m & below)] = (uint32_t)lane | ((uint32_t)p << 5);
          if (certain()) alive = false;   // the entry just queued is dropped with the row's others at the flush
        }
        n0 += __popc(m);
        if (n0 + 32 > L0_QCAP) {
          __syncwarp();
          l0_flush(a, q0, n0, __ballot_sync(0xffffffffu, alive), coll_lanes, row_base);
          n0 = 0;
        }
      }
    };
    // the tables are read through explicit shared-window addresses held in registers (the generic pointers
    // were rebuilt from %cluster_ctaid in every trip of the loops below: three instructions and a stall per pair)
    const uint32_t gp_sh = (uint32_t)__cvta_generic_to_shared(s_gp), sg_sh = (uint32_t)__cvta_generic_to_shared(s_sg);
    const float *cl = cen + lane;
    auto centre = [&](int g) { const float *cc = cl + g * 96; return mk<float>(cc[0], cc[32], cc[64]); };
    int p = 0;
    {  // moving sphere against moving sphere (pairs sorted by ga: its centre is fetched once per run)
      int cached = -1;
      V3<float> cA = mk<float>(0.f, 0.f, 0.f);
VK_UNROLL(VK_L0_UNROLL)
      for (; p < a.gp_kind_end[0]; p++) {
        const uint4 g = lds128(gp_sh + p * 16);   // GroupPair: ga | gb << 16, first | n << 16 | kind << 24, lim, lim_in
        const int ga = (int)(g.x & 0xffffu), gb = (int)(g.x >> 16);
        if (ga != cached) { cached = ga; cA = centre(ga); }
        const V3<float> d = cA - centre(gb);
        const float d2 = dot(d, d);
        push(alive && d2 <= __uint_as_float(g.z), p, [&]() { return d2 < __uint_as_float(g.w); });
      }
    }
    {  // moving sphere against a world-fixed capsule
      int cached = -1;
      V3<float> cA = mk<float>(0.f, 0.f, 0.f);
VK_UNROLL(VK_L0_UNROLL)
      for (; p < a.gp_kind_end[1]; p++) {
        const uint4 g = lds128(gp_sh + p * 16);
        const int ga = (int)(g.x & 0xffffu), gb = (int)(g.x >> 16);
        if (ga != cached) { cached = ga; cA = centre(ga); }
        const uint4 s0 = lds128(sg_sh + gb * 32), s1 = lds128(sg_sh + gb * 32 + 16);   // StaticGroup: a, len | u, th
        const V3<float> e = cA - mk<float>(__uint_as_float(s0.x), __uint_as_float(s0.y), __uint_as_float(s0.z));
        const float len = __uint_as_float(s0.w);
        float sc;
        const float d2 = point_segment_d2(e, mk<float>(__uint_as_float(s1.x), __uint_as_float(s1.y), __uint_as_float(s1.z)), len, &sc);
        push(alive && d2 <= __uint_as_float(g.z), p, [&]() { return d2 < __uint_as_float(g.w) && fabsf(sc - 0.5f * len) <= __uint_as_float(s1.w); });
      }
    }
    for (; p < a.gp_kind_end[2]; p++) {  // moving sphere against a plane
      const GroupPair g = s_gp[p];
      const StaticGroup S = s_sg[g.gb];
      const V3<float> e = centre(g.ga) - mk<float>(S.a[0], S.a[1], S.a[2]);
      const float h = dot(e, mk<float>(S.u[0], S.u[1], S.u[2]));
      push(alive && h <= g.lim, p, [&]() { return h < g.lim_in; });
    }
    __syncwarp();
    if (n0) l0_flush(a, q0, n0, __ballot_sync(0xffffffffu, alive), coll_lanes, row_base);
    if (do_coll && !alive) {
      if ((a.mode == MODE_EDGES || a.mode == MODE_CHAINS) && !(a.flags & F_ROWMASK)) atomicMin(&a.first_bad[e_idx], e_k);
      else a.valid[row] = 0;
    }
  }
  if (lane == 0 && rows_total) atomicAdd(&a.counters[C_ROWS], (unsigned long long)rows_total);
}

// ---------------------------------------------------------------------------------------------------
struct MidLayout { size_t shapes, pairs, gpairs, member, q1, q2, stage, stage_cnt, total; };
__host__ __device__ inline MidLayout mid_layout(int nshape, int npair, int ngpair, int nmember) {
  MidLayout L;
  const int W = MID_THREADS / 32;
  size_t o = 0;
  L.shapes = o; o = align_up(o + (size_t)nshape * sizeof(Shape<float>), 128);
  L.pairs = o; o = align_up(o + (size_t)npair * sizeof(Pair), 128);
  L.gpairs = o; o = align_up(o + (size_t)ngpair * sizeof(GroupPair), 128);
  L.member = o; o = align_up(o + (size_t)nmember * sizeof(uint16_t), 128);
  L.q1 = o; o = align_up(o + (size_t)W * MID_Q1CAP * sizeof(uint32_t), 128);
  L.q2 = o; o = align_up(o + (size_t)W * MID_Q2CAP * sizeof(unsigned long long), 128);
  L.stage = o; o = align_up(o + (size_t)W * NBIN * MID_STAGE * sizeof(unsigned long long), 128);
  L.stage_cnt = o; o = align_up(o + (size_t)W * NBIN * sizeof(int), 128);
  L.total = o;
  return L;
}

// The OBB cull of the last `count` (<= 32) survivors queued in q2, then the bins (lane = survivor).
// Out of line: it is called from two places of mid_kernel and must not be duplicated there.
// one warp's staged items of bin b -> the global bin (a full bin: decided on the spot, see vk_split.cuh)
__device__ __forceinline__ void mid_flush_bin(const KArgs &a, const unsigned long long *stage, int c, int b) {
  const int lane = threadIdx.x & 31;
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(&a.counters[C_BIN + b], (unsigned long long)c);
  base = __shfl_sync(0xffffffffu, base, 0);
  for (int i = lane; i < c; i += 32) {
    const unsigned long long it = stage[b * MID_STAGE + i];
    if (base + i < a.bin_capv[b]) a.bin_items[a.bin_off[b] + base + i] = it;
    else broad_overflow_item(a, (int)(it >> 44), (long long)(it & ((1ull << 44) - 1ull)));
  }
  __syncwarp();
}

__device__ __noinline__ int mid_drain(const KArgs &a, const Shape<float> *s_shapes, const Pair *s_pairs,
                                      const unsigned long long *q2, int n2, int count, unsigned long long *stage, int *stage_cnt) {
  const int lane = threadIdx.x & 31;
  const bool use_obb = !(a.flags & F_NO_OBB);
  const float slack = 1e-4f;
  const unsigned below = (1u << lane) - 1u;
  int items = 0;
  int bin = -1;
  unsigned long long it = 0;
  __syncwarp();
  if (lane < count) {
    it = q2[n2 - count + lane];
    const int ip = (int)(it >> 44);
    const long long irow = (long long)(it & ((1ull << 44) - 1ull));
    const Pair pr = s_pairs[ip];
    const Shape<float> &A = s_shapes[pr.sa];
    const Shape<float> &B = s_shapes[pr.sb];
    bool keep = true;
    if (use_obb && pr.kind != PK_SEGSEG) {
      const Pose<float> PB = load_pose8(a.pose8, a.nslot, irow, B.slot);
      const Pose<float> PA = load_pose8(a.pose8, a.nslot, irow, A.slot);   // identity for a world-fixed shape (planes)
      // inner capsules overlap: a certain contact, the row is settled without a narrow phase (here, with
      // full lanes, not right after the capsule cull where half of them have already dropped out)
      if (inner_contact(pr, A, B, PA, PB)) { mark_contact(a, irow); keep = false; }
      else if (pr.flags & PF_OBB) {
        if (pr.kind == PK_PLANE) keep = !obb_above_plane(A, B, PB, pr.rsum - swept_radius(B) + slack);
        else keep = !obb_disjoint(A, B, relative_pose(PA, PB), pr.rsum - swept_radius(A) - swept_radius(B) + slack);
      }
    }
    if (keep) bin = item_bin(pr, A, B);
  }
  __syncwarp();
  // Survivors are staged per warp and per bin in shared memory; a bin's stage goes to the global bin
  // with ONE atomic when it is full (every warp of the grid appends to the same eight counters).
  unsigned todo = __ballot_sync(0xffffffffu, bin >= 0);
  while (todo) {
    const int b = __shfl_sync(0xffffffffu, bin, __ffs(todo) - 1);
    const unsigned m = __ballot_sync(0xffffffffu, bin == b);
    int c = stage_cnt[b];
    __syncwarp();
    if (c + __popc(m) > MID_STAGE) { mid_flush_bin(a, stage, c, b); c = 0; }
    if (bin == b) { stage[b * MID_STAGE + c + __popc(m & below)] = it; items += 1; }
    __syncwarp();
    if (lane == 0) stage_cnt[b] = c + __popc(m);
    __syncwarp();
    todo &= ~m;
  }
  return items;
}

__global__ void __launch_bounds__(MID_THREADS, MID_CTAS) mid_kernel(const __grid_constant__ KArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const MidLayout L = mid_layout(a.nshape, a.npair, a.ngpair, a.nmember);
  Shape<float> *s_shapes = reinterpret_cast<Shape<float> *>(smem + L.shapes);
  Pair *s_pairs = reinterpret_cast<Pair *>(smem + L.pairs);
  GroupPair *s_gp = reinterpret_cast<GroupPair *>(smem + L.gpairs);
  uint16_t *s_member = reinterpret_cast<uint16_t *>(smem + L.member);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t *q1 = reinterpret_cast<uint32_t *>(smem + L.q1) + (size_t)warp * MID_Q1CAP;
  unsigned long long *q2 = reinterpret_cast<unsigned long long *>(smem + L.q2) + (size_t)warp * MID_Q2CAP;
  unsigned long long *stage = reinterpret_cast<unsigned long long *>(smem + L.stage) + (size_t)warp * NBIN * MID_STAGE;
  int *stage_cnt = reinterpret_cast<int *>(smem + L.stage_cnt) + warp * NBIN;
  if (lane < NBIN) stage_cnt[lane] = 0;
  __shared__ uint64_t s_bar;

  const uint32_t bytes_s = (uint32_t)(a.nshape * sizeof(Shape<float>));
  const uint32_t bytes_p = (uint32_t)(a.npair * sizeof(Pair));
  const uint32_t bytes_g = (uint32_t)(a.ngpair * sizeof(GroupPair));
  const uint32_t bytes_m = (uint32_t)align_up((size_t)a.nmember * sizeof(uint16_t), 16);   // device array is padded
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&s_bar, bytes_s + bytes_p + bytes_g + bytes_m);
    if (bytes_s) bulk_g2s(s_shapes, a.shapes, bytes_s, &s_bar);
    if (bytes_p) bulk_g2s(s_pairs, a.pairs, bytes_p, &s_bar);
    if (bytes_g) bulk_g2s(s_gp, a.gpairs, bytes_g, &s_bar);
    if (bytes_m) bulk_g2s(s_member, a.gp_member, bytes_m, &s_bar);
  }
  mbar_wait(&s_bar, 0);
  __syncthreads();

  unsigned long long total = a.counters[C_L0];
  if (total > a.l0_cap) total = a.l0_cap;
  const bool use_obb = !(a.flags & F_NO_OBB);
  const float slack = 1e-4f;
  const unsigned below = (1u << lane) - 1u;
  long long items_total = 0;
  int n2 = 0;   // warp-uniform: capsule survivors waiting in q2

  // The level-0 list is consumed in batches of 32 entries (lane = entry), a warp claiming MID_CHUNK
  // entries per ticket.  Everything with a long latency is issued ONE BATCH AHEAD: the next batch's
  // entries are loaded (and the next chunk's ticket claimed) while the current batch is processed, and
  // the pose blocks of the next batch's rows are prefetched into L1 -- the kernel was bound by exactly
  // these dependent global loads (ncu: 35 % of the stall samples on the long scoreboard).
  const size_t pose_row_bytes = (size_t)a.nslot * 8 * sizeof(float);
  // entries per ticket: MID_CHUNK for large lists, fewer (a multiple of 32) when the list would otherwise
  // be shared out among a handful of warps (small batches run as long as their busiest warp)
  unsigned long long chunk_size = total / ((unsigned long long)gridDim.x * (MID_THREADS / 32)) / 32 * 32;
  chunk_size = chunk_size < 32 ? 32 : (chunk_size > MID_CHUNK ? MID_CHUNK : chunk_size);
  // The ticket of the next chunk is claimed one batch before it is needed and only read (the shuffle) when
  // it is: every warp of the grid hits the same word, the answer takes a microsecond to come back.
  unsigned long long ticket = 0;   // lane 0: the ticket claimed ahead
  bool ticket_out = false;
  auto claim_ahead = [&]() {
    if (lane == 0) ticket = atomicAdd(&a.counters[C_L0TICKET], 1ull);
    ticket_out = true;
  };
  auto claim = [&]() {
    if (!ticket_out) claim_ahead();
    ticket_out = false;
    return __shfl_sync(0xffffffffu, ticket, 0) * chunk_size;
  };
  auto load_entry = [&](unsigned long long pos) {   // entry of this lane in the batch starting at pos (or ~0)
    const unsigned long long ei = pos + lane;
    unsigned long long e = ~0ull;
    if (pos < total && ei < total) e = a.l0_items[ei];
    return e;
  };
  // the pose block of an entry's row -> L1.  Issued for the NEXT batch's entries once the current batch has been
  // expanded: by then their load has landed (issued right behind the load, the prefetch made the warp wait for
  // it: 7 % of the stall samples), and the capsule cull of the current batch still lies ahead of their use.
  auto prefetch_poses = [&](unsigned long long e) {
    if (e != ~0ull) {
      const char *pb = reinterpret_cast<const char *>(a.pose8) + (size_t)(e & ((1ull << 40) - 1ull)) * pose_row_bytes;
      for (size_t o = 0; o < pose_row_bytes; o += 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(pb + o));
    }
  };
  unsigned long long pos = claim(), chunk_end = pos + chunk_size;
  unsigned long long e_next = load_entry(pos);
  prefetch_poses(e_next);
  while (pos < total) {
    const unsigned long long e_cur = e_next;
    // the batch after this one: same chunk, or the first batch of a freshly claimed chunk
    unsigned long long pos_next = pos + 32;
    if (pos_next >= chunk_end || pos_next >= total) { pos_next = claim(); chunk_end = pos_next + chunk_size; }
    e_next = load_entry(pos_next);
    if (!ticket_out && pos_next < total && (pos_next + 32 * MID_AHEAD >= chunk_end || pos_next + 32 * MID_AHEAD >= total)) claim_ahead();
    {
      // ---- expand: every entry (row, group pair) -> its shape pairs ----------------------------------
      long long row = 0;
      int first = 0, n = 0;
      if (e_cur != ~0ull) {
        row = (long long)(e_cur & ((1ull << 40) - 1ull));
        const GroupPair g = s_gp[(int)(e_cur >> 40)];
        first = g.first; n = g.n;
      }
      int off = n;   // inclusive warp scan of n
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, off, o);
        if (lane >= o) off += v;
      }
      const int T = __shfl_sync(0xffffffffu, off, 31);
      off -= n;
      for (int i = 0; i < n; i++)
        if (off + i < MID_Q1CAP) q1[off + i] = ((uint32_t)lane << 16) | (uint32_t)s_member[first + i];
      __syncwarp();
      const int Tq = T < MID_Q1CAP ? T : MID_Q1CAP;
      // entries whose shape pairs did not fit (T > MID_Q1CAP: never with the shipped models, whose group
      // pairs have <= 18 members; 32 x 32 = 1024) fall back to whole-row fp64 re-evaluation
      if (T > MID_Q1CAP && off + n > MID_Q1CAP && n > 0) pipe_row_overflow(a, row);
      prefetch_poses(e_next);
      // ---- capsule cull, lane = shape pair of some entry -------------------------------------------------
#pragma unroll 1
      for (int s0 = 0; s0 < Tq; s0 += 32) {
        const int s = s0 + lane;
        const uint32_t subit = s < Tq ? q1[s] : 0u;
        const long long irow = __shfl_sync(0xffffffffu, row, (int)(subit >> 16));
        bool keep = false;
        const int ip = (int)(subit & 0xffffu);
        if (s < Tq) {
          const Pair pr = s_pairs[ip];
          keep = true;
          if (use_obb && pr.kind != PK_SEGSEG) {
            const Shape<float> &A = s_shapes[pr.sa];
            const Shape<float> &B = s_shapes[pr.sb];
            const Pose<float> PA = load_pose8(a.pose8, a.nslot, irow, A.slot);
            const Pose<float> PB = load_pose8(a.pose8, a.nslot, irow, B.slot);
            keep = !capsule_cull(pr, A, B, PA, PB, pr.rsum - swept_radius(A) - swept_radius(B) + slack);
          }
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) q2[n2 + __popc(m & below)] = (unsigned long long)irow | ((unsigned long long)ip << 44);
        n2 += __popc(m);
        __syncwarp();
        while (n2 >= 32) { items_total += mid_drain(a, s_shapes, s_pairs, q2, n2, 32, stage, stage_cnt); n2 -= 32; }   // keeps n2 + 32 <= MID_Q2CAP
      }
    }
    pos = pos_next;
  }
  while (n2 > 0) { const int c = n2 < 32 ? n2 : 32; items_total += mid_drain(a, s_shapes, s_pairs, q2, n2, c, stage, stage_cnt); n2 -= c; }
  __syncwarp();
  for (int b = 0; b < NBIN; b++) { const int c = stage_cnt[b]; if (c) mid_flush_bin(a, stage, c, b); }
  if (items_total) atomicAdd(&a.counters[C_ITEMS], (unsigned long long)items_total);
}

}  // namespace vk
