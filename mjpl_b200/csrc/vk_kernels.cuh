// vk_kernels.cuh -- sm_100a kernels of the configuration-validity path.
//
//   validity_kernel<TILE>  persistent CTAs; per tile of TILE rows:
//       P0  rows -> shared memory (dense rows: one cp.async.bulk (TMA) copy per tile, mbarrier
//           completion; edge waypoints / sweep rows are generated in registers)
//       P1  one lane per row: joint-limit mask (fp64 compare, reference semantics) + forward
//           kinematics; body poses -> per-CTA scratch (L2 resident), world bounding-sphere
//           centres of every moving geom -> shared memory (SoA, conflict free)
//       P2  one lane per row: broad phase over the static pair list (bounding spheres, then an
//           OBB separating-axis mid-phase); survivors are compacted into a shared-memory work
//           queue with warp ballots (one shared atomic per warp and pair)
//       P3  one lane per (row, pair) item: plane / segment-segment / GJK narrow phase with
//           certified three-way verdicts; first certain contact flags the row and later
//           items of that row are skipped (early exit)
//       P4  valid mask out; rows with an uncertain item and no certain contact go to the
//           fp64 list
//   recheck_kernel         fp64 re-evaluation of listed rows, one lane per row.
//   fk_kernel              mj_kinematics only (parity / debugging entry point).
//
// Model tables reach the CTA once (persistent kernel): hull vertices, shapes and pairs are
// copied global -> shared with cp.async.bulk; the kinematic tree sits in the kernel parameter
// (constant bank).  No tensor cores: nothing here is a dense contraction.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "vk_core.cuh"

namespace vk {

constexpr int MODE_DENSE = 0, MODE_EDGES = 1, MODE_SWEEP = 2, MODE_CHAINS = 3;
constexpr uint32_t F_LIMITS = 1u, F_COLLISION = 2u, F_NO_OBB = 4u, F_NO_RECHECK = 8u, F_LIMITS_OUTWARD = 16u;
// internal (pipeline, edges): the kernels keep one byte per WAYPOINT in `valid` (scratch) exactly as for dense rows
// -- a store per contact and an early exit per row -- and edge_rowmask_kernel folds it into first_bad afterwards.
// (Without it every contact of every waypoint searched the 100k-entry edge prefix: the UR5e edge benchmark ran at
// 8.9 ms in the pipeline against 4.4 ms in the single kernel.)
constexpr uint32_t F_ROWMASK = 1u << 30;
#ifndef VK_Q1
#define VK_Q1 8
#endif
#ifndef VK_Q2
#define VK_Q2 4
#endif
constexpr int Q1_PER_ROW = VK_Q1;   // sphere-survivor queue capacity per round = Q1_PER_ROW * TILE items
constexpr int Q2_PER_ROW = VK_Q2;   // narrow-phase queue capacity per round = Q2_PER_ROW * TILE items

struct KArgs {
  FkTables<float> fk;
  double jnt_lo[MAX_JNT], jnt_hi[MAX_JNT];
  int slot_shape_adr[MAX_BODY], slot_shape_num[MAX_BODY];
  const Shape<float> *shapes;
  const Vtx<float> *verts;
  const Pair *pairs;
  const uint16_t *adj_start;            // hull graphs: nvert+1 offsets into adj
  const uint8_t *adj;                   // local neighbour ids
  int nadj;
  int nshape, nmoving, nvert, npair, nslot;
  int nrounds;
  int round_start[MAX_ROUNDS + 1];
  int round_gjk[MAX_ROUNDS];
  int mode;
  // row sources
  const float *q; int ldq; long long n;                       // dense
  const unsigned long long *rows_ready;   // dense, optional: rows [0, *rows_ready) have landed (host copy in flight)
  const float *q0, *q1; const long long *edge_prefix; long long nedge; float step;  // edges
  unsigned long long seed; long long row0;                    // sweep
  const double *c0, *c1; double ceps;                         // chains: fp64 end points (n,nq) and step
  // outputs
  uint8_t *valid; int *first_bad; uint32_t flags;
  // per-handle scratch
  float *pose_scratch;                  // [grid][nslot*7][TILE]
  unsigned long long *counters;         // see C_* below
  long long *recheck_rows;              // rows to re-evaluate whole; capacity >= number of rows
  unsigned long long *recheck_items;    // (row | pair << 44) items to re-evaluate; capacity item_cap
  unsigned long long item_cap;
  // two-kernel pipeline (broad_kernel -> narrow_kernel): poses and (row, pair) items of the whole batch
  float *pose8;                         // [row][slot][8]: px py pz qw qx qy qz 0
  unsigned long long *bin_items;        // NBIN regions: row | pair << 44
  unsigned long long bin_off[NBIN], bin_capv[NBIN];   // start and capacity of each bin's region
  uint32_t *row_flags;                  // 1 byte per row (bit 0: queued for whole-row fp64 re-evaluation)
  // cull groups (vk_pipe.cuh)
  int slot_group_adr[MAX_BODY], slot_group_num[MAX_BODY];   // pose slot -> its moving groups
  float group_c[MAX_GROUP][3];          // bounding-sphere centre of a moving group, body frame
  const GroupPair *gpairs; const StaticGroup *sgroups; const uint16_t *gp_member;
  int ngpair, nsgroup, ngroup_moving, nmember;
  int gp_kind_end[3];
  unsigned long long *l0_items;         // level-0 survivors: row | group pair << 40
  unsigned long long l0_cap;
  // small launches (vk_row.cuh): row_kernel takes launches of at most rowk_max rows, validity_kernel the rest
  // (both are enqueued when only the device knows the row count; -1: no such dispatch)
  long long rowk_max;
  // support maps (vk_core.cuh): global memory, read through L1
  const uint32_t *smap_cells; const uint8_t *smap_ids;
};

// counters layout
constexpr int C_TICKET = 0, C_RECHECK = 1, C_RTICKET = 2, C_RITEMS = 3, C_RITICKET = 4;   // per launch
constexpr int C_BIN = 5, C_BTICKET = 13;   // per launch: fill and consumer ticket of each bin
constexpr int C_L0 = 21, C_L0TICKET = 22;  // per launch: fill of the level-0 list (vk_pipe.cuh) and its consumer ticket
constexpr int C_ITEMS = 23, C_OVERFLOW = 24, C_UNCERTAIN = 25, C_ROWS = 26, C_TRIPS = 27, C_HIST = 28, C_NCOUNTERS = 38;  // statistics
constexpr int C_PER_LAUNCH = 23;  // counters [0, C_PER_LAUNCH) are cleared before every launch

// ---------------------------------------------------------------------------- PTX helpers (sm_90+/sm_100a)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// bounded spin: a copy that never lands (bad size / alignment) traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); spins++) {
    if (spins > (1u << 26)) __trap();
  }
}
// 1-D bulk async copy global -> shared (TMA engine; SASS: UBLKCP); bytes % 16 == 0, 16B aligned
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// dynamic shared memory carve-up (host and device agree through this one function)
struct SmemLayout {
  size_t verts, shapes, pairs, adjs, adj, cen, qtile, queue1, queue2, hit, bars, total;
};
template <int TILE>
__host__ __device__ inline SmemLayout smem_layout(int nvert, int nshape, int npair, int nmoving, int nq, int nadj) {
  SmemLayout L;
  size_t o = 0;
  L.verts = o; o = align_up(o + (size_t)nvert * sizeof(Vtx<float>), 128);
  L.shapes = o; o = align_up(o + (size_t)nshape * sizeof(Shape<float>), 128);
  L.pairs = o; o = align_up(o + (size_t)npair * sizeof(Pair), 128);
  L.adjs = o; o = align_up(o + (size_t)(nvert + 1) * sizeof(uint16_t), 128);
  L.adj = o; o = align_up(o + (size_t)nadj, 128);
  L.cen = o; o = align_up(o + (size_t)(nmoving + 1) * 3 * TILE * sizeof(float), 128);  // + one block of static centres
  L.qtile = o; o = align_up(o + (size_t)TILE * nq * sizeof(float), 128);
  L.queue1 = o; o = align_up(o + (size_t)Q1_PER_ROW * TILE * sizeof(uint32_t), 128);
  L.queue2 = o; o = align_up(o + (size_t)Q2_PER_ROW * TILE * sizeof(uint32_t), 128);
  L.hit = o; o = align_up(o + (size_t)TILE * sizeof(uint32_t), 128);  // per row: uncertain-item count << 16 | first such pair
  L.bars = o; o = align_up(o + 64 + 8 * (TILE / 32), 128);  // [0] tables, then one row-load barrier per warp
  L.total = o;
  return L;
}

// row -> (edge, k) by binary search in the exclusive prefix sums of waypoint counts
__device__ __forceinline__ void edge_lookup(const long long *prefix, long long nedge, long long u, long long &e, int &k) {
  long long lo = 0, hi = nedge;  // prefix has nedge+1 entries; find e with prefix[e] <= u < prefix[e+1]
  while (hi - lo > 1) {
    long long mid = (lo + hi) >> 1;
    if (__ldg(prefix + mid) <= u) lo = mid; else hi = mid;
  }
  e = lo;
  k = (int)(u - __ldg(prefix + lo));
}

// waypoint k (0-based interior index) of edge e:  q0 + (k+1)*step*(q1-q0)/|q1-q0|
// reference: _valid_collision_interval / _step (src/mjpl/planning/utils.py:167-216)
template <typename T>
__device__ __forceinline__ void edge_row(const float *q0, const float *q1, int ldq, int nq, float step, long long e, int k, T *q) {
  double d2 = 0;
  for (int j = 0; j < nq; j++) {
    double d = (double)q1[e * ldq + j] - (double)q0[e * ldq + j];
    d2 += d * d;
  }
  double dist = sqrt(d2);
  double s = dist > 0 ? (double)(k + 1) * (double)step / dist : 0.0;
  for (int j = 0; j < nq; j++) {
    double x0 = q0[e * ldq + j], x1 = q1[e * ldq + j];
    q[j] = (T)(x0 + s * (x1 - x0));
  }
}

// step k (0-based) of the extend chain of query e: near + (k+1)*eps*(target-near)/|target-near|,
// and the target itself when that step covers the remaining distance.  Evaluated in fp64 exactly
// like chain_append_kernel stores it, so the checked row is the fp32 rounding of the stored node.
// reference: _constrained_extend / _step (src/mjpl/planning/utils.py:139-185)
// `lo` / `hi` (optional): joint limits; the return value is the reference's JointLimitConstraint
// answer for the fp64 chain point itself (joint_limit_constraint.py:19-20), decided BEFORE the point
// is rounded to fp32 for the collision check.
template <typename TO>
__device__ __forceinline__ bool chain_point(const double *c0, const double *c1, int nq, double eps, long long e, int k, TO *q,
                                            const double *lo = nullptr, const double *hi = nullptr) {
  double d2 = 0;
  for (int j = 0; j < nq; j++) {
    double d = c1[e * nq + j] - c0[e * nq + j];
    d2 += d * d;
  }
  const double dist = sqrt(d2);
  const double reach = (double)(k + 1) * eps;
  const double s = reach >= dist ? 1.0 : reach / dist;
  bool ok = true;
  for (int j = 0; j < nq; j++) {
    const double x = reach >= dist ? c1[e * nq + j] : c0[e * nq + j] + s * (c1[e * nq + j] - c0[e * nq + j]);
    if (lo) ok = ok && (x >= lo[j]) && (x <= hi[j]);
    q[j] = (TO)x;
  }
  return ok;
}

// JointLimitConstraint on the fp32 row (reference: np.all((q >= lower) & (q <= upper)) in fp64,
// joint_limit_constraint.py:19-20).  For a caller whose rows ARE fp32 this is the reference's answer
// for that row.  A caller whose rows were fp64 decides the limits itself on the fp64 values and asks
// for OUTWARD-rounded limits here, so that a row sitting exactly on a limit that fp32 cannot represent
// is not thrown away by the cast (the exact mask is AND-ed by the caller: mjpl_b200/engine.py).
__device__ __forceinline__ bool limits_ok(const float *q, int njnt, const double *lo, const double *hi, bool outward) {
  bool ok = true;
#pragma unroll 1
  for (int j = 0; j < njnt; j++) {
    const double x = (double)q[j];
    const double l = outward ? (double)__double2float_rd(lo[j]) : lo[j];
    const double h = outward ? (double)__double2float_ru(hi[j]) : hi[j];
    ok = ok && (x >= l) && (x <= h);
  }
  return ok;
}

__device__ __forceinline__ Pose<float> load_pose(const float *ps, int slot, int cfg, int tile) {
  Pose<float> P;
  if (slot < 0) { P.p = mk<float>(0, 0, 0); P.q.w = 1; P.q.x = P.q.y = P.q.z = 0; return P; }
  const float *b = ps + (size_t)slot * 7 * tile + cfg;
  P.p.x = b[0]; P.p.y = b[tile]; P.p.z = b[2 * tile];
  P.q.w = b[3 * tile]; P.q.x = b[4 * tile]; P.q.y = b[5 * tile]; P.q.z = b[6 * tile];
  return P;
}

// ---------------------------------------------------------------------------- main kernel
// Work-queue items: row-in-warp in the low 16 bits, pair index in the high 16 bits.
// Queues are WARP-LOCAL (each warp owns 32 rows and a private slice of shared memory), so a
// push is a ballot + popc with the fill count held in a warp-uniform register: no atomics, no
// CTA barriers between stages; warps drift apart and hide each other's latency.
// An item the fp32 path could not certify goes to the fp64 pass as (row, pair): all the other
// pairs of its row were certified by the fast path, so one pair is all there is to redo.  If the
// item list is full the row is flagged and re-evaluated whole (the row list holds every row).
__device__ __forceinline__ void note_uncertain(unsigned long long *counters, unsigned long long *items, unsigned long long cap,
                                               long long row, int pair, uint32_t *row_flag) {
  const unsigned long long slot = atomicAdd(&counters[C_RITEMS], 1ull);
  if (slot < cap) items[slot] = (unsigned long long)row | ((unsigned long long)pair << 44);
  else atomicOr(row_flag, 1u);
}

// The caller guarantees count + 32 <= cap before the push, so nothing can be dropped.
__device__ __forceinline__ void warp_push(bool want, uint32_t item, uint32_t *queue, int &count, int cap, int lane) {
  const unsigned m = __ballot_sync(0xffffffffu, want);
  if (m == 0) return;  // warp-uniform: the common case in the sphere stage
  const int idx = count + __popc(m & ((1u << lane) - 1u));
  count += __popc(m);
  if (want && idx < cap) queue[idx] = item;
}

// Lanes-per-item of the narrow phase: GRP consecutive lanes cooperate on one (row, pair) item
// (measured on B200, 1M Franka rows, same box: GRP=1 with the scan unrolled x4 3.34 ms, GRP=2 3.52 ms,
// GRP=4 3.8 ms, GRP=8 4.4 ms per step; hulls here have 41-152 vertices -- larger hulls favour GRP > 1).
// They split every support scan (hull vertices) GRP ways and butterfly-reduce the arg-max, then
// run the (cheap, identical) simplex update redundantly, so a warp works on 32/GRP items at
// once with every lane busy during the scans that dominate the cost.
#ifndef VK_GRP
#define VK_GRP 1
#endif
#ifndef VK_HILL
#define VK_HILL 0   // hill-climbing support queries on the hull graph: compiled out (slower than scanning on hulls of <= 152
#endif              // vertices, superseded by the support maps, and its code costs instruction-cache misses in every kernel)
#ifndef VK_SCAN_UNROLL
#define VK_SCAN_UNROLL 4
#endif
#ifndef VK_SMALL_TILES
#define VK_SMALL_TILES 1
#endif
#ifndef VK_MIN_TILE_ROWS
#define VK_MIN_TILE_ROWS 4   // smallest tile the adaptive choice may pick (4: 4,096 rows 154 us, 8: 166 us)
#endif
#ifndef VK_A_UNROLL
#define VK_A_UNROLL 1
#endif
#ifndef VK_SUSPEND
#define VK_SUSPEND 12   // leave stage C with at most this many unfinished items when the queue is empty
#endif
#ifndef VK_FLUSH_EARLY
#define VK_FLUSH_EARLY 1   // rounds whose narrow-phase items are processed at once (early exit); B200: 1 -> 3.15 ms, 2 -> 3.35, 0 -> 3.26
#endif
#ifndef VK_FLUSH_FILL
#define VK_FLUSH_FILL (3 * Q2CAP / 4)   // later rounds: run the narrow phase once this many items wait
#endif
#define VK_PRAGMA(x) _Pragma(#x)
#define VK_UNROLL(n) VK_PRAGMA(unroll n)
constexpr int GRP = VK_GRP;           // lanes per item of the large-batch instance of validity_kernel
#ifndef VK_GRP_SMALL
#define VK_GRP_SMALL 8
#endif
constexpr int GRP_SMALL = VK_GRP_SMALL;  // ... of the small-batch instance: an item's hull scans are split over 8 lanes, which
                                         // shortens the dependent chain a small launch waits for (planner extends)

// `warm` carries the last support vertex of this shape within one GJK run (-1 = cold start).
template <int G, bool HILL = (VK_HILL != 0)>
__device__ __forceinline__ V3<float> group_support(const Shape<float> &s, const Vtx<float> *__restrict__ verts,
                                                  const uint16_t *__restrict__ adjs, const uint8_t *__restrict__ adj,
                                                  V3<float> d, int gl, unsigned gmask, int &warm,
                                                  const uint32_t *__restrict__ smap_cells = nullptr, const uint8_t *__restrict__ smap_ids = nullptr) {
  if (s.kind == SK_CYL) return support_cyl(s, d);
  const Vtx<float> *__restrict__ v = verts + s.vadr;
  int bi = 0;
  if (smap_cells && s.map >= 0) {
    // support map: one cell, a handful of candidate vertices (every lane of a group reads the same ones)
    bi = support_mapped(v, s.nvert, smap_cells + s.map, smap_ids, d);
  } else if (HILL && s.graph) {
    // hill-climbing on the hull graph; the lanes of the group split each neighbour list
    const uint16_t *__restrict__ as = adjs + s.vadr;
    bi = warm >= 0 ? warm : hill_start(s, d);
    const Vtx<float> p0 = v[bi];
    float best = p0.x * d.x + p0.y * d.y + p0.z * d.z;
    for (;;) {
      int cj = bi;
      float cb = best;
      const int e1 = as[bi + 1];
      for (int e = as[bi] + gl; e < e1; e += G) {
        const int j = adj[e];
        const Vtx<float> p = v[j];
        const float t = p.x * d.x + p.y * d.y + p.z * d.z;
        if (t > cb) { cb = t; cj = j; }
      }
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(gmask, cb, o);
        const int oj = __shfl_xor_sync(gmask, cj, o);
        const bool take = (ob > cb) || (ob == cb && oj < cj);
        cb = take ? ob : cb;
        cj = take ? oj : cj;
      }
      if (cj == bi) break;
      bi = cj;
      best = cb;
    }
    warm = bi;
  } else {
    const int n = s.nvert;
    float best = -3.0e38f;
    VK_UNROLL(VK_SCAN_UNROLL)
    for (int i = gl; i < n; i += G) {
      const Vtx<float> p = v[i];
      const float t = p.x * d.x + p.y * d.y + p.z * d.z;
      const bool g = t > best;
      best = g ? t : best;
      bi = g ? i : bi;
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(gmask, best, o);
      const int oi = __shfl_xor_sync(gmask, bi, o);
      const bool take = (ob > best) || (ob == best && oi < bi);
      best = take ? ob : best;
      bi = take ? oi : bi;
    }
  }
  const Vtx<float> w = v[bi];
  return mk<float>(w.x, w.y, w.z);
}

template <int TILE, int G>
__global__ void __launch_bounds__(TILE) validity_kernel(const __grid_constant__ KArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int nq = a.fk.nq;
  const SmemLayout L = smem_layout<TILE>(a.nvert, a.nshape, a.npair, a.nmoving, nq, a.nadj);
  Vtx<float> *s_verts = reinterpret_cast<Vtx<float> *>(smem + L.verts);
  Shape<float> *s_shapes = reinterpret_cast<Shape<float> *>(smem + L.shapes);
  Pair *s_pairs = reinterpret_cast<Pair *>(smem + L.pairs);
  uint16_t *s_adjs = reinterpret_cast<uint16_t *>(smem + L.adjs);
  uint8_t *s_adj = smem + L.adj;
  float *s_cen = reinterpret_cast<float *>(smem + L.cen);
  float *s_q = reinterpret_cast<float *>(smem + L.qtile);
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem + L.bars);       // [0] tables, [8 + w] rows of warp w
  uint32_t *s_unc = reinterpret_cast<uint32_t *>(smem + L.hit);
  constexpr int Q1CAP = Q1_PER_ROW * 32, Q2CAP = Q2_PER_ROW * 32;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int wrow0 = tid & ~31;  // first row (within the tile) owned by this warp
  if (a.rowk_max >= 0 && (a.mode == MODE_EDGES || a.mode == MODE_CHAINS) && a.edge_prefix[a.nedge] <= a.rowk_max) return;   // row_kernel's launch
  uint32_t *q1 = reinterpret_cast<uint32_t *>(smem + L.queue1) + (tid >> 5) * Q1CAP;  // sphere-cull survivors
  uint32_t *q2 = reinterpret_cast<uint32_t *>(smem + L.queue2) + (tid >> 5) * Q2CAP;  // narrow-phase items
  float *pose = a.pose_scratch + (size_t)blockIdx.x * a.nslot * 7 * TILE;

  // ---- one-time: model tables -> shared memory through the bulk-copy engine ---------------------
  const uint32_t bytes_v = (uint32_t)(a.nvert * sizeof(Vtx<float>));
  const uint32_t bytes_s = (uint32_t)(a.nshape * sizeof(Shape<float>));
  const uint32_t bytes_p = (uint32_t)(a.npair * sizeof(Pair));
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    fence_barrier_init();
  }
  __syncthreads();
  const uint32_t bytes_as = (uint32_t)align_up((size_t)(a.nvert + 1) * sizeof(uint16_t), 16);  // device arrays are padded
  const uint32_t bytes_a = (uint32_t)align_up((size_t)a.nadj, 16);
  if (tid == 0) {
    mbar_expect_tx(&s_bar[0], bytes_v + bytes_s + bytes_p + bytes_as + bytes_a);
    if (bytes_v) bulk_g2s(s_verts, a.verts, bytes_v, &s_bar[0]);
    if (bytes_s) bulk_g2s(s_shapes, a.shapes, bytes_s, &s_bar[0]);
    if (bytes_p) bulk_g2s(s_pairs, a.pairs, bytes_p, &s_bar[0]);
    bulk_g2s(s_adjs, a.adj_start, bytes_as, &s_bar[0]);
    if (bytes_a) bulk_g2s(s_adj, a.adj, bytes_a, &s_bar[0]);
  }
  mbar_wait(&s_bar[0], 0);
  // static (world-fixed) shapes: their centres never change; they live in one extra [xyz][TILE]
  // block of the centre array, indexed by (shape - nmoving), so the sphere stage addresses moving
  // and static centres the same way
  for (int k = tid; k < a.nshape - a.nmoving; k += TILE) {
    const Shape<float> &S = s_shapes[a.nmoving + k];
    float *cc = s_cen + (size_t)a.nmoving * 3 * TILE + k;
    cc[0] = S.bc[0]; cc[TILE] = S.bc[1]; cc[2 * TILE] = S.bc[2];
  }
  __syncthreads();

  // total number of rows (edges: read from the device-side prefix sums)
  long long nrows = a.n;
  if (a.mode == MODE_EDGES || a.mode == MODE_CHAINS) nrows = a.edge_prefix[a.nedge];
  // A tile = the rows one warp owns: 32, or fewer when the batch is too small to give every
  // resident warp a full tile (the planner's extends: a few thousand rows).  Fewer rows per warp
  // means fewer items per warp in the lane = item stages, i.e. a shorter critical path, at no cost
  // while warps would otherwise sit idle.
  const long long warps_total = (long long)gridDim.x * (TILE / 32);
  // (B200, Franka rows, same box: 4,096 rows 224 -> 164 us, 16,384 rows 258 -> 189 us with 8 / 16
  // rows per warp; batched bi-RRT 1,475 -> 1,623 plans/s; 65k rows and more unchanged.)
  const int rpt = !VK_SMALL_TILES ? 32
                  : (nrows <= warps_total * VK_MIN_TILE_ROWS ? VK_MIN_TILE_ROWS : (nrows <= warps_total * 8 ? 8 : (nrows <= warps_total * 16 ? 16 : 32)));
  const long long ntiles = (nrows + rpt - 1) / rpt;
  uint32_t row_parity = 0;
  const bool dense_bulk = (a.mode == MODE_DENSE) && (a.ldq == nq) && ((reinterpret_cast<uintptr_t>(a.q) & 15) == 0);
  long long items_total = 0, rows_total = 0;
#ifdef VK_STATS
  long long st_trips = 0, st_busy = 0, st_flushes = 0, st_bbatches = 0, st_bbusy = 0;
  long long st_hist[4] = {0, 0, 0, 0}, st_histb[4] = {0, 0, 0, 0};
#endif
  const bool use_obb = !(a.flags & F_NO_OBB);
  const float slack = 1e-4f;
  uint64_t *wbar = s_bar + 8 + (tid >> 5);       // this warp's row-load barrier
  float *wq = s_q + (size_t)wrow0 * nq;          // this warp's 32 rows
  if (lane == 0) mbar_init(wbar, 1);
  fence_barrier_init();
  __syncwarp();

  for (;;) {
    // ---- tile ticket: warps are fully independent from here on (no CTA-wide barrier) ------------
    long long tile = 0;
    if (lane == 0) tile = (long long)atomicAdd(&a.counters[C_TICKET], 1ull);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile >= ntiles) break;
    const long long row_base = tile * rpt;
    const int rows_here = (int)((nrows - row_base) < rpt ? (nrows - row_base) : rpt);
    const long long row = row_base + lane;
    const bool active = lane < rows_here;
    rows_total += rows_here;

    // ---- P0: the warp's rows -> shared (one TMA bulk copy per warp tile) -----------------------------
    if (a.mode == MODE_DENSE) {
      if (a.rows_ready) {
        // rows are still being copied host -> device on another stream (mjb_check_configs_host): the
        // copy publishes its progress after every chunk; tiles are handed out in row order, so a
        // warp only ever waits for the chunk that holds its own rows.  Bounded: a copy that never
        // lands traps instead of hanging the GPU.
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

    // ---- P1: limits + FK, one lane per row ---------------------------------------------------------------
    float *q = s_q + tid * nq;  // this lane's row (stride nq words: conflict free for odd nq)
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
    }
    const bool do_coll = active && lim_ok && (a.flags & F_COLLISION);
    if (do_coll) {
      Pose<float> prev;
      prev.p = mk<float>(0, 0, 0); prev.q.w = 1; prev.q.x = prev.q.y = prev.q.z = 0;
      int prev_slot = -1;
#pragma unroll 1
      for (int s = 0; s < a.nslot; s++) {
        const int ps = a.fk.body_parent[s];
        Pose<float> P = (ps == prev_slot) ? prev : load_pose(pose, ps, tid, TILE);
        Pose<float> B = fk_body(a.fk, s, P, q);
        prev = B; prev_slot = s;
        float *b = pose + (size_t)s * 7 * TILE + tid;
        b[0] = B.p.x; b[TILE] = B.p.y; b[2 * TILE] = B.p.z;
        b[3 * TILE] = B.q.w; b[4 * TILE] = B.q.x; b[5 * TILE] = B.q.y; b[6 * TILE] = B.q.z;
        const int sa = a.slot_shape_adr[s], sn = a.slot_shape_num[s];
        for (int k = 0; k < sn; k++) {
          const Shape<float> &S = s_shapes[sa + k];
          V3<float> c = B.p + qrot(B.q, mk<float>(S.bc[0], S.bc[1], S.bc[2]));
          float *cc = s_cen + (size_t)(sa + k) * 3 * TILE + tid;
          cc[0] = c.x; cc[TILE] = c.y; cc[2 * TILE] = c.z;
        }
      }
    }
    __syncwarp();  // poses (global scratch) and centres of this warp's rows are visible to its lanes

    // ---- rounds over the (contact-likelihood ordered) pair list, all warp-local --------------------
    // Stage A fills q1 (sphere-cull survivors), stage B drains q1 into q2 (mid-phase survivors),
    // stage C consumes q2 (narrow phase).  Every stage stops while the next queue still has room
    // for a full warp of pushes, so NO item is ever dropped, whatever the rows look like
    // (correlated rows of a planner chain fill the queues far beyond the calibrated average).
    unsigned hit_mask = 0;  // warp-uniform: bit r = row r of this warp has a certain contact
    unsigned unc_mask = 0;  //               bit r = row r has an uncertain item
    int n1 = 0, b_pos = 0;  // q1 fill and the next q1 index stage B will take (warp-uniform)
    int n2 = 0;             // q2 fill (warp-uniform)
    const unsigned coll_mask = __ballot_sync(0xffffffffu, do_coll);
    // narrow-phase state of this lane's current GJK item.  It outlives one pass of stage C: when
    // the item queue is empty and only a few lanes still iterate (the long tail of slow items),
    // the warp goes back to produce more items and the unfinished ones resume with full lanes.
    GjkState<float> gs;
    Rel<float> rel;
    const Shape<float> *SA = s_shapes, *SB = s_shapes;
    float R = 0.f;
    int r = 0;
    int pidx = 0;          // pair index of the current item (reported if it ends uncertain)
    int wa = -1, wb = -1;  // warm-start vertices of the current item's two shapes
    bool have = false;
    s_unc[tid] = 0;
    const int gl = lane & (G - 1);
    const unsigned gmask = ((1u << G) - 1u) << (lane & ~(G - 1));
#pragma unroll 1
    for (int rd = 0; rd < a.nrounds && (coll_mask & ~hit_mask); rd++) {
      int p = a.round_start[rd];
      const int p1 = a.round_start[rd + 1];
      // early rounds (most likely contacts) are flushed right away so that hit rows stop
      // generating work; later rounds accumulate items for better lane balance
      const bool flush_round = (rd < VK_FLUSH_EARLY) || (rd + 1 == a.nrounds);
#pragma unroll 1
      for (;;) {
        // A: sphere cull, lane = row.  Rows that already have a certain contact drop out.
        if (b_pos >= n1) {
          n1 = 0;
          b_pos = 0;
          const bool live = do_coll && !((hit_mask >> lane) & 1u);
          int cached_sa = -1;          // pairs of a round are sorted by shape A: its centre is
          V3<float> cA = mk<float>(0.f, 0.f, 0.f);  // fetched once per run of pairs (warp-uniform test)
          const int stat_off = a.nmoving * 3 * TILE - a.nmoving;  // static centre k sits at stat_off + shape index
          VK_UNROLL(VK_A_UNROLL)
          for (; p < p1 && n1 + 32 <= Q1CAP; p++) {
            const Pair pr = s_pairs[p];
            if ((int)pr.sa != cached_sa) {
              cached_sa = pr.sa;
              const float *cc = s_cen + ((pr.flags & PF_A_STATIC) ? stat_off + (int)pr.sa : (int)pr.sa * 3 * TILE + tid);
              cA = mk<float>(cc[0], cc[TILE], cc[2 * TILE]);
            }
            const float *cb = s_cen + ((pr.flags & PF_B_STATIC) ? stat_off + (int)pr.sb : (int)pr.sb * 3 * TILE + tid);
            const V3<float> cB = mk<float>(cb[0], cb[TILE], cb[2 * TILE]);
            const float lim = pr.bsum + slack;
            bool survive;
            if (pr.kind == PK_PLANE) {
              const Shape<float> &A = s_shapes[pr.sa];
              const float d = A.ax[0] * (cB.x - A.c[0]) + A.ax[1] * (cB.y - A.c[1]) + A.ax[2] * (cB.z - A.c[2]);
              survive = live && d <= lim;
            } else {
              const V3<float> d = cA - cB;
              survive = live && dot(d, d) <= lim * lim;
            }
            warp_push(survive, (uint32_t)lane | ((uint32_t)p << 16), q1, n1, Q1CAP, lane);
          }
          __syncwarp();
        }
        // B: mid-phase cull (OBB-OBB separating axes / OBB above plane), lane = surviving (row, pair)
#pragma unroll 1
        for (; b_pos < n1 && n2 + 32 <= Q2CAP; b_pos += 32) {
          const int i = b_pos + lane;
          bool keep = false;
          uint32_t it = 0;
          unsigned certain = 0;
          if (i < n1) {
            it = q1[i];
            const int r = it & 0xffff;
            const Pair pr = s_pairs[it >> 16];
            keep = !((hit_mask >> r) & 1u);
            if (keep && use_obb) {
              const Shape<float> &A = s_shapes[pr.sa];
              const Shape<float> &B = s_shapes[pr.sb];
              Pose<float> PA = load_pose(pose, A.slot, wrow0 + r, TILE);
              Pose<float> PB = load_pose(pose, B.slot, wrow0 + r, TILE);
              keep = !midphase_cull(pr, A, B, PA, PB, pr.rsum - swept_radius(A) - swept_radius(B), slack);
              // inner capsules overlap: a certain contact, the row is settled without a narrow phase
              if (keep && pr.kind != PK_SEGSEG && inner_contact(pr, A, B, PA, PB)) { certain = 1u << r; keep = false; }
            }
          }
          hit_mask |= __reduce_or_sync(0xffffffffu, certain);
          warp_push(keep, it, q2, n2, Q2CAP, lane);
        }
        __syncwarp();
        const bool a_done = (p >= p1) && (b_pos >= n1);
        const bool run_c = (b_pos < n1) /* q2 has no room */ || (a_done && (flush_round || n2 >= VK_FLUSH_FILL));
        if (run_c) {
          // C: narrow phase with persistent lane groups: every trip runs ONE GJK iteration per
          // group; a group whose item is decided fetches the next one, so lanes do not idle
          // while the slowest item of the warp converges.  Plane and segment items are decided
          // in the fetch step.
          items_total += n2;
#ifdef VK_STATS
          st_flushes++;
#endif
          int head = 0;  // warp-uniform
          const bool drain = a_done && (rd + 1 == a.nrounds);
          if (have && ((hit_mask >> r) & 1u)) have = false;  // decided while the item was parked
#pragma unroll 1
          for (;;) {
            const unsigned need = __ballot_sync(0xffffffffu, !have);   // group-uniform bits
            if (head >= n2 && (need == 0xffffffffu || (!drain && __popc(~need) <= VK_SUSPEND * G))) break;
            unsigned hb = 0, ub = 0;
            if (!have) {
              const int i = head + __popc(need & ((1u << (lane & ~(G - 1))) - 1u)) / G;
              if (i < n2) {
                const uint32_t it = q2[i];
                r = it & 0xffff;
                if (!((hit_mask >> r) & 1u)) {
                  pidx = (int)(it >> 16);
                  const Pair pr = s_pairs[pidx];
                  SA = s_shapes + pr.sa;
                  SB = s_shapes + pr.sb;
                  R = pr.rsum;
                  Pose<float> PA = load_pose(pose, SA->slot, wrow0 + r, TILE);
                  Pose<float> PB = load_pose(pose, SB->slot, wrow0 + r, TILE);
                  if (pr.kind == PK_GJK) {
                    rel = relative_pose(PA, PB);
                    gjk_init(gs, *SA, *SB, rel);
                    wa = wb = -1;
                    have = true;
                  } else {
                    int v;
                    if (pr.kind == PK_PLANE) {
                      const Shape<float> &Bs = *SB;
                      int cold = -1;
                      v = plane_classify(*SA, Bs, PB, R, [&](V3<float> d) { return group_support<G>(Bs, s_verts, s_adjs, s_adj, d, gl, gmask, cold); });
                    } else {
                      v = segseg_item(*SA, *SB, s_verts, PA, PB, R);
                    }
                    if (v == V_PEN) hb = 1u << r;
                    else if (v == V_UNC) { ub = 1u << r; if (!(a.flags & F_NO_RECHECK) && gl == 0) note_uncertain(a.counters, a.recheck_items, a.item_cap, row_base + r, pidx, s_unc + wrow0 + r); }
                  }
                }
              }
            }
            head += __popc(need) / G;
#ifdef VK_STATS
            { const int nb = __popc(__ballot_sync(0xffffffffu, have)) / G; st_trips++; st_busy += nb;
              const int bin = nb <= 4 ? 0 : (nb <= 8 ? 1 : (nb <= 16 ? 2 : 3)); st_hist[bin]++; st_histb[bin] += nb; }
#endif
            if (have) {
              const Shape<float> &As = *SA, &Bs = *SB;
              const int v = gjk_step_impl(
                  gs, rel, R, [&](V3<float> d) { return group_support<G>(As, s_verts, s_adjs, s_adj, d, gl, gmask, wa); },
                  [&](V3<float> d) { return group_support<G>(Bs, s_verts, s_adjs, s_adj, d, gl, gmask, wb); });
              if (v >= 0) {
                if (v == V_PEN) hb = 1u << r;
                else if (v == V_UNC) { ub = 1u << r; if (!(a.flags & F_NO_RECHECK) && gl == 0) note_uncertain(a.counters, a.recheck_items, a.item_cap, row_base + r, pidx, s_unc + wrow0 + r); }
                have = false;
              }
            }
            hit_mask |= __reduce_or_sync(0xffffffffu, hb);
            unc_mask |= __reduce_or_sync(0xffffffffu, ub);
            if (have && ((hit_mask >> r) & 1u)) have = false;  // another pair already decided this row
          }
          n2 = 0;
          __syncwarp();
        }
        if (a_done) break;
      }
    }

    // ---- P4: results, lane = row ---------------------------------------------------------------------------------
    if (active) {
      const bool hit = (hit_mask >> lane) & 1u;
      const bool unc = (unc_mask >> lane) & 1u;
      bool ok = lim_ok && !hit;
      bool pending = lim_ok && !hit && unc;
      if (pending) atomicAdd(&a.counters[C_UNCERTAIN], 1ull);
      if (pending && !(a.flags & F_NO_RECHECK)) {
        if (s_unc[tid]) {  // its items did not fit the item list
          unsigned long long slot = atomicAdd(&a.counters[C_RECHECK], 1ull);
          a.recheck_rows[slot] = row;
        }
      }
      if (a.mode == MODE_EDGES || a.mode == MODE_CHAINS) {
        if (!ok && !pending) atomicMin(&a.first_bad[e_idx], e_k);
        else if (pending && (a.flags & F_NO_RECHECK)) atomicMin(&a.first_bad[e_idx], e_k);
      } else {
        a.valid[row] = pending ? (uint8_t)((a.flags & F_NO_RECHECK) ? 2 : 1) : (uint8_t)(ok ? 1 : 0);
      }
    }
  }
  if (lane == 0 && items_total) atomicAdd(&a.counters[C_ITEMS], (unsigned long long)items_total);
  if (lane == 0 && rows_total) atomicAdd(&a.counters[C_ROWS], (unsigned long long)rows_total);
#ifdef VK_STATS
  if (lane == 0) { atomicAdd(&a.counters[C_TRIPS], (unsigned long long)st_trips); atomicAdd(&a.counters[C_OVERFLOW], (unsigned long long)st_busy); atomicAdd(&a.counters[C_HIST + 8], (unsigned long long)st_flushes);
    for (int b = 0; b < 4; b++) { atomicAdd(&a.counters[C_HIST + b], (unsigned long long)st_hist[b]); atomicAdd(&a.counters[C_HIST + 4 + b], (unsigned long long)st_histb[b]); } }
#endif
}

// ---------------------------------------------------------------------------- fp64 re-evaluation
struct RArgs {
  const FkTables<double> *fk;
  const Shape<double> *shapes;
  const Vtx<double> *verts;
  const Pair *pairs;
  const double *pair_rsum;   // fp64 copies of Pair::rsum / bsum
  const double *pair_bsum;
  int npair, nslot;
  int mode;
  const float *q; int ldq;
  const float *q0, *q1; const long long *edge_prefix; long long nedge; float step;
  const double *c0, *c1; double ceps;
  unsigned long long seed; long long row0;
  uint8_t *valid; int *first_bad;
  unsigned long long *counters;
  const long long *recheck_rows;
  const unsigned long long *recheck_items;
  unsigned long long item_cap;
};

// support point of a shape with the 32 lanes of a warp splitting the vertex scan (fp64); every
// lane returns the same vertex (ties go to the lower index)
__device__ __forceinline__ V3<double> warp_support64(const Shape<double> &s, const Vtx<double> *__restrict__ verts, V3<double> d,
                                                    int lane) {
  if (s.kind == SK_CYL) return support_cyl(s, d);
  const Vtx<double> *__restrict__ v = verts + s.vadr;
  double best = -1.0e300;
  int bi = 0;
  for (int i = lane; i < s.nvert; i += 32) {
    const Vtx<double> p = v[i];
    const double t = p.x * d.x + p.y * d.y + p.z * d.z;
    if (t > best) { best = t; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    const bool take = (ob > best) || (ob == best && oi < bi);
    best = take ? ob : best;
    bi = take ? oi : bi;
  }
  const Vtx<double> p = v[bi];
  return mk<double>(p.x, p.y, p.z);
}

// the fp64 row the fast path saw (as fp32) and its place in an edge / chain, for any row source
__device__ __forceinline__ void recheck_row(const RArgs &a, const FkTables<double> &fk, long long row, double *q, long long &e_idx,
                                            int &e_k) {
  e_idx = 0; e_k = 0;
  if (a.mode == MODE_DENSE) {
    for (int j = 0; j < fk.nq; j++) q[j] = (double)a.q[row * a.ldq + j];
  } else if (a.mode == MODE_EDGES) {
    edge_lookup(a.edge_prefix, a.nedge, row, e_idx, e_k);
    float qf[MAX_JNT];
    edge_row<float>(a.q0, a.q1, a.ldq, fk.nq, a.step, e_idx, e_k, qf);
    for (int j = 0; j < fk.nq; j++) q[j] = (double)qf[j];
  } else if (a.mode == MODE_CHAINS) {
    edge_lookup(a.edge_prefix, a.nedge, row, e_idx, e_k);
    chain_point<double>(a.c0, a.c1, fk.nq, a.ceps, e_idx, e_k, q);
    for (int j = 0; j < fk.nq; j++) q[j] = (double)(float)q[j];
  } else {
    for (int j = 0; j < fk.nq; j++)
      q[j] = (double)sweep_value(a.seed, (uint64_t)(a.row0 + row), (uint32_t)j, (float)fk.jnt_lo[j], (float)fk.jnt_hi[j]);
  }
}

// fp64 re-evaluation, one WARP per unit of work, two passes:
//  1. uncertain ITEMS (row, pair): every lane rebuilds the row and its fp64 poses (cheap,
//     redundant).  A convex pair near touching needs many GJK iterations in fp64, so the whole
//     warp runs ONE GJK instance with each support scan split 32 ways; other kinds are closed
//     form.  "Uncertain" in fp64 means touching to within rounding: MuJoCo reports a contact for
//     distance <= margin, so it counts as one.
//  2. whole ROWS (only when the item list overflowed): the 32 lanes split the static pair list.
// Both only ever turn a tentatively valid row invalid.
#ifndef VK_RECHECK_CTAS
#define VK_RECHECK_CTAS 2
#endif
__global__ void __launch_bounds__(128, VK_RECHECK_CTAS) recheck_kernel(const RArgs a) {
  const FkTables<double> &fk = *a.fk;
  const int lane = threadIdx.x & 31;
  // (q and P as per-warp shared arrays instead of per-thread local ones: the fp64 pass of a 1M-row launch is as
  // long either way, 0.053 ms, and the planner's iterations got slower and jittery -- measured, dropped)
  double q[MAX_JNT];
  Pose<double> P[MAX_BODY];
  Pose<double> ident; ident.p = mk<double>(0, 0, 0); ident.q.w = 1; ident.q.x = ident.q.y = ident.q.z = 0;
  long long e_idx;
  int e_k;
  unsigned long long nitems = a.counters[C_RITEMS];
  if (nitems > a.item_cap) nitems = a.item_cap;
  for (;;) {
    unsigned long long t = 0;
    if (lane == 0) t = atomicAdd(&a.counters[C_RITICKET], 1ull);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= nitems) break;
    const unsigned long long entry = a.recheck_items[t];
    const long long row = (long long)(entry & ((1ull << 44) - 1ull));
    const int p = (int)(entry >> 44);
    recheck_row(a, fk, row, q, e_idx, e_k);
    const bool edges = a.mode == MODE_EDGES || a.mode == MODE_CHAINS;
    if (!edges && a.valid[row] == 0) continue;  // the row already has a certain contact
    for (int s = 0; s < a.nslot; s++) {
      int ps = fk.body_parent[s];
      P[s] = fk_body(fk, s, ps < 0 ? ident : P[ps], q);
    }
    const Pair pr = a.pairs[p];
    const Shape<double> &A = a.shapes[pr.sa];
    const Shape<double> &B = a.shapes[pr.sb];
    const Pose<double> &PA = A.slot < 0 ? ident : P[A.slot];
    const Pose<double> &PB = B.slot < 0 ? ident : P[B.slot];
    int v;
    if (pr.kind == PK_GJK) {
      const Rel<double> rel = relative_pose(PA, PB);
      GjkState<double> gs;
      gjk_init(gs, A, B, rel);
      do {
        v = gjk_step_impl(gs, rel, a.pair_rsum[p], [&](V3<double> d) { return warp_support64(A, a.verts, d, lane); },
                          [&](V3<double> d) { return warp_support64(B, a.verts, d, lane); });
      } while (v < 0);
    } else {
      v = narrow_item<double>(pr.kind, A, B, a.verts, PA, PB, a.pair_rsum[p]);
    }
    if (lane == 0 && v != V_SEP) {
      if (edges) atomicMin(&a.first_bad[e_idx], e_k);
      else a.valid[row] = 0;
    }
  }
  const unsigned long long total = a.counters[C_RECHECK];
  for (;;) {
    unsigned long long t = 0;
    if (lane == 0) t = atomicAdd(&a.counters[C_RTICKET], 1ull);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= total) break;
    const long long row = a.recheck_rows[t];
    recheck_row(a, fk, row, q, e_idx, e_k);
    for (int s = 0; s < a.nslot; s++) {
      int ps = fk.body_parent[s];
      P[s] = fk_body(fk, s, ps < 0 ? ident : P[ps], q);
    }
    bool contact = false;
    for (int base = 0; base < a.npair; base += 32) {
      const int p = base + lane;
      if (p < a.npair) {
        const Pair pr = a.pairs[p];
        const Shape<double> &A = a.shapes[pr.sa];
        const Shape<double> &B = a.shapes[pr.sb];
        const Pose<double> &PA = A.slot < 0 ? ident : P[A.slot];
        const Pose<double> &PB = B.slot < 0 ? ident : P[B.slot];
        V3<double> cB = PB.p + qrot(PB.q, mk<double>(B.bc[0], B.bc[1], B.bc[2]));
        const double bsum = a.pair_bsum[p];
        bool near;
        if (pr.kind == PK_PLANE) {
          double d = A.ax[0] * (cB.x - A.c[0]) + A.ax[1] * (cB.y - A.c[1]) + A.ax[2] * (cB.z - A.c[2]);
          near = d <= bsum + 1e-6;
        } else {
          V3<double> cA = PA.p + qrot(PA.q, mk<double>(A.bc[0], A.bc[1], A.bc[2]));
          V3<double> d = cA - cB;
          near = dot(d, d) <= (bsum + 1e-6) * (bsum + 1e-6);
        }
        if (near && narrow_item<double>(pr.kind, A, B, a.verts, PA, PB, a.pair_rsum[p]) != V_SEP) contact = true;
      }
      if (__any_sync(0xffffffffu, contact)) { contact = true; break; }
    }
    if (lane == 0 && contact) {
      if (a.mode == MODE_EDGES || a.mode == MODE_CHAINS) atomicMin(&a.first_bad[e_idx], e_k);
      else a.valid[row] = 0;
    }
  }
}

// ---------------------------------------------------------------------------- FK only
struct FArgs {
  FkTables<float> fk;
  const float *q; int ldq; long long n;
  int nbody_all;
  const int *body_slot;      // [nbody_all] slot or -1
  const float *static_pose;  // [nbody_all][7]
  float *xpos, *xquat;
};

__global__ void __launch_bounds__(128) fk_kernel(const __grid_constant__ FArgs a) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= a.n) return;
  float q[MAX_JNT];
  for (int j = 0; j < a.fk.nq; j++) q[j] = a.q[row * a.ldq + j];
  Pose<float> P[MAX_BODY];
  Pose<float> ident; ident.p = mk<float>(0, 0, 0); ident.q.w = 1; ident.q.x = ident.q.y = ident.q.z = 0;
  for (int s = 0; s < a.fk.nbody; s++) {
    int ps = a.fk.body_parent[s];
    P[s] = fk_body(a.fk, s, ps < 0 ? ident : P[ps], q);
  }
  for (int b = 0; b < a.nbody_all; b++) {
    int s = a.body_slot[b];
    float *xp = a.xpos + (row * a.nbody_all + b) * 3, *xq = a.xquat + (row * a.nbody_all + b) * 4;
    if (s >= 0) {
      xp[0] = P[s].p.x; xp[1] = P[s].p.y; xp[2] = P[s].p.z;
      xq[0] = P[s].q.w; xq[1] = P[s].q.x; xq[2] = P[s].q.y; xq[3] = P[s].q.z;
    } else {
      const float *sp = a.static_pose + 7 * b;
      xp[0] = sp[0]; xp[1] = sp[1]; xp[2] = sp[2];
      xq[0] = sp[3]; xq[1] = sp[4]; xq[2] = sp[5]; xq[3] = sp[6];
    }
  }
}

// ---------------------------------------------------------------------------- small helpers
// waypoint counts per edge: K = max(0, ceil(|q1-q0|/step) - 1)
__global__ void edge_count_kernel(const float *q0, const float *q1, long long ne, int nq, int ldq, float step,
                                  long long *count, int *first_bad) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  double d2 = 0;
  for (int j = 0; j < nq; j++) {
    double d = (double)q1[e * ldq + j] - (double)q0[e * ldq + j];
    d2 += d * d;
  }
  double m = sqrt(d2) / (double)step;
  long long k = (long long)ceil(m) - 1;
  count[e] = k > 0 ? k : 0;
  first_bad[e] = 0x7fffffff;
}

// one warp per edge: the first waypoint whose byte of the row mask is 0 -> first_bad (min with what is there)
__global__ void __launch_bounds__(256) edge_rowmask_kernel(long long ne, const long long *prefix, const uint8_t *rowmask, int *first_bad) {
  const long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (e >= ne) return;
  const long long r0 = prefix[e], K = prefix[e + 1] - r0;
  for (long long k0 = 0; k0 < K; k0 += 32) {
    const long long k = k0 + lane;
    const unsigned bad = __ballot_sync(0xffffffffu, k < K && rowmask[r0 + k] == 0);
    if (bad) {
      if (lane == 0) atomicMin(&first_bad[e], (int)(k0 + __ffs(bad) - 1));
      return;
    }
  }
}

__global__ void edge_finalize_kernel(long long ne, int *first_bad_tmp, uint8_t *valid, int *first_bad) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  int fb = first_bad_tmp[e];
  valid[e] = fb == 0x7fffffff ? 1 : 0;
  if (first_bad) first_bad[e] = fb == 0x7fffffff ? -1 : fb;
}

__global__ void sweep_rows_kernel(FkTables<float> fk, unsigned long long seed, long long row0, long long n, float *q) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * fk.nq) return;
  long long r = i / fk.nq;
  int j = (int)(i - r * fk.nq);
  q[i] = sweep_value(seed, (uint64_t)(row0 + r), (uint32_t)j, fk.jnt_lo[j], fk.jnt_hi[j]);
}

// ---------------------------------------------------------------------------- pose constraint kernels (math in vk_core.cuh)
// world pose of a site for a block of rows (fp64) -- site_pose (reference: src/mjpl/utils.py:60-75)
__global__ void site_pose_kernel(const FkTables<double> *fkp, int nslot, PoseSpec spec, const double *q, long long n,
                                 double *pos, double *quat) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  const FkTables<double> &fk = *fkp;
  Pose<double> P[MAX_BODY];
  V3<double> anchor[MAX_JNT], axis[MAX_JNT];
  fk_with_joints(fk, nslot, q + row * fk.nq, P, anchor, axis);
  Pose<double> ident; ident.p = mk<double>(0, 0, 0); ident.q.w = 1; ident.q.x = ident.q.y = ident.q.z = 0;
  const Pose<double> &B = spec.site_slot < 0 ? ident : P[spec.site_slot];
  Q4<double> sq; sq.w = spec.site_quat[0]; sq.x = spec.site_quat[1]; sq.y = spec.site_quat[2]; sq.z = spec.site_quat[3];
  V3<double> p = B.p + qrot(B.q, mk<double>(spec.site_pos[0], spec.site_pos[1], spec.site_pos[2]));
  Q4<double> r = qnormalize(qmul(B.q, sq));
  pos[row * 3] = p.x; pos[row * 3 + 1] = p.y; pos[row * 3 + 2] = p.z;
  quat[row * 4] = r.w; quat[row * 4 + 1] = r.x; quat[row * 4 + 2] = r.y; quat[row * 4 + 3] = r.z;
}

struct PoseArgs {
  const FkTables<double> *fk;
  int nslot;
  PoseSpec spec;
  const double *q_old, *q;      // (n,nq)
  long long n;
  int project;                  // 0: valid_config only, 1: apply (projection)
  int max_iters;
  double *q_out;                // (n,nq) projected rows (project) -- untouched rows when !ok
  uint8_t *ok;                  // valid / projection succeeded
  int *iters;                   // optional
  const uint8_t *mask;          // optional: only rows with mask[row] == 1 are evaluated (others: ok = 0)
};

__global__ void __launch_bounds__(64) pose_kernel(const PoseArgs a) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= a.n) return;
  const FkTables<double> &fk = *a.fk;
  const int nq = fk.nq;
  if (a.mask && a.mask[row] != 1) { a.ok[row] = 0; return; }
  double q[MAX_JNT];
  for (int j = 0; j < nq; j++) q[j] = a.q[row * nq + j];
  if (!a.project) {
    a.ok[row] = pose_valid_row(fk, a.nslot, a.spec, q) ? 1 : 0;
    return;
  }
  int it = 0;
  const bool ok = pose_project_row(fk, a.nslot, a.spec, a.q_old + row * nq, q, a.max_iters, &it);
  a.ok[row] = ok ? 1 : 0;
  if (a.iters) a.iters[row] = it;
  if (ok) for (int j = 0; j < nq; j++) a.q_out[row * nq + j] = q[j];
}

// ---------------------------------------------------------------------------- inverse kinematics
// one lane per (target, initial guess) row; math in vk_core.cuh (ik_row)
struct IkArgs {
  const FkTables<double> *fk;
  int nslot;
  IkSpec spec;
  const double *tpos, *tquat;   // (n,3), (n,4 wxyz) world-frame targets
  const double *q_init;         // (n,nq)
  long long n;
  double *q_out;                // (n,nq): last iterate (a solution when ok)
  uint8_t *ok;
  int *iters;                   // optional
  double *err;                  // optional (n,2): position / orientation error of q_out
};

__global__ void __launch_bounds__(64) ik_kernel(const IkArgs a) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= a.n) return;
  const FkTables<double> &fk = *a.fk;
  const int nq = fk.nq;
  double q[MAX_JNT];
  for (int j = 0; j < nq; j++) q[j] = a.q_init[row * nq + j];
  int it = 0;
  double pe = 0, oe = 0;
  const bool ok = ik_row(fk, a.nslot, a.spec, a.tpos + row * 3, a.tquat + row * 4, q, &it, &pe, &oe);
  a.ok[row] = ok ? 1 : 0;
  if (a.iters) a.iters[row] = it;
  if (a.err) { a.err[row * 2] = pe; a.err[row * 2 + 1] = oe; }
  for (int j = 0; j < nq; j++) a.q_out[row * nq + j] = q[j];
}

// ---------------------------------------------------------------------------- RRT extend chains
// near[i] = nodes[slot_i][nn[i]]; chain length K = min(ceil(|target-near|/eps), kcap) (0 if equal)
__global__ void chain_setup_kernel(const double *nodes, long long cap, int nq, const long long *slots, const long long *nn,
                                   const double *targets, long long n, double eps, int kcap, double *near, long long *count,
                                   int *first_bad, const uint8_t *active) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long slot = slots ? slots[i] : i;
  const double *src = nodes + (slot * cap + nn[i]) * nq;
  double d2 = 0;
  for (int j = 0; j < nq; j++) {
    double x = src[j];
    near[i * nq + j] = x;
    double d = targets[i * nq + j] - x;
    d2 += d * d;
  }
  const double dist = sqrt(d2);
  long long k = dist > 0 ? (long long)ceil(dist / eps) : 0;   // NaN targets: no chain
  if (k > kcap) k = kcap;
  if (active && !active[i]) k = 0;                            // masked query: no chain, nothing appended
  count[i] = k;
  first_bad[i] = 0x7fffffff;
}

// Append the valid prefix of every chain to its tree (reference stop rules,
// src/mjpl/planning/utils.py:151-160: constraints failed -> first_bad; a final step shorter than
// 1e-8 does not count), report the reached configuration and its node index.
__global__ void chain_append_kernel(double *nodes, long long *parent, long long *tree_count, long long cap, int nq,
                                    const long long *slots, const long long *nn, const double *near, const double *targets,
                                    const long long *count, const int *first_bad, long long n, double eps, double *reached,
                                    long long *last, unsigned long long *overflow) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n) return;
  const long long slot = slots ? slots[w] : w;
  const long long K = count[w];
  long long good = first_bad[w] == 0x7fffffff ? K : (long long)first_bad[w];
  if (good > K) good = K;
  double d2 = 0;
  for (int j = 0; j < nq; j++) {
    double d = targets[w * nq + j] - near[w * nq + j];
    d2 += d * d;
  }
  const double dist = sqrt(d2);
  // stop rule: the last step lands on the target; if that step is shorter than 1e-8 it is rejected
  if (good == K && K > 0 && (double)K * eps >= dist && dist - (double)(K - 1) * eps < 1e-8) good = K - 1;
  const long long base = tree_count[slot];
  if (base + good > cap) {  // never write past the tree (the host grows trees ahead of time)
    if (lane == 0) atomicAdd(overflow, 1ull);
    good = cap - base;
  }
  double *dst = nodes + (slot * cap + base) * nq;
  for (long long idx = lane; idx < good * nq; idx += 32) {
    const long long k = idx / nq;
    const int j = (int)(idx - k * nq);
    const double reach = (double)(k + 1) * eps;
    const double x0 = near[w * nq + j], x1 = targets[w * nq + j];
    dst[idx] = reach >= dist ? x1 : x0 + (reach / dist) * (x1 - x0);
  }
  for (long long k = lane; k < good; k += 32) parent[slot * cap + base + k] = k == 0 ? nn[w] : base + k - 1;
  __syncwarp();
  const long long li = good > 0 ? base + good - 1 : nn[w];
  if (lane == 0) {
    tree_count[slot] = base + good;
    last[w] = li;
  }
  for (int j = lane; j < nq; j += 32) {
    double v;
    if (good > 0) {
      const double reach = (double)good * eps;
      const double x0 = near[w * nq + j], x1 = targets[w * nq + j];
      v = reach >= dist ? x1 : x0 + (reach / dist) * (x1 - x0);
    } else v = near[w * nq + j];
    reached[w * nq + j] = v;
  }
}

// ---------------------------------------------------------------------------- batched nearest neighbour
// Tree.nearest_neighbor (reference: src/mjpl/planning/tree.py:57-66) for many trees at once:
// trees are rows of a padded (B, cap, nq) fp64 array; one warp scans one tree and keeps the
// arg-min of the squared distance (lowest index wins ties).  Nodes with non-finite entries (the
// +inf sink root of the reference's goal tree) can never win.
// One CTA of 128 threads per tree (a warp per tree ran as long as its 19 dependent trips through a 600-node
// tree took: 33 us per call in the planner's loop, as much as the validity check of the extension itself).
constexpr int NEAREST_THREADS = 128;
__global__ void __launch_bounds__(NEAREST_THREADS) nearest_kernel(const double *nodes, long long cap, int nq, const long long *count,
                                                                  const long long *rows, const double *targets, long long n,
                                                                  long long *out, const uint8_t *active = nullptr) {
  const long long w = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (w >= n) return;
  if (active && !active[w]) { if (tid == 0) out[w] = 0; return; }   // masked query: no scan (its chain is empty anyway)
  __shared__ double s_t[MAX_JNT];
  __shared__ double s_best[NEAREST_THREADS / 32];
  __shared__ long long s_bi[NEAREST_THREADS / 32];
  const long long tree = rows ? rows[w] : w;
  const double *base = nodes + tree * cap * nq;
  if (tid < nq) s_t[tid] = targets[w * nq + tid];
  __syncthreads();
  const long long cnt = count[tree];
  double best = 1.0e300;
  long long bi = 0;
  for (long long i = tid; i < cnt; i += NEAREST_THREADS) {
    double d2 = 0;
    for (int j = 0; j < nq; j++) {
      double d = base[i * nq + j] - s_t[j];
      d2 += d * d;
    }
    if (d2 < best) { best = d2; bi = i; }   // NaN / inf never pass; within a thread the indices ascend: the first minimum stays
  }
  for (int o = 16; o > 0; o >>= 1) {
    double ob = __shfl_xor_sync(0xffffffffu, best, o);
    long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0) { s_best[warp] = best; s_bi[warp] = bi; }
  __syncthreads();
  if (tid == 0) {
    for (int k = 1; k < NEAREST_THREADS / 32; k++)
      if (s_best[k] < best || (s_best[k] == best && s_bi[k] < bi)) { best = s_best[k]; bi = s_bi[k]; }
    out[w] = bi;
  }
}

// ---------------------------------------------------------------------------- parent chains of many trees
// Tree.get_path (reference: src/mjpl/planning/tree.py:68-81): one thread per tree follows the parent links from
// `first` to the root (a few hundred dependent loads: microseconds; the host-side loop it replaces launched four
// kernels per level).
__global__ void tree_paths_kernel(const long long *parent, long long cap, const long long *rows, const long long *first, long long n,
                                  long long max_depth, long long *steps, long long *len) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long *par = parent + (rows ? rows[i] : i) * cap;
  long long *out = steps + i * max_depth;
  long long idx = first[i], d = 0;
  while (idx >= 0 && idx < cap && d < max_depth) { out[d++] = idx; idx = par[idx]; }
  len[i] = (idx >= 0) ? -1 : d;
  for (long long k = d; k < max_depth; k++) out[k] = -1;
}

// ---------------------------------------------------------------------------- signed distance per row (band accounting)
struct MArgs {
  const FkTables<double> *fk;
  const Shape<double> *shapes;
  const Vtx<double> *verts;
  const Pair *pairs;
  const double *pair_rsum;
  int npair, nslot;
  const float *q; int ldq; long long n;
  double far_cap, depth_cap;
  double *dist; int *pair;
};
__global__ void __launch_bounds__(64) min_distance_kernel(const MArgs a) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= a.n) return;
  const FkTables<double> &fk = *a.fk;
  double q[MAX_JNT];
  for (int j = 0; j < fk.nq; j++) q[j] = (double)a.q[row * a.ldq + j];
  double best;
  int bestp;
  row_min_distance(fk, a.nslot, a.shapes, a.verts, a.pairs, a.pair_rsum, a.npair, q, a.far_cap, a.depth_cap, best, bestp);
  a.dist[row] = best;
  if (a.pair) a.pair[row] = bestp;
}

// ---------------------------------------------------------------------------- bi-RRT iteration on the device
// Sampling step of RRT.plan_to_configs for S queries at once (reference: src/mjpl/planning/rrt.py:206-215):
// with probability goal_bias the target is the other tree's root (the goal, or q_init once the trees
// have been swapped), otherwise q_init with the planning joints drawn uniformly within the joint
// limits.  The random stream is counter based (slot, iteration, joint), so a CUDA graph of iterations
// replays without host-side state; the iteration counter lives on the device (counters[0]) and its
// parity is the reference's `swapped` flag.  Masked (finished) slots get NaN targets.
__global__ void rrt_sample_kernel(unsigned long long seed, const long long *counters, long long nslots, int nq,
                                  const double *q_init, const double *q_goal, const uint8_t *plan_mask, const double *lo,
                                  const double *hi, double goal_bias, const uint8_t *active, double *targets) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nslots) return;
  const unsigned long long it = (unsigned long long)counters[0];
  const bool swapped = it & 1ull;
  double *t = targets + s * nq;
  if (!active[s]) {
    for (int j = 0; j < nq; j++) t[j] = __longlong_as_double(0x7ff8000000000000ll);
    return;
  }
  const unsigned long long key = seed + 0xD1B54A32D192ED03ull * (it + 1ull);
  const double u = (double)sweep_bits(key, (uint64_t)s, 63u) * (1.0 / 16777216.0);
  if (u <= goal_bias) {
    const double *src = (swapped ? q_init : q_goal) + s * nq;
    for (int j = 0; j < nq; j++) t[j] = src[j];
    return;
  }
  for (int j = 0; j < nq; j++) {
    double v = q_init[s * nq + j];
    if (plan_mask[j]) {
      const double r = ((double)sweep_bits(key, (uint64_t)s, (uint32_t)(2 * j)) * 16777216.0 + (double)sweep_bits(key, (uint64_t)s, (uint32_t)(2 * j + 1))) *
                       (1.0 / 281474976710656.0);   // 48 random bits in [0, 1)
      v = lo[j] + r * (hi[j] - lo[j]);
    }
    t[j] = v;
  }
}

// End of one iteration (rrt.py:217-235): a query whose two extends reached the same configuration is
// solved -- the connecting node of each tree is recorded and the slot is retired; a query that has
// used up its iteration budget is retired unsolved.  Advances the iteration counter.
// counters: [0] iteration, [1] solved, [2] gave up, [3] still active after this iteration.
__global__ void rrt_meet_kernel(long long nslots, int nq, const double *qa, const double *qb, const long long *ia, const long long *ib,
                                long long max_age, uint8_t *active, long long *age, long long *res_start, long long *res_goal,
                                long long *counters) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool swapped = counters[0] & 1ll;
  int still = 0;
  if (s < nslots && active[s]) {
    bool met = true;
    for (int j = 0; j < nq; j++) met = met && (qa[s * nq + j] == qb[s * nq + j]);
    age[s] += 1;
    if (met) {
      res_start[s] = swapped ? ib[s] : ia[s];
      res_goal[s] = swapped ? ia[s] : ib[s];
      active[s] = 0;
      atomicAdd((unsigned long long *)&counters[1], 1ull);
    } else if (age[s] >= max_age) {
      active[s] = 0;
      atomicAdd((unsigned long long *)&counters[2], 1ull);
    } else {
      still = 1;
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, still);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd((unsigned long long *)&counters[4], (unsigned long long)__popc(m));
}
// one thread, after rrt_meet_kernel: iteration += 1, active count of this iteration published
__global__ void rrt_advance_kernel(long long *counters) {
  counters[0] += 1;
  counters[3] = counters[4];
  counters[4] = 0;
}

// ---------------------------------------------------------------------------- CBiRRT with a projecting constraint: ticks
// RRT.plan_to_configs (reference: src/mjpl/planning/rrt.py:195-235) with the step-by-step
// _constrained_extend (planning/utils.py:139-164) that a projecting constraint needs, for S queries
// that advance ASYNCHRONOUSLY: every slot is a small state machine, and one TICK moves every slot by
// one projected step of whatever extend it is in (or sets up its next extend).  No slot waits for the
// longest chain of another one, and nothing returns to the host in between.
//   phase 0  start of an iteration: sample a target, nearest node of tree A      -> 1
//   phase 1  stepping in tree A            (extend ended: remember q_a)            -> 2
//   phase 2  nearest node of tree B to q_a                                        -> 3
//   phase 3  stepping in tree B            (extend ended: connection test; swap)   -> 0 / 4
//   phase 4  retired (solved, or out of iterations)
// tick_begin (one warp per slot): setups and the proposal of one step (`_step`, and the non-projecting
// constraints that come before the projecting one see the unprojected configuration: joint limits).
// The projection (pose_kernel, masked) and the validity launch on the projected rows follow; tick_end
// (one thread per slot) applies the reference's stop rules (:151-160), appends, and moves the phase.
struct TickState {
  long long nslots, cap;
  int nq;
  double eps, goal_bias;
  unsigned long long seed;
  long long max_age;
  int check_limits_before;            // a JointLimitConstraint precedes the projecting constraint
  const double *q_init, *q_goal;      // (S,nq)
  const uint8_t *plan_mask;           // (nq)
  const double *lo, *hi;              // (nq)
  double *nodes[2]; long long *parent[2]; long long *count[2];   // [0] start trees, [1] goal trees: (S,cap,nq) (S,cap) (S)
  int *phase; uint8_t *swapped; long long *age;
  double *target, *tip, *qa; long long *last, *ia;
  double *cand; float *cand32; double *proj; uint8_t *proj_ok, *valid, *stepping;
  long long *res_start, *res_goal;
  long long *counters;                // [0] ticks, [1] solved, [2] gave up, [3] slots not yet retired, [4] scratch, [5] trees at capacity
};

__global__ void __launch_bounds__(128) tick_begin_kernel(const TickState t) {
  const long long s = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= t.nslots) return;
  const int nq = t.nq;
  int ph = t.phase[s];
  if (ph == 4) { if (lane == 0) t.stepping[s] = 0; return; }
  const int A = t.swapped[s] ? 1 : 0;
  double *tgt = t.target + s * nq, *tip = t.tip + s * nq;
  if (ph == 0 || ph == 2) {
    // ---- set up an extend: the target, then Tree.nearest_neighbor (tree.py:57-66) with the warp
    if (ph == 0) {
      if (lane == 0) {
        const unsigned long long key = t.seed + 0xD1B54A32D192ED03ull * ((unsigned long long)t.age[s] + 1ull);
        const double u = (double)sweep_bits(key, (uint64_t)s, 63u) * (1.0 / 16777216.0);
        const double *src = (A ? t.q_init : t.q_goal) + s * nq;   // the other tree's root (rrt.py:207-212)
        for (int j = 0; j < nq; j++) {
          double v = src[j];
          if (!(u <= t.goal_bias)) {
            v = t.q_init[s * nq + j];
            if (t.plan_mask[j]) {
              const double r = ((double)sweep_bits(key, (uint64_t)s, (uint32_t)(2 * j)) * 16777216.0 +
                                (double)sweep_bits(key, (uint64_t)s, (uint32_t)(2 * j + 1))) * (1.0 / 281474976710656.0);
              v = t.lo[j] + r * (t.hi[j] - t.lo[j]);
            }
          }
          tgt[j] = v;
        }
      }
    } else if (lane < nq) {
      tgt[lane] = t.qa[s * nq + lane];
    }
    __syncwarp();
    const int T = ph == 0 ? A : 1 - A;
    const double *base = t.nodes[T] + s * t.cap * nq;
    const long long cnt = t.count[T][s];
    double best = 1.0e300;
    long long bi = 0;
    for (long long i = lane; i < cnt; i += 32) {
      double d2 = 0;
      for (int j = 0; j < nq; j++) { const double d = base[i * nq + j] - tgt[j]; d2 += d * d; }
      if (d2 < best) { best = d2; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane < nq) tip[lane] = base[bi * nq + lane];
    if (lane == 0) t.last[s] = bi;
    ph += 1;
    __syncwarp();
  }
  // ---- propose one step (ph is 1 or 3): _step (planning/utils.py:167-185), target reached => extend over
  if (lane == 0) {
    bool same = true;
    double d2 = 0;
    for (int j = 0; j < nq; j++) { const double d = tgt[j] - tip[j]; same = same && (tgt[j] == tip[j]); d2 += d * d; }
    int st = 2;   // extend ends without a step
    if (!same) {
      const double dist = sqrt(d2);
      bool lim = true;
      for (int j = 0; j < nq; j++) {
        const double x = dist <= t.eps ? tgt[j] : tip[j] + (tgt[j] - tip[j]) * (t.eps / dist);
        t.cand[s * nq + j] = x;
        lim = lim && (x >= t.lo[j]) && (x <= t.hi[j]);
      }
      st = (!t.check_limits_before || lim) ? 1 : 2;
    }
    t.stepping[s] = (uint8_t)st;
    t.phase[s] = ph;
  }
}

// projected rows of the stepping slots -> fp32 rows for the validity launch (other slots keep a copy
// of their tip: a configuration that is already known to be valid)
__global__ void tick_rows_kernel(const TickState t) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= t.nslots * t.nq) return;
  const long long s = i / t.nq;
  t.cand32[i] = (float)((t.stepping[s] == 1 && t.proj_ok[s]) ? t.proj[i] : t.tip[i]);
}

__global__ void tick_end_kernel(const TickState t) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int alive = 0;
  if (s < t.nslots && t.phase[s] != 4) {
    const int nq = t.nq;
    const int ph = t.phase[s], st = t.stepping[s];
    const int A = t.swapped[s] ? 1 : 0, T = ph == 1 ? A : 1 - A;
    double *tip = t.tip + s * nq;
    const double *tgt = t.target + s * nq, *pr = t.proj + s * nq;
    bool ended = st == 2;
    if (st == 1) {
      double moved = 0, dnew = 0, dold = 0;
      for (int j = 0; j < nq; j++) {
        moved += (pr[j] - tip[j]) * (pr[j] - tip[j]);
        dnew += (tgt[j] - pr[j]) * (tgt[j] - pr[j]);
        dold += (tgt[j] - tip[j]) * (tgt[j] - tip[j]);
      }
      // planning/utils.py:151-160: constraints failed, no progress, or further from the target than before
      bool ok = t.proj_ok[s] && t.valid[s] && sqrt(moved) >= 1e-8 && sqrt(dnew) <= sqrt(dold);
      const long long idx = t.count[T][s];
      if (ok && idx >= t.cap) { ok = false; atomicAdd((unsigned long long *)&t.counters[5], 1ull); }
      if (ok) {
        double *dst = t.nodes[T] + (s * t.cap + idx) * nq;
        for (int j = 0; j < nq; j++) { dst[j] = pr[j]; tip[j] = pr[j]; }
        t.parent[T][s * t.cap + idx] = t.last[s];
        t.count[T][s] = idx + 1;
        t.last[s] = idx;
      } else {
        ended = true;
      }
    }
    alive = 1;
    if (ended) {
      if (ph == 1) {
        for (int j = 0; j < nq; j++) t.qa[s * nq + j] = tip[j];
        t.ia[s] = t.last[s];
        t.phase[s] = 2;
      } else {
        bool met = true;
        for (int j = 0; j < nq; j++) met = met && (tip[j] == t.qa[s * nq + j]);
        if (met) {   // rrt.py:223-229
          t.res_start[s] = A ? t.last[s] : t.ia[s];
          t.res_goal[s] = A ? t.ia[s] : t.last[s];
          t.phase[s] = 4; alive = 0;
          atomicAdd((unsigned long long *)&t.counters[1], 1ull);
        } else {
          t.swapped[s] ^= 1;
          t.age[s] += 1;
          if (t.age[s] >= t.max_age) { t.phase[s] = 4; alive = 0; atomicAdd((unsigned long long *)&t.counters[2], 1ull); }
          else t.phase[s] = 0;
        }
      }
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, alive);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd((unsigned long long *)&t.counters[4], (unsigned long long)__popc(m));
}
__global__ void tick_advance_kernel(long long *counters) {
  counters[0] += 1;
  counters[3] = counters[4];
  counters[4] = 0;
}

}  // namespace vk
