// vk_split.cuh -- narrow phase of the multi-kernel pipeline (its broad phase is vk_pipe.cuh).
//
//   narrow_kernel  consumes the item bins with persistent lanes: one lane per (row, pair) item, one GJK
//                  iteration per trip, a lane whose item is decided takes the next one of the same bin.
//
// Why a pipeline: inside the single validity_kernel a warp only has the items of its own 32 rows to
// feed its lanes (9 of 32 active in the narrow phase: few items per flush, hulls of different sizes
// side by side, a tail of slow items), and 6k SASS instructions shared by 16 warps in different stages
// thrash the instruction cache (22 % of the stall samples).  Here every narrow-phase lane always has an
// item, neighbouring lanes scan hulls of similar size, and each kernel is small.  What is given up is
// the per-warp early exit; its place is taken by the row's byte in the output mask, which a lane looks
// at before it starts an item (bins are consumed cheapest-first: closed-form kinds sit in bin 0).
//
// Same arithmetic, same certified verdicts and the same fp64 item pass as the single kernel
// (vk_kernels.cuh); results are identical (tests/test_gpu_parity.py runs both).
#pragma once

#include "vk_kernels.cuh"

namespace vk {

__device__ __forceinline__ Pose<float> load_pose8(const float *pose8, int nslot, long long row, int slot) {
  Pose<float> P;
  if (slot < 0) { P.p = mk<float>(0, 0, 0); P.q.w = 1; P.q.x = P.q.y = P.q.z = 0; return P; }
  const float4 *b = reinterpret_cast<const float4 *>(pose8 + ((size_t)row * nslot + slot) * 8);
  const float4 u = b[0], v = b[1];
  P.p.x = u.x; P.p.y = u.y; P.p.z = u.z; P.q.w = u.w;
  P.q.x = v.x; P.q.y = v.y; P.q.z = v.z;
  return P;
}

// a certain contact of `row`: the reference's answer for the row is "invalid"
__device__ __forceinline__ void mark_contact(const KArgs &a, long long row) {
  if ((a.mode == MODE_EDGES || a.mode == MODE_CHAINS) && !(a.flags & F_ROWMASK)) {
    long long e;
    int k;
    edge_lookup(a.edge_prefix, a.nedge, row, e, k);
    atomicMin(&a.first_bad[e], k);
  } else {
    a.valid[row] = 0;
  }
}

// an item the fp32 path could not certify -> fp64 item list; if that is full, the row goes to the
// whole-row list exactly once (row_flags de-duplicates, so the row list cannot overflow either)
__device__ __forceinline__ void mark_uncertain(const KArgs &a, long long row, int pair) {
  atomicAdd(&a.counters[C_UNCERTAIN], 1ull);
  if (a.flags & F_NO_RECHECK) {  // debugging aid: leave the row marked 2 unless it has a contact
    if (a.mode == MODE_EDGES || a.mode == MODE_CHAINS) mark_contact(a, row);
    else if (a.valid[row] == 1) a.valid[row] = 2;
    return;
  }
  const unsigned long long slot = atomicAdd(&a.counters[C_RITEMS], 1ull);
  if (slot < a.item_cap) {
    a.recheck_items[slot] = (unsigned long long)row | ((unsigned long long)pair << 44);
  } else {
    const uint32_t bit = 1u << ((unsigned)(row & 3) * 8u);
    if (!(atomicOr(&a.row_flags[row >> 2], bit) & bit)) {
      const unsigned long long s = atomicAdd(&a.counters[C_RECHECK], 1ull);
      a.recheck_rows[s] = row;
    }
  }
}

// a bin is full (far more items than the calibrated average): decide the item on the spot.
// Out of line and on the global tables: it is almost never executed and must not sit in the hot code.
__device__ __noinline__ void broad_overflow_item(const KArgs &a, int ip, long long irow) {
  const Pair pr = a.pairs[ip];
  const Shape<float> &A = a.shapes[pr.sa];
  const Shape<float> &B = a.shapes[pr.sb];
  Pose<float> PA = load_pose8(a.pose8, a.nslot, irow, A.slot);
  Pose<float> PB = load_pose8(a.pose8, a.nslot, irow, B.slot);
  const int v = narrow_item<float>(pr.kind, A, B, a.verts, PA, PB, pr.rsum);
  if (v == V_PEN) mark_contact(a, irow);
  else if (v == V_UNC) mark_uncertain(a, irow, ip);
}

struct NarrowLayout {
  size_t verts, shapes, pairs, adjs, adj, bars, total;
};
__host__ __device__ inline NarrowLayout narrow_layout(int nvert, int nshape, int npair, int nadj) {
  NarrowLayout L;
  size_t o = 0;
  L.verts = o; o = align_up(o + (size_t)nvert * sizeof(Vtx<float>), 128);
  L.shapes = o; o = align_up(o + (size_t)nshape * sizeof(Shape<float>), 128);
  L.pairs = o; o = align_up(o + (size_t)npair * sizeof(Pair), 128);
  L.adjs = o; o = align_up(o + (size_t)(nvert + 1) * sizeof(uint16_t), 128);
  L.adj = o; o = align_up(o + (size_t)nadj, 128);
  L.bars = o; o = align_up(o + 64, 128);
  L.total = o;
  return L;
}

#ifndef VK_NARROW_THREADS
#define VK_NARROW_THREADS 256
#endif
constexpr int NARROW_THREADS = VK_NARROW_THREADS;
#ifndef VK_NARROW_HILL
#define VK_NARROW_HILL 1
#endif
// the hull-graph branch of the support query stays compiled into narrow_kernel although no shape of the shipped models
// takes it: without it ptxas lays the kernel out 8 % slower (0.570 -> 0.614 ms per 1M rows, same box, twice)
constexpr bool NARROW_HILL = VK_NARROW_HILL != 0;
#ifndef VK_NARROW_CLAIM
#define VK_NARROW_CLAIM 64
#endif
constexpr int NARROW_CLAIM = VK_NARROW_CLAIM;   // items a warp claims per ticket
#ifndef VK_NARROW_CTAS
#define VK_NARROW_CTAS 2   // resident CTAs per SM the register budget is set for (B200, 1M Franka rows, narrow_kernel alone: 1 -> 0.274 ms, 2 -> 0.273, 3 -> 0.303, 4 -> 0.356; 128 registers, no spills)
#endif

__global__ void __launch_bounds__(NARROW_THREADS, VK_NARROW_CTAS) narrow_kernel(const __grid_constant__ KArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const NarrowLayout L = narrow_layout(a.nvert, a.nshape, a.npair, a.nadj);
  Vtx<float> *s_verts = reinterpret_cast<Vtx<float> *>(smem + L.verts);
  Shape<float> *s_shapes = reinterpret_cast<Shape<float> *>(smem + L.shapes);
  Pair *s_pairs = reinterpret_cast<Pair *>(smem + L.pairs);
  uint16_t *s_adjs = reinterpret_cast<uint16_t *>(smem + L.adjs);
  uint8_t *s_adj = smem + L.adj;
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem + L.bars);
  const int tid = threadIdx.x;
  const int lane = tid & 31;

  const uint32_t bytes_v = (uint32_t)(a.nvert * sizeof(Vtx<float>));
  const uint32_t bytes_s = (uint32_t)(a.nshape * sizeof(Shape<float>));
  const uint32_t bytes_p = (uint32_t)(a.npair * sizeof(Pair));
  const uint32_t bytes_as = (uint32_t)align_up((size_t)(a.nvert + 1) * sizeof(uint16_t), 16);
  const uint32_t bytes_a = (uint32_t)align_up((size_t)a.nadj, 16);
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&s_bar[0], bytes_v + bytes_s + bytes_p + bytes_as + bytes_a);
    if (bytes_v) bulk_g2s(s_verts, a.verts, bytes_v, &s_bar[0]);
    if (bytes_s) bulk_g2s(s_shapes, a.shapes, bytes_s, &s_bar[0]);
    if (bytes_p) bulk_g2s(s_pairs, a.pairs, bytes_p, &s_bar[0]);
    bulk_g2s(s_adjs, a.adj_start, bytes_as, &s_bar[0]);
    if (bytes_a) bulk_g2s(s_adj, a.adj, bytes_a, &s_bar[0]);
  }
  mbar_wait(&s_bar[0], 0);
  __syncthreads();

  const bool by_row = !(a.mode == MODE_EDGES || a.mode == MODE_CHAINS) || (a.flags & F_ROWMASK);  // the output mask doubles as the early-exit flag
  const int gl = 0;
  const unsigned gmask = 1u << lane;
  GjkState<float> gs;
  Rel<float> rel;
  const Shape<float> *SA = s_shapes, *SB = s_shapes;
  float R = 0.f;
  long long row = 0;
  int pidx = 0;
  int wa = -1, wb = -1;
  bool have = false;

  // Bins are consumed in order, but a warp does not wait for a bin to drain: when the current bin
  // has no unclaimed item left, idle lanes move on to the next bin while the others finish theirs
  // (one tail at the very end instead of one per bin).
  int b = -1;
  unsigned long long count = 0;
  const unsigned long long *items = a.bin_items;
  bool more = false;  // warp-uniform: the current bin may still hold unclaimed items
  unsigned long long blk_next = 0, blk_end = 0, claim = 1;   // warp-uniform: the block of items this warp has claimed, block size
#pragma unroll 1
  for (;;) {
    // fill: every lane without an item takes the next one; items of rows that already have a
    // contact are dropped, plane items are decided on the spot -- keep taking until every lane
    // holds a GJK item or all bins are exhausted.  (Refilling only once 8..24 lanes are idle, so
    // that the fetch code runs with more lanes active, measured 2.5-4.5 % slower.)
#pragma unroll 1
    for (;;) {
      const unsigned need = __ballot_sync(0xffffffffu, !have);
      if (!need) break;
      if (!more) {
        if (++b >= NBIN) break;
        count = a.counters[C_BIN + b];
        if (count > a.bin_capv[b]) count = a.bin_capv[b];
        items = a.bin_items + a.bin_off[b];
        more = count > 0;
        blk_next = blk_end = 0;
        // block size: NARROW_CLAIM for large bins, smaller when the bin would otherwise be shared out
        // among a handful of warps (small batches: the launch then runs as long as its busiest warp)
        claim = count / ((unsigned long long)gridDim.x * (NARROW_THREADS / 32));
        claim = claim < 1 ? 1 : (claim > NARROW_CLAIM ? NARROW_CLAIM : claim);
        continue;
      }
      // Items are claimed in blocks of NARROW_CLAIM per warp (one atomic per block, not per refill:
      // every warp of the grid hits the same ticket word, and a same-address atomic with a return
      // value is served at well under one per nanosecond -- at one ticket per refill the whole kernel
      // ran at the speed of that one word).  A block is private to the warp until it is used up.
      if (blk_next >= blk_end) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&a.counters[C_BTICKET + b], claim);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= count) { more = false; continue; }
        blk_next = base;
        blk_end = base + claim < count ? base + claim : count;
      }
      const unsigned long long left = blk_end - blk_next;
      const unsigned take = (unsigned long long)__popc(need) < left ? (unsigned)__popc(need) : (unsigned)left;
      const unsigned rank = __popc(need & ((1u << lane) - 1u));
      const unsigned long long idx = blk_next + rank;
      blk_next += take;
      if (blk_next >= blk_end && blk_end >= count) more = false;   // that was the bin's last block
      if (!have) {
        if (rank < take) {
          const unsigned long long it = items[idx];
          row = (long long)(it & ((1ull << 44) - 1ull));
          pidx = (int)(it >> 44);
          // the row's flag is loaded together with the poses, not ahead of them: one round trip to L2 instead of two
          const uint8_t alive = by_row ? *reinterpret_cast<volatile const uint8_t *>(a.valid + row) : (uint8_t)1;
          const Pair pr = s_pairs[pidx];
          SA = s_shapes + pr.sa;
          SB = s_shapes + pr.sb;
          R = pr.rsum;
          Pose<float> PA = load_pose8(a.pose8, a.nslot, row, SA->slot);
          Pose<float> PB = load_pose8(a.pose8, a.nslot, row, SB->slot);
          if (alive) {
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
                v = plane_classify(*SA, Bs, PB, R, [&](V3<float> d) { return group_support<1, NARROW_HILL>(Bs, s_verts, s_adjs, s_adj, d, gl, gmask, cold, a.smap_cells, a.smap_ids); });
              } else {
                v = segseg_item(*SA, *SB, s_verts, PA, PB, R);
              }
              if (v == V_PEN) mark_contact(a, row);
              else if (v == V_UNC) mark_uncertain(a, row, pidx);
            }
          }
        }
      }
    }
    if (__ballot_sync(0xffffffffu, have) == 0) {
      if (b >= NBIN) break;
      continue;
    }
    if (have) {
      const Shape<float> &As = *SA, &Bs = *SB;
      const int v = gjk_step_impl(
          gs, rel, R, [&](V3<float> d) { return group_support<1, NARROW_HILL>(As, s_verts, s_adjs, s_adj, d, gl, gmask, wa, a.smap_cells, a.smap_ids); },
          [&](V3<float> d) { return group_support<1, NARROW_HILL>(Bs, s_verts, s_adjs, s_adj, d, gl, gmask, wb, a.smap_cells, a.smap_ids); });
      if (v >= 0) {
        if (v == V_PEN) mark_contact(a, row);
        else if (v == V_UNC) mark_uncertain(a, row, pidx);
        have = false;
      }
    }
  }
}

}  // namespace vk
