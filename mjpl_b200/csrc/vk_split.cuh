// vk_split.cuh -- the validity path as a two-kernel pipeline.
//
//   broad_kernel   rows -> limits, forward kinematics (poses to an L2-resident [row][slot][8]
//                  array), sphere cull, OBB mid-phase; every surviving (row, pair) item is appended
//                  to a global bin chosen by the size of the two hulls.
//   narrow_kernel  consumes the bins with persistent lanes: one lane per item, one GJK iteration
//                  per trip, a lane whose item is decided takes the next one of the same bin.
//
// Why: inside the single validity_kernel a warp only has the items of its own 32 rows to feed its
// lanes (9 of 32 active in the narrow phase: few items per flush, hulls of different sizes side by
// side, a tail of slow items), and 6k SASS instructions shared by 16 warps in different stages
// thrash the instruction cache (22 % of the stall samples).  Here every narrow-phase lane always
// has an item, neighbouring lanes scan hulls of similar size, and each kernel is small.  What is
// given up is the per-warp early exit; its place is taken by the row's byte in the output mask,
// which a lane looks at before it starts an item (bins are consumed cheapest-first).
//
// Same arithmetic, same certified verdicts and the same fp64 item pass as the single kernel
// (vk_kernels.cuh); results are identical (tests/test_gpu_parity.py runs both).
#pragma once

#include "vk_kernels.cuh"

#ifndef VK_BROAD_PAIRS2
#define VK_BROAD_PAIRS2 1   // sphere stage of broad_kernel: two pairs per trip (interleaved dependency chains)
#endif

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
  if (a.mode == MODE_EDGES || a.mode == MODE_CHAINS) {
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

struct BroadLayout {
  size_t shapes, pairs, cen, qtile, queue1, bars, total;
};
template <int TILE>
__host__ __device__ inline BroadLayout broad_layout(int nshape, int npair, int nmoving, int nq) {
  BroadLayout L;
  size_t o = 0;
  L.shapes = o; o = align_up(o + (size_t)nshape * sizeof(Shape<float>), 128);
  L.pairs = o; o = align_up(o + (size_t)npair * sizeof(Pair), 128);
  L.cen = o; o = align_up(o + (size_t)(nmoving + 1) * 3 * TILE * sizeof(float), 128);
  L.qtile = o; o = align_up(o + (size_t)TILE * nq * sizeof(float), 128);
  L.queue1 = o; o = align_up(o + (size_t)Q1_PER_ROW * TILE * sizeof(uint32_t), 128);
  L.bars = o; o = align_up(o + 64 + 8 * (TILE / 32), 128);
  L.total = o;
  return L;
}

template <int TILE>
__global__ void __launch_bounds__(TILE) broad_kernel(const __grid_constant__ KArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int nq = a.fk.nq;
  const BroadLayout L = broad_layout<TILE>(a.nshape, a.npair, a.nmoving, nq);
  Shape<float> *s_shapes = reinterpret_cast<Shape<float> *>(smem + L.shapes);
  Pair *s_pairs = reinterpret_cast<Pair *>(smem + L.pairs);
  float *s_cen = reinterpret_cast<float *>(smem + L.cen);
  float *s_q = reinterpret_cast<float *>(smem + L.qtile);
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem + L.bars);
  constexpr int Q1CAP = Q1_PER_ROW * 32;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int wrow0 = tid & ~31;
  uint32_t *q1 = reinterpret_cast<uint32_t *>(smem + L.queue1) + (tid >> 5) * Q1CAP;

  // ---- one-time: shapes and pairs -> shared memory through the bulk-copy engine --------------------
  const uint32_t bytes_s = (uint32_t)(a.nshape * sizeof(Shape<float>));
  const uint32_t bytes_p = (uint32_t)(a.npair * sizeof(Pair));
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&s_bar[0], bytes_s + bytes_p);
    if (bytes_s) bulk_g2s(s_shapes, a.shapes, bytes_s, &s_bar[0]);
    if (bytes_p) bulk_g2s(s_pairs, a.pairs, bytes_p, &s_bar[0]);
  }
  mbar_wait(&s_bar[0], 0);
  for (int k = tid; k < a.nshape - a.nmoving; k += TILE) {
    const Shape<float> &S = s_shapes[a.nmoving + k];
    float *cc = s_cen + (size_t)a.nmoving * 3 * TILE + k;
    cc[0] = S.bc[0]; cc[TILE] = S.bc[1]; cc[2 * TILE] = S.bc[2];
  }
  __syncthreads();

  long long nrows = a.n;
  if (a.mode == MODE_EDGES || a.mode == MODE_CHAINS) nrows = a.edge_prefix[a.nedge];
  const long long ntiles = (nrows + 31) / 32;
  uint32_t row_parity = 0;
  const bool dense_bulk = (a.mode == MODE_DENSE) && (a.ldq == nq) && ((reinterpret_cast<uintptr_t>(a.q) & 15) == 0);
  long long items_total = 0, rows_total = 0;
  const bool use_obb = !(a.flags & F_NO_OBB);
  const float slack = 1e-4f;
  uint64_t *wbar = s_bar + 8 + (tid >> 5);
  float *wq = s_q + (size_t)wrow0 * nq;
  if (lane == 0) mbar_init(wbar, 1);
  fence_barrier_init();
  __syncwarp();
  const int stat_off = a.nmoving * 3 * TILE - a.nmoving;  // static centre k sits at stat_off + shape index

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

    // ---- P0: the warp's rows -> shared --------------------------------------------------------------
    if (a.mode == MODE_DENSE) {
      if (a.rows_ready) {
        if (lane == 0) {
          const unsigned long long need = (unsigned long long)(row_base + rows_here);
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

    // ---- P1: limits + FK, one lane per row; the row's preliminary answer ---------------------------
    float *q = s_q + tid * nq;
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
      if (a.mode == MODE_EDGES || a.mode == MODE_CHAINS) {
        if (!lim_ok) atomicMin(&a.first_bad[e_idx], e_k);
      } else {
        a.valid[row] = lim_ok ? 1 : 0;   // narrow_kernel / the fp64 pass can only turn it to 0
      }
    }
    const bool do_coll = active && lim_ok && (a.flags & F_COLLISION);
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
        const int sa = a.slot_shape_adr[s], sn = a.slot_shape_num[s];
        for (int k = 0; k < sn; k++) {
          const Shape<float> &S = s_shapes[sa + k];
          V3<float> c = B.p + qrot(B.q, mk<float>(S.bc[0], S.bc[1], S.bc[2]));
          float *cc = s_cen + (size_t)(sa + k) * 3 * TILE + tid;
          cc[0] = c.x; cc[TILE] = c.y; cc[2 * TILE] = c.z;
        }
      }
    }
    __syncwarp();  // poses (global) and centres (shared) of this warp's rows are visible to its lanes

    // ---- one pass over the pair list: A fills q1, B drains it ------------------------------------------
    // Closed-form items (plane / sphere / capsule / box corners) are decided right here, lane = item;
    // they come first in the pair order, so a row they prove invalid stops producing work.  Items
    // that need hull scans go to the global bins.
    if (__ballot_sync(0xffffffffu, do_coll) == 0) continue;
    unsigned hit_mask = 0;  // warp-uniform: bit r = row r of this warp has a certain contact
    int p = 0;
    const int p1 = a.npair;
#pragma unroll 1
    for (;;) {
      // A: sphere cull, lane = row
      int n1 = 0;
      {
        const bool live = do_coll && !((hit_mask >> lane) & 1u);
        if (__ballot_sync(0xffffffffu, live) == 0) break;
        int cached_sa = -1;
        V3<float> cA = mk<float>(0.f, 0.f, 0.f);
        auto centre = [&](int shape, bool is_static) {
          const float *cc = s_cen + (is_static ? stat_off + shape : shape * 3 * TILE + tid);
          return mk<float>(cc[0], cc[TILE], cc[2 * TILE]);
        };
        auto test = [&](const Pair &pr, const V3<float> &a_c, const V3<float> &b_c) {
          const float lim = pr.bsum + slack;
          if (pr.kind == PK_PLANE) {
            const Shape<float> &A = s_shapes[pr.sa];
            const float d = A.ax[0] * (b_c.x - A.c[0]) + A.ax[1] * (b_c.y - A.c[1]) + A.ax[2] * (b_c.z - A.c[2]);
            return live && d <= lim;
          }
          const V3<float> d = a_c - b_c;
          return live && dot(d, d) <= lim * lim;
        };
#if VK_BROAD_PAIRS2
        // Two pairs per trip: the loop is bound by the dependent chain load -> address -> load ->
        // test -> vote of ONE pair (fixed-latency stalls), so two independent chains are
        // interleaved.  B200, 1M Franka rows, same box: 2.77 -> 2.69 ms; a generic N-pair form
        // with small arrays measured 2.72 (N = 2, 3) and 2.74 (N = 4).
        for (; p + 1 < p1 && n1 + 64 <= Q1CAP; p += 2) {
          const Pair pr0 = s_pairs[p], pr1 = s_pairs[p + 1];
          if ((int)pr0.sa != cached_sa) cA = centre(pr0.sa, pr0.flags & PF_A_STATIC);
          const V3<float> cA1 = (pr1.sa == pr0.sa) ? cA : centre(pr1.sa, pr1.flags & PF_A_STATIC);
          const V3<float> cB0 = centre(pr0.sb, pr0.flags & PF_B_STATIC);
          const V3<float> cB1 = centre(pr1.sb, pr1.flags & PF_B_STATIC);
          const bool s0 = test(pr0, cA, cB0), s1 = test(pr1, cA1, cB1);
          const unsigned m0 = __ballot_sync(0xffffffffu, s0), m1 = __ballot_sync(0xffffffffu, s1);
          if (m0 | m1) {
            const unsigned below = (1u << lane) - 1u;
            if (s0) q1[n1 + __popc(m0 & below)] = (uint32_t)lane | ((uint32_t)p << 16);
            n1 += __popc(m0);
            if (s1) q1[n1 + __popc(m1 & below)] = (uint32_t)lane | ((uint32_t)(p + 1) << 16);
            n1 += __popc(m1);
          }
          cached_sa = pr1.sa;
          cA = cA1;
        }
#endif
        for (; p < p1 && n1 + 32 <= Q1CAP; p++) {
          const Pair pr = s_pairs[p];
          if ((int)pr.sa != cached_sa) {
            cached_sa = pr.sa;
            cA = centre(pr.sa, pr.flags & PF_A_STATIC);
          }
          const V3<float> cB = centre(pr.sb, pr.flags & PF_B_STATIC);
          warp_push(test(pr, cA, cB), (uint32_t)lane | ((uint32_t)p << 16), q1, n1, Q1CAP, lane);
        }
        __syncwarp();
      }
      // B: mid-phase cull, lane = surviving (row, pair)
#pragma unroll 1
      for (int b_pos = 0; b_pos < n1; b_pos += 32) {
        const int i = b_pos + lane;
        int bin = -1;
        long long irow = 0;
        int ip = 0;
        unsigned hb = 0;
        if (i < n1) {
          const uint32_t it = q1[i];
          const int r = it & 0xffff;
          ip = (int)(it >> 16);
          irow = row_base + r;
          if (!((hit_mask >> r) & 1u)) {
            const Pair pr = s_pairs[ip];
            const Shape<float> &A = s_shapes[pr.sa];
            const Shape<float> &B = s_shapes[pr.sb];
            const bool scan = item_needs_scan(pr, B);
            bool keep = true;
            Pose<float> PA, PB;
            if ((use_obb && (pr.flags & PF_OBB)) || !scan) {
              PA = load_pose8(a.pose8, a.nslot, irow, A.slot);
              PB = load_pose8(a.pose8, a.nslot, irow, B.slot);
            }
            if (use_obb && (pr.flags & PF_OBB))
              keep = !midphase_cull(pr, A, B, PA, PB, pr.rsum - swept_radius(A) - swept_radius(B), slack);
            if (keep) {
              if (scan) {
                bin = item_bin(pr, A, B);
              } else {
                int v;
                if (pr.kind == PK_PLANE) v = plane_classify(A, B, PB, pr.rsum, [&](V3<float> d) { return support_shape(B, a.verts, d); });
                else v = segseg_item(A, B, a.verts, PA, PB, pr.rsum);
                if (v == V_PEN) { hb = 1u << r; mark_contact(a, irow); }
                else if (v == V_UNC) mark_uncertain(a, irow, ip);
                items_total += 1;
              }
            }
          }
        }
        hit_mask |= __reduce_or_sync(0xffffffffu, hb);
        if (bin >= 0 && ((hit_mask >> (int)(irow - row_base)) & 1u)) bin = -1;  // decided in this very batch
        unsigned todo = __ballot_sync(0xffffffffu, bin >= 0);
        while (todo) {
          const int b = __shfl_sync(0xffffffffu, bin, __ffs(todo) - 1);   // bin of the first pending lane
          const unsigned m = __ballot_sync(0xffffffffu, bin == b);
          unsigned long long base = 0;
          if (lane == __ffs(m) - 1) base = atomicAdd(&a.counters[C_BIN + b], (unsigned long long)__popc(m));
          base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
          if (bin == b) {
            const unsigned long long idx = base + __popc(m & ((1u << lane) - 1u));
            if (idx < a.bin_capv[b]) a.bin_items[a.bin_off[b] + idx] = (unsigned long long)irow | ((unsigned long long)ip << 44);
            else broad_overflow_item(a, ip, irow);
            items_total += 1;
          }
          todo &= ~m;
        }
      }
      __syncwarp();
      if (p >= p1) break;
    }
  }
  if (items_total) atomicAdd(&a.counters[C_ITEMS], (unsigned long long)items_total);
  if (lane == 0 && rows_total) atomicAdd(&a.counters[C_ROWS], (unsigned long long)rows_total);
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
#ifndef VK_NARROW_CTAS
#define VK_NARROW_CTAS 3   // resident CTAs per SM the register budget is set for (B200, 1M Franka rows: 2 -> 2.91 ms, 3 -> 2.77, 4 -> 2.92)
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

  static_assert(GRP == 1, "narrow_kernel assigns one lane per item");
  const bool by_row = !(a.mode == MODE_EDGES || a.mode == MODE_CHAINS);  // the output mask doubles as the early-exit flag
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
        continue;
      }
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(&a.counters[C_BTICKET + b], (unsigned long long)__popc(need));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base + __popc(need) >= count) more = false;
      if (!have) {
        const unsigned long long idx = base + __popc(need & ((1u << lane) - 1u));
        if (idx < count) {
          const unsigned long long it = items[idx];
          row = (long long)(it & ((1ull << 44) - 1ull));
          pidx = (int)(it >> 44);
          if (!(by_row && *reinterpret_cast<volatile const uint8_t *>(a.valid + row) == 0)) {
            const Pair pr = s_pairs[pidx];
            SA = s_shapes + pr.sa;
            SB = s_shapes + pr.sb;
            R = pr.rsum;
            Pose<float> PA = load_pose8(a.pose8, a.nslot, row, SA->slot);
            Pose<float> PB = load_pose8(a.pose8, a.nslot, row, SB->slot);
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
                v = plane_classify(*SA, Bs, PB, R, [&](V3<float> d) { return group_support(Bs, s_verts, s_adjs, s_adj, d, gl, gmask, cold); });
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
          gs, rel, R, [&](V3<float> d) { return group_support(As, s_verts, s_adjs, s_adj, d, gl, gmask, wa); },
          [&](V3<float> d) { return group_support(Bs, s_verts, s_adjs, s_adj, d, gl, gmask, wb); });
      if (v >= 0) {
        if (v == V_PEN) mark_contact(a, row);
        else if (v == V_UNC) mark_uncertain(a, row, pidx);
        have = false;
      }
    }
  }
}

}  // namespace vk
