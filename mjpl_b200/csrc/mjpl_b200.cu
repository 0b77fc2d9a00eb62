// mjpl_b200.cu -- C-ABI entry points (include/mjpl_b200.h) and launch logic.
//
// No CPU fallback: every compute entry point needs a CUDA device and fails with
// MJB_ERR_CUDA otherwise.
#include <cuda_runtime.h>

#include <cub/device/device_scan.cuh>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mjpl_b200.h"
#include "vk_build.h"
#include "vk_core.cuh"
#include "vk_kernels.cuh"
#include "vk_split.cuh"
#include "vk_pipe.cuh"
#include "vk_row.cuh"

using namespace vk;

static const size_t TEV = 5;   // timing events per validity launch (mjb_kernel_timing)

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define CU(call)                                                                                 \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess)                                                                      \
      return fail(MJB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));            \
  } while (0)

struct mjb_model {
  vkb::HostModel H;
  int device = 0;
  int num_sms = 0;
  int tile = 0;         // rows per tile (= threads per CTA)
  int ctas_per_sm = 1;
  int grid = 0;
  size_t smem_bytes = 0;
  KArgs kargs;          // constant part pre-filled
  RArgs rargs;
  FArgs fargs;
  // device tables
  Shape<float> *d_shapes32 = nullptr; Vtx<float> *d_verts32 = nullptr; Pair *d_pairs = nullptr;
  Shape<double> *d_shapes64 = nullptr; Vtx<double> *d_verts64 = nullptr; FkTables<double> *d_fk64 = nullptr;
  double *d_rsum64 = nullptr, *d_bsum64 = nullptr;
  int *d_body_slot = nullptr; float *d_static_pose = nullptr;
  uint16_t *d_adj_start = nullptr; uint8_t *d_adj = nullptr;
  // scratch
  float *d_pose = nullptr;
  unsigned long long *d_counters = nullptr;
  long long *d_recheck = nullptr; size_t recheck_cap = 0;
  // multi-kernel pipeline (vk_pipe.cuh, large batches): poses / level-0 list / item bins / row flags of one batch
  bool split = false; size_t fk_smem = 0, mid_smem = 0, narrow_smem = 0; int fk_grid = 0, mid_grid = 0, narrow_grid = 0;
  float *d_pose8 = nullptr; unsigned long long *d_bins = nullptr; uint32_t *d_row_flags = nullptr; size_t split_cap = 0;
  uint8_t *d_rowmask = nullptr;   // edges through the pipeline: one byte per waypoint (F_ROWMASK)
  size_t chain_hint = 0;          // mjb_set_chain_hint
  bool light_narrow = false;      // no hull of more than 8 vertices: edge batches stay in the single kernel
  unsigned long long *d_l0 = nullptr; size_t l0_cap = 0;
  bool rowk = false; size_t rowk_smem = 0; int rowk_grid = 0; long long rowk_rows = 0;   // one-warp-per-row kernel for small launches
  GroupPair *d_gpairs = nullptr; StaticGroup *d_sgroups = nullptr; uint16_t *d_gp_member = nullptr;
  uint32_t *d_smap_cells = nullptr; uint8_t *d_smap_ids = nullptr;
  size_t cur_rows = 0, cur_rows_hint = 0, grp_small_rows = 0, split_min = 0, bin_cap_override = 0, l0_cap_override = 0; bool use_split = false;
  // optional per-kernel timing (mjb_kernel_timing): 4 events per validity launch
  bool timing = false; std::vector<cudaEvent_t> tev; std::vector<uint8_t> tev_split; size_t tev_used = 0;   // decided per launch from the row count
  long long *d_edge_count = nullptr, *d_edge_prefix = nullptr; int *d_first_bad = nullptr; size_t edge_cap = 0;
  void *d_cub = nullptr; size_t cub_bytes = 0;
  double *d_chain_near = nullptr; long long *d_chain_nn = nullptr; size_t chain_cap = 0;
  float *d_stage_q = nullptr; uint8_t *d_stage_v = nullptr; size_t stage_rows = 0;
  float *h_pin_q = nullptr; uint8_t *h_pin_v = nullptr; size_t pin_rows = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t copy_stream = nullptr; cudaEvent_t ev_ready_reset = nullptr;   // streamed host entry point
  unsigned long long *d_rows_ready = nullptr, *h_progress = nullptr; size_t progress_cap = 0;
  // per-handle ordering of calls that arrive on different streams (see enter_stream)
  cudaStream_t last_stream = nullptr; cudaEvent_t ev_last = nullptr; bool have_last = false;
  long long rows_total = 0, launches = 0;
};

extern "C" const char *mjb_last_error(void) { return g_err.c_str(); }

extern "C" int mjb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

template <typename T> static int upload(T **dst, const std::vector<T> &src) {
  size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
  CU(cudaMalloc((void **)dst, align_up(bytes, 256)));
  if (!src.empty()) CU(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  return MJB_OK;
}

// rows resident per SM with TILE rows per CTA, or -1 if the tables do not fit
template <int TILE> static int try_tile(const vkb::HostModel &H, int max_smem_optin, int *ctas, size_t *smem) {
  SmemLayout L = smem_layout<TILE>((int)H.verts.size(), (int)H.shapes.size(), (int)H.pairs.size(), H.nmoving_shapes, H.nq,
                                   (int)H.adj.size());
  if ((int)L.total > max_smem_optin) return -1;
  if ((int)H.shapes.size() - H.nmoving_shapes > TILE) return -1;  // static centres share one [xyz][TILE] block
  // the attribute is per kernel function, shared by every handle in the process: always ask for
  // the device maximum so that handles with different table sizes can coexist
  if (cudaFuncSetAttribute(validity_kernel<TILE, GRP>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin) != cudaSuccess ||
      cudaFuncSetAttribute(validity_kernel<TILE, GRP_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, validity_kernel<TILE, GRP>, TILE, L.total) != cudaSuccess || occ < 1) {
    cudaGetLastError();
    return -1;
  }
  *ctas = occ;
  *smem = L.total;
  return occ * TILE;
}

extern "C" int mjb_model_create(const mjb_model_desc *desc, mjb_model **out) {
  if (!desc || !out) return fail(MJB_ERR_ARG, "null argument");
  *out = nullptr;
  mjb_model *m = new mjb_model();
  if (!vkb::build_host_model(desc, m->H)) {
    std::string e = m->H.err;
    delete m;
    return fail(MJB_ERR_MODEL, e);
  }
  auto bail = [&](int rc) { mjb_model_destroy(m); return rc; };
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
    cudaGetLastError();
    delete m;
    return fail(MJB_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  }
  cudaError_t ce = cudaGetDevice(&m->device);
  if (ce != cudaSuccess) { delete m; return fail(MJB_ERR_CUDA, cudaGetErrorString(ce)); }
  cudaDeviceProp prop;
  ce = cudaGetDeviceProperties(&prop, m->device);
  if (ce != cudaSuccess) { delete m; return fail(MJB_ERR_CUDA, cudaGetErrorString(ce)); }
  m->num_sms = prop.multiProcessorCount;
  const auto &H = m->H;

  // device tables (fp32 fast path + fp64 re-evaluation path)
  std::vector<Shape<float>> s32; std::vector<Vtx<float>> v32;
  for (auto &s : H.shapes) s32.push_back(vkb::convert_shape<float>(s));
  for (auto &v : H.verts) { Vtx<float> f; f.x = (float)v.x; f.y = (float)v.y; f.z = (float)v.z; f.w = 0; v32.push_back(f); }
  int rc;
  if ((rc = upload(&m->d_shapes32, s32))) return bail(rc);
  if ((rc = upload(&m->d_verts32, v32))) return bail(rc);
  if ((rc = upload(&m->d_pairs, H.pairs))) return bail(rc);
  if ((rc = upload(&m->d_shapes64, H.shapes))) return bail(rc);
  if ((rc = upload(&m->d_verts64, H.verts))) return bail(rc);
  if ((rc = upload(&m->d_rsum64, H.pair_rsum64))) return bail(rc);
  if ((rc = upload(&m->d_bsum64, H.pair_bsum64))) return bail(rc);
  std::vector<FkTables<double>> fk1(1, H.fk);
  if ((rc = upload(&m->d_fk64, fk1))) return bail(rc);
  std::vector<float> sp(7 * (size_t)H.nbody);
  for (int b = 0; b < H.nbody; b++) {
    const auto &P = H.static_pose[b];
    float v[7] = {(float)P.p.x, (float)P.p.y, (float)P.p.z, (float)P.q.w, (float)P.q.x, (float)P.q.y, (float)P.q.z};
    if (H.body_slot[b] >= 0) { v[0] = v[1] = v[2] = 0; v[3] = 1; v[4] = v[5] = v[6] = 0; }
    memcpy(&sp[7 * b], v, sizeof v);
  }
  if ((rc = upload(&m->d_static_pose, sp))) return bail(rc);
  if ((rc = upload(&m->d_body_slot, H.body_slot))) return bail(rc);
  {
    std::vector<uint16_t> mem = H.gp_member; mem.resize(align_up(std::max<size_t>(mem.size(), 1), 8), 0);   // 16-byte units
    if ((rc = upload(&m->d_gpairs, H.group_pairs))) return bail(rc);
    if ((rc = upload(&m->d_sgroups, H.static_groups))) return bail(rc);
    if ((rc = upload(&m->d_gp_member, mem))) return bail(rc);
  }
  if (!H.smap_cells.empty() && !(getenv("MJB_SMAP") && atoi(getenv("MJB_SMAP")) == 0)) {   // MJB_SMAP=0: scan all vertices (A/B, tests)
    if ((rc = upload(&m->d_smap_cells, H.smap_cells))) return bail(rc);
    if ((rc = upload(&m->d_smap_ids, H.smap_ids))) return bail(rc);
  }
  {  // padded to 16-byte multiples: the kernel bulk-copies whole 16-byte units
    std::vector<uint16_t> as = H.adj_start; as.resize(align_up(as.size(), 8), 0);
    std::vector<uint8_t> ad = H.adj; ad.resize(align_up(std::max<size_t>(ad.size(), 1), 16), 0);
    if ((rc = upload(&m->d_adj_start, as))) return bail(rc);
    if ((rc = upload(&m->d_adj, ad))) return bail(rc);
  }

  // kernel configuration: the tile (rows per CTA) that keeps most rows resident per SM
  int optin = 0;
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, m->device);
  int best = -1, r, ctas = 0;
  size_t smem = 0;
  const char *force = getenv("MJB_TILE");  // tuning knob: force the rows-per-CTA choice
  const int forced = force ? atoi(force) : 0;
  if ((!forced || forced == 512) && (r = try_tile<512>(H, optin, &ctas, &smem)) > best) { best = r; m->tile = 512; m->ctas_per_sm = ctas; m->smem_bytes = smem; }
  if ((!forced || forced == 256) && (r = try_tile<256>(H, optin, &ctas, &smem)) > best) { best = r; m->tile = 256; m->ctas_per_sm = ctas; m->smem_bytes = smem; }
  if ((!forced || forced == 128) && (r = try_tile<128>(H, optin, &ctas, &smem)) > best) { best = r; m->tile = 128; m->ctas_per_sm = ctas; m->smem_bytes = smem; }
  if (best < 0) return bail(fail(MJB_ERR_MODEL, "model tables do not fit in shared memory"));
  m->grid = m->num_sms * m->ctas_per_sm;

  {  // Multi-kernel pipeline (vk_pipe.cuh) for large batches.  Measured on B200 against the single
     // kernel (tools/split_crossover.py, Franka rows, raw call): 1k..64k rows 1.5-1.9x slower (five
     // launches and three persistent grids that each stage their tables cost ~0.2 ms before the first
     // row), 131k rows 7 % slower, 262k 24 % faster, 524k 31 % faster, 1M 43 % faster.  MJB_SPLIT=0
     // never, =1 always, default: batches of at least MJB_SPLIT_MIN rows (200000).
    const char *sp = getenv("MJB_SPLIT");
    const char *smin = getenv("MJB_SPLIT_MIN");
    const int mode = sp ? atoi(sp) : -1;
    m->split_min = mode == 1 ? 0 : (smin ? (size_t)atoll(smin) : (size_t)200000);
    m->light_narrow = mode != 1;
    for (const auto &sh : H.shapes) if (sh.kind == SK_VERTS && sh.nvert > 8) m->light_narrow = false;
    const char *bc = getenv("MJB_BIN_CAP");   // testing: tiny bins force the on-the-spot path of full bins
    m->bin_cap_override = bc ? (size_t)atoll(bc) : 0;
    // launches of at most this many rows use the GRP_SMALL-lanes-per-item instance of validity_kernel.  B200, Franka rows, raw
    // call (tools/grp_crossover.py): 1k rows 132 -> 121 us, 8k 149 -> 130, 16k 183 -> 175, 32k 210 -> 228, 131k 439 -> 550.
    const char *gr = getenv("MJB_GRP_ROWS");
    m->grp_small_rows = gr ? (size_t)atoll(gr) : (size_t)12288;
    const char *lc = getenv("MJB_L0_CAP");    // testing: a tiny level-0 list forces the whole-row fp64 path
    m->l0_cap_override = lc ? (size_t)atoll(lc) : 0;
    if (mode != 0) {
      const PipeFkLayout FL = pipe_fk_layout((int)H.group_pairs.size(), (int)H.static_groups.size(), H.ngroup_moving, H.nq);
      const MidLayout ML = mid_layout((int)H.shapes.size(), (int)H.pairs.size(), (int)H.group_pairs.size(), (int)H.gp_member.size());
      const NarrowLayout NL = narrow_layout((int)H.verts.size(), (int)H.shapes.size(), (int)H.pairs.size(), (int)H.adj.size());
      int focc = 0, mocc = 0, nocc = 0;
      const bool dbg = getenv("MJB_DEBUG") != nullptr;
      auto ok = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess && dbg) fprintf(stderr, "[mjb] %s: %s\n", what, cudaGetErrorString(e));
        if (e != cudaSuccess) cudaGetLastError();
        return e == cudaSuccess;
      };
      if ((int)FL.total <= optin && (int)ML.total <= optin && (int)NL.total <= optin && !H.group_pairs.empty() &&
          ok(cudaFuncSetAttribute(fk_cull_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin), "attr fk_cull_kernel") &&
          ok(cudaFuncSetAttribute(mid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 1024), "attr mid_kernel") &&
          ok(cudaFuncSetAttribute(narrow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin), "attr narrow_kernel") &&
          ok(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&focc, fk_cull_kernel, PIPE_FK_THREADS, FL.total), "occupancy fk_cull_kernel") && focc >= 1 &&
          ok(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&mocc, mid_kernel, MID_THREADS, ML.total), "occupancy mid_kernel") && mocc >= 1 &&
          ok(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nocc, narrow_kernel, NARROW_THREADS, NL.total), "occupancy narrow_kernel") && nocc >= 1) {
        m->split = true;
        m->fk_smem = FL.total; m->mid_smem = ML.total; m->narrow_smem = NL.total;
        m->fk_grid = m->num_sms * focc; m->mid_grid = m->num_sms * mocc; m->narrow_grid = m->num_sms * nocc;
      }
      if (getenv("MJB_DEBUG"))
        fprintf(stderr, "[mjb] pipeline %s: smem fk %zu mid %zu narrow %zu (optin %d), CTAs/SM fk %d mid %d narrow %d, group pairs %zu\n",
                m->split ? "on" : "OFF", FL.total, ML.total, NL.total, optin, focc, mocc, nocc, H.group_pairs.size());
    }
  }
  {  // small launches: one warp per row (vk_row.cuh)
    const RowLayout RL = row_layout((int)H.verts.size(), (int)H.shapes.size(), (int)H.pairs.size(), (int)H.group_pairs.size(),
                                    (int)H.static_groups.size(), (int)H.gp_member.size(), (int)H.adj.size(), H.nslot, H.ngroup_moving, H.nq);
    int rocc = 0;
    const char *rk = getenv("MJB_ROWK_ROWS");   // launches of at most this many rows take row_kernel (0: never)
    m->rowk_rows = rk ? atoll(rk) : 6144;   // B200, Franka rows, raw call (tools/rowk_crossover.py): 64 rows 101 -> 39 us, 1k 121 -> 49, 4k 128 -> 72, 8k 130 -> 121, 16k 183 -> 244
    if (m->rowk_rows > 0 && (int)RL.total <= optin && !H.group_pairs.empty() && H.pairs.size() <= 65535 &&
        cudaFuncSetAttribute(row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&rocc, row_kernel, ROWK_THREADS, RL.total) == cudaSuccess && rocc >= 1) {
      m->rowk = true; m->rowk_smem = RL.total; m->rowk_grid = m->num_sms * rocc;
    } else {
      cudaGetLastError();
    }
    if (getenv("MJB_DEBUG")) fprintf(stderr, "[mjb] row kernel %s: smem %zu, CTAs/SM %d, up to %lld rows\n", m->rowk ? "on" : "OFF", RL.total, rocc, m->rowk_rows);
  }
  CU(cudaMalloc((void **)&m->d_pose, (size_t)m->grid * std::max(H.nslot, 1) * 7 * m->tile * sizeof(float)));
  CU(cudaMalloc((void **)&m->d_counters, C_NCOUNTERS * sizeof(unsigned long long)));
  CU(cudaMemset(m->d_counters, 0, C_NCOUNTERS * sizeof(unsigned long long)));
  CU(cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&m->ev_last, cudaEventDisableTiming));

  // constant kernel arguments
  KArgs &k = m->kargs;
  memset(&k, 0, sizeof k);
  k.fk = vkb::convert_fk<float>(H.fk);
  for (int j = 0; j < H.njnt; j++) { k.jnt_lo[j] = H.jnt_lo[j]; k.jnt_hi[j] = H.jnt_hi[j]; }
  for (int s = 0; s < H.nslot; s++) { k.slot_shape_adr[s] = H.slot_shape_adr[s]; k.slot_shape_num[s] = H.slot_shape_num[s]; }
  k.shapes = m->d_shapes32; k.verts = m->d_verts32; k.pairs = m->d_pairs;
  k.adj_start = m->d_adj_start; k.adj = m->d_adj; k.nadj = (int)H.adj.size();
  k.nshape = (int)H.shapes.size(); k.nmoving = H.nmoving_shapes; k.nvert = (int)H.verts.size();
  k.npair = (int)H.pairs.size(); k.nslot = H.nslot;
  k.nrounds = H.nrounds;
  for (int r = 0; r <= H.nrounds; r++) k.round_start[r] = H.round_start[r];
  for (int r = 0; r < H.nrounds; r++) k.round_gjk[r] = H.round_gjk[r];
  k.pose_scratch = m->d_pose; k.counters = m->d_counters;
  for (int sl = 0; sl < MAX_BODY; sl++) { k.slot_group_adr[sl] = H.slot_group_adr[sl]; k.slot_group_num[sl] = H.slot_group_num[sl]; }
  for (int g = 0; g < H.ngroup_moving; g++) for (int a = 0; a < 3; a++) k.group_c[g][a] = (float)H.group_c[g][a];
  k.gpairs = m->d_gpairs; k.sgroups = m->d_sgroups; k.gp_member = m->d_gp_member;
  k.smap_cells = m->d_smap_cells; k.smap_ids = m->d_smap_ids;
  k.ngpair = (int)H.group_pairs.size(); k.nsgroup = (int)H.static_groups.size(); k.ngroup_moving = H.ngroup_moving;
  k.nmember = (int)H.gp_member.size();
  for (int a = 0; a < 3; a++) k.gp_kind_end[a] = H.gp_kind_end[a];
  RArgs &ra = m->rargs;
  memset(&ra, 0, sizeof ra);
  ra.fk = m->d_fk64; ra.shapes = m->d_shapes64; ra.verts = m->d_verts64; ra.pairs = m->d_pairs;
  ra.pair_rsum = m->d_rsum64; ra.pair_bsum = m->d_bsum64; ra.npair = k.npair; ra.nslot = H.nslot;
  ra.counters = m->d_counters;
  FArgs &fa = m->fargs;
  memset(&fa, 0, sizeof fa);
  fa.fk = k.fk; fa.nbody_all = H.nbody; fa.body_slot = m->d_body_slot; fa.static_pose = m->d_static_pose;
  *out = m;
  return MJB_OK;
}

extern "C" void mjb_model_destroy(mjb_model *m) {
  if (!m) return;
  // At interpreter exit the CUDA runtime may already be unloading when a handle is finalised:
  // its resources are gone with the context, and touching them (thousands of events) can crash.
  if (cudaSetDevice(m->device) != cudaSuccess) {
    cudaGetLastError();
    delete m;
    return;
  }
  cudaDeviceSynchronize();
  cudaFree(m->d_shapes32); cudaFree(m->d_verts32); cudaFree(m->d_pairs); cudaFree(m->d_shapes64);
  cudaFree(m->d_verts64); cudaFree(m->d_fk64); cudaFree(m->d_rsum64); cudaFree(m->d_bsum64);
  cudaFree(m->d_body_slot); cudaFree(m->d_static_pose); cudaFree(m->d_adj_start); cudaFree(m->d_adj); cudaFree(m->d_pose); cudaFree(m->d_counters);
  cudaFree(m->d_pose8); cudaFree(m->d_bins); cudaFree(m->d_row_flags); cudaFree(m->d_l0);
  cudaFree(m->d_rowmask); cudaFree(m->d_gpairs); cudaFree(m->d_sgroups); cudaFree(m->d_gp_member); cudaFree(m->d_smap_cells); cudaFree(m->d_smap_ids);
  cudaFree(m->d_recheck); cudaFree(m->d_edge_count); cudaFree(m->d_edge_prefix); cudaFree(m->d_first_bad);
  cudaFree(m->d_cub); cudaFree(m->d_stage_q); cudaFree(m->d_stage_v); cudaFree(m->d_chain_near); cudaFree(m->d_chain_nn);
  if (m->h_pin_q) cudaFreeHost(m->h_pin_q);
  if (m->h_pin_v) cudaFreeHost(m->h_pin_v);
  if (m->own_stream) cudaStreamDestroy(m->own_stream);
  if (m->copy_stream) { cudaStreamDestroy(m->copy_stream); cudaEventDestroy(m->ev_ready_reset); cudaFree(m->d_rows_ready); }
  if (m->h_progress) cudaFreeHost(m->h_progress);
  if (m->ev_last) cudaEventDestroy(m->ev_last);
  for (cudaEvent_t e : m->tev) cudaEventDestroy(e);   // timing events still alive (timing left on)
  m->tev.clear();
  cudaGetLastError();
  delete m;
}

extern "C" int32_t mjb_model_npair(const mjb_model *m) { return m ? (int32_t)m->H.pairs.size() : 0; }
extern "C" int mjb_model_pairs(const mjb_model *m, int32_t *g1, int32_t *g2) {
  if (!m || !g1 || !g2) return fail(MJB_ERR_ARG, "null argument");
  for (size_t i = 0; i < m->H.pairs.size(); i++) { g1[i] = m->H.pair_g1[i]; g2[i] = m->H.pair_g2[i]; }
  return MJB_OK;
}

static int ensure_edge_buffers(mjb_model *m, size_t ne, cudaStream_t st);

// capacity of bin b for `rows` rows: twice the calibrated average plus a floor (uniform random rows
// are what the calibration saw; planner chains near obstacles produce more, and a full bin only
// costs speed -- its items are decided on the spot)
static size_t bin_capacity(const vkb::HostModel &H, int b, size_t rows) {
  return (size_t)((2.0 * H.bin_expect[b] + 0.25) * (double)rows) + 1024;
}

static int ensure_recheck(mjb_model *m, size_t rows, cudaStream_t st, bool may_split = true, size_t rows_hint = 0) {
  m->cur_rows = rows;
  m->cur_rows_hint = rows_hint ? rows_hint : rows;   // expected rows when `rows` is only a worst case (chains)
  m->use_split = may_split && m->split && rows >= m->split_min;
  if (m->use_split && rows > m->split_cap) {
    CU(cudaStreamSynchronize(st));
    cudaFree(m->d_pose8); cudaFree(m->d_bins); cudaFree(m->d_row_flags); cudaFree(m->d_l0); cudaFree(m->d_rowmask);
    m->d_pose8 = nullptr; m->d_bins = nullptr; m->d_row_flags = nullptr; m->d_l0 = nullptr; m->d_rowmask = nullptr;
    size_t cap = std::max<size_t>(rows, 1 << 16);
    // level-0 list: 3x the calibrated survivors per row + 4 (uniform rows are what the calibration saw;
    // rows whose entries do not fit are re-evaluated whole in fp64, so a full list costs speed only)
    m->l0_cap = (size_t)((3.0 * m->H.calib_l0_per_row + 4.0) * (double)cap) + 4096;
    CU(cudaMalloc((void **)&m->d_l0, m->l0_cap * sizeof(unsigned long long)));
    CU(cudaMalloc((void **)&m->d_pose8, cap * std::max(m->H.nslot, 1) * 8 * sizeof(float)));
    size_t tot = 0;
    for (int b = 0; b < NBIN; b++) tot += bin_capacity(m->H, b, cap);
    CU(cudaMalloc((void **)&m->d_bins, tot * sizeof(unsigned long long)));
    CU(cudaMalloc((void **)&m->d_row_flags, align_up(cap, 4)));
    CU(cudaMalloc((void **)&m->d_rowmask, align_up(cap, 256)));
    m->split_cap = cap;
  }
  if (rows <= m->recheck_cap) return MJB_OK;
  if (m->d_recheck) { CU(cudaStreamSynchronize(st)); CU(cudaFree(m->d_recheck)); m->d_recheck = nullptr; }
  size_t cap = std::max<size_t>(rows, 1 << 20);
  CU(cudaMalloc((void **)&m->d_recheck, 2 * cap * sizeof(long long)));  // row list, then item list
  m->recheck_cap = cap;
  return MJB_OK;
}

#ifndef VK_RECHECK_GRID
#define VK_RECHECK_GRID 4
#endif
static int launch_validity(mjb_model *m, KArgs &k, RArgs &r, cudaStream_t st) {
  CU(cudaMemsetAsync(m->d_counters, 0, C_PER_LAUNCH * sizeof(unsigned long long), st));
  k.recheck_rows = m->d_recheck;
  r.recheck_rows = m->d_recheck;
  k.recheck_items = (unsigned long long *)(m->d_recheck + m->recheck_cap);
  r.recheck_items = k.recheck_items;
  k.item_cap = r.item_cap = m->recheck_cap;
  cudaEvent_t *ev = nullptr;   // TEV events per launch: start, after each of up to four kernels
  const bool pipe = m->use_split && (k.flags & F_COLLISION);
  if (m->timing && m->tev_used + TEV <= m->tev.size()) {
    ev = &m->tev[m->tev_used];
    m->tev_split[m->tev_used / TEV] = pipe ? 1 : 0;
    m->tev_used += TEV;
  }
  if (ev) CU(cudaEventRecord(ev[0], st));
  if (pipe) {
    // multi-kernel pipeline (vk_pipe.cuh): FK + group cull -> expansion, capsule and OBB culls, bins -> narrow phase
    k.pose8 = m->d_pose8; k.bin_items = m->d_bins; k.row_flags = m->d_row_flags;
    size_t off = 0;
    for (int b = 0; b < NBIN; b++) {
      const size_t c = bin_capacity(m->H, b, m->split_cap);
      k.bin_off[b] = off; k.bin_capv[b] = m->bin_cap_override ? std::min(c, m->bin_cap_override) : c;
      off += c;
    }
    k.l0_items = m->d_l0; k.l0_cap = m->l0_cap_override ? std::min(m->l0_cap, m->l0_cap_override) : m->l0_cap;
    CU(cudaMemsetAsync(m->d_row_flags, 0, align_up(std::min(m->cur_rows, m->split_cap), 4), st));
    fk_cull_kernel<<<m->fk_grid, PIPE_FK_THREADS, m->fk_smem, st>>>(k);
    CU(cudaGetLastError());
    if (ev) CU(cudaEventRecord(ev[1], st));
    mid_kernel<<<m->mid_grid, MID_THREADS, m->mid_smem, st>>>(k);
    CU(cudaGetLastError());
    if (ev) CU(cudaEventRecord(ev[2], st));
    narrow_kernel<<<m->narrow_grid, NARROW_THREADS, m->narrow_smem, st>>>(k);
    CU(cudaGetLastError());
    if (ev) CU(cudaEventRecord(ev[3], st));
    m->launches += 3;
  } else {
    // small launches: one warp per row.  Dense / sweep launches know their size here; for edges and chains
    // only the device does: both kernels are enqueued and the one out of its regime returns at once.
    const bool device_count = k.mode == MODE_EDGES || k.mode == MODE_CHAINS;
    // For chains cur_rows is a worst case and cur_rows_hint the expected count: when the expectation is small only
    // row_kernel is enqueued, without a cap -- it is correct for any number of rows, merely slower than the
    // throughput kernel if the chains turn out much longer than expected (an empty launch of that kernel's
    // persistent grid costs 8 us, a quarter of a small extension).
    const bool rowk_only = m->rowk && (long long)(device_count ? m->cur_rows_hint : m->cur_rows) <= m->rowk_rows;
    k.rowk_max = (m->rowk && device_count && !rowk_only) ? m->rowk_rows : -1;
    if (rowk_only || (m->rowk && device_count)) {
      const long long want = device_count ? (long long)m->rowk_grid : std::min<long long>(m->rowk_grid, ((long long)m->cur_rows + 7) / 8);
      row_kernel<<<(unsigned)std::max<long long>(want, 1), ROWK_THREADS, m->rowk_smem, st>>>(k);
      CU(cudaGetLastError());
      m->launches++;
    }
    // (the planner's extends): the instance whose narrow phase puts GRP_SMALL lanes on one item
    const bool small = m->cur_rows_hint <= m->grp_small_rows;
    if (!rowk_only)
    switch (m->tile) {
      case 512: if (small) validity_kernel<512, GRP_SMALL><<<m->grid, 512, m->smem_bytes, st>>>(k); else validity_kernel<512, GRP><<<m->grid, 512, m->smem_bytes, st>>>(k); break;
      case 256: if (small) validity_kernel<256, GRP_SMALL><<<m->grid, 256, m->smem_bytes, st>>>(k); else validity_kernel<256, GRP><<<m->grid, 256, m->smem_bytes, st>>>(k); break;
      default: if (small) validity_kernel<128, GRP_SMALL><<<m->grid, 128, m->smem_bytes, st>>>(k); else validity_kernel<128, GRP><<<m->grid, 128, m->smem_bytes, st>>>(k); break;
    }
    CU(cudaGetLastError());
    if (ev) { CU(cudaEventRecord(ev[1], st)); CU(cudaEventRecord(ev[2], st)); CU(cudaEventRecord(ev[3], st)); }
    m->launches++;
  }
  if ((k.flags & F_COLLISION) && !(k.flags & F_NO_RECHECK)) {
    // one warp per fp64 item, tickets for the rest: a small launch does not need the full grid
    const long long rgrid = std::min<long long>((long long)m->num_sms * VK_RECHECK_GRID, std::max<long long>(((long long)m->cur_rows_hint + 3) / 4, 16));
    recheck_kernel<<<(unsigned)rgrid, 128, 0, st>>>(r);
    CU(cudaGetLastError());
    m->launches++;
  }
  if (ev) CU(cudaEventRecord(ev[4], st));
  return MJB_OK;
}

static int check_common(mjb_model *m, uint32_t flags) {
  if (!m) return fail(MJB_ERR_ARG, "null model");
  if (!(flags & (MJB_CHECK_LIMITS | MJB_CHECK_COLLISION))) return fail(MJB_ERR_ARG, "flags select no check");
  int dev = -1;
  CU(cudaGetDevice(&dev));
  if (dev != m->device) CU(cudaSetDevice(m->device));
  return MJB_OK;
}

// Every call on a handle uses the handle's scratch (counters with the tile ticket, pose arrays, item
// bins, fp64 work lists, edge prefix sums).  Calls are therefore ordered per handle: each one leaves
// an event behind on its stream, and a call that arrives on a DIFFERENT stream first makes its stream
// wait for that event (also before any scratch buffer is grown and the old one freed).  Calls from
// several host threads still have to be serialised by the caller (mjpl_b200.engine holds a lock).
// While `st` is being captured into a CUDA graph (the planner replays whole iterations as graphs) no
// cross-stream event is waited for or left behind: the graph's own edges order the work, and the
// caller keeps other calls on the handle away until its replays are done.
static bool capturing(cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); return false; }
  return cs != cudaStreamCaptureStatusNone;
}
static int enter_stream(mjb_model *m, cudaStream_t st) {
  if (capturing(st)) return MJB_OK;
  if (m->have_last && m->last_stream != st) CU(cudaStreamWaitEvent(st, m->ev_last, 0));
  return MJB_OK;
}
static int leave_stream(mjb_model *m, cudaStream_t st) {
  if (capturing(st)) return MJB_OK;
  CU(cudaEventRecord(m->ev_last, st));
  m->last_stream = st;
  m->have_last = true;
  return MJB_OK;
}

extern "C" int mjb_check_configs(mjb_model *m, const float *d_q, int64_t n, int32_t ldq, uint8_t *d_valid,
                                 uint32_t flags, void *stream) {
  int rc = check_common(m, flags);
  if (rc) return rc;
  if (n < 0 || ldq < m->H.nq) return fail(MJB_ERR_ARG, "bad n / ldq");
  if (n == 0) return MJB_OK;
  if (!d_q || !d_valid) return fail(MJB_ERR_ARG, "null device pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = enter_stream(m, st))) return rc;
  if ((rc = ensure_recheck(m, (size_t)n, st))) return rc;
  KArgs k = m->kargs;
  k.mode = MODE_DENSE; k.q = d_q; k.ldq = ldq; k.n = n; k.valid = d_valid; k.flags = flags;
  RArgs r = m->rargs;
  r.mode = MODE_DENSE; r.q = d_q; r.ldq = ldq; r.valid = d_valid;
  m->rows_total += n;
  if ((rc = launch_validity(m, k, r, st))) return rc;
  return leave_stream(m, st);
}

// Host buffers in, host mask out.  The rows are copied in chunks on a second stream while ONE
// validity launch is already consuming them: after every chunk the copy stream publishes the number
// of rows that have landed (an 8-byte copy from a pinned progress table), and a warp whose tile lies
// beyond that mark waits for it (validity_kernel, P0).  Copy and compute overlap without cutting
// the batch into several launches (each of which would end in a tail of idle SMs).
static const int64_t HOST_CHUNK_ROWS = 65536;

// cuStreamWriteValue64 through the runtime's driver entry point table (the library does not link libcuda): a
// stream memory operation publishes the copy's progress in about a microsecond, where an 8-byte cudaMemcpyAsync
// through the copy engine cost 5 (sixteen of them per million rows: 0.82 -> 0.7x ms for the 36 MB).
typedef int (*stream_write64_fn)(void *stream, unsigned long long dptr, unsigned long long value, unsigned int flags);
static stream_write64_fn stream_write64() {
  static stream_write64_fn fn = []() -> stream_write64_fn {
    if (getenv("MJB_NO_STREAM_WRITE")) return nullptr;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuStreamWriteValue64", &p, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    return (stream_write64_fn)p;
  }();
  return fn;
}

extern "C" int mjb_check_configs_host(mjb_model *m, const float *h_q, int64_t n, uint8_t *h_valid, uint32_t flags) {
  int rc = check_common(m, flags);
  if (rc) return rc;
  if (n < 0) return fail(MJB_ERR_ARG, "bad n");
  if (n == 0) return MJB_OK;
  if (!h_q || !h_valid) return fail(MJB_ERR_ARG, "null host pointer");
  const int nq = m->H.nq;
  cudaStream_t st = m->own_stream;
  if ((rc = enter_stream(m, st))) return rc;
  if ((size_t)n > m->stage_rows) {
    CU(cudaStreamSynchronize(st));
    cudaFree(m->d_stage_q); cudaFree(m->d_stage_v);
    m->d_stage_q = nullptr; m->d_stage_v = nullptr;
    size_t cap = std::max<size_t>((size_t)n, 4096);
    CU(cudaMalloc((void **)&m->d_stage_q, cap * nq * sizeof(float)));
    CU(cudaMalloc((void **)&m->d_stage_v, cap));
    m->stage_rows = cap;
  }
  const int64_t nchunk = (n + HOST_CHUNK_ROWS - 1) / HOST_CHUNK_ROWS;
  if (nchunk < 2) {  // small batch: nothing to overlap
    CU(cudaMemcpyAsync(m->d_stage_q, h_q, (size_t)n * nq * sizeof(float), cudaMemcpyHostToDevice, st));
    rc = mjb_check_configs(m, m->d_stage_q, n, nq, m->d_stage_v, flags, st);
    if (rc) return rc;
    CU(cudaMemcpyAsync(h_valid, m->d_stage_v, (size_t)n, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return MJB_OK;
  }
  if (!m->copy_stream) {
    CU(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&m->ev_ready_reset, cudaEventDisableTiming));
    CU(cudaMalloc((void **)&m->d_rows_ready, sizeof(unsigned long long)));
  }
  if ((size_t)nchunk > m->progress_cap) {
    CU(cudaStreamSynchronize(m->copy_stream));
    if (m->h_progress) cudaFreeHost(m->h_progress);
    m->h_progress = nullptr;
    size_t cap = std::max<size_t>((size_t)nchunk, 64);
    CU(cudaMallocHost((void **)&m->h_progress, cap * sizeof(unsigned long long)));
    m->progress_cap = cap;
  }
  if ((rc = ensure_recheck(m, (size_t)n, st))) return rc;
  // compute stream: reset the progress word.  Copy stream: every chunk in row order, each followed
  // by its progress mark.  Only then is the kernel launched -- with every copy already queued, a
  // launch that blocks (CUDA_LAUNCH_BLOCKING, compute-sanitizer, ncu replay) still sees its rows
  // arrive, and no error path can leave a kernel polling for rows that will never be sent.
  CU(cudaMemsetAsync(m->d_rows_ready, 0, sizeof(unsigned long long), st));
  CU(cudaEventRecord(m->ev_ready_reset, st));
  CU(cudaStreamWaitEvent(m->copy_stream, m->ev_ready_reset, 0));
  for (int64_t c = 0; c < nchunk; c++) {
    const int64_t r0 = c * HOST_CHUNK_ROWS, r1 = std::min<int64_t>(n, r0 + HOST_CHUNK_ROWS);
    m->h_progress[c] = (unsigned long long)r1;
    CU(cudaMemcpyAsync(m->d_stage_q + (size_t)r0 * nq, h_q + (size_t)r0 * nq, (size_t)(r1 - r0) * nq * sizeof(float),
                       cudaMemcpyHostToDevice, m->copy_stream));
    stream_write64_fn w64 = stream_write64();
    if (!w64 || w64((void *)m->copy_stream, (unsigned long long)(uintptr_t)m->d_rows_ready, (unsigned long long)r1, 0) != 0)
      CU(cudaMemcpyAsync(m->d_rows_ready, &m->h_progress[c], sizeof(unsigned long long), cudaMemcpyHostToDevice, m->copy_stream));
  }
  // Only the first kernel of a launch can work on rows as they arrive (it polls the progress word); the
  // pipeline's later kernels start when ALL rows of their launch are on the device.  Cutting the batch into a
  // few launches lets slice s be culled and decided while slices s+1.. are still on the bus.
  // (Measured on B200, 1M Franka rows, 36 MB from pinned memory in 0.77 ms, kernels 1.02 ms: 1 launch 1.73 ms,
  // 2 equal launches 1.74 ms, 4 launches 2.07 ms -- every launch of the pipeline costs 0.25 ms of ramp and tail,
  // which eats what the overlap gives.  What does pay is a SHORT first launch, see MJB_HOST_FIRST below; from 4M
  // rows on equal launches of about 2M rows do too (4M rows: 6.38 -> 5.62 ms).  MJB_HOST_SLICES forces a count.)
  int64_t nslice = (m->split && m->split_min > 0) ? std::min<int64_t>(std::max<int64_t>(n / 2000000, 1), 4) : 1;
  if (m->split && m->split_min > 0 && getenv("MJB_HOST_SLICES"))
    nslice = std::min<int64_t>(std::max<int64_t>(atoi(getenv("MJB_HOST_SLICES")), 1), std::max<int64_t>(n / (int64_t)m->split_min, 1));
  // slice boundaries (multiples of the chunk size).  MJB_HOST_FIRST=f: two launches, the first with the fraction f
  // of the rows -- a short first launch is done before the copy is, and the second starts on rows that have
  // mostly landed (an equal split finishes its first half just AFTER the copy and gains nothing).
  std::vector<int64_t> cuts;
  const char *hf = getenv("MJB_HOST_FIRST");
  const double first = hf ? atof(hf) : 0.25;   // B200, 1M Franka rows: mask on the host after 1.55 ms (one launch), 1.41 (0.25), 1.46 (0.3-0.4), 1.62 (0.5)
  if (m->split && m->split_min > 0 && first > 0.0 && first < 1.0 && nslice == 1 && n >= 2 * (int64_t)m->split_min) {
    int64_t c0 = std::max<int64_t>((int64_t)(first * (double)nchunk + 0.5), 1) * HOST_CHUNK_ROWS;
    c0 = std::max<int64_t>(c0, (int64_t)align_up(m->split_min, (size_t)HOST_CHUNK_ROWS));
    if (n - c0 >= (int64_t)m->split_min) cuts.push_back(c0);
  } else {
    const int64_t chunks_per_slice = (nchunk + nslice - 1) / nslice;
    for (int64_t r = chunks_per_slice * HOST_CHUNK_ROWS; r < n; r += chunks_per_slice * HOST_CHUNK_ROWS) cuts.push_back(r);
  }
  cuts.push_back(n);
  m->rows_total += n;
  const bool trace = getenv("MJB_HOST_TRACE") != nullptr;
  std::vector<cudaEvent_t> tev;
  auto mark = [&](cudaStream_t s) { if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); tev.push_back(e); } };
  mark(st);
  mark(m->copy_stream);
  int64_t r0 = 0;
  for (size_t ci = 0; ci < cuts.size(); r0 = cuts[ci], ci++) {
    const int64_t r1 = cuts[ci];
    KArgs k = m->kargs;
    k.mode = MODE_DENSE; k.q = m->d_stage_q + (size_t)r0 * nq; k.ldq = nq; k.n = r1 - r0; k.valid = m->d_stage_v + r0; k.flags = flags;
    k.rows_ready = m->d_rows_ready; k.row0 = r0;
    RArgs r = m->rargs;
    r.mode = MODE_DENSE; r.q = k.q; r.ldq = nq; r.valid = k.valid;
    if ((rc = ensure_recheck(m, (size_t)(r1 - r0), st))) return rc;   // (sets the launch's row count; the buffers already fit)
    if ((rc = launch_validity(m, k, r, st))) return rc;
    mark(st);
  }
  CU(cudaMemcpyAsync(h_valid, m->d_stage_v, (size_t)n, cudaMemcpyDeviceToHost, st));
  mark(st);
  CU(cudaStreamSynchronize(st));
  CU(cudaStreamSynchronize(m->copy_stream));
  if (trace) {
    fprintf(stderr, "[mjb] host batch %lld rows, %lld slices:", (long long)n, (long long)cuts.size());
    for (size_t i = 1; i < tev.size(); i++) { float ms = 0; cudaEventElapsedTime(&ms, tev[0], tev[i]); fprintf(stderr, " %.3f", ms); }
    fprintf(stderr, " ms (copies done, each slice done, mask on host)\n");
    for (auto e : tev) cudaEventDestroy(e);
  }
  return leave_stream(m, st);
}

extern "C" int mjb_fk(mjb_model *m, const float *d_q, int64_t n, int32_t ldq, float *d_xpos, float *d_xquat, void *stream) {
  int rc = check_common(m, MJB_CHECK_COLLISION);
  if (rc) return rc;
  if (n < 0 || ldq < m->H.nq) return fail(MJB_ERR_ARG, "bad n / ldq");
  if (n == 0) return MJB_OK;
  if (!d_q || !d_xpos || !d_xquat) return fail(MJB_ERR_ARG, "null device pointer");
  FArgs f = m->fargs;
  f.q = d_q; f.ldq = ldq; f.n = n; f.xpos = d_xpos; f.xquat = d_xquat;
  cudaStream_t st = (cudaStream_t)stream;
  fk_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(f);
  CU(cudaGetLastError());
  m->launches++;
  return MJB_OK;
}

extern "C" int mjb_check_edges(mjb_model *m, const float *d_q0, const float *d_q1, int64_t ne, int32_t ldq, float step,
                               uint8_t *d_valid, int32_t *d_first_bad, uint32_t flags, void *stream) {
  int rc = check_common(m, flags);
  if (rc) return rc;
  // reference: raise ValueError("`step_dist` must be > 0") (src/mjpl/planning/utils.py:203-204)
  if (!(step > 0.0f)) return fail(MJB_ERR_ARG, "`step_dist` must be > 0");
  if (ne < 0 || ldq < m->H.nq) return fail(MJB_ERR_ARG, "bad ne / ldq");
  if (ne == 0) return MJB_OK;
  if (!d_q0 || !d_q1 || !d_valid) return fail(MJB_ERR_ARG, "null device pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = enter_stream(m, st))) return rc;
  if ((rc = ensure_edge_buffers(m, (size_t)ne, st))) return rc;
  const int nq = m->H.nq;
  edge_count_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, st>>>(d_q0, d_q1, ne, nq, ldq, step, m->d_edge_count, m->d_first_bad);
  CU(cudaGetLastError());
  CU(cudaMemsetAsync(m->d_edge_count + ne, 0, sizeof(long long), st));
  size_t tb = m->cub_bytes;
  CU(cub::DeviceScan::ExclusiveSum(m->d_cub, tb, m->d_edge_count, m->d_edge_prefix, (int)(ne + 1), st));
  // size the fp64 work list for the worst case (every waypoint uncertain): one small D2H read
  // of the waypoint total -- this entry point synchronises `stream` once here.
  long long total = 0;
  CU(cudaMemcpyAsync(&total, m->d_edge_prefix + ne, sizeof total, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  // Waypoints of one edge are neighbours in configuration space: the lanes of a warp agree on which pairs are near,
  // and the single kernel skips a pair no lane of the warp needs, while the pipeline's level 0 tests every group
  // pair for every row.  For models whose narrow phase is cheap (primitives and boxes only) that wins: UR5e, 100k
  // edges / 12.1M waypoints 4.4 ms in the single kernel, 7.0 ms in the pipeline; with hulls to scan it is the other
  // way round (Franka obstacle scene 23.4 vs 6.8 ms).
  if ((rc = ensure_recheck(m, (size_t)std::max<long long>(total, 1), st, /*may_split=*/!m->light_narrow))) return rc;
  m->rows_total += total;
  KArgs k = m->kargs;
  k.mode = MODE_EDGES; k.q0 = d_q0; k.q1 = d_q1; k.ldq = ldq; k.edge_prefix = m->d_edge_prefix; k.nedge = ne;
  k.step = step; k.first_bad = m->d_first_bad; k.flags = flags; k.n = 0;
  RArgs r = m->rargs;
  r.mode = MODE_EDGES; r.q0 = d_q0; r.q1 = d_q1; r.ldq = ldq; r.edge_prefix = m->d_edge_prefix; r.nedge = ne;
  r.step = step; r.first_bad = m->d_first_bad;
  m->launches += 2;
  const bool rowmask = m->use_split && (flags & F_COLLISION);   // the pipeline keeps a byte per waypoint, see F_ROWMASK
  if (rowmask) { k.valid = m->d_rowmask; k.flags |= F_ROWMASK; }
  if ((rc = launch_validity(m, k, r, st))) return rc;
  if (rowmask) {
    edge_rowmask_kernel<<<(unsigned)((ne * 32 + 255) / 256), 256, 0, st>>>(ne, m->d_edge_prefix, m->d_rowmask, m->d_first_bad);
    CU(cudaGetLastError());
    m->launches++;
  }
  edge_finalize_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, st>>>(ne, m->d_first_bad, d_valid, d_first_bad);
  CU(cudaGetLastError());
  m->launches++;
  return leave_stream(m, st);
}

extern "C" int mjb_check_sweep(mjb_model *m, uint64_t seed, int64_t row0, int64_t n, uint8_t *d_valid, uint32_t flags, void *stream) {
  int rc = check_common(m, flags);
  if (rc) return rc;
  if (n < 0 || row0 < 0) return fail(MJB_ERR_ARG, "bad n / row0");
  if (n == 0) return MJB_OK;
  if (!d_valid) return fail(MJB_ERR_ARG, "null device pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = enter_stream(m, st))) return rc;
  if ((rc = ensure_recheck(m, (size_t)n, st))) return rc;
  KArgs k = m->kargs;
  k.mode = MODE_SWEEP; k.seed = seed; k.row0 = row0; k.n = n; k.valid = d_valid; k.flags = flags; k.ldq = m->H.nq;
  RArgs r = m->rargs;
  r.mode = MODE_SWEEP; r.seed = seed; r.row0 = row0; r.valid = d_valid;
  m->rows_total += n;
  if ((rc = launch_validity(m, k, r, st))) return rc;
  return leave_stream(m, st);
}

extern "C" int mjb_sweep_rows(mjb_model *m, uint64_t seed, int64_t row0, int64_t n, float *d_q, void *stream) {
  int rc = check_common(m, MJB_CHECK_COLLISION);
  if (rc) return rc;
  if (n < 0 || row0 < 0) return fail(MJB_ERR_ARG, "bad n / row0");
  if (n == 0) return MJB_OK;
  if (!d_q) return fail(MJB_ERR_ARG, "null device pointer");
  cudaStream_t st = (cudaStream_t)stream;
  long long total = n * m->H.nq;
  sweep_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(m->kargs.fk, seed, row0, n, d_q);
  CU(cudaGetLastError());
  m->launches++;
  return MJB_OK;
}

extern "C" int mjb_kernel_timing(mjb_model *m, int enable, double *ms4, int64_t *launches) {
  if (!m) return fail(MJB_ERR_ARG, "null model");
  CU(cudaSetDevice(m->device));
  if (ms4) {
    CU(cudaDeviceSynchronize());
    ms4[0] = ms4[1] = ms4[2] = ms4[3] = 0.0;
    for (size_t i = 0; i + TEV <= m->tev_used; i += TEV) {
      for (int k = 0; k < 4; k++) {
        float t = 0;
        CU(cudaEventElapsedTime(&t, m->tev[i + k], m->tev[i + k + 1]));
        ms4[k] += t;
      }
    }
    if (launches) *launches = (int64_t)(m->tev_used / TEV);
    m->tev_used = 0;
  }
  m->timing = enable != 0;
  if (m->timing && m->tev.empty()) {
    m->tev.resize(TEV * 2048);
    m->tev_split.assign(2048, 0);
    for (auto &e : m->tev) CU(cudaEventCreate(&e));
  }
  if (!m->timing && !m->tev.empty()) {   // events only live while timing is on
    CU(cudaDeviceSynchronize());
    for (cudaEvent_t e : m->tev) cudaEventDestroy(e);
    cudaGetLastError();
    m->tev.clear();
    m->tev_used = 0;
  }
  return MJB_OK;
}

extern "C" int mjb_get_stats(mjb_model *m, mjb_stats *out) {
  if (!m || !out) return fail(MJB_ERR_ARG, "null argument");
  CU(cudaSetDevice(m->device));
  CU(cudaDeviceSynchronize());
  unsigned long long c[C_NCOUNTERS];
  CU(cudaMemcpy(c, m->d_counters, sizeof c, cudaMemcpyDeviceToHost));
  out->rows = (int64_t)c[C_ROWS];
  out->narrow_items = (int64_t)c[C_ITEMS];
  out->uncertain_rows = (int64_t)c[C_UNCERTAIN];
  out->queue_overflow = (int64_t)c[C_OVERFLOW];
  out->launches = m->launches;
#ifdef VK_STATS
  out->launches = (int64_t)c[C_TRIPS];  // debug build: GJK loop trips
  out->uncertain_rows = (int64_t)c[C_HIST + 8];  // debug build: narrow-phase passes
  fprintf(stderr, "[vk_stats] trips by busy lanes <=4 / <=8 / <=16 / <=32: %llu %llu %llu %llu; lane-trips: %llu %llu %llu %llu\n",
          c[C_HIST], c[C_HIST + 1], c[C_HIST + 2], c[C_HIST + 3], c[C_HIST + 4], c[C_HIST + 5], c[C_HIST + 6], c[C_HIST + 7]);
#endif
  return MJB_OK;
}

extern "C" int mjb_reset_stats(mjb_model *m) {
  if (!m) return fail(MJB_ERR_ARG, "null argument");
  CU(cudaSetDevice(m->device));
  CU(cudaDeviceSynchronize());
  CU(cudaMemset(m->d_counters, 0, C_NCOUNTERS * sizeof(unsigned long long)));
  m->rows_total = 0;
  m->launches = 0;
  return MJB_OK;
}

extern "C" int mjb_set_chain_hint(mjb_model *m, int64_t expected_rows) {
  if (!m || expected_rows < 0) return fail(MJB_ERR_ARG, "null model / negative hint");
  m->chain_hint = (size_t)expected_rows;
  return MJB_OK;
}

extern "C" int mjb_tree_paths(const int64_t *d_parent, int64_t cap, const int64_t *d_rows, const int64_t *d_first, int64_t n,
                              int64_t max_depth, int64_t *d_steps, int64_t *d_len, void *stream) {
  if (n < 0 || cap < 1 || max_depth < 1) return fail(MJB_ERR_ARG, "bad n / cap / max_depth");
  if (n == 0) return MJB_OK;
  if (!d_parent || !d_first || !d_steps || !d_len) return fail(MJB_ERR_ARG, "null device pointer");
  tree_paths_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      (const long long *)d_parent, (long long)cap, (const long long *)d_rows, (const long long *)d_first, (long long)n,
      (long long)max_depth, (long long *)d_steps, (long long *)d_len);
  CU(cudaGetLastError());
  return MJB_OK;
}

extern "C" int mjb_nearest_batch(const double *d_nodes, int64_t cap, int32_t nq, const int64_t *d_count, const int64_t *d_rows,
                                 const double *d_targets, int64_t n, int64_t *d_out, void *stream) {
  if (n < 0 || cap < 1 || nq < 1 || nq > MAX_JNT) return fail(MJB_ERR_ARG, "bad n / cap / nq");
  if (n == 0) return MJB_OK;
  if (!d_nodes || !d_count || !d_targets || !d_out) return fail(MJB_ERR_ARG, "null device pointer");
  cudaStream_t st = (cudaStream_t)stream;
  nearest_kernel<<<(unsigned)n, NEAREST_THREADS, 0, st>>>(d_nodes, (long long)cap, nq, (const long long *)d_count,
                                                                  (const long long *)d_rows, d_targets, (long long)n,
                                                                  (long long *)d_out);
  CU(cudaGetLastError());
  return MJB_OK;
}

// grow the per-handle edge / chain buffers (count, prefix, first_bad, cub scratch) to `ne` entries
static int ensure_edge_buffers(mjb_model *m, size_t ne, cudaStream_t st) {
  if (ne + 1 <= m->edge_cap) return MJB_OK;
  CU(cudaStreamSynchronize(st));
  cudaFree(m->d_edge_count); cudaFree(m->d_edge_prefix); cudaFree(m->d_first_bad); cudaFree(m->d_cub);
  m->d_edge_count = m->d_edge_prefix = nullptr; m->d_first_bad = nullptr; m->d_cub = nullptr;
  size_t cap = std::max<size_t>(ne + 1, 1 << 16);
  CU(cudaMalloc((void **)&m->d_edge_count, cap * sizeof(long long)));
  CU(cudaMalloc((void **)&m->d_edge_prefix, cap * sizeof(long long)));
  CU(cudaMalloc((void **)&m->d_first_bad, cap * sizeof(int)));
  size_t tb = 0;
  CU(cub::DeviceScan::ExclusiveSum(nullptr, tb, m->d_edge_count, m->d_edge_prefix, (int)cap, st));
  CU(cudaMalloc(&m->d_cub, tb + 256));
  m->cub_bytes = tb + 256;
  m->edge_cap = cap;
  return MJB_OK;
}

extern "C" int mjb_rrt_extend(mjb_model *m, double *d_nodes, int64_t *d_parent, int64_t *d_count, int64_t cap,
                              const int64_t *d_slots, const double *d_targets, int64_t n, double eps, int32_t kcap,
                              uint32_t flags, double *d_reached, int64_t *d_last, void *stream) {
  return mjb_rrt_extend_masked(m, d_nodes, d_parent, d_count, cap, d_slots, d_targets, nullptr, n, eps, kcap, flags, d_reached,
                               d_last, stream);
}

extern "C" int mjb_rrt_extend_masked(mjb_model *m, double *d_nodes, int64_t *d_parent, int64_t *d_count, int64_t cap,
                                     const int64_t *d_slots, const double *d_targets, const uint8_t *d_active, int64_t n,
                                     double eps, int32_t kcap, uint32_t flags, double *d_reached, int64_t *d_last, void *stream) {
  int rc = check_common(m, flags);
  if (rc) return rc;
  // reference: ValueError("`max_step_dist` must be > 0.0") (src/mjpl/planning/utils.py:179-180)
  if (!(eps > 0.0)) return fail(MJB_ERR_ARG, "`max_step_dist` must be > 0.0");
  if (n < 0 || cap < 1 || kcap < 1) return fail(MJB_ERR_ARG, "bad n / cap / kcap");
  if (n == 0) return MJB_OK;
  if (!d_nodes || !d_parent || !d_count || !d_targets || !d_reached || !d_last) return fail(MJB_ERR_ARG, "null device pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int nq = m->H.nq;
  if ((rc = enter_stream(m, st))) return rc;
  if ((rc = ensure_edge_buffers(m, (size_t)n, st))) return rc;
  if ((size_t)n > m->chain_cap) {
    CU(cudaStreamSynchronize(st));
    cudaFree(m->d_chain_near); cudaFree(m->d_chain_nn);
    m->d_chain_near = nullptr; m->d_chain_nn = nullptr;
    size_t c = std::max<size_t>((size_t)n, 4096);
    CU(cudaMalloc((void **)&m->d_chain_near, c * nq * sizeof(double)));
    CU(cudaMalloc((void **)&m->d_chain_nn, c * sizeof(long long)));
    m->chain_cap = c;
  }
  // worst case, no host read-back; the actual chains are short, so always the single kernel
  // (expected rows: ~24 chain steps per query on the Franka scenes; only picks the kernel instance)
  if ((rc = ensure_recheck(m, (size_t)n * (size_t)kcap, st, false, m->chain_hint ? m->chain_hint : (size_t)n * 24))) return rc;
  // 1. nearest node of every query's tree   2. chain lengths + prefix sums
  nearest_kernel<<<(unsigned)n, NEAREST_THREADS, 0, st>>>(d_nodes, (long long)cap, nq, (const long long *)d_count,
                                                                  (const long long *)d_slots, d_targets, (long long)n, m->d_chain_nn, d_active);
  chain_setup_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d_nodes, (long long)cap, nq, (const long long *)d_slots, m->d_chain_nn,
                                                                 d_targets, (long long)n, eps, kcap, m->d_chain_near, m->d_edge_count,
                                                                 m->d_first_bad, d_active);
  CU(cudaGetLastError());
  CU(cudaMemsetAsync(m->d_edge_count + n, 0, sizeof(long long), st));
  size_t tb = m->cub_bytes;
  CU(cub::DeviceScan::ExclusiveSum(m->d_cub, tb, m->d_edge_count, m->d_edge_prefix, (int)(n + 1), st));
  // 3. validity of every chain step (rows generated on the device), first failing step per chain
  KArgs k = m->kargs;
  k.mode = MODE_CHAINS; k.c0 = m->d_chain_near; k.c1 = d_targets; k.ceps = eps; k.edge_prefix = m->d_edge_prefix; k.nedge = n;
  k.first_bad = m->d_first_bad; k.flags = flags; k.ldq = nq;
  RArgs r = m->rargs;
  r.mode = MODE_CHAINS; r.c0 = m->d_chain_near; r.c1 = d_targets; r.ceps = eps; r.edge_prefix = m->d_edge_prefix; r.nedge = n;
  r.first_bad = m->d_first_bad; r.ldq = nq;
  if ((rc = launch_validity(m, k, r, st))) return rc;
  // 4. append the valid prefixes
  chain_append_kernel<<<(unsigned)((n * 32 + 127) / 128), 128, 0, st>>>(
      d_nodes, (long long *)d_parent, (long long *)d_count, (long long)cap, nq, (const long long *)d_slots, m->d_chain_nn,
      m->d_chain_near, d_targets, m->d_edge_count, m->d_first_bad, (long long)n, eps, d_reached, (long long *)d_last,
      m->d_counters + C_OVERFLOW);
  CU(cudaGetLastError());
  m->launches += 4;
  return leave_stream(m, st);
}

// ---- PoseConstraint --------------------------------------------------------------------------------
static int fill_pose_spec(mjb_model *m, const mjb_pose_spec *in, PoseSpec &sp) {
  std::string err;
  if (!vkb::make_pose_spec(m->H, in, sp, err)) return fail(MJB_ERR_ARG, err);
  return MJB_OK;
}

static int launch_pose(mjb_model *m, const mjb_pose_spec *spec, const double *d_q_old, const double *d_q, int64_t n, int project,
                       int32_t max_iters, double *d_q_out, uint8_t *d_ok, int32_t *d_iters, void *stream, const uint8_t *d_mask = nullptr) {
  int rc = check_common(m, MJB_CHECK_LIMITS);
  if (rc) return rc;
  if (n < 0) return fail(MJB_ERR_ARG, "bad n");
  PoseArgs a;
  memset(&a, 0, sizeof a);
  if ((rc = fill_pose_spec(m, spec, a.spec))) return rc;
  if (n == 0) return MJB_OK;
  if (!d_q || !d_ok || (project && (!d_q_old || !d_q_out))) return fail(MJB_ERR_ARG, "null device pointer");
  a.fk = m->d_fk64; a.nslot = m->H.nslot; a.q_old = d_q_old; a.q = d_q; a.n = n; a.project = project;
  a.max_iters = max_iters > 0 ? max_iters : 1000; a.q_out = d_q_out; a.ok = d_ok; a.iters = d_iters; a.mask = d_mask;
  cudaStream_t st = (cudaStream_t)stream;
  pose_kernel<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(a);
  CU(cudaGetLastError());
  m->launches++;
  return MJB_OK;
}

extern "C" int mjb_pose_valid(mjb_model *m, const mjb_pose_spec *spec, const double *d_q, int64_t n, uint8_t *d_valid, void *stream) {
  return launch_pose(m, spec, nullptr, d_q, n, 0, 0, nullptr, d_valid, nullptr, stream);
}

extern "C" int mjb_pose_project(mjb_model *m, const mjb_pose_spec *spec, const double *d_q_old, const double *d_q, int64_t n,
                                int32_t max_iters, double *d_q_out, uint8_t *d_ok, int32_t *d_iters, void *stream) {
  return launch_pose(m, spec, d_q_old, d_q, n, 1, max_iters, d_q_out, d_ok, d_iters, stream);
}

extern "C" int mjb_ik_solve(mjb_model *m, const mjb_ik_spec *spec, const double *d_target_pos, const double *d_target_quat,
                            const double *d_q_init, int64_t n, double *d_q_out, uint8_t *d_ok, int32_t *d_iters, double *d_err,
                            void *stream) {
  int rc = check_common(m, MJB_CHECK_LIMITS);
  if (rc) return rc;
  if (n < 0) return fail(MJB_ERR_ARG, "bad n");
  IkArgs a;
  memset(&a, 0, sizeof a);
  std::string err;
  if (!vkb::make_ik_spec(m->H, spec, a.spec, err)) return fail(MJB_ERR_ARG, err);
  if (n == 0) return MJB_OK;
  if (!d_target_pos || !d_target_quat || !d_q_init || !d_q_out || !d_ok) return fail(MJB_ERR_ARG, "null device pointer");
  a.fk = m->d_fk64; a.nslot = m->H.nslot; a.tpos = d_target_pos; a.tquat = d_target_quat; a.q_init = d_q_init; a.n = n;
  a.q_out = d_q_out; a.ok = d_ok; a.iters = d_iters; a.err = d_err;
  cudaStream_t st = (cudaStream_t)stream;
  ik_kernel<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(a);
  CU(cudaGetLastError());
  m->launches++;
  return MJB_OK;
}

extern "C" int mjb_site_pose(mjb_model *m, int32_t site_bodyid, const double *site_pos, const double *site_quat, const double *d_q,
                             int64_t n, double *d_pos, double *d_quat, void *stream) {
  int rc = check_common(m, MJB_CHECK_LIMITS);
  if (rc) return rc;
  if (n < 0 || !site_pos || !site_quat) return fail(MJB_ERR_ARG, "bad argument");
  mjb_pose_spec in;
  memset(&in, 0, sizeof in);
  in.site_bodyid = site_bodyid;
  memcpy(in.site_pos, site_pos, sizeof in.site_pos);
  memcpy(in.site_quat, site_quat, sizeof in.site_quat);
  in.ref_quat[0] = 1.0; in.q_step = 1.0;
  PoseSpec sp;
  if ((rc = fill_pose_spec(m, &in, sp))) return rc;
  if (n == 0) return MJB_OK;
  if (!d_q || !d_pos || !d_quat) return fail(MJB_ERR_ARG, "null device pointer");
  cudaStream_t st = (cudaStream_t)stream;
  site_pose_kernel<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(m->d_fk64, m->H.nslot, sp, d_q, (long long)n, d_pos, d_quat);
  CU(cudaGetLastError());
  m->launches++;
  return MJB_OK;
}

// ---- FP32 FMA throughput of the device (roofline denominator) ---------------
// Every thread runs 8 independent FFMA chains; nothing else is in the loop, so the kernel runs at the
// FMA pipes' issue rate.  The result is written out so the compiler cannot drop the chains.
__global__ void __launch_bounds__(256) fma_peak_kernel(float *out, int iters, float a, float b) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

extern "C" int mjb_fma_peak(double *tflops, double *ms_out) {
  if (!tflops) return fail(MJB_ERR_ARG, "null argument");
  int dev = 0, sms = 0;
  CU(cudaGetDevice(&dev));
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int ctas = sms * 8, threads = 256, iters = 4096;
  float *out = nullptr;
  CU(cudaMalloc((void **)&out, (size_t)ctas * threads * sizeof(float)));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  double best = 1e30;
  for (int rep = 0; rep < 5; rep++) {   // first pass warms up; best of the rest
    CU(cudaEventRecord(e0, 0));
    fma_peak_kernel<<<ctas, threads>>>(out, iters, 0.999f, 1e-3f);
    CU(cudaEventRecord(e1, 0));
    CU(cudaEventSynchronize(e1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  CU(cudaGetLastError());
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  const double flops = 2.0 * 128.0 * (double)iters * (double)ctas * (double)threads;
  *tflops = flops / (best * 1e-3) / 1e12;
  if (ms_out) *ms_out = best;
  return MJB_OK;
}

// ---- bi-RRT iteration helpers (device-resident planner state; mjpl_b200/planning/batched_rrt.py) ----
extern "C" int mjb_rrt_sample(uint64_t seed, const int64_t *d_counters, int64_t nslots, int32_t nq, const double *d_q_init,
                              const double *d_q_goal, const uint8_t *d_plan_mask, const double *d_lo, const double *d_hi,
                              double goal_bias, const uint8_t *d_active, double *d_targets, void *stream) {
  if (nslots < 0 || nq < 1 || nq > MAX_JNT) return fail(MJB_ERR_ARG, "bad nslots / nq");
  // reference: ValueError("`goal_biasing_probability` must be within [0.0, 1.0].") (src/mjpl/planning/rrt.py:57-58)
  if (!(goal_bias >= 0.0 && goal_bias <= 1.0)) return fail(MJB_ERR_ARG, "`goal_biasing_probability` must be within [0.0, 1.0].");
  if (nslots == 0) return MJB_OK;
  if (!d_counters || !d_q_init || !d_q_goal || !d_plan_mask || !d_lo || !d_hi || !d_active || !d_targets)
    return fail(MJB_ERR_ARG, "null device pointer");
  rrt_sample_kernel<<<(unsigned)((nslots + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      seed, (const long long *)d_counters, (long long)nslots, nq, d_q_init, d_q_goal, d_plan_mask, d_lo, d_hi, goal_bias, d_active,
      d_targets);
  CU(cudaGetLastError());
  return MJB_OK;
}

extern "C" int mjb_rrt_meet(int64_t nslots, int32_t nq, const double *d_qa, const double *d_qb, const int64_t *d_ia,
                            const int64_t *d_ib, int64_t max_age, uint8_t *d_active, int64_t *d_age, int64_t *d_res_start,
                            int64_t *d_res_goal, int64_t *d_counters, void *stream) {
  if (nslots < 0 || nq < 1) return fail(MJB_ERR_ARG, "bad nslots / nq");
  if (nslots == 0) return MJB_OK;
  if (!d_qa || !d_qb || !d_ia || !d_ib || !d_active || !d_age || !d_res_start || !d_res_goal || !d_counters)
    return fail(MJB_ERR_ARG, "null device pointer");
  cudaStream_t st = (cudaStream_t)stream;
  rrt_meet_kernel<<<(unsigned)((nslots + 127) / 128), 128, 0, st>>>((long long)nslots, nq, d_qa, d_qb, (const long long *)d_ia,
                                                                    (const long long *)d_ib, (long long)max_age, d_active,
                                                                    (long long *)d_age, (long long *)d_res_start,
                                                                    (long long *)d_res_goal, (long long *)d_counters);
  rrt_advance_kernel<<<1, 1, 0, st>>>((long long *)d_counters);
  CU(cudaGetLastError());
  return MJB_OK;
}

// ---- CBiRRT with a projecting constraint: one tick for every slot (vk_kernels.cuh: TickState) ----------
extern "C" int mjb_cbirrt_tick(mjb_model *m, const mjb_cbirrt_state *cs, const mjb_pose_spec *pose, int32_t pose_max_iters,
                               uint32_t flags, void *stream) {
  if (!m || !cs || !pose) return fail(MJB_ERR_ARG, "null argument");
  if (cs->nq != m->H.nq || cs->nslots < 0 || cs->cap < 2) return fail(MJB_ERR_ARG, "bad nq / nslots / cap");
  // reference: ValueError texts of RRT.__init__ (src/mjpl/planning/rrt.py:53-58)
  if (!(cs->eps > 0.0)) return fail(MJB_ERR_ARG, "`epsilon` must be > 0.0");
  if (!(cs->goal_bias >= 0.0 && cs->goal_bias <= 1.0)) return fail(MJB_ERR_ARG, "`goal_biasing_probability` must be within [0.0, 1.0].");
  if (cs->nslots == 0) return MJB_OK;
  TickState t;
  memset(&t, 0, sizeof t);
  t.nslots = cs->nslots; t.cap = cs->cap; t.nq = cs->nq; t.eps = cs->eps; t.goal_bias = cs->goal_bias; t.seed = cs->seed;
  t.max_age = cs->max_age; t.check_limits_before = cs->check_limits_before;
  t.q_init = cs->q_init; t.q_goal = cs->q_goal; t.plan_mask = cs->plan_mask; t.lo = cs->lo; t.hi = cs->hi;
  for (int k = 0; k < 2; k++) { t.nodes[k] = cs->nodes[k]; t.parent[k] = (long long *)cs->parent[k]; t.count[k] = (long long *)cs->count[k]; }
  t.phase = cs->phase; t.swapped = cs->swapped; t.age = (long long *)cs->age;
  t.target = cs->target; t.tip = cs->tip; t.qa = cs->qa; t.last = (long long *)cs->last; t.ia = (long long *)cs->ia;
  t.cand = cs->cand; t.cand32 = cs->cand32; t.proj = cs->proj; t.proj_ok = cs->proj_ok; t.valid = cs->valid; t.stepping = cs->stepping;
  t.res_start = (long long *)cs->res_start; t.res_goal = (long long *)cs->res_goal; t.counters = (long long *)cs->counters;
  const void *need[] = {t.q_init, t.q_goal, t.plan_mask, t.lo, t.hi, t.nodes[0], t.nodes[1], t.parent[0], t.parent[1], t.count[0],
                        t.count[1], t.phase, t.swapped, t.age, t.target, t.tip, t.qa, t.last, t.ia, t.cand, t.cand32, t.proj,
                        t.proj_ok, t.valid, t.stepping, t.res_start, t.res_goal, t.counters};
  for (const void *p : need) if (!p) return fail(MJB_ERR_ARG, "null device pointer in mjb_cbirrt_state");
  cudaStream_t st = (cudaStream_t)stream;
  const long long S = cs->nslots;
  tick_begin_kernel<<<(unsigned)((S * 32 + 127) / 128), 128, 0, st>>>(t);
  CU(cudaGetLastError());
  // the projecting constraint's apply() for the slots that proposed a step (pose_constraint.py:78-91)
  int rc = launch_pose(m, pose, t.tip, t.cand, S, 1, pose_max_iters, t.proj, t.proj_ok, nullptr, stream, t.stepping);
  if (rc) return rc;
  tick_rows_kernel<<<(unsigned)((S * cs->nq + 255) / 256), 256, 0, st>>>(t);
  CU(cudaGetLastError());
  // the constraints after it, and the re-validation of apply_constraints (constraint/utils.py:38-43): the
  // projection has enforced the joint limits and the pose in fp64; what is left is the collision check
  if (flags & MJB_CHECK_COLLISION) {
    rc = mjb_check_configs(m, t.cand32, S, cs->nq, t.valid, MJB_CHECK_COLLISION, stream);
    if (rc) return rc;
  } else {
    CU(cudaMemsetAsync(t.valid, 1, (size_t)S, st));
  }
  tick_end_kernel<<<(unsigned)((S + 127) / 128), 128, 0, st>>>(t);
  tick_advance_kernel<<<1, 1, 0, st>>>(t.counters);
  CU(cudaGetLastError());
  m->launches += 5;
  return MJB_OK;
}

// ---- signed distance per row --------------------------------------------------------------------------
extern "C" int mjb_min_distance(mjb_model *m, const float *d_q, int64_t n, int32_t ldq, double far_cap, double *d_dist,
                                int32_t *d_pair, void *stream) {
  int rc = check_common(m, MJB_CHECK_COLLISION);
  if (rc) return rc;
  if (n < 0 || ldq < m->H.nq) return fail(MJB_ERR_ARG, "bad n / ldq");
  if (n == 0) return MJB_OK;
  if (!d_q || !d_dist) return fail(MJB_ERR_ARG, "null device pointer");
  MArgs a;
  memset(&a, 0, sizeof a);
  a.fk = m->d_fk64; a.shapes = m->d_shapes64; a.verts = m->d_verts64; a.pairs = m->d_pairs; a.pair_rsum = m->d_rsum64;
  a.npair = (int)m->H.pairs.size(); a.nslot = m->H.nslot;
  a.q = d_q; a.ldq = ldq; a.n = n;
  a.far_cap = far_cap > 0.0 ? far_cap : MJB_DIST_FAR_DEFAULT; a.depth_cap = MJB_DEPTH_CAP;
  a.dist = d_dist; a.pair = d_pair;
  min_distance_kernel<<<(unsigned)((n + 63) / 64), 64, 0, (cudaStream_t)stream>>>(a);
  CU(cudaGetLastError());
  m->launches++;
  return MJB_OK;
}
