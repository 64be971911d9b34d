// C ABI (include/ccsdt_b200.h) and the per-GPU driver of the fused CCSD(T) triples path:
// block store in HBM, per-task panel staging, the task loop with staging/compute overlap, the
// deterministic energy reduction.  Replaces CCSD_T_Fused_Driver<T>::execute's loop
// (exachem/cc/ccsd_t/ccsd_t_fused_driver.hpp:368-483) and ccsd_t_fully_fused_none_df_none_task
// (exachem/cc/ccsd_t/ccsd_t_all_fused.hpp:77-286).
#include "../../include/ccsdt_b200.h"
#include "ccsdt_device.hpp"
#include "ccsdt_host.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <array>
#include <map>
#include <string>
#include <vector>

#include <dlfcn.h>

using namespace ccsdt;

namespace {

std::string g_create_error;

struct BlockRef {
  const double* ptr = nullptr;
  int64_t       stride[4] = {0, 0, 0, 0};
};

struct BlockKey {
  int      tensor;
  uint32_t b[4];
  bool     operator<(const BlockKey& o) const {
    if(tensor != o.tensor) return tensor < o.tensor;
    return std::lexicographical_compare(b, b + 4, o.b, o.b + 4);
  }
};

struct BlockEntry {
  double* dev;
  size_t  bytes;
  int64_t last_use;
};

// which index kind ('o'/'v') each dimension of a tensor has
const char* kKinds[5] = {"vo", "vvoo", "oovv", "ooov", "ovvv"};

struct StageBuf {
  PoolGeom     geom{};
  double*      s1_a = nullptr; // [9][THp][TPp]
  double*      s1_b = nullptr; // [9][THp][THp][TPp][TPp]
  GatherDesc*  d_descs = nullptr;
  GatherDesc*  h_descs = nullptr; // pinned
  int          desc_cap = 0;
  double*      d_partial = nullptr;
  int64_t      partial_cap = 0;
  uint32_t*    d_counter = nullptr; // dynamic box scheduler of the fused kernel
  int64_t      nparts = 0;          // entries of d_partial the reduction reads
  CUtensorMap  tmap[4];
  cudaEvent_t  staged = nullptr, done = nullptr, k0 = nullptr, k1 = nullptr, g0 = nullptr, g1 = nullptr;
  bool         timing_pending = false;
  TaskParams   params;
  int          grid = 0, consumer_warps = 0;
  size_t       smem = 0;
  double       eval_fraction = 1.0; // boxes evaluated / boxes of the tile (symmetry)
  cudaStream_t cs = nullptr;        // compute stream of this buffer (see run_task_list)
};

} // namespace

struct ccsdt_ctx {
  int            device = 0;
  cudaDeviceProp prop{};
  std::string    err;
  ccsdt_options  opt{};
  Space          sp;
  bool           have_space = false;
  std::vector<Task>    tasks;
  std::vector<int32_t> owner;
  int64_t              n_outer = 0;

  double*                        dense[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t                         dense_elems[5] = {0, 0, 0, 0, 0};
  bool                           dense_sparse_ok[5] = {false, false, false, false, false}; // spin-forbidden blocks are zero
  std::map<BlockKey, BlockEntry> blocks;
  size_t                         block_bytes = 0, block_budget = 0;
  int64_t                        use_clock = 0;
  ccsdt_fetch_fn                 fetch = nullptr;
  void*                          fetch_user = nullptr;
  double*                        h_fetch = nullptr; // pinned bounce buffer
  size_t                         h_fetch_cap = 0;
  bool                           synthetic = false;
  uint64_t                       seed = 0;

  double*      d_evl = nullptr;
  StageBuf     buf[2];
  bool         pools_ready = false;
  double*      d_task_energy = nullptr;
  int64_t      task_energy_cap = 0;
  uint32_t*    d_error = nullptr;
  cudaStream_t s_compute = nullptr, s_compute2 = nullptr, s_stage = nullptr;
  cudaEvent_t  ev_base = nullptr;   // start of the current run: kernel intervals are placed on its time line
  double       kernel_busy_until = 0.0; // end (ms after ev_base) of the union of fused-kernel intervals so far
  void*        encode_fn = nullptr;
  int64_t*     task_counter = nullptr; // process-shared dynamic task counter (NULL = static split)
  ccsdt_stats  stats{};
  int64_t      pending_h2d = 0; // bytes uploaded by ccsdt_put_* since the last run
  // asynchronous dense uploads (ccsdt_put_dense_async): the all-alpha blocks of every tensor travel on s_copy_a,
  // the other spin patterns on s_copy_b; tasks whose six tiles are all alpha only wait for the first
  cudaStream_t s_copy_a = nullptr, s_copy_b = nullptr;
  cudaEvent_t  ev_alpha[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}, ev_full[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  bool         upload_pending[5] = {false, false, false, false, false};
  // symmetry-reduced box lists (device), keyed by (nbox, brick, nbrick, sym): a handful per job
  struct BoxList {
    int32_t* dev = nullptr;
    int32_t  n   = 0;
  };
  std::map<std::array<int, 19>, BoxList> box_lists;

  int fail(const std::string& m, int code = 1) {
    err = m;
    return code;
  }
};

namespace {

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if(e_ != cudaSuccess)                                                                          \
      return ctx->fail(std::string(#call) + ": " + cudaGetErrorName(e_) + " " + cudaGetErrorString(e_), 2); \
  } while(0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int round_up(int64_t x, int64_t m) { return (int) ((x + m - 1) / m * m); }

void free_pools(ccsdt_ctx* ctx) {
  for(auto& b: ctx->buf) {
    for(int pl = 0; pl < 2; pl++) {
      if(b.geom.hpp[pl]) cudaFree(b.geom.hpp[pl]);
      if(b.geom.hhp[pl]) cudaFree(b.geom.hhp[pl]);
      b.geom.hpp[pl] = b.geom.hhp[pl] = nullptr;
    }
    if(b.s1_a) cudaFree(b.s1_a);
    if(b.s1_b) cudaFree(b.s1_b);
    if(b.d_descs) cudaFree(b.d_descs);
    if(b.h_descs) cudaFreeHost(b.h_descs);
    if(b.d_partial) cudaFree(b.d_partial);
    if(b.d_counter) cudaFree(b.d_counter);
    b.d_counter = nullptr;
    b.s1_a = b.s1_b = nullptr;
    b.d_descs = b.h_descs = nullptr;
    b.d_partial           = nullptr;
    b.desc_cap            = 0;
    b.partial_cap         = 0;
  }
  ctx->pools_ready = false;
}

void free_operands(ccsdt_ctx* ctx) {
  for(auto& d: ctx->dense) {
    if(d) cudaFree(d);
    d = nullptr;
  }
  for(auto& kv: ctx->blocks) cudaFree(kv.second.dev);
  ctx->blocks.clear();
  ctx->block_bytes = 0;
}

int64_t dim_full(const Space& sp, char kind) { return kind == 'o' ? sp.n_occ() : sp.n_virt(); }
int     tile_of(const Space& sp, char kind, uint32_t bid) { return kind == 'o' ? (int) bid : (int) bid + sp.noab(); }
// orbital number inside the occupied or the virtual range of the first orbital of a tile
int64_t tile_start(const Space& sp, int tile) {
  return tile < sp.noab() ? sp.k_offset[tile] : sp.k_offset[tile] - sp.k_offset[sp.noab()];
}

size_t block_elems(const Space& sp, int tensor, const uint32_t bid[4]) {
  const char* kinds = kKinds[tensor];
  size_t      n     = 1;
  for(int d = 0; kinds[d]; d++) n *= (size_t) sp.k_range[tile_of(sp, kinds[d], bid[d])];
  return n;
}

int make_tmaps(ccsdt_ctx* ctx, StageBuf& b) {
  auto enc = (EncodeTiledFn) ctx->encode_fn;
  for(int pl = 0; pl < 2; pl++)
    for(int kind = 0; kind < 2; kind++) {
      const PoolGeom& g   = b.geom;
      const int64_t   Kp  = g.Kp[pl];
      void*           ptr = kind == 0 ? (void*) g.hpp[pl] : (void*) g.hhp[pl];
      // dims innermost first: (k, inner particle, o1, o2 x 9 panels)
      cuuint64_t dims[4]    = {(cuuint64_t) Kp, (cuuint64_t) g.TPp, (cuuint64_t) (kind == 0 ? g.TPp : g.THp),
                               (cuuint64_t) g.THp * 9};
      cuuint64_t strides[3] = {(cuuint64_t) Kp * 8, (cuuint64_t) g.TPp * Kp * 8,
                               (cuuint64_t) (kind == 0 ? g.TPp : g.THp) * g.TPp * Kp * 8};
      cuuint32_t box[4]     = {KSLAB, 8, (cuuint32_t) (kind == 0 ? 8 : 2), 1};
      cuuint32_t es[4]      = {1, 1, 1, 1};
      CUresult   r = enc(&b.tmap[pl * 2 + kind], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, ptr, dims, strides, box, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if(r != CUDA_SUCCESS) return ctx->fail("cuTensorMapEncodeTiled failed with code " + std::to_string((int) r), 3);
    }
  return 0;
}

int ensure_pools(ccsdt_ctx* ctx) {
  if(ctx->pools_ready) return 0;
  const Space& sp = ctx->sp;
  int          cmax = 2;
  for(int i = 0; i < 3; i++) cmax = std::max(cmax, 2 * ctx->opt.sub[i]);
  int64_t th = sp.max_hole_tile(), tp = sp.max_particle_tile();
  int     THp = 0;
  for(int i = 0; i < 3; i++) THp = std::max(THp, round_up(th, 2 * ctx->opt.sub[i]));
  THp           = round_up(THp, 2);
  const int TPp = round_up(tp, 8);
  int       tb, te;
  int64_t   na, nb;
  sp.spin_range(false, 1, tb, te, na);
  sp.spin_range(false, 2, tb, te, nb);
  const int Kp1 = round_up(std::max<int64_t>(std::max(na, nb), 1), KSLAB);
  sp.spin_range(true, 1, tb, te, na);
  sp.spin_range(true, 2, tb, te, nb);
  const int Kp2 = round_up(std::max<int64_t>(std::max(na, nb), 1), KSLAB);
  (void) cmax;

  const int nbuf = ctx->opt.overlap ? 2 : 1;
  for(int i = 0; i < nbuf; i++) {
    StageBuf& b = ctx->buf[i];
    b.geom.THp  = THp;
    b.geom.TPp  = TPp;
    b.geom.Kp[0] = Kp1;
    b.geom.Kp[1] = Kp2;
    for(int pl = 0; pl < 2; pl++) {
      const size_t hpp = (size_t) 9 * b.geom.hpp_stride_panel(pl) * 8, hhp = (size_t) 9 * b.geom.hhp_stride_panel(pl) * 8;
      CK(cudaMalloc(&b.geom.hpp[pl], hpp));
      CK(cudaMalloc(&b.geom.hhp[pl], hhp));
      CK(cudaMemset(b.geom.hpp[pl], 0, hpp));
      CK(cudaMemset(b.geom.hhp[pl], 0, hhp));
    }
    const size_t s1a = (size_t) 9 * THp * TPp * 8, s1b = (size_t) 9 * TPp * TPp * THp * THp * 8;
    CK(cudaMalloc(&b.s1_a, s1a));
    CK(cudaMalloc(&b.s1_b, s1b));
    CK(cudaMemset(b.s1_a, 0, s1a));
    CK(cudaMemset(b.s1_b, 0, s1b));
    b.desc_cap = 64 + 9 * 2 * 2 * (sp.noab() + sp.nvab() + 2);
    CK(cudaMalloc(&b.d_descs, sizeof(GatherDesc) * b.desc_cap));
    CK(cudaMallocHost(&b.h_descs, sizeof(GatherDesc) * b.desc_cap));
    CK(cudaMalloc(&b.d_counter, 4 * COUNTER_WORDS));
    if(int rc = make_tmaps(ctx, b)) return rc;
    if(!b.staged) {
      CK(cudaEventCreateWithFlags(&b.staged, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&b.done, cudaEventDisableTiming));
      CK(cudaEventCreate(&b.k0));
      CK(cudaEventCreate(&b.k1));
      CK(cudaEventCreate(&b.g0));
      CK(cudaEventCreate(&b.g1));
    }
  }
  ctx->pools_ready = true;
  return 0;
}

// ---- operand resolution -------------------------------------------------------------------------
int resolve_block(ccsdt_ctx* ctx, int tensor, const uint32_t bid[4], BlockRef& out) {
  const Space& sp    = ctx->sp;
  const char*  kinds = kKinds[tensor];
  const int    nd    = (int) strlen(kinds);
  int64_t      ext[4] = {1, 1, 1, 1};
  for(int d = 0; d < nd; d++) ext[d] = sp.k_range[tile_of(sp, kinds[d], bid[d])];
  if(ctx->dense[tensor]) {
    int64_t st = 1, off = 0;
    for(int d = nd - 1; d >= 0; d--) {
      out.stride[d] = st;
      off += tile_start(sp, tile_of(sp, kinds[d], bid[d])) * st;
      st *= dim_full(sp, kinds[d]);
    }
    out.ptr = ctx->dense[tensor] + off;
    return 0;
  }
  BlockKey key{tensor, {bid[0], bid[1], nd > 2 ? bid[2] : 0u, nd > 3 ? bid[3] : 0u}};
  auto     it = ctx->blocks.find(key);
  if(it == ctx->blocks.end()) {
    if(!ctx->fetch)
      return ctx->fail("block {" + std::to_string(bid[0]) + "," + std::to_string(bid[1]) + ",..} of tensor " +
                       std::to_string(tensor) + " was never supplied (no dense tensor, block or fetch callback)", 4);
    size_t n = 1;
    for(int d = 0; d < nd; d++) n *= (size_t) ext[d];
    // evict blocks that the task being staged and the task in flight do not use
    while(ctx->block_budget && ctx->block_bytes + n * 8 > ctx->block_budget && !ctx->blocks.empty()) {
      auto victim = ctx->blocks.end();
      for(auto jt = ctx->blocks.begin(); jt != ctx->blocks.end(); ++jt)
        if(jt->second.last_use + 2 <= ctx->use_clock && (victim == ctx->blocks.end() || jt->second.last_use < victim->second.last_use))
          victim = jt;
      if(victim == ctx->blocks.end()) break;
      CK(cudaFreeAsync(victim->second.dev, ctx->s_stage));
      ctx->block_bytes -= victim->second.bytes;
      ctx->blocks.erase(victim);
    }
    if(n > ctx->h_fetch_cap) {
      // the bounce buffer is reused: previous copies must have left it
      CK(cudaStreamSynchronize(ctx->s_stage));
      if(ctx->h_fetch) cudaFreeHost(ctx->h_fetch);
      ctx->h_fetch_cap = std::max(n, (size_t) 1 << 20);
      CK(cudaMallocHost(&ctx->h_fetch, ctx->h_fetch_cap * 8));
    }
    else { CK(cudaStreamSynchronize(ctx->s_stage)); }
    if(ctx->fetch(ctx->fetch_user, tensor, bid, ctx->h_fetch, n) != 0)
      return ctx->fail("fetch callback failed for tensor " + std::to_string(tensor), 5);
    double* dev = nullptr;
    CK(cudaMallocAsync(&dev, n * 8, ctx->s_stage));
    CK(cudaMemcpyAsync(dev, ctx->h_fetch, n * 8, cudaMemcpyHostToDevice, ctx->s_stage));
    ctx->stats.h2d_bytes += (int64_t) n * 8;
    ctx->stats.blocks_fetched++;
    ctx->block_bytes += n * 8;
    it = ctx->blocks.emplace(key, BlockEntry{dev, n * 8, ctx->use_clock}).first;
  }
  it->second.last_use = ctx->use_clock;
  int64_t st          = 1;
  for(int d = nd - 1; d >= 0; d--) {
    out.stride[d] = st;
    st *= ext[d];
  }
  out.ptr = it->second.dev;
  return 0;
}

// one panel-build work item.  dims_of_dst[j] = which dimension of the source block feeds dst dim j
// (dst dims ordered o2, o1, in, k); ext taken from the block.
struct SrcSpec {
  int      tensor;
  uint32_t bid[4];
  int      dim_of[4]; // -1: dst dim has extent 1 and no source dim
  double   scale;
};

int add_gather(ccsdt_ctx* ctx, StageBuf& b, int& nd, int64_t& max_elems, double* dst, const int64_t ds[4],
               const SrcSpec& s) {
  const Space& sp = ctx->sp;
  if(nd >= b.desc_cap) return ctx->fail("internal: gather descriptor capacity exceeded", 6);
  GatherDesc& g     = b.h_descs[nd];
  const char* kinds = kKinds[s.tensor];
  memset(&g, 0, sizeof(g));
  g.dst          = dst;
  g.scale        = s.scale;
  g.synth_tensor = -1;
  int64_t total  = 1;
  BlockRef ref;
  if(!ctx->synthetic) {
    if(int rc = resolve_block(ctx, s.tensor, s.bid, ref)) return rc;
    g.src = ref.ptr;
  }
  else g.synth_tensor = s.tensor;
  for(int j = 0; j < 4; j++) {
    g.ds[j] = ds[j];
    const int d = s.dim_of[j];
    if(d < 0) {
      g.n[j]     = 1;
      g.ss[j]    = 0;
      g.gpos[j]  = 3; // harmless: adds 0 below
      g.gbase[j] = 0;
      continue;
    }
    const int tile = tile_of(sp, kinds[d], s.bid[d]);
    g.n[j]         = (int32_t) sp.k_range[tile];
    g.ss[j]        = ref.stride[d];
    g.gpos[j]      = d;
    g.gbase[j]     = tile_start(sp, tile);
    total *= g.n[j];
  }
  if(ctx->synthetic) {
    // unused tensor positions (T1 has two) must stay 0: point extent-1 dims at distinct spare slots
    int spare = (int) strlen(kinds);
    for(int j = 0; j < 4; j++)
      if(s.dim_of[j] < 0) g.gpos[j] = std::min(spare++, 3);
  }
  max_elems = std::max(max_elems, total);
  nd++;
  return 0;
}

int add_zero(ccsdt_ctx* ctx, StageBuf& b, int& nd, int64_t& max_elems, double* dst, const int64_t ds[4],
             const int n[4]) {
  if(nd >= b.desc_cap) return ctx->fail("internal: gather descriptor capacity exceeded", 6);
  GatherDesc& g = b.h_descs[nd];
  memset(&g, 0, sizeof(g));
  g.dst          = dst;
  g.synth_tensor = -1;
  int64_t total  = 1;
  for(int j = 0; j < 4; j++) {
    g.ds[j] = ds[j];
    g.n[j]  = n[j];
    total *= n[j];
  }
  if(total <= 0) return 0;
  max_elems = std::max(max_elems, total);
  nd++;
  return 0;
}

// ---- staging of one task ------------------------------------------------------------------------
int stage_task(ccsdt_ctx* ctx, StageBuf& b, const Task& t) {
  const Space&    sp   = ctx->sp;
  const PoolGeom& g    = b.geom;
  const int       noab = sp.noab();
  TaskParams&     P    = b.params;
  memset(&P, 0, sizeof(P));
  memcpy(P.tmap, b.tmap, sizeof(P.tmap));
  P.geom = g;
  ctx->use_clock++;

  int ext[6];
  for(int i = 0; i < 6; i++) ext[i] = P.ext[i] = (int) sp.k_range[t.t[i]];
  for(int i = 0; i < 3; i++) {
    P.sub[i]  = ctx->opt.sub[i];
    P.c[i]    = 2 * ctx->opt.sub[i];
    P.nbox[i] = (ext[i] + P.c[i] - 1) / P.c[i];
  }
  for(int i = 3; i < 6; i++) P.nbox[i] = (ext[i] + PBOX - 1) / PBOX;
  int64_t nboxes = 1;
  for(int i = 0; i < 6; i++) nboxes *= P.nbox[i];
  if(nboxes > 0x7fffffff) return ctx->fail("task has too many CTA boxes", 7);
  P.nboxes = (int) nboxes;
  for(int i = 0; i < 6; i++) P.evl[i] = ctx->d_evl + sp.k_offset[t.t[i]];
  P.error_flag = ctx->d_error;

  int     nd = 0;
  int64_t max_elems = 1;
  TermDev terms[MAX_TERMS];
  int     nterms = 0;

  auto tile = [&](int idx) { return (uint32_t) t.t[idx]; };
  auto vt   = [&](int idx) { return (uint32_t) (t.t[idx] - noab); };

  // ---- d1 terms ----
  for(int k = 0; k < 9; k++) {
    const int s7 = d1_contracted_spin(sp, t, k);
    if(!s7) continue;
    const D1Term& T = kD1[k];
    int           tb, te;
    int64_t       K;
    sp.spin_range(false, s7, tb, te, K);
    if(K <= 0) continue;
    TermDev td{};
    td.pool      = 0;
    td.kslabs    = (int) ((K + KSLAB - 1) / KSLAB);
    td.ksteps_last = (int) ((K - (int64_t) (td.kslabs - 1) * KSLAB + 3) / 4);
    td.hpp_hole  = T.hx;
    td.qt        = T.pa;
    td.inner_hpp = T.pb;
    td.hhp_o2    = T.hy;
    td.hhp_o1    = T.hz;
    td.inner_hhp = T.pc;
    td.a_is_hpp  = (T.pb == 5);
    td.layout_y  = (T.pa == 4);
    td.hpp_panel = td.hhp_panel = k;
    double*       hpp = g.hpp[0] + k * g.hpp_stride_panel(0);
    double*       hhp = g.hhp[0] + k * g.hhp_stride_panel(0);
    const int64_t ds_hpp[4] = {g.hpp_stride_o2(0), g.hpp_stride_o1(0), g.hpp_stride_in(0), 1};
    const int64_t ds_hhp[4] = {g.hhp_stride_o2(0), g.hhp_stride_o1(0), g.hhp_stride_in(0), 1};
    int64_t       koff = 0;
    for(int h7 = tb; h7 < te; h7++) {
      // T2[pa,pb,hx,l]: canonical block order (min(hx,h7), max(hx,h7)), sign -1 when transposed
      SrcSpec a{};
      a.tensor = CCSDT_T2;
      if(h7 < t.t[T.hx]) {
        const uint32_t bid[4] = {vt(T.pa), vt(T.pb), (uint32_t) h7, tile(T.hx)};
        memcpy(a.bid, bid, sizeof(bid));
        const int dim_of[4] = {3, 0, 1, 2}; // dst (hx, pa, pb, l)
        memcpy(a.dim_of, dim_of, sizeof(dim_of));
        a.scale = -1.0 * T.sign;
      }
      else {
        const uint32_t bid[4] = {vt(T.pa), vt(T.pb), tile(T.hx), (uint32_t) h7};
        memcpy(a.bid, bid, sizeof(bid));
        const int dim_of[4] = {2, 0, 1, 3};
        memcpy(a.dim_of, dim_of, sizeof(dim_of));
        a.scale = 1.0 * T.sign;
      }
      if(int rc = add_gather(ctx, b, nd, max_elems, hpp + koff, ds_hpp, a)) return rc;
      // v2ijka[hy,hz,l,pc]
      SrcSpec        v{};
      v.tensor              = CCSDT_V_IJKA;
      const uint32_t vbid[4] = {tile(T.hy), tile(T.hz), (uint32_t) h7, vt(T.pc)};
      memcpy(v.bid, vbid, sizeof(vbid));
      const int vdim[4] = {0, 1, 3, 2}; // dst (hy, hz, pc, l)
      memcpy(v.dim_of, vdim, sizeof(vdim));
      v.scale = 1.0;
      if(int rc = add_gather(ctx, b, nd, max_elems, hhp + koff, ds_hhp, v)) return rc;
      koff += sp.k_range[h7];
    }
    const int tail = td.kslabs * KSLAB - (int) K;
    if(tail > 0) {
      const int n_hpp[4] = {ext[T.hx], ext[T.pa], ext[T.pb], tail};
      const int n_hhp[4] = {ext[T.hy], ext[T.hz], ext[T.pc], tail};
      if(int rc = add_zero(ctx, b, nd, max_elems, hpp + K, ds_hpp, n_hpp)) return rc;
      if(int rc = add_zero(ctx, b, nd, max_elems, hhp + K, ds_hhp, n_hhp)) return rc;
    }
    terms[nterms++] = td;
  }
  // ---- d2 terms ----
  for(int k = 0; k < 9; k++) {
    const int s7 = d2_contracted_spin(sp, t, k);
    if(!s7) continue;
    const D2Term& T = kD2[k];
    int           tb, te;
    int64_t       K;
    sp.spin_range(true, s7, tb, te, K);
    if(K <= 0) continue;
    TermDev td{};
    td.pool      = 1;
    td.kslabs    = (int) ((K + KSLAB - 1) / KSLAB);
    td.ksteps_last = (int) ((K - (int64_t) (td.kslabs - 1) * KSLAB + 3) / 4);
    td.hpp_hole  = T.hz;
    td.qt        = T.pb;
    td.inner_hpp = T.pc;
    td.hhp_o2    = T.hx;
    td.hhp_o1    = T.hy;
    td.inner_hhp = T.pa;
    td.a_is_hpp  = (T.pc == 5);
    td.layout_y  = (T.pb == 4);
    td.hpp_panel = td.hhp_panel = k;
    double*       hpp = g.hpp[1] + k * g.hpp_stride_panel(1);
    double*       hhp = g.hhp[1] + k * g.hhp_stride_panel(1);
    const int64_t ds_hpp[4] = {g.hpp_stride_o2(1), g.hpp_stride_o1(1), g.hpp_stride_in(1), 1};
    const int64_t ds_hhp[4] = {g.hhp_stride_o2(1), g.hhp_stride_o1(1), g.hhp_stride_in(1), 1};
    int64_t       koff = 0;
    for(int p7 = tb; p7 < te; p7++) {
      // T2[pa,d,hx,hy] -> HHP[hx][hy][pa][d]
      SrcSpec a{};
      a.tensor = CCSDT_T2;
      if(p7 < t.t[T.pa]) {
        const uint32_t bid[4] = {(uint32_t) (p7 - noab), vt(T.pa), tile(T.hx), tile(T.hy)};
        memcpy(a.bid, bid, sizeof(bid));
        const int dim_of[4] = {2, 3, 1, 0};
        memcpy(a.dim_of, dim_of, sizeof(dim_of));
        a.scale = -1.0 * T.sign;
      }
      else {
        const uint32_t bid[4] = {vt(T.pa), (uint32_t) (p7 - noab), tile(T.hx), tile(T.hy)};
        memcpy(a.bid, bid, sizeof(bid));
        const int dim_of[4] = {2, 3, 0, 1};
        memcpy(a.dim_of, dim_of, sizeof(dim_of));
        a.scale = 1.0 * T.sign;
      }
      if(int rc = add_gather(ctx, b, nd, max_elems, hhp + koff, ds_hhp, a)) return rc;
      // v2iabc[hz,d,pb,pc] -> HPP[hz][pb][pc][d]
      SrcSpec        v{};
      v.tensor               = CCSDT_V_IABC;
      const uint32_t vbid[4] = {tile(T.hz), (uint32_t) (p7 - noab), vt(T.pb), vt(T.pc)};
      memcpy(v.bid, vbid, sizeof(vbid));
      const int vdim[4] = {0, 2, 3, 1};
      memcpy(v.dim_of, vdim, sizeof(vdim));
      v.scale = 1.0;
      if(int rc = add_gather(ctx, b, nd, max_elems, hpp + koff, ds_hpp, v)) return rc;
      koff += sp.k_range[p7];
    }
    const int tail = td.kslabs * KSLAB - (int) K;
    if(tail > 0) {
      const int n_hpp[4] = {ext[T.hz], ext[T.pb], ext[T.pc], tail};
      const int n_hhp[4] = {ext[T.hx], ext[T.hy], ext[T.pa], tail};
      if(int rc = add_zero(ctx, b, nd, max_elems, hpp + K, ds_hpp, n_hpp)) return rc;
      if(int rc = add_zero(ctx, b, nd, max_elems, hhp + K, ds_hhp, n_hhp)) return rc;
    }
    terms[nterms++] = td;
  }
  // X terms first, then Y (stable)
  P.nterms = nterms;
  int w    = 0;
  for(int i = 0; i < nterms; i++)
    if(!terms[i].layout_y) P.term[w++] = terms[i];
  P.nterms_x = w;
  for(int i = 0; i < nterms; i++)
    if(terms[i].layout_y) P.term[w++] = terms[i];

  // ---- s1 terms: a = sign*T1[pa,hx] as [hx][pa], b = v2ijab[hz,hy,pc,pb] as [hy][hz][pb][pc] ----
  P.ns1 = 0;
  for(int k = 0; k < 9; k++) {
    if(!s1_enabled(sp, t, k)) continue;
    const S1Term& T  = kS1[k];
    double*       pa = b.s1_a + (int64_t) k * g.THp * g.TPp;
    double*       pb = b.s1_b + (int64_t) k * g.TPp * g.TPp * g.THp * g.THp;
    S1Dev&        sd = P.s1[P.ns1++];
    memset(&sd, 0, sizeof(sd));
    sd.a        = pa;
    sd.b        = pb;
    sd.hx       = T.hx;
    sd.pa       = T.pa;
    sd.sa[T.hx] = g.TPp;
    sd.sa[T.pa] = 1;
    // particle indices innermost: the 32 lanes of a warp differ in particle offsets only, so one warp-wide
    // load of b touches at most 8 rows of 64 bytes (it touched 32 lines with the holes innermost)
    sd.sb[T.hy] = g.THp * g.TPp * g.TPp;
    sd.sb[T.hz] = g.TPp * g.TPp;
    sd.sb[T.pb] = g.TPp;
    sd.sb[T.pc] = 1;
    SrcSpec a{};
    a.tensor              = CCSDT_T1;
    const uint32_t abid[4] = {vt(T.pa), tile(T.hx), 0, 0};
    memcpy(a.bid, abid, sizeof(abid));
    const int adim[4] = {-1, -1, 1, 0}; // dst (1, 1, hx, pa)
    memcpy(a.dim_of, adim, sizeof(adim));
    a.scale              = (double) T.sign;
    const int64_t ds_a[4] = {0, 0, g.TPp, 1};
    if(int rc = add_gather(ctx, b, nd, max_elems, pa, ds_a, a)) return rc;
    SrcSpec v{};
    v.tensor               = CCSDT_V_IJAB;
    const uint32_t vbid[4] = {tile(T.hz), tile(T.hy), vt(T.pc), vt(T.pb)};
    memcpy(v.bid, vbid, sizeof(vbid));
    const int vdim[4] = {1, 0, 3, 2}; // dst (hy, hz, pb, pc)
    memcpy(v.dim_of, vdim, sizeof(vdim));
    v.scale               = 1.0;
    const int64_t ds_b[4] = {(int64_t) g.THp * g.TPp * g.TPp, (int64_t) g.TPp * g.TPp, g.TPp, 1};
    if(int rc = add_gather(ctx, b, nd, max_elems, pb, ds_b, v)) return rc;
  }

  // ---- launch geometry, box order, partial buffer ----
  int64_t need_partial;
  if(ctx->opt.kernel == CCSDT_KERNEL_SIMPLE) {
    int64_t total = 1;
    for(int i = 0; i < 6; i++) total *= ext[i];
    need_partial = (total + 255) / 256;
  }
  else {
    const int ncw  = 4 * P.sub[0] * P.sub[1] * P.sub[2];
    int       rows = 0;
    for(int hh = 0; hh < 3; hh++) {
      const int a = hh == 0 ? 1 : 0, c2 = hh == 2 ? 1 : 2;
      rows = std::max(rows, P.c[hh] * 64 + P.c[a] * P.c[c2] * 8);
    }
    P.stage_bytes = rows * ROW_BYTES;
    // 128 registers per thread: 3 CTAs of 4+1 warps or 1 CTA of 8+1 / 12+1 warps per SM
    int ctas = ctx->opt.ctas_per_sm > 0 ? ctx->opt.ctas_per_sm : (ncw <= 4 ? 3 : 1);
    ctas     = std::min(ctas, ncw <= 4 ? 3 : 1);
    const size_t smem_total = (size_t) ctx->prop.sharedMemPerMultiprocessor;
    const size_t per_cta    = std::min((size_t) ctx->prop.sharedMemPerBlockOptin,
                                       smem_total / ctas - 1024 /*driver reserve*/) - 2048 /*static + slack*/;
    int stages = (int) ((per_cta - 1024) / P.stage_bytes);
    stages     = std::min(stages, MAX_STAGES);
    if(ctx->opt.stages > 0) stages = std::min(stages, ctx->opt.stages);
    const int64_t box_elems = (int64_t) P.c[0] * P.c[1] * P.c[2] * 512;
    if(stages < 2 || (int64_t) stages * P.stage_bytes < box_elems * 8)
      return ctx->fail("shared memory too small for the requested CTA box", 8);
    P.stages         = stages;
    b.smem           = (size_t) stages * P.stage_bytes + 1024;
    b.consumer_warps = ncw;
    const int64_t in_flight = (int64_t) ctx->prop.multiProcessorCount * ctas;
    // (A staging block cannot start on an SM that holds three fused CTAs: the register file is partitioned per
    // scheduler and three of the four partitions are full -- tools/probes/coresidency.cu.  Leaving CTA slots free
    // for the panel build was measured: it hides the build but costs the same time in the fused kernel.)
    b.grid           = (int) std::min<int64_t>(nboxes, in_flight);
    // bricks: grow the index with the smallest element extent until one brick holds about as many
    // boxes as there are CTAs in flight, then even the bricks out over each index
    const int cext[6] = {P.c[0], P.c[1], P.c[2], PBOX, PBOX, PBOX};
    int64_t   vol     = 1;
    for(int i = 0; i < 6; i++) P.brick[i] = 1;
    while(vol < in_flight) {
      int best = -1;
      for(int i = 0; i < 6; i++)
        if(P.brick[i] < P.nbox[i] && (best < 0 || cext[i] * P.brick[i] < cext[best] * P.brick[best])) best = i;
      if(best < 0) break;
      vol = vol / P.brick[best] * (P.brick[best] + 1);
      P.brick[best]++;
    }
    int64_t padded = 1;
    for(int i = 0; i < 6; i++) {
      P.nbrick[i] = (P.nbox[i] + P.brick[i] - 1) / P.brick[i];
      P.brick[i]  = (P.nbox[i] + P.nbrick[i] - 1) / P.nbrick[i];
      padded *= (int64_t) P.nbrick[i] * P.brick[i];
    }
    if(padded > 0x7fffffff) return ctx->fail("task has too many CTA boxes", 7);
    P.nboxes_padded = (int) padded;
    P.box_counter   = b.d_counter;
    // permutational symmetry of coinciding tiles: evaluate ascending box coordinates only (weights in the kernel)
    P.sym = 0, P.nlist = 0, P.box_list = nullptr;
    if(ctx->opt.symmetry) {
      if(t.t[0] == t.t[1] && P.c[0] == P.c[1]) P.sym |= 1;
      if(t.t[1] == t.t[2] && P.c[1] == P.c[2]) P.sym |= 2;
      if(t.t[3] == t.t[4]) P.sym |= 4;
      if(t.t[4] == t.t[5]) P.sym |= 8;
      if((P.sym & 3) == 3 && P.c[0] == 2) P.sym |= 16;
    }
    if(P.sym) {
      std::array<int, 19> key;
      for(int i = 0; i < 6; i++) key[i] = P.nbox[i], key[6 + i] = P.brick[i], key[12 + i] = P.nbrick[i];
      key[18]  = P.sym;
      auto& bl = ctx->box_lists[key];
      if(!bl.dev) {
        std::vector<int32_t> ids;
        const int            order[6] = {2, 1, 0, 5, 4, 3}; // same decoding as decode_box (ccsdt_kernel_common.cuh)
        for(int64_t id = 0; id < padded; id++) {
          int     in[6], bx[6];
          int64_t r     = id;
          bool    valid = true;
          for(int i = 0; i < 6; i++) {
            const int d = order[i];
            in[d]       = (int) (r % P.brick[d]);
            r /= P.brick[d];
          }
          for(int i = 0; i < 6; i++) {
            const int d = order[i];
            bx[d]       = (int) (r % P.nbrick[d]) * P.brick[d] + in[d];
            r /= P.nbrick[d];
            valid &= bx[d] < P.nbox[d];
          }
          if(valid && box_weight(P.sym, bx) > 0) ids.push_back((int32_t) id);
        }
        bl.n = (int32_t) ids.size();
        CK(cudaMalloc(&bl.dev, sizeof(int32_t) * std::max<size_t>(ids.size(), 1)));
        CK(cudaMemcpy(bl.dev, ids.data(), sizeof(int32_t) * ids.size(), cudaMemcpyHostToDevice));
      }
      P.box_list = bl.dev;
      P.nlist    = bl.n;
      nboxes     = bl.n;
      b.grid     = (int) std::max<int64_t>(1, std::min<int64_t>(nboxes, in_flight));
    }
    b.eval_fraction = (double) nboxes / (double) P.nboxes;
    // one box keeps the tensor pipe of an SM busy for (k-steps x 16 DMMA x 16 cycles) / 4 sub-partitions x
    // 4 warps = k-steps x 256 cycles; co-resident CTAs start that far apart (options.stagger, default on)
    int64_t ksteps = 0;
    for(int t = 0; t < P.nterms; t++) ksteps += (int64_t) (P.term[t].kslabs - 1) * 4 + P.term[t].ksteps_last;
    P.ctas_per_sm    = ctas;
    P.stagger_cycles = (ctas > 1 && ctx->opt.stagger && nboxes >= 4 * in_flight)
                         ? (int) std::min<int64_t>(ksteps * 256 + 8192, 50000000)
                         : 0;
    need_partial    = P.box_list ? std::max<int64_t>(P.nlist, 1) : padded; // one partial per evaluated box
  }
  b.nparts = need_partial;
  if(need_partial > b.partial_cap) {
    CK(cudaStreamSynchronize(b.cs));
    if(b.d_partial) CK(cudaFree(b.d_partial));
    b.partial_cap = need_partial + need_partial / 4 + 64;
    CK(cudaMalloc(&b.d_partial, (size_t) b.partial_cap * 16));
  }
  P.partial = b.d_partial;

  // ---- launch the panel build on the staging stream ----
  static_assert(sizeof(GatherDesc) % 4 == 0, "descriptor copy works in 32-bit words");
  CK(launch_copy_from_pinned(b.h_descs, b.d_descs, sizeof(GatherDesc) * nd, ctx->s_stage));
  SynthInfo si{ctx->seed, 0, 0, 0, 0};
  {
    int     tb, te;
    int64_t n;
    sp.spin_range(false, 1, tb, te, n); si.noa = (int) n;
    sp.spin_range(false, 2, tb, te, n); si.nob = (int) n;
    sp.spin_range(true, 1, tb, te, n);  si.nva = (int) n;
    sp.spin_range(true, 2, tb, te, n);  si.nvb = (int) n;
  }
  // asynchronous dense uploads still in flight: an all-alpha task only needs the all-alpha blocks
  {
    bool all_alpha = true;
    for(int i = 0; i < 6; i++) all_alpha &= sp.k_spin[t.t[i]] == 1;
    for(int tn = 0; tn < 5; tn++)
      if(ctx->upload_pending[tn]) {
        CK(cudaStreamWaitEvent(ctx->s_stage, ctx->ev_alpha[tn], 0));
        if(!all_alpha) CK(cudaStreamWaitEvent(ctx->s_stage, ctx->ev_full[tn], 0));
      }
  }
  // the buffer's partials and box-scheduler words are zeroed here, on the staging stream, long before the launch
  // (ids of the padded brick grid that are not boxes are never written: their partials stay zero)
  CK(launch_zero(b.d_partial, 2 * b.nparts, b.d_counter, COUNTER_WORDS, ctx->s_stage));
  CK(cudaEventRecord(b.g0, ctx->s_stage));
  CK(launch_gather(b.d_descs, nd, max_elems, si, ctx->s_stage));
  CK(cudaEventRecord(b.g1, ctx->s_stage));
  CK(cudaEventRecord(b.staged, ctx->s_stage));
  ctx->stats.kernel_launches += 2 + (nd + 65534) / 65535;
  return 0;
}

int launch_task(ccsdt_ctx* ctx, StageBuf& b, int64_t slot) {
  CK(cudaStreamWaitEvent(b.cs, b.staged, 0));
  if(ctx->opt.kernel == CCSDT_KERNEL_SIMPLE) {
    int nparts = 0;
    CK(cudaEventRecord(b.k0, b.cs));
    CK(launch_fused_simple(b.params, b.cs, &nparts));
    CK(cudaEventRecord(b.k1, b.cs));
  }
  else if(b.params.nterms == 0) {
    // no doubles contraction is enabled: d = 0 for every element, so E[T] and E(T) of the task are exactly 0
    CK(cudaEventRecord(b.k0, b.cs)); // the staging stream zeroed the partials
    CK(cudaEventRecord(b.k1, b.cs));
  }
  else {
    CK(cudaEventRecord(b.k0, b.cs));
    CK(launch_fused_dmma(b.params, b.grid, b.consumer_warps, b.smem, b.cs));
    CK(cudaEventRecord(b.k1, b.cs));
  }
  CK(launch_reduce_partials(b.d_partial, (int) b.nparts, ctx->d_task_energy + 2 * slot, b.cs));
  CK(cudaEventRecord(b.done, b.cs));
  b.timing_pending = true;
  ctx->stats.kernel_launches += 2;
  return 0;
}

int collect_timing(ccsdt_ctx* ctx, StageBuf& b) {
  if(!b.timing_pending) return 0;
  CK(cudaEventSynchronize(b.done));
  // Consecutive tasks run on two compute streams, so the [k0,k1] intervals of neighbours overlap (a task's
  // CTAs fill the SMs the previous task's tail frees): seconds_kernel is the length of the UNION of the
  // intervals, accumulated in launch order on the time line of ev_base.
  float ms = 0, ms0 = 0, ms1 = 0;
  CK(cudaEventElapsedTime(&ms0, ctx->ev_base, b.k0));
  CK(cudaEventElapsedTime(&ms1, ctx->ev_base, b.k1));
  const double start = std::max((double) ms0, ctx->kernel_busy_until);
  if(ms1 > start) ctx->stats.seconds_kernel += (ms1 - start) * 1e-3;
  ctx->kernel_busy_until = std::max(ctx->kernel_busy_until, (double) ms1);
  CK(cudaEventElapsedTime(&ms, b.g0, b.g1));
  ctx->stats.seconds_staging += ms * 1e-3;
  b.timing_pending = false;
  return 0;
}

} // namespace

static int run_task_list(ccsdt_ctx* ctx, const std::vector<int64_t>& ids, bool use_global_split, double energies[2],
                         double* per_task, ccsdt_stats* stats_out) {
  cudaSetDevice(ctx->device);
  const auto t0 = std::chrono::high_resolution_clock::now();
  ctx->stats    = ccsdt_stats{};
  ctx->stats.h2d_bytes = ctx->pending_h2d; // uploads (ccsdt_put_*) since the previous run belong to this one
  ctx->pending_h2d     = 0;
  energies[0] = energies[1] = 0.0;
  const int64_t nids = (int64_t) ids.size();

  // Task hand-out.  Static: the tasks the cost-balanced split gave this rank, in canonical order.
  // Dynamic (a process-shared counter was set): every rank walks the SAME list, the tasks of the range in
  // descending cost order, and claims the next unclaimed entry with one atomic fetch-add -- the role of
  // the reference's AtomicCounterGA (ccsd_t_fused_driver.hpp:169-172, 456), with longest-task-first order.
  std::vector<int64_t> mine, order;
  std::vector<int64_t> pos_of(ctx->tasks.size(), -1); // position in ids (for per_task)
  for(int64_t k = 0; k < nids; k++) pos_of[ids[k]] = k;
  if(ctx->task_counter) {
    order = ids;
    std::vector<long double> cost(ctx->tasks.size(), 0);
    for(int64_t id: ids) cost[id] = task_cost(ctx->sp, ctx->tasks[id], ctx->opt.symmetry != 0);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return cost[a] > cost[b]; });
  }
  else if(use_global_split || ctx->opt.nranks <= 1) {
    for(int64_t id: ids)
      if(ctx->owner.empty() || ctx->owner[id] == ctx->opt.rank) order.push_back(id);
  }
  else {
    // static split of this list alone: longest-processing-time greedy, identical on every rank
    std::vector<Task> sub;
    for(int64_t id: ids) sub.push_back(ctx->tasks[id]);
    const std::vector<int32_t> own = partition_tasks(ctx->sp, sub, ctx->opt.nranks, ctx->opt.symmetry != 0);
    for(int64_t k = 0; k < nids; k++)
      if(own[k] == ctx->opt.rank) order.push_back(ids[k]);
  }
  {
    // uploads in flight: the tasks that only need the all-alpha blocks (the first to arrive) go first; per-task energies
    // are summed in canonical order whatever the execution order
    bool pending = false;
    for(int tn = 0; tn < 5; tn++) pending |= ctx->upload_pending[tn];
    if(pending && !ctx->task_counter)
      std::stable_partition(order.begin(), order.end(), [&](int64_t id) {
        bool a = true;
        for(int i = 0; i < 6; i++) a &= ctx->sp.k_spin[ctx->tasks[id].t[i]] == 1;
        return a;
      });
  }
  int64_t cursor = 0;
  auto    next_task = [&]() -> int64_t {
    const int64_t k = ctx->task_counter ? __atomic_fetch_add(ctx->task_counter, (int64_t) 1, __ATOMIC_RELAXED) : cursor++;
    return k < (int64_t) order.size() ? order[k] : -1;
  };
  if(per_task) std::fill(per_task, per_task + 2 * nids, 0.0);

  if(!order.empty()) {
    if(int rc = ensure_pools(ctx)) return rc;
    const int64_t cap = (int64_t) order.size();
    if(cap > ctx->task_energy_cap) {
      if(ctx->d_task_energy) CK(cudaFree(ctx->d_task_energy));
      ctx->task_energy_cap = cap + 16;
      CK(cudaMalloc(&ctx->d_task_energy, (size_t) ctx->task_energy_cap * 16));
    }
    const int nbuf = ctx->opt.overlap ? 2 : 1;
    ctx->buf[0].cs = ctx->s_compute;
    ctx->buf[1].cs = ctx->opt.overlap ? ctx->s_compute2 : ctx->s_compute;
    CK(cudaMemsetAsync(ctx->d_error, 0, 4, ctx->s_compute));
    CK(cudaEventRecord(ctx->ev_base, ctx->s_compute));
    CK(cudaStreamWaitEvent(ctx->s_compute2, ctx->ev_base, 0));
    ctx->kernel_busy_until = 0.0;
    for(int64_t j = 0;; j++) {
      const int64_t ti = next_task();
      if(ti < 0) break;
      mine.push_back(ti);
      StageBuf& b = ctx->buf[j % nbuf];
      // the buffer's previous task must have finished computing before its panels are rebuilt
      if(b.timing_pending) {
        if(int rc = collect_timing(ctx, b)) return rc;
      }
      CK(cudaStreamWaitEvent(ctx->s_stage, b.done, 0));
      if(int rc = stage_task(ctx, b, ctx->tasks[ti])) return rc;
      if(int rc = launch_task(ctx, b, j)) return rc;
      const double ops = (double) task_ops(ctx->sp, ctx->tasks[ti]);
      ctx->stats.counted_flops += ops;
      ctx->stats.evaluated_flops += ops * (ctx->opt.kernel == CCSDT_KERNEL_DMMA ? b.eval_fraction : 1.0);
    }
    const int64_t n = (int64_t) mine.size();
    for(int64_t k = std::max<int64_t>(0, n - nbuf); k < n; k++) // launch order (the interval union needs it)
      if(int rc = collect_timing(ctx, ctx->buf[k % nbuf])) return rc;
    cudaError_t e = cudaStreamSynchronize(ctx->s_compute);
    if(e == cudaSuccess) e = cudaStreamSynchronize(ctx->s_compute2);
    if(e != cudaSuccess) {
      return ctx->fail(std::string("fused kernel failed: ") + cudaGetErrorName(e) + " " + cudaGetErrorString(e) +
                       " (a pipeline timeout traps instead of hanging)", 9);
    }
    CK(cudaStreamSynchronize(ctx->s_stage));
    CK(cudaStreamSynchronize(ctx->s_copy_a));
    CK(cudaStreamSynchronize(ctx->s_copy_b));
    for(int tn = 0; tn < 5; tn++) ctx->upload_pending[tn] = false;
    uint32_t flag = 0;
    CK(cudaMemcpy(&flag, ctx->d_error, 4, cudaMemcpyDeviceToHost));
    if(flag) return ctx->fail("device error flag " + std::to_string(flag), 9);
    std::vector<double> e_host((size_t) 2 * std::max<int64_t>(n, 1));
    if(n) CK(cudaMemcpy(e_host.data(), ctx->d_task_energy, (size_t) n * 16, cudaMemcpyDeviceToHost));
    ctx->stats.d2h_bytes += n * 16;
    // reduction order: boxes in box-id order inside a task (fixed tree), then this rank's tasks in
    // canonical task order (whatever order they were claimed in)
    std::vector<int64_t> slot_of(n);
    for(int64_t j = 0; j < n; j++) slot_of[j] = j;
    std::sort(slot_of.begin(), slot_of.end(), [&](int64_t a, int64_t b) { return mine[a] < mine[b]; });
    for(int64_t jj = 0; jj < n; jj++) {
      const int64_t j  = slot_of[jj];
      const double  f  = ctx->tasks[mine[j]].factor;
      const double  e1 = f * e_host[2 * j], e2 = f * e_host[2 * j + 1];
      energies[0] += e1;
      energies[1] += e2;
      if(per_task) {
        per_task[2 * pos_of[mine[j]]]     = e1;
        per_task[2 * pos_of[mine[j]] + 1] = e2;
      }
    }
    ctx->stats.tasks_run = n;
  }
  const auto t1            = std::chrono::high_resolution_clock::now();
  ctx->stats.seconds_total = std::chrono::duration<double>(t1 - t0).count();
  if(stats_out) *stats_out = ctx->stats;
  return 0;
}

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

CCSDT_API const char* ccsdt_last_error(const ccsdt_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int ccsdt_box_weight(int sym, const int32_t box[6]) {
  const int bi[6] = {box[0], box[1], box[2], box[3], box[4], box[5]};
  return box_weight(sym, bi);
}

int ccsdt_default_options(ccsdt_options* o) {
  if(!o) return 1;
  memset(o, 0, sizeof(*o));
  o->kernel = CCSDT_KERNEL_DMMA;
  o->sub[0] = 1, o->sub[1] = 1, o->sub[2] = 1;
  o->nranks  = 1;
  o->overlap = 1;
  o->stagger = 1;
  o->symmetry = 1;
  return 0;
}

int ccsdt_create(ccsdt_ctx** out, int device) {
  if(!out) return 1;
  *out = nullptr;
  int         ndev = 0;
  cudaError_t e    = cudaGetDeviceCount(&ndev);
  if(e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no usable CUDA device (") + cudaGetErrorString(e) +
                     "); this library has no CPU fallback";
    return 2;
  }
  if(device < 0 && cudaGetDevice(&device) != cudaSuccess) device = 0; // negative: the caller's current device
  if(device >= ndev) {
    g_create_error = "device index out of range";
    return 2;
  }
  auto* ctx   = new ccsdt_ctx;
  ctx->device = device;
  ccsdt_default_options(&ctx->opt);
  auto bail = [&](const std::string& m) {
    g_create_error = m;
    delete ctx;
    return 2;
  };
  if((e = cudaSetDevice(device)) != cudaSuccess) return bail(cudaGetErrorString(e));
  if((e = cudaGetDeviceProperties(&ctx->prop, device)) != cudaSuccess) return bail(cudaGetErrorString(e));
  if(ctx->prop.major < 10)
    return bail("device is sm_" + std::to_string(ctx->prop.major) + std::to_string(ctx->prop.minor) +
                "; this library is built for sm_100a (B200) only");
  cudaDriverEntryPointQueryResult qres;
  if((e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ctx->encode_fn, cudaEnableDefault, &qres)) != cudaSuccess ||
     !ctx->encode_fn)
    return bail("cuTensorMapEncodeTiled not available from the driver");
  if((e = cudaStreamCreateWithFlags(&ctx->s_compute, cudaStreamNonBlocking)) != cudaSuccess) return bail(cudaGetErrorString(e));
  if((e = cudaStreamCreateWithFlags(&ctx->s_compute2, cudaStreamNonBlocking)) != cudaSuccess) return bail(cudaGetErrorString(e));
  if((e = cudaEventCreate(&ctx->ev_base)) != cudaSuccess) return bail(cudaGetErrorString(e));
  if((e = cudaStreamCreateWithFlags(&ctx->s_copy_a, cudaStreamNonBlocking)) != cudaSuccess) return bail(cudaGetErrorString(e));
  if((e = cudaStreamCreateWithFlags(&ctx->s_copy_b, cudaStreamNonBlocking)) != cudaSuccess) return bail(cudaGetErrorString(e));
  for(int t = 0; t < 5; t++) {
    if((e = cudaEventCreateWithFlags(&ctx->ev_alpha[t], cudaEventDisableTiming)) != cudaSuccess) return bail(cudaGetErrorString(e));
    if((e = cudaEventCreateWithFlags(&ctx->ev_full[t], cudaEventDisableTiming)) != cudaSuccess) return bail(cudaGetErrorString(e));
  }
  {
    // staging blocks (128 threads x 32 registers) slot in next to the resident fused CTAs; give them priority
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if((e = cudaStreamCreateWithPriority(&ctx->s_stage, cudaStreamNonBlocking, hi)) != cudaSuccess) return bail(cudaGetErrorString(e));
  }
  if((e = cudaMalloc(&ctx->d_error, 4)) != cudaSuccess) return bail(cudaGetErrorString(e));
  cudaMemset(ctx->d_error, 0, 4);
  if((e = fused_dmma_configure((size_t) ctx->prop.sharedMemPerBlockOptin - 2048)) != cudaSuccess)
    return bail(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
  {
    // fetched blocks come from the stream-ordered allocator: keep its memory across synchronisations (the default
    // pool hands everything back to the driver at every sync, and each block upload is followed by one)
    cudaMemPool_t pool = nullptr;
    if(cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  ctx->block_budget = total_b / 2;
  *out              = ctx;
  return 0;
}

int ccsdt_destroy(ccsdt_ctx* ctx) {
  if(!ctx) return 0;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  free_pools(ctx);
  free_operands(ctx);
  for(auto& b: ctx->buf)
    for(cudaEvent_t ev: {b.staged, b.done, b.k0, b.k1, b.g0, b.g1})
      if(ev) cudaEventDestroy(ev);
  if(ctx->d_evl) cudaFree(ctx->d_evl);
  if(ctx->d_task_energy) cudaFree(ctx->d_task_energy);
  if(ctx->d_error) cudaFree(ctx->d_error);
  for(auto& kv: ctx->box_lists)
    if(kv.second.dev) cudaFree(kv.second.dev);
  if(ctx->h_fetch) cudaFreeHost(ctx->h_fetch);
  if(ctx->s_compute) cudaStreamDestroy(ctx->s_compute);
  if(ctx->s_compute2) cudaStreamDestroy(ctx->s_compute2);
  if(ctx->ev_base) cudaEventDestroy(ctx->ev_base);
  if(ctx->s_copy_a) cudaStreamDestroy(ctx->s_copy_a);
  if(ctx->s_copy_b) cudaStreamDestroy(ctx->s_copy_b);
  for(int t = 0; t < 5; t++) {
    if(ctx->ev_alpha[t]) cudaEventDestroy(ctx->ev_alpha[t]);
    if(ctx->ev_full[t]) cudaEventDestroy(ctx->ev_full[t]);
  }
  if(ctx->s_stage) cudaStreamDestroy(ctx->s_stage);
  delete ctx;
  return 0;
}

int ccsdt_set_options(ccsdt_ctx* ctx, const ccsdt_options* o) {
  if(!ctx || !o) return 1;
  ccsdt_options n = *o;
  if(n.sub[0] == 0 && n.sub[1] == 0 && n.sub[2] == 0) n.sub[0] = 1, n.sub[1] = 1, n.sub[2] = 1;
  int prod = 1;
  for(int i = 0; i < 3; i++) {
    if(n.sub[i] < 1 || n.sub[i] > 3) return ctx->fail("options.sub entries must be 1, 2 or 3");
    prod *= n.sub[i];
  }
  if(prod > 3) return ctx->fail("options.sub product must be <= 3 (at most 12 consumer warps)");
  if(n.kernel != CCSDT_KERNEL_DMMA && n.kernel != CCSDT_KERNEL_SIMPLE)
    return ctx->fail("unknown kernel id");
  if(n.nranks < 1) n.nranks = 1;
  if(n.rank < 0 || n.rank >= n.nranks) return ctx->fail("rank out of range");
  if(n.stages < 0 || n.stages > MAX_STAGES) return ctx->fail("stages out of range");
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  free_pools(ctx); // box shape and overlap change the pool geometry
  ctx->opt = n;
  if(ctx->have_space) ctx->owner = partition_tasks(ctx->sp, ctx->tasks, ctx->opt.nranks, ctx->opt.symmetry != 0);
  return 0;
}

int ccsdt_tiles(int64_t n_occ_alpha, int64_t n_occ_beta, int64_t n_vir_alpha, int64_t n_vir_beta, int64_t tilesize,
                int64_t* k_range, int32_t* k_spin, int32_t counts[4], int cap) {
  if(tilesize <= 0 || !k_range || !k_spin || !counts) return -1;
  std::vector<int64_t> kr;
  std::vector<int32_t> ks;
  const int            n = make_tiles(n_occ_alpha, n_occ_beta, n_vir_alpha, n_vir_beta, tilesize, kr, ks, counts);
  if(n > cap) return -n;
  std::copy(kr.begin(), kr.end(), k_range);
  std::copy(ks.begin(), ks.end(), k_spin);
  return n;
}

int64_t ccsdt_enumerate(int noab, int nvab, const int32_t* k_spin, int is_restricted, int64_t* tasks7,
                        double* factors, int64_t cap, int64_t* n_outer) {
  auto tasks = enumerate_tasks(noab, nvab, k_spin, is_restricted != 0, n_outer);
  for(int64_t i = 0; i < (int64_t) tasks.size() && i < cap; i++) {
    if(tasks7) {
      for(int j = 0; j < 6; j++) tasks7[7 * i + j] = tasks[i].t[j];
      tasks7[7 * i + 6] = tasks[i].outer;
    }
    if(factors) factors[i] = tasks[i].factor;
  }
  return (int64_t) tasks.size();
}

static Space space_from_args(int noab, int nvab, const int32_t* k_spin, const int64_t* k_range, int restricted) {
  int noa = 0, nva = 0;
  for(int i = 0; i < noab; i++) noa += k_spin[i] == 1;
  for(int i = noab; i < noab + nvab; i++) nva += k_spin[i] == 1;
  return make_space(noa, noab - noa, nva, nvab - nva, k_range, k_spin, nullptr, restricted != 0);
}

int ccsdt_task_terms(int noab, int nvab, const int32_t* k_spin, const int64_t* k_range, int is_restricted,
                     const int64_t task[6], uint8_t* s1_on, uint8_t* d1_on, uint8_t* d2_on) {
  Space sp = space_from_args(noab, nvab, k_spin, k_range, is_restricted);
  Task  t{};
  for(int i = 0; i < 6; i++) t.t[i] = (int32_t) task[i];
  task_terms(sp, t, s1_on, d1_on, d2_on);
  return 0;
}

int ccsdt_count_ops(int noab, int nvab, const int32_t* k_spin, const int64_t* k_range, int is_restricted,
                    long double* total) {
  if(!total) return 1;
  *total = count_ops(space_from_args(noab, nvab, k_spin, k_range, is_restricted));
  return 0;
}

int ccsdt_partition(int noab, int nvab, const int32_t* k_spin, const int64_t* k_range, int is_restricted, int nranks,
                    int32_t* owner, int64_t cap) {
  Space sp    = space_from_args(noab, nvab, k_spin, k_range, is_restricted);
  auto  tasks = enumerate_tasks(noab, nvab, k_spin, is_restricted != 0, nullptr);
  auto  own   = partition_tasks(sp, tasks, nranks);
  for(int64_t i = 0; i < (int64_t) own.size() && i < cap; i++) owner[i] = own[i];
  return 0;
}

int ccsdt_check_memory(int tilesize, int nbf, size_t gpu_bytes, size_t* required) {
  // same bound as the reference's pre-check: 9 * (T^2 + T^4 + 4*nbf*T^3) doubles must fit
  const double T = tilesize;
  const double need = 9.0 * (T * T + T * T * T * T + 4.0 * nbf * T * T * T) * 8.0;
  if(required) *required = (size_t) need;
  return need < (double) gpu_bytes ? 0 : 1;
}

int ccsdt_set_space(ccsdt_ctx* ctx, int noa, int nob, int nva, int nvb, const int64_t* k_range, const int32_t* k_spin,
                    const double* evl, int is_restricted) {
  if(!ctx || !k_range || !k_spin || !evl) return 1;
  Space       sp = make_space(noa, nob, nva, nvb, k_range, k_spin, evl, is_restricted != 0);
  std::string v  = sp.validate();
  if(!v.empty()) return ctx->fail("invalid space: " + v);
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  free_pools(ctx);
  free_operands(ctx);
  ctx->sp         = sp;
  ctx->have_space = true;
  ctx->tasks      = enumerate_tasks(sp.noab(), sp.nvab(), sp.k_spin.data(), sp.restricted, &ctx->n_outer);
  ctx->owner      = partition_tasks(sp, ctx->tasks, ctx->opt.nranks, ctx->opt.symmetry != 0);
  if(ctx->d_evl) cudaFree(ctx->d_evl);
  CK(cudaMalloc(&ctx->d_evl, sp.evl.size() * 8));
  CK(cudaMemcpy(ctx->d_evl, sp.evl.data(), sp.evl.size() * 8, cudaMemcpyHostToDevice));
  return 0;
}

// =================================================================================================
// V2 from Cholesky vectors (row f2 of SURVEY.md 8f): replaces setupV2Tensors
// (exachem/cholesky/v2tensors.cpp:52-90, called at exachem/cc/ccsd_t/ccsd_t.cpp:168-193)
//   v2ijab(h1,h2,p1,p2) = L(h1,p1,c) L(h2,p2,c) - L(h1,p2,c) L(h2,p1,c)      v2tensors.cpp:68-69
//   v2ijka(h1,h2,h3,p1) = L(h1,h3,c) L(h2,p1,c) - L(h2,h3,c) L(h1,p1,c)      v2tensors.cpp:77-78
//   v2iabc(h1,p1,p2,p3) = L(h1,p2,c) L(p1,p3,c) - L(h1,p3,c) L(p1,p2,c)      v2tensors.cpp:85-86
// Each is one plain FP64 GEMM over the Cholesky index (cuBLAS, loaded on first use) followed by an
// antisymmetrising gather; v2iabc is formed one occupied row at a time, so the GEMM scratch is V^3, not O V^3.
// =================================================================================================
} // extern "C"

namespace {

// rows (p in [p0,p0+np), q in [q0,q0+nq)) of chol[N][N][ncv] packed as out[(p,q)][ncv]
__global__ void __launch_bounds__(256) pack_pairs_kernel(const double* __restrict__ chol, int64_t N, int64_t ncv, int64_t p0,
                                                         int64_t np, int64_t q0, int64_t nq, double* __restrict__ out) {
  const int64_t total = np * nq * ncv;
  for(int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t) gridDim.x * blockDim.x) {
    const int64_t c = e % ncv, pq = e / ncv, q = pq % nq, pp = pq / nq;
    out[e]          = chol[((p0 + pp) * N + (q0 + q)) * ncv + c];
  }
}
// v2ijab[h1,h2,p1,p2] from G[(h1,p1),(h2,p2)]  (G is OV x OV)
__global__ void __launch_bounds__(256) v2ijab_kernel(const double* __restrict__ G, int64_t O, int64_t V, double* __restrict__ out) {
  const int64_t total = O * O * V * V, OV = O * V;
  for(int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t) gridDim.x * blockDim.x) {
    int64_t       r  = e;
    const int64_t p2 = r % V; r /= V;
    const int64_t p1 = r % V; r /= V;
    const int64_t h2 = r % O, h1 = r / O;
    out[e] = G[(h1 * V + p1) * OV + h2 * V + p2] - G[(h1 * V + p2) * OV + h2 * V + p1];
  }
}
// v2ijka[h1,h2,h3,p1] from G[(h1,h3),(h2,p1)]  (G is OO x OV)
__global__ void __launch_bounds__(256) v2ijka_kernel(const double* __restrict__ G, int64_t O, int64_t V, double* __restrict__ out) {
  const int64_t total = O * O * O * V, OV = O * V;
  for(int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t) gridDim.x * blockDim.x) {
    int64_t       r  = e;
    const int64_t p1 = r % V; r /= V;
    const int64_t h3 = r % O; r /= O;
    const int64_t h2 = r % O, h1 = r / O;
    out[e] = G[(h1 * O + h3) * OV + h2 * V + p1] - G[(h2 * O + h3) * OV + h1 * V + p1];
  }
}
// one occupied row: v2iabc[h1,p1,p2,p3] from G[(p2),(p1,p3)]  (G is V x VV for this h1)
__global__ void __launch_bounds__(256) v2iabc_row_kernel(const double* __restrict__ G, int64_t V, double* __restrict__ out) {
  const int64_t total = V * V * V, VV = V * V;
  for(int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t) gridDim.x * blockDim.x) {
    int64_t       r  = e;
    const int64_t p3 = r % V; r /= V;
    const int64_t p2 = r % V, p1 = r / V;
    out[e] = G[p2 * VV + p1 * V + p3] - G[p3 * VV + p1 * V + p2];
  }
}

// the four cuBLAS entry points this file needs, resolved from libcublas.so.12 on first use
struct Cublas {
  void* lib = nullptr;
  void* handle = nullptr;
  int (*create)(void**) = nullptr;
  int (*destroy)(void*) = nullptr;
  int (*set_stream)(void*, cudaStream_t) = nullptr;
  int (*dgemm)(void*, int, int, int, int, int, const double*, const double*, int, const double*, int, const double*, double*,
               int) = nullptr;
  std::string open() {
    if(handle) return "";
    for(const char* name: {"libcublas.so.12", "libcublas.so"}) {
      lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if(lib) break;
    }
    if(!lib) return std::string("cannot load libcublas: ") + dlerror();
    create     = (int (*)(void**)) dlsym(lib, "cublasCreate_v2");
    destroy    = (int (*)(void*)) dlsym(lib, "cublasDestroy_v2");
    set_stream = (int (*)(void*, cudaStream_t)) dlsym(lib, "cublasSetStream_v2");
    dgemm      = (decltype(dgemm)) dlsym(lib, "cublasDgemm_v2");
    if(!create || !destroy || !set_stream || !dgemm) return "libcublas lacks cublasCreate/Destroy/SetStream/Dgemm";
    if(create(&handle) != 0) return "cublasCreate failed";
    return "";
  }
  // row-major C[m][n] = A[m][k] * B[n][k]^T
  int abt(cudaStream_t st, int m, int n, int k, const double* A, const double* B, double* C) {
    const double one = 1.0, zero = 0.0;
    set_stream(handle, st);
    return dgemm(handle, /*CUBLAS_OP_T*/ 1, /*CUBLAS_OP_N*/ 0, n, m, k, &one, B, k, A, k, &zero, C, n);
  }
  ~Cublas() {
    if(handle && destroy) destroy(handle);
  }
};
Cublas g_cublas;

inline unsigned grid_for(int64_t n) { return (unsigned) std::min<int64_t>((n + 255) / 256, 148 * 16); }

int dense_alloc(ccsdt_ctx* ctx, int tensor, size_t n) {
  if(ctx->dense[tensor] && ctx->dense_elems[tensor] != n) {
    CK(cudaFree(ctx->dense[tensor]));
    ctx->dense[tensor] = nullptr;
  }
  if(!ctx->dense[tensor]) {
    CK(cudaMalloc(&ctx->dense[tensor], n * 8));
    ctx->dense_elems[tensor] = n;
  }
  ctx->dense_sparse_ok[tensor] = false; // written in full by the caller (ccsdt_put_cholesky)
  return 0;
}

} // namespace

extern "C" {

int ccsdt_put_cholesky(ccsdt_ctx* ctx, const double* host_chol, int64_t ncv) {
  if(!ctx || !host_chol || ncv <= 0) return 1;
  if(!ctx->have_space) return ctx->fail("ccsdt_set_space must be called first");
  cudaSetDevice(ctx->device);
  const std::string cerr = g_cublas.open();
  if(!cerr.empty()) return ctx->fail(cerr, 10);
  const int64_t O = ctx->sp.n_occ(), V = ctx->sp.n_virt(), N = O + V;
  if(O * V * ncv > 0x7fffffffll || V * V > 0x7fffffffll) return ctx->fail("ccsdt_put_cholesky: GEMM dimension exceeds int32", 10);
  CK(cudaStreamSynchronize(ctx->s_compute));
  CK(cudaStreamSynchronize(ctx->s_compute2));
  cudaStream_t st = ctx->s_stage;
  double *d_chol = nullptr, *Loo = nullptr, *Lov = nullptr, *Lvv = nullptr, *G = nullptr;
  struct Scratch { // the temporaries are released on every way out (the CK macro returns on failure)
    double** p[5];
    ~Scratch() {
      for(double** q: p)
        if(*q) cudaFree(*q);
    }
  } scratch{{&d_chol, &Loo, &Lov, &Lvv, &G}};
  const size_t nchol = (size_t) N * N * ncv;
  CK(cudaMalloc(&d_chol, nchol * 8));
  CK(cudaMemcpyAsync(d_chol, host_chol, nchol * 8, cudaMemcpyHostToDevice, st));
  ctx->pending_h2d += (int64_t) nchol * 8;
  CK(cudaMalloc(&Loo, (size_t) O * O * ncv * 8));
  CK(cudaMalloc(&Lov, (size_t) O * V * ncv * 8));
  CK(cudaMalloc(&Lvv, (size_t) V * V * ncv * 8));
  pack_pairs_kernel<<<grid_for(O * O * ncv), 256, 0, st>>>(d_chol, N, ncv, 0, O, 0, O, Loo);
  pack_pairs_kernel<<<grid_for(O * V * ncv), 256, 0, st>>>(d_chol, N, ncv, 0, O, O, V, Lov);
  pack_pairs_kernel<<<grid_for(V * V * ncv), 256, 0, st>>>(d_chol, N, ncv, O, V, O, V, Lvv);
  CK(cudaGetLastError());
  // scratch for the largest product: (OV x OV), (OO x OV) or one row of (V x VV)
  const size_t gmax = std::max({(size_t) O * V * O * V, (size_t) O * O * O * V, (size_t) V * V * V});
  CK(cudaMalloc(&G, gmax * 8));
  if(int rc = dense_alloc(ctx, CCSDT_V_IJAB, (size_t) O * O * V * V)) return rc;
  if(int rc = dense_alloc(ctx, CCSDT_V_IJKA, (size_t) O * O * O * V)) return rc;
  if(int rc = dense_alloc(ctx, CCSDT_V_IABC, (size_t) O * V * V * V)) return rc;
  int brc = 0;
  brc |= g_cublas.abt(st, (int) (O * V), (int) (O * V), (int) ncv, Lov, Lov, G);
  v2ijab_kernel<<<grid_for(O * O * V * V), 256, 0, st>>>(G, O, V, ctx->dense[CCSDT_V_IJAB]);
  brc |= g_cublas.abt(st, (int) (O * O), (int) (O * V), (int) ncv, Loo, Lov, G);
  v2ijka_kernel<<<grid_for(O * O * O * V), 256, 0, st>>>(G, O, V, ctx->dense[CCSDT_V_IJKA]);
  for(int64_t h1 = 0; h1 < O; h1++) {
    brc |= g_cublas.abt(st, (int) V, (int) (V * V), (int) ncv, Lov + h1 * V * ncv, Lvv, G);
    v2iabc_row_kernel<<<grid_for(V * V * V), 256, 0, st>>>(G, V, ctx->dense[CCSDT_V_IABC] + h1 * V * V * V);
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  ctx->stats.kernel_launches += 5 + 2 * O;
  if(brc) return ctx->fail("cublasDgemm failed", 10);
  ctx->synthetic = false;
  return 0;
}

static int put_dense_impl(ccsdt_ctx* ctx, int tensor, const double* host, bool async) {
  if(!ctx || tensor < 0 || tensor > 4 || !host) return 1;
  if(!ctx->have_space) return ctx->fail("ccsdt_set_space must be called first");
  cudaSetDevice(ctx->device);
  size_t n = 1;
  for(const char* k = kKinds[tensor]; *k; k++) n *= (size_t) dim_full(ctx->sp, *k);
  // re-uploads (same space) reuse the allocation; the copy is stream-ordered before the next panel build
  if(ctx->dense[tensor] && ctx->dense_elems[tensor] != n) {
    CK(cudaStreamSynchronize(ctx->s_compute));
    CK(cudaFree(ctx->dense[tensor]));
    ctx->dense[tensor] = nullptr;
  }
  CK(cudaStreamSynchronize(ctx->s_compute)); // a running task may still read the old contents
  CK(cudaStreamSynchronize(ctx->s_compute2));
  if(!ctx->dense[tensor]) {
    CK(cudaMalloc(&ctx->dense[tensor], n * 8));
    ctx->dense_elems[tensor]     = n;
    ctx->dense_sparse_ok[tensor] = false;
  }
  cudaStream_t sa = async ? ctx->s_copy_a : ctx->s_stage, sb = async ? ctx->s_copy_b : ctx->s_stage;
  if(!ctx->dense_sparse_ok[tensor]) {
    CK(cudaMemsetAsync(ctx->dense[tensor], 0, n * 8, sa)); // spin-forbidden blocks stay zero for good
    CK(cudaStreamSynchronize(sa));
    ctx->dense_sparse_ok[tensor] = true;
  }
  // Only the spin-conserving blocks cross the bus: T1[a,i] with s_a = s_i, four-index tensors with
  // s_0 + s_1 = s_2 + s_3 (6 of the 16 spin patterns).  These are the only blocks any enabled term of any task
  // reads -- and the only ones the reference ever requests through Tensor::get -- so the rest of the dense host
  // array is never looked at.  One strided 3-d copy per (leading index, spin pattern).
  const char*  kinds = kKinds[tensor];
  const int    nd    = (int) strlen(kinds);
  int64_t      full[4] = {1, 1, 1, 1}, lo[4][2], len[4][2];
  for(int d = 0; d < nd; d++) {
    const bool virt = kinds[d] == 'v';
    full[d]         = dim_full(ctx->sp, kinds[d]);
    int     tb, te;
    int64_t na = 0, nbeta = 0;
    ctx->sp.spin_range(virt, 1, tb, te, na);
    ctx->sp.spin_range(virt, 2, tb, te, nbeta);
    lo[d][0] = 0, len[d][0] = na, lo[d][1] = na, len[d][1] = nbeta;
  }
  int64_t sent = 0;
  if(nd == 2) {
    for(int s0 = 0; s0 < 2; s0++) {
      if(len[0][s0] <= 0 || len[1][s0] <= 0) continue;
      const int64_t off = lo[0][s0] * full[1] + lo[1][s0];
      CK(cudaMemcpy2DAsync(ctx->dense[tensor] + off, (size_t) full[1] * 8, host + off, (size_t) full[1] * 8,
                           (size_t) len[1][s0] * 8, (size_t) len[0][s0], cudaMemcpyHostToDevice, s0 == 0 ? sa : sb));
      sent += len[0][s0] * len[1][s0] * 8;
    }
  }
  else {
    for(int pat = 0; pat < 16; pat++) {
      const int sp_[4] = {(pat >> 3) & 1, (pat >> 2) & 1, (pat >> 1) & 1, pat & 1};
      if(sp_[0] + sp_[1] != sp_[2] + sp_[3]) continue;
      bool empty = false;
      for(int d = 0; d < 4; d++) empty |= len[d][sp_[d]] <= 0;
      if(empty) continue;
      // one 3-d copy per value of the SHORTEST of the three leading indices (fewest calls); rows run along the last
      // index, the copy's y and z are the two other leading indices A < B with their natural strides
      const int64_t stride[4] = {full[1] * full[2] * full[3], full[2] * full[3], full[3], 1};
      int           L = 0;
      for(int d = 1; d < 3; d++)
        if(len[d][sp_[d]] < len[L][sp_[L]]) L = d;
      const int A = L == 0 ? 1 : 0, B = L == 2 ? 1 : 2;
      int64_t   base = 0;
      for(int d = 0; d < 4; d++) base += lo[d][sp_[d]] * stride[d];
      for(int64_t il = 0; il < len[L][sp_[L]]; il++) {
        const int64_t     off = base + il * stride[L];
        cudaMemcpy3DParms cp{};
        cp.srcPtr = make_cudaPitchedPtr((void*) (host + off), (size_t) stride[B] * 8, (size_t) full[3] * 8,
                                        (size_t) (stride[A] / stride[B]));
        cp.dstPtr = make_cudaPitchedPtr((void*) (ctx->dense[tensor] + off), (size_t) stride[B] * 8, (size_t) full[3] * 8,
                                        (size_t) (stride[A] / stride[B]));
        cp.extent = make_cudaExtent((size_t) len[3][sp_[3]] * 8, (size_t) len[B][sp_[B]], (size_t) len[A][sp_[A]]);
        cp.kind   = cudaMemcpyHostToDevice;
        CK(cudaMemcpy3DAsync(&cp, pat == 0 ? sa : sb));
        sent += len[A][sp_[A]] * len[B][sp_[B]] * len[3][sp_[3]] * 8;
      }
    }
  }
  if(async) {
    CK(cudaEventRecord(ctx->ev_alpha[tensor], sa));
    CK(cudaEventRecord(ctx->ev_full[tensor], sb));
    ctx->upload_pending[tensor] = true;
  }
  else {
    CK(cudaStreamSynchronize(ctx->s_stage)); // the caller may reuse `host` on return
    ctx->upload_pending[tensor] = false;
  }
  ctx->pending_h2d += sent;                  // reported by the next run's stats
  ctx->synthetic = false;
  return 0;
}

int ccsdt_put_dense(ccsdt_ctx* ctx, int tensor, const double* host) { return put_dense_impl(ctx, tensor, host, false); }
int ccsdt_put_dense_async(ccsdt_ctx* ctx, int tensor, const double* host) { return put_dense_impl(ctx, tensor, host, true); }

int ccsdt_put_block(ccsdt_ctx* ctx, int tensor, const uint32_t bid[4], const double* host) {
  if(!ctx || tensor < 0 || tensor > 4 || !host || !bid) return 1;
  if(!ctx->have_space) return ctx->fail("ccsdt_set_space must be called first");
  cudaSetDevice(ctx->device);
  const int nd = (int) strlen(kKinds[tensor]);
  for(int d = 0; d < nd; d++) {
    const int lim = kKinds[tensor][d] == 'o' ? ctx->sp.noab() : ctx->sp.nvab();
    if((int) bid[d] >= lim) return ctx->fail("block id out of range");
  }
  const size_t n = block_elems(ctx->sp, tensor, bid);
  BlockKey     key{tensor, {bid[0], bid[1], nd > 2 ? bid[2] : 0u, nd > 3 ? bid[3] : 0u}};
  auto         it = ctx->blocks.find(key);
  double*      dev;
  if(it == ctx->blocks.end()) {
    CK(cudaMalloc(&dev, n * 8));
    ctx->blocks.emplace(key, BlockEntry{dev, n * 8, (int64_t) 1 << 60 /* never evicted */});
    ctx->block_bytes += n * 8;
  }
  else dev = it->second.dev;
  CK(cudaMemcpy(dev, host, n * 8, cudaMemcpyHostToDevice));
  ctx->pending_h2d += (int64_t) n * 8;
  ctx->synthetic = false;
  return 0;
}

int ccsdt_set_fetch(ccsdt_ctx* ctx, ccsdt_fetch_fn fn, void* user) {
  if(!ctx) return 1;
  ctx->fetch      = fn;
  ctx->fetch_user = user;
  if(fn) ctx->synthetic = false;
  return 0;
}

int ccsdt_set_synthetic(ccsdt_ctx* ctx, uint64_t seed) {
  if(!ctx) return 1;
  ctx->synthetic = true;
  ctx->seed      = seed;
  return 0;
}

int ccsdt_set_task_counter(ccsdt_ctx* ctx, int64_t* counter) {
  if(!ctx) return 1;
  ctx->task_counter = counter;
  return 0;
}

int ccsdt_run(ccsdt_ctx* ctx, int64_t task_begin, int64_t task_end, double energies[2], double* per_task,
              ccsdt_stats* stats_out) {
  if(!ctx || !energies) return 1;
  if(!ctx->have_space) return ctx->fail("ccsdt_set_space must be called first");
  const int64_t nt = (int64_t) ctx->tasks.size();
  if(task_end < 0 || task_end > nt) task_end = nt;
  if(task_begin < 0) task_begin = 0;
  if(task_begin > task_end) task_begin = task_end;
  std::vector<int64_t> ids;
  for(int64_t i = task_begin; i < task_end; i++) ids.push_back(i);
  // the whole list uses the split computed at set_space/set_options; a sub-range re-balances itself
  return run_task_list(ctx, ids, task_begin == 0 && task_end == nt, energies, per_task, stats_out);
}

int ccsdt_run_tasks(ccsdt_ctx* ctx, const int64_t* task_ids, int64_t n, double energies[2], double* per_task,
                    ccsdt_stats* stats_out) {
  if(!ctx || !energies || (n > 0 && !task_ids)) return 1;
  if(!ctx->have_space) return ctx->fail("ccsdt_set_space must be called first");
  std::vector<int64_t> ids(task_ids, task_ids + (n > 0 ? n : 0));
  for(int64_t id: ids)
    if(id < 0 || id >= (int64_t) ctx->tasks.size()) return ctx->fail("task id out of range");
  return run_task_list(ctx, ids, false, energies, per_task, stats_out);
}

// ---- diagnostics --------------------------------------------------------------------------------
static int probe_device(int device) {
  int ndev = 0;
  if(cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    g_create_error = "no usable CUDA device";
    return 2;
  }
  cudaSetDevice(device);
  return 0;
}

int ccsdt_probe_fp64_peak(int device, int use_dmma, int iters, double* tflops, double* ms) {
  if(int rc = probe_device(device)) return rc;
  cudaError_t e = probe_fp64_peak(use_dmma, iters, tflops, ms);
  if(e != cudaSuccess) g_create_error = cudaGetErrorString(e);
  return e == cudaSuccess ? 0 : 2;
}
int ccsdt_probe_mainloop(int device, int ta, int tb, int warps_per_cta, int ctas_per_sm, int iters, double* tflops) {
  if(int rc = probe_device(device)) return rc;
  cudaError_t e = probe_mainloop(ta, tb, warps_per_cta, ctas_per_sm, iters, tflops);
  if(e != cudaSuccess) g_create_error = cudaGetErrorString(e);
  return e == cudaSuccess ? 0 : 2;
}
int ccsdt_probe_dmma_layout(int device, double* c_out, const double* a, const double* b) {
  if(int rc = probe_device(device)) return rc;
  cudaError_t e = probe_dmma_layout(c_out, a, b);
  if(e != cudaSuccess) g_create_error = cudaGetErrorString(e);
  return e == cudaSuccess ? 0 : 2;
}
int ccsdt_probe_tma_swizzle(int device, double* smem_dump, int rows) {
  if(int rc = probe_device(device)) return rc;
  cudaError_t e = probe_tma_swizzle(smem_dump, rows);
  if(e != cudaSuccess) g_create_error = cudaGetErrorString(e);
  return e == cudaSuccess ? 0 : 2;
}
int ccsdt_synth_block(int device, uint64_t seed, int tensor, int noa, int nob, int nva, int nvb, const int64_t lo[4],
                      const int64_t n[4], double* host_out) {
  if(int rc = probe_device(device)) return rc;
  SynthInfo   si{seed, noa, nob, nva, nvb};
  cudaError_t e = synth_block_device(si, tensor, lo, n, host_out);
  if(e != cudaSuccess) g_create_error = cudaGetErrorString(e);
  return e == cudaSuccess ? 0 : 2;
}

} // extern "C"
