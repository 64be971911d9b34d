// C ABI (include/ccsdt_b200.h) and the per-GPU driver of the fused CCSD(T) triples path: context life cycle, the
// execution tiling, per-task panel staging, the task loop with staging / fetch / compute overlap, the deterministic
// energy reduction.  Replaces CCSD_T_Fused_Driver<T>::execute's loop
// (exachem/cc/ccsd_t/ccsd_t_fused_driver.hpp:368-483) and ccsd_t_fully_fused_none_df_none_task
// (exachem/cc/ccsd_t/ccsd_t_all_fused.hpp:77-286).  The operand store is ccsdt_store.cu.
#include "ccsdt_ctx.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

using namespace ccsdt;

namespace {

std::string g_create_error;

// contexts parked by ccsdt_destroy, one per device (see the header: execute creates and destroys one per call)
std::mutex                g_park_mutex;
std::map<int, ccsdt_ctx*> g_parked;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int round_up(int64_t x, int64_t m) { return (int) ((x + m - 1) / m * m); }

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

void free_pools(ccsdt_ctx* ctx) {
  for(auto& b: ctx->buf) {
    for(int pl = 0; pl < 2; pl++) {
      if(b.geom.hpp[pl]) cudaFree(b.geom.hpp[pl]);
      if(b.geom.hhp[pl]) cudaFree(b.geom.hhp[pl]);
      b.geom.hpp[pl] = b.geom.hhp[pl] = nullptr;
      b.hpp_cap[pl] = b.hhp_cap[pl] = 0;
    }
    if(b.s1_a) cudaFree(b.s1_a);
    if(b.s1_b) cudaFree(b.s1_b);
    if(b.d_descs) cudaFree(b.d_descs);
    if(b.h_descs) cudaFreeHost(b.h_descs);
    if(b.d_partial) cudaFree(b.d_partial);
    if(b.d_counter) cudaFree(b.d_counter);
    b.d_counter = nullptr;
    b.s1_a = b.s1_b = nullptr;
    b.s1a_cap = b.s1b_cap = 0;
    b.d_descs = b.h_descs = nullptr;
    b.d_partial           = nullptr;
    b.desc_cap            = 0;
    b.partial_cap         = 0;
  }
  ctx->pools_ready = false;
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

// (re)allocates *p when it is smaller than `need`; new memory is zeroed: padding rows of the panels are read by TMA
// boxes that overhang a ragged tile and must hold finite numbers (their elements are masked in the epilogue)
int ensure_dev(ccsdt_ctx* ctx, double** p, size_t* cap, size_t need, bool* changed) {
  if(*p && *cap >= need) return 0;
  if(*p) CK(cudaFree(*p));
  *p = nullptr;
  CK(cudaMalloc(p, need));
  CK(cudaMemset(*p, 0, need));
  *cap     = need;
  *changed = true;
  return 0;
}

// Panel pools of the two staging buffers.  Allocations are kept (also across ccsdt_destroy / ccsdt_create of a parked
// context) and only grow; the geometry and the four tensor maps follow the execution tiling in force.
int ensure_pools(ccsdt_ctx* ctx) {
  if(ctx->pools_ready) return 0;
  const Space& sp = ctx->sp;
  int64_t      th = sp.max_hole_tile(), tp = sp.max_particle_tile();
  int          THp = 0;
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

  const int nbuf    = ctx->opt.overlap ? 2 : 1;
  bool      changed = false;
  for(int i = 0; i < nbuf; i++) {
    StageBuf& b = ctx->buf[i];
    if(b.geom.THp != THp || b.geom.TPp != TPp || b.geom.Kp[0] != Kp1 || b.geom.Kp[1] != Kp2) changed = true;
    b.geom.THp   = THp;
    b.geom.TPp   = TPp;
    b.geom.Kp[0] = Kp1;
    b.geom.Kp[1] = Kp2;
    for(int pl = 0; pl < 2; pl++) {
      const size_t hpp = (size_t) 9 * b.geom.hpp_stride_panel(pl) * 8, hhp = (size_t) 9 * b.geom.hhp_stride_panel(pl) * 8;
      if(int rc = ensure_dev(ctx, &b.geom.hpp[pl], &b.hpp_cap[pl], hpp, &changed)) return rc;
      if(int rc = ensure_dev(ctx, &b.geom.hhp[pl], &b.hhp_cap[pl], hhp, &changed)) return rc;
    }
    const size_t s1a = (size_t) 9 * THp * TPp * 8, s1b = (size_t) 9 * TPp * TPp * THp * THp * 8;
    if(int rc = ensure_dev(ctx, &b.s1_a, &b.s1a_cap, s1a, &changed)) return rc;
    if(int rc = ensure_dev(ctx, &b.s1_b, &b.s1b_cap, s1b, &changed)) return rc;
    if(!b.d_counter) CK(cudaMalloc(&b.d_counter, 4 * COUNTER_WORDS));
    if(!b.staged) {
      CK(cudaEventCreateWithFlags(&b.staged, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&b.done, cudaEventDisableTiming));
      CK(cudaEventCreate(&b.k0));
      CK(cudaEventCreate(&b.k1));
      CK(cudaEventCreate(&b.g0));
      CK(cudaEventCreate(&b.g1));
    }
  }
  if(changed) {
    CK(cudaDeviceSynchronize()); // the zero fills ran on the legacy stream; ours are non-blocking streams
    for(int i = 0; i < nbuf; i++)
      if(int rc = make_tmaps(ctx, ctx->buf[i])) return rc;
  }
  ctx->pools_ready = true;
  return 0;
}

// ---- the sources of one task ---------------------------------------------------------------------
struct TaskLayout {
  TermDev terms[MAX_TERMS];
  int     nterms = 0;
  S1Dev   s1[9];
  int     ns1 = 0;
};

// Walks every operand source of task t: on_src(dst, dst strides, SrcSpec) for each (term, contracted tile) pair and
// the two s1 operands, on_zero(dst, strides, extents) for the K tails.  Block ids are EXECUTION tiles inside the
// tensor's own sub-space.  Used by the panel staging (descriptors) and by the prefetch (block keys only); lay
// (optional) receives the term tables of the fused kernel.
// Reference: ccsd_t_data_d1_new / _d2_new / _s1_new, exachem/cc/ccsd_t/ccsd_t_all_fused_{doubles1,doubles2,singles}.hpp.
template<class OnSrc, class OnZero>
int walk_task_sources(const Space& sp, const PoolGeom& g, double* s1_a, double* s1_b, const Task& t, TaskLayout* lay,
                      OnSrc&& on_src, OnZero&& on_zero) {
  const int noab = sp.noab();
  int       ext[6];
  for(int i = 0; i < 6; i++) ext[i] = (int) sp.k_range[t.t[i]];
  auto tile = [&](int idx) { return (uint32_t) t.t[idx]; };
  auto vt   = [&](int idx) { return (uint32_t) (t.t[idx] - noab); };
  auto spec = [](int tensor, uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3, int d0, int d1, int d2, int d3, double scale) {
    SrcSpec s{};
    s.tensor = tensor;
    s.bid[0] = b0, s.bid[1] = b1, s.bid[2] = b2, s.bid[3] = b3;
    s.dim_of[0] = d0, s.dim_of[1] = d1, s.dim_of[2] = d2, s.dim_of[3] = d3;
    s.scale = scale;
    return s;
  };
  TaskLayout local;
  if(!lay) lay = &local;
  lay->nterms = lay->ns1 = 0;

  // ---- d1 terms: HPP = sign * T2[pa,pb,hx,l] as [hx][pa][pb][l], HHP = v2ijka[hy,hz,l,pc] as [hy][hz][pc][l] ----
  for(int k = 0; k < 9; k++) {
    const int s7 = d1_contracted_spin(sp, t, k);
    if(!s7) continue;
    const D1Term& T = kD1[k];
    int           tb, te;
    int64_t       K;
    sp.spin_range(false, s7, tb, te, K);
    if(K <= 0) continue;
    TermDev td{};
    td.pool        = 0;
    td.kslabs      = (int) ((K + KSLAB - 1) / KSLAB);
    td.ksteps_last = (int) ((K - (int64_t) (td.kslabs - 1) * KSLAB + 3) / 4);
    td.hpp_hole    = T.hx;
    td.qt          = T.pa;
    td.inner_hpp   = T.pb;
    td.hhp_o2      = T.hy;
    td.hhp_o1      = T.hz;
    td.inner_hhp   = T.pc;
    td.a_is_hpp    = (T.pb == 5);
    td.layout_y    = (T.pa == 4);
    td.hpp_panel = td.hhp_panel = k;
    double*       hpp = g.hpp[0] + k * g.hpp_stride_panel(0);
    double*       hhp = g.hhp[0] + k * g.hhp_stride_panel(0);
    const int64_t ds_hpp[4] = {g.hpp_stride_o2(0), g.hpp_stride_o1(0), g.hpp_stride_in(0), 1};
    const int64_t ds_hhp[4] = {g.hhp_stride_o2(0), g.hhp_stride_o1(0), g.hhp_stride_in(0), 1};
    int64_t       koff = 0;
    for(int h7 = tb; h7 < te; h7++) {
      // T2[pa,pb,hx,l]: canonical block order (min(hx,h7), max(hx,h7)), sign -1 when transposed
      const SrcSpec a = h7 < t.t[T.hx]
                          ? spec(CCSDT_T2, vt(T.pa), vt(T.pb), (uint32_t) h7, tile(T.hx), 3, 0, 1, 2, -1.0 * T.sign)
                          : spec(CCSDT_T2, vt(T.pa), vt(T.pb), tile(T.hx), (uint32_t) h7, 2, 0, 1, 3, 1.0 * T.sign);
      if(int rc = on_src(hpp + koff, ds_hpp, a)) return rc;
      const SrcSpec v = spec(CCSDT_V_IJKA, tile(T.hy), tile(T.hz), (uint32_t) h7, vt(T.pc), 0, 1, 3, 2, 1.0);
      if(int rc = on_src(hhp + koff, ds_hhp, v)) return rc;
      koff += sp.k_range[h7];
    }
    const int tail = td.kslabs * KSLAB - (int) K;
    if(tail > 0) {
      const int n_hpp[4] = {ext[T.hx], ext[T.pa], ext[T.pb], tail};
      const int n_hhp[4] = {ext[T.hy], ext[T.hz], ext[T.pc], tail};
      if(int rc = on_zero(hpp + K, ds_hpp, n_hpp)) return rc;
      if(int rc = on_zero(hhp + K, ds_hhp, n_hhp)) return rc;
    }
    lay->terms[lay->nterms++] = td;
  }
  // ---- d2 terms: HHP = sign * T2[pa,d,hx,hy] as [hx][hy][pa][d], HPP = v2iabc[hz,d,pb,pc] as [hz][pb][pc][d] ----
  for(int k = 0; k < 9; k++) {
    const int s7 = d2_contracted_spin(sp, t, k);
    if(!s7) continue;
    const D2Term& T = kD2[k];
    int           tb, te;
    int64_t       K;
    sp.spin_range(true, s7, tb, te, K);
    if(K <= 0) continue;
    TermDev td{};
    td.pool        = 1;
    td.kslabs      = (int) ((K + KSLAB - 1) / KSLAB);
    td.ksteps_last = (int) ((K - (int64_t) (td.kslabs - 1) * KSLAB + 3) / 4);
    td.hpp_hole    = T.hz;
    td.qt          = T.pb;
    td.inner_hpp   = T.pc;
    td.hhp_o2      = T.hx;
    td.hhp_o1      = T.hy;
    td.inner_hhp   = T.pa;
    td.a_is_hpp    = (T.pc == 5);
    td.layout_y    = (T.pb == 4);
    td.hpp_panel = td.hhp_panel = k;
    double*       hpp = g.hpp[1] + k * g.hpp_stride_panel(1);
    double*       hhp = g.hhp[1] + k * g.hhp_stride_panel(1);
    const int64_t ds_hpp[4] = {g.hpp_stride_o2(1), g.hpp_stride_o1(1), g.hpp_stride_in(1), 1};
    const int64_t ds_hhp[4] = {g.hhp_stride_o2(1), g.hhp_stride_o1(1), g.hhp_stride_in(1), 1};
    int64_t       koff = 0;
    for(int p7 = tb; p7 < te; p7++) {
      const uint32_t v7 = (uint32_t) (p7 - noab);
      const SrcSpec  a  = p7 < t.t[T.pa] ? spec(CCSDT_T2, v7, vt(T.pa), tile(T.hx), tile(T.hy), 2, 3, 1, 0, -1.0 * T.sign)
                                         : spec(CCSDT_T2, vt(T.pa), v7, tile(T.hx), tile(T.hy), 2, 3, 0, 1, 1.0 * T.sign);
      if(int rc = on_src(hhp + koff, ds_hhp, a)) return rc;
      const SrcSpec v = spec(CCSDT_V_IABC, tile(T.hz), v7, vt(T.pb), vt(T.pc), 0, 2, 3, 1, 1.0);
      if(int rc = on_src(hpp + koff, ds_hpp, v)) return rc;
      koff += sp.k_range[p7];
    }
    const int tail = td.kslabs * KSLAB - (int) K;
    if(tail > 0) {
      const int n_hpp[4] = {ext[T.hz], ext[T.pb], ext[T.pc], tail};
      const int n_hhp[4] = {ext[T.hx], ext[T.hy], ext[T.pa], tail};
      if(int rc = on_zero(hpp + K, ds_hpp, n_hpp)) return rc;
      if(int rc = on_zero(hhp + K, ds_hhp, n_hhp)) return rc;
    }
    lay->terms[lay->nterms++] = td;
  }
  // ---- s1 terms: a = sign*T1[pa,hx] as [hx][pa], b = v2ijab[hz,hy,pc,pb] as [hy][hz][pb][pc] ----
  for(int k = 0; k < 9; k++) {
    if(!s1_enabled(sp, t, k)) continue;
    const S1Term& T  = kS1[k];
    double*       pa = s1_a + (int64_t) k * g.THp * g.TPp;
    double*       pb = s1_b + (int64_t) k * g.TPp * g.TPp * g.THp * g.THp;
    S1Dev&        sd = lay->s1[lay->ns1++];
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
    const SrcSpec a       = spec(CCSDT_T1, vt(T.pa), tile(T.hx), 0, 0, -1, -1, 1, 0, (double) T.sign); // dst (1, 1, hx, pa)
    const int64_t ds_a[4] = {0, 0, g.TPp, 1};
    if(int rc = on_src(pa, ds_a, a)) return rc;
    const SrcSpec v       = spec(CCSDT_V_IJAB, tile(T.hz), tile(T.hy), vt(T.pc), vt(T.pb), 1, 0, 3, 2, 1.0); // dst (hy, hz, pb, pc)
    const int64_t ds_b[4] = {(int64_t) g.THp * g.TPp * g.TPp, (int64_t) g.TPp * g.TPp, g.TPp, 1};
    if(int rc = on_src(pb, ds_b, v)) return rc;
  }
  return 0;
}

// one source -> panel-build work items: a single item for a dense or procedural tensor, one item per storage block
// (cut to the overlap with the execution tile) for the block store
int add_gather(ccsdt_ctx* ctx, std::vector<GatherDesc>& descs, int64_t& max_elems, double* dst, const int64_t ds[4],
               const SrcSpec& s, std::vector<SrcPiece>& scratch) {
  const Space& sp    = ctx->sp;
  const char*  kinds = kKinds[s.tensor];
  GatherDesc   g;
  memset(&g, 0, sizeof(g));
  g.synth_tensor = -1;
  if(ctx->synthetic || ctx->dense[s.tensor]) {
    BlockRef ref;
    if(ctx->synthetic) g.synth_tensor = s.tensor;
    else {
      if(int rc = resolve_dense(ctx, s.tensor, s.bid, ref)) return rc;
      g.src = ref.ptr;
    }
    g.dst         = dst;
    g.scale       = s.scale;
    int64_t total = 1;
    for(int j = 0; j < 4; j++) {
      g.ds[j]     = ds[j];
      const int d = s.dim_of[j];
      if(d < 0) {
        g.n[j] = 1, g.ss[j] = 0, g.gpos[j] = 3 /* harmless: adds 0 */, g.gbase[j] = 0;
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
    descs.push_back(g);
    return 0;
  }
  scratch.clear();
  source_pieces(ctx, s, scratch);
  for(const SrcPiece& p: scratch) {
    BlockRef ref;
    if(int rc = resolve_block(ctx, p.key, p.elems, ctx->use_clock, ref)) return rc;
    // strides of the REQUESTED dims inside the canonical storage block
    int64_t st[4] = {0, 0, 0, 0};
    for(int d = 0; kinds[d]; d++) st[d] = ref.stride[p.perm[d]];
    const double* src  = ref.ptr;
    double*       out  = dst;
    int64_t       total = 1;
    for(int d = 0; kinds[d]; d++) src += p.store_off[d] * st[d];
    g.scale = s.scale * p.sign;
    for(int j = 0; j < 4; j++) {
      g.ds[j]     = ds[j];
      const int d = s.dim_of[j];
      if(d < 0) {
        g.n[j] = 1, g.ss[j] = 0;
        continue;
      }
      g.n[j]  = (int32_t) p.len[d];
      g.ss[j] = st[d];
      out += p.exec_off[d] * ds[j];
      total *= g.n[j];
    }
    g.src = src;
    g.dst = out;
    max_elems = std::max(max_elems, total);
    descs.push_back(g);
  }
  return 0;
}

// ---- staging of one task ------------------------------------------------------------------------
int stage_task(ccsdt_ctx* ctx, StageBuf& b, const Task& t, std::vector<GatherDesc>& descs, std::vector<SrcPiece>& scratch) {
  const Space&    sp = ctx->sp;
  const PoolGeom& g  = b.geom;
  TaskParams&     P  = b.params;
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

  descs.clear();
  int64_t    max_elems = 1;
  TaskLayout lay;
  int        rc = walk_task_sources(
    sp, g, b.s1_a, b.s1_b, t, &lay,
    [&](double* dst, const int64_t ds[4], const SrcSpec& s) { return add_gather(ctx, descs, max_elems, dst, ds, s, scratch); },
    [&](double* dst, const int64_t ds[4], const int n[4]) {
      GatherDesc z;
      memset(&z, 0, sizeof(z));
      z.dst          = dst;
      z.synth_tensor = -1;
      int64_t total  = 1;
      for(int j = 0; j < 4; j++) z.ds[j] = ds[j], z.n[j] = n[j], total *= n[j];
      if(total > 0) {
        max_elems = std::max(max_elems, total);
        descs.push_back(z);
      }
      return 0;
    });
  if(rc) return rc;
  // X terms first, then Y (stable)
  P.nterms = lay.nterms;
  int w    = 0;
  for(int i = 0; i < lay.nterms; i++)
    if(!lay.terms[i].layout_y) P.term[w++] = lay.terms[i];
  P.nterms_x = w;
  for(int i = 0; i < lay.nterms; i++)
    if(lay.terms[i].layout_y) P.term[w++] = lay.terms[i];
  P.ns1 = lay.ns1;
  memcpy(P.s1, lay.s1, sizeof(S1Dev) * lay.ns1);

  // ---- launch geometry, box order, partial buffer ----
  int64_t need_partial;
  double  executed = 0.0;
  if(ctx->opt.kernel == CCSDT_KERNEL_SIMPLE) {
    int64_t total = 1;
    for(int i = 0; i < 6; i++) total *= ext[i];
    need_partial    = (total + 255) / 256;
    b.eval_fraction = 1.0;
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
    b.grid = (int) std::min<int64_t>(nboxes, in_flight);
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
    // Boundary layers: a hole tile of ODD extent ends in boxes that hold one valid hole index where the sub-box has two
    // slots; the kernel runs those boxes through half-size instantiations of its k-loops (consume_term<.., SKIP>).  The
    // boxes are therefore handed out by class -- interior first, then the boundary layer of h1, h2, h3 -- so that the CTAs
    // of an SM run the same few instantiations at any time (instruction cache); inside a class the brick-major order stays.
    int odd = 0;
    for(int j = 0; j < 3; j++)
      if(P.c[j] == 2 && (ext[j] & 1)) odd |= 1 << j;
    auto warps_with_work = [](int extent, int off) { return std::max(0, std::min(4, (extent - off + 1) / 2)); };
    double units_x = 0, units_y = 0;
    if(P.sym || odd) {
      std::array<int, 19> key;
      for(int i = 0; i < 6; i++) key[i] = P.nbox[i], key[6 + i] = P.brick[i], key[12 + i] = P.nbrick[i];
      key[18]  = P.sym | (odd << 8) | ((ext[3] & 7) << 12) | ((ext[4] & 7) << 16);
      auto& bl = ctx->box_lists[key];
      if(!bl.dev) {
        std::vector<int32_t> ids[4]; // class: 0 = interior, 1 + j = boundary layer of hole j
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
          if(!valid || (P.sym && box_weight(P.sym, bx) <= 0)) continue;
          int cls = 0;
          for(int j = 2; j >= 0; j--)
            if(((odd >> j) & 1) && bx[j] == P.nbox[j] - 1) cls = 1 + j;
          ids[cls].push_back((int32_t) id);
          const double per_warp = cls ? 8.0 : 16.0;
          bl.units_x += per_warp * warps_with_work(ext[3], bx[3] * PBOX);
          bl.units_y += per_warp * warps_with_work(ext[4], bx[4] * PBOX);
        }
        std::vector<int32_t> all;
        for(auto& v: ids) all.insert(all.end(), v.begin(), v.end());
        bl.n = (int32_t) all.size();
        CK(cudaMalloc(&bl.dev, sizeof(int32_t) * std::max<size_t>(all.size(), 1)));
        CK(cudaMemcpy(bl.dev, all.data(), sizeof(int32_t) * all.size(), cudaMemcpyHostToDevice));
      }
      P.box_list = bl.dev;
      P.nlist    = bl.n;
      nboxes     = bl.n;
      b.grid     = (int) std::max<int64_t>(1, std::min<int64_t>(nboxes, in_flight));
      units_x = bl.units_x, units_y = bl.units_y;
    }
    else {
      // no list: every box runs the full loops; warps whose tile-particle slots are all outside the tile idle
      double wx = 0, wy = 0;
      for(int bq = 0; bq < P.nbox[3]; bq++) wx += warps_with_work(ext[3], bq * PBOX);
      for(int bq = 0; bq < P.nbox[4]; bq++) wy += warps_with_work(ext[4], bq * PBOX);
      units_x = 16.0 * (double) P.nboxes / P.nbox[3] * wx;
      units_y = 16.0 * (double) P.nboxes / P.nbox[4] * wy;
    }
    b.eval_fraction = (P.sym ? (double) nboxes : (double) P.nboxes) / (double) P.nboxes;
    // one box keeps the tensor pipe of an SM busy for (k-steps x 16 DMMA x 16 cycles) / 4 sub-partitions x
    // 4 warps = k-steps x 256 cycles; co-resident CTAs start that far apart (options.stagger, default on)
    int64_t ksteps = 0;
    for(int i = 0; i < P.nterms; i++) ksteps += (int64_t) (P.term[i].kslabs - 1) * 4 + P.term[i].ksteps_last;
    P.ctas_per_sm    = ctas;
    P.stagger_cycles = (ctas > 1 && ctx->opt.stagger && nboxes >= 4 * in_flight)
                         ? (int) std::min<int64_t>(ksteps * 256 + 8192, 50000000)
                         : 0;
    need_partial = P.box_list ? std::max<int64_t>(P.nlist, 1) : padded; // one partial per evaluated box
    // DMMA flops issued (512 per DMMA.8x8x4): per k-step a warp with work issues 16 DMMAs, 8 in the boundary layer of an odd
    // hole tile; warps of a ragged last tile-particle box without a valid slot issue none
    int64_t ksteps_x = 0;
    for(int i = 0; i < P.nterms_x; i++) ksteps_x += (int64_t) (P.term[i].kslabs - 1) * 4 + P.term[i].ksteps_last;
    const double sub_groups = (double) (P.sub[0] * P.sub[1] * P.sub[2]);
    executed = 512.0 * sub_groups * (units_x * (double) ksteps_x + units_y * (double) (ksteps - ksteps_x));
  }
  b.nparts = need_partial;
  if(need_partial > b.partial_cap) {
    CK(cudaStreamSynchronize(b.cs));
    if(b.d_partial) CK(cudaFree(b.d_partial));
    b.partial_cap = need_partial + need_partial / 4 + 64;
    CK(cudaMalloc(&b.d_partial, (size_t) b.partial_cap * 16));
  }
  P.partial = b.d_partial;
  ctx->stats.executed_flops += executed;

  // ---- launch the panel build on the staging stream ----
  const int nd = (int) descs.size();
  if(nd > b.desc_cap) {
    // the buffer is idle here: its previous task has finished (run_task_list waited for b.done)
    if(b.d_descs) CK(cudaFree(b.d_descs));
    if(b.h_descs) CK(cudaFreeHost(b.h_descs));
    b.desc_cap = nd + nd / 2 + 256;
    CK(cudaMalloc(&b.d_descs, sizeof(GatherDesc) * b.desc_cap));
    CK(cudaMallocHost(&b.h_descs, sizeof(GatherDesc) * b.desc_cap));
  }
  memcpy(b.h_descs, descs.data(), sizeof(GatherDesc) * nd);
  static_assert(sizeof(GatherDesc) % 4 == 0, "descriptor copy works in 32-bit words");
  if(nd) CK(launch_copy_from_pinned(b.h_descs, b.d_descs, sizeof(GatherDesc) * nd, ctx->s_stage));
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
  // blocks fetched for this task (and any prefetched earlier) are still on their way: one event, no host wait
  const bool fenced = ctx->fetch_dirty || !ctx->sym_check_pending.empty();
  if(int rc2 = fetch_fence(ctx, ctx->s_stage)) return rc2;
  if(fenced)
    if(int rc2 = flush_block_symmetry_checks(ctx)) return rc2; // behind the fence: the blocks have landed
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
  const double w0 = now_s();
  CK(cudaEventSynchronize(b.done));
  ctx->stats.seconds_host_wait += now_s() - w0;
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
  if(b.trace_slot >= 0 && b.trace_slot < (int64_t) ctx->trace.size()) {
    TraceRow& r = ctx->trace[b.trace_slot];
    float     g0 = 0, g1 = 0;
    CK(cudaEventElapsedTime(&g0, ctx->ev_base, b.g0));
    CK(cudaEventElapsedTime(&g1, ctx->ev_base, b.g1));
    r.gather0 = g0, r.gather1 = g1, r.k0 = ms0, r.k1 = ms1;
  }
  b.timing_pending = false;
  return 0;
}

// the storage blocks task t reads that are not resident yet, in staging order
void missing_blocks(ccsdt_ctx* ctx, const Task& t, std::vector<SrcPiece>& out, std::vector<SrcPiece>& scratch) {
  walk_task_sources(
    ctx->sp, ctx->buf[0].geom, ctx->buf[0].s1_a, ctx->buf[0].s1_b, t, nullptr,
    [&](double*, const int64_t*, const SrcSpec& s) {
      if(ctx->dense[s.tensor]) return 0;
      scratch.clear();
      source_pieces(ctx, s, scratch);
      for(const SrcPiece& p: scratch)
        if(!ctx->blocks.count(p.key)) out.push_back(p);
      return 0;
    },
    [](double*, const int64_t*, const int*) { return 0; });
}

// Execution order of a rank's tasks when the operands come through the fetch callback.  The host thread and the GPU
// form a two-stage flow shop: stage 1 pulls a task's missing blocks through the callback (the caller's Tensor::get, on
// this thread), stage 2 is the GPU.  Units are the particle-tile triples (p4b,p5b,p6b): the tasks of one triple share
// their v2iabc and most of their T2 blocks, so they run back to back and those blocks are fetched (and stay resident)
// once per triple.  Units run in ascending order of the bytes they need, in octaves (units of similar size keep their
// canonical order, hence the block reuse between neighbouring triples): the GPU starts after a SMALL fetch, and the
// data of the big units arrive under the kernels of the smaller ones before them (measured on the benzene shape: the
// 180 MB of the first mixed-spin task cost 32 ms of idle GPU at the head of a 215 ms job when it came first, 9 ms with
// this order).  Per-task energies are summed in canonical order, so the result does not depend on the order.
void order_for_fetch(ccsdt_ctx* ctx, std::vector<int64_t>& order) {
  struct Unit {
    std::array<int32_t, 3> key;
    std::vector<int64_t>   tasks;
    double                 bytes = 0, flops = 0;
  };
  std::map<std::array<int32_t, 3>, size_t> index;
  std::vector<Unit>                        units;
  std::vector<SrcPiece>                    pieces, scratch;
  for(int64_t id: order) {
    const Task&                  t = ctx->tasks[id];
    const std::array<int32_t, 3> key{t.t[3], t.t[4], t.t[5]};
    auto                         it = index.find(key);
    if(it == index.end()) {
      it = index.emplace(key, units.size()).first;
      units.push_back(Unit{key, {}, 0, 0});
    }
    units[it->second].tasks.push_back(id);
    units[it->second].flops += (double) task_cost(ctx->sp, t, ctx->opt.symmetry != 0);
  }
  for(Unit& u: units) {
    std::map<BlockKey, size_t> seen;
    for(int64_t id: u.tasks) {
      pieces.clear();
      missing_blocks(ctx, ctx->tasks[id], pieces, scratch);
      for(const SrcPiece& p: pieces) seen[p.key] = p.elems * 8;
    }
    for(auto& kv: seen) u.bytes += (double) kv.second;
  }
  auto octave = [](double bytes) {
    int o = 0;
    for(double x = bytes; x >= 2.0; x *= 0.5) o++;
    return o;
  };
  std::stable_sort(units.begin(), units.end(), [&](const Unit& x, const Unit& y) { return octave(x.bytes) < octave(y.bytes); });
  order.clear();
  for(const Unit& u: units) order.insert(order.end(), u.tasks.begin(), u.tasks.end());
}

} // namespace

static int run_task_list(ccsdt_ctx* ctx, const std::vector<int64_t>& ids, bool use_global_split, double energies[2],
                         double* per_task, ccsdt_stats* stats_out) {
  cudaSetDevice(ctx->device);
  const double t0 = now_s();
  ctx->stats           = ccsdt_stats{};
  ctx->trace.clear();
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
  const bool fetching = ctx->fetch && !ctx->synthetic;
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
    const double m0 = now_s();
    if(int rc = ensure_pools(ctx)) return rc;
    const double m1 = now_s();
    if(int rc = update_block_budget(ctx)) return rc;
    const double m2 = now_s();
    if(!ctx->task_counter && fetching) order_for_fetch(ctx, order);
    if(ctx->opt.verbose >= 2)
      fprintf(stderr, "[ccsdt trace] before the loop: hand-out %.3f ms, pools %.3f ms, budget %.3f ms, order %.3f ms\n", (m0 - t0) * 1e3,
              (m1 - m0) * 1e3, (m2 - m1) * 1e3, (now_s() - m2) * 1e3);
    const int64_t cap = (int64_t) order.size();
    if(cap > ctx->task_energy_cap) {
      if(ctx->d_task_energy) CK(cudaFree(ctx->d_task_energy));
      ctx->task_energy_cap = cap + 16;
      CK(cudaMalloc(&ctx->d_task_energy, (size_t) ctx->task_energy_cap * 16));
    }
    const int nbuf = ctx->opt.overlap ? 2 : 1;
    ctx->buf[0].cs = ctx->s_compute;
    ctx->buf[1].cs = ctx->opt.overlap ? ctx->s_compute2 : ctx->s_compute;
    {
      // error word + the pipeline watchdog (nanoseconds of %globaltimer; 0 = never)
      const int64_t  ms      = ctx->opt.watchdog_ms < 0 ? 0 : (ctx->opt.watchdog_ms > 0 ? ctx->opt.watchdog_ms : 20000);
      const uint64_t ns      = (uint64_t) ms * 1000000ull;
      const uint32_t init[4] = {0u, (uint32_t) (ns & 0xffffffffu), (uint32_t) (ns >> 32), 0u};
      CK(cudaMemcpyAsync(ctx->d_error, init, 16, cudaMemcpyHostToDevice, ctx->s_compute));
    }
    CK(cudaEventRecord(ctx->ev_base, ctx->s_compute));
    CK(cudaStreamWaitEvent(ctx->s_compute2, ctx->ev_base, 0));
    ctx->kernel_busy_until = 0.0;
    if(int rc = check_dense_symmetry(ctx)) return rc;

    // Prefetch (static hand-out, blocks through the callback): while the host would only wait for the GPU to
    // release a staging buffer, it pulls the blocks of the tasks to come into HBM -- the callback (the caller's
    // Tensor::get) runs on this thread, its copies travel on s_fetch.  It runs as far ahead as the block budget
    // allows (options.prefetch_tasks bounds it in tasks): order_for_fetch starts with the units that need the fewest
    // bytes, and the blocks of the bigger ones arrive under the kernels before them.  Blocks of future tasks carry
    // their future clock, so the LRU does not take them back before they are used.
    const int64_t lookahead = (!ctx->task_counter && fetching && ctx->opt.prefetch_tasks >= 0)
                                ? (ctx->opt.prefetch_tasks > 0 ? ctx->opt.prefetch_tasks : ((int64_t) 1 << 40)) : 0;
    // Dynamic hand-out: the next task is not known before it is claimed, so a rank claims ONE task ahead and fetches
    // that task's blocks while it waits for the GPU (the tail of the list is the cheap tasks: holding one of them does
    // not unbalance the ranks).
    const bool claim_ahead = ctx->task_counter && fetching && ctx->opt.prefetch_tasks >= 0;
    int64_t               pf_task = 0;      // next entry of `order` whose blocks have not been listed
    std::vector<SrcPiece> pf_queue, scratch;
    size_t                pf_pos   = 0;
    int64_t               pf_clock = 0;     // use_clock the tasks of pf_queue will have
    const int64_t         clock0   = ctx->use_clock;
    std::vector<GatherDesc> descs;
    auto prefetch_one = [&](int64_t j) -> int { // returns 1 when there is nothing (more) to prefetch
      while(pf_pos >= pf_queue.size()) {
        pf_task = std::max(pf_task, j);
        if(pf_task >= (int64_t) order.size() || pf_task > j + lookahead) return 1;
        pf_queue.clear();
        pf_pos   = 0;
        pf_clock = clock0 + pf_task + 1;
        missing_blocks(ctx, ctx->tasks[order[pf_task]], pf_queue, scratch);
        pf_task++;
      }
      const SrcPiece& p = pf_queue[pf_pos++];
      if(ctx->blocks.count(p.key)) return 0;
      // a prefetch never pushes out blocks of the tasks in flight or of tasks still to come (the LRU only gives up
      // blocks whose last reader has passed); when that does not make room it pauses until tasks have retired
      if(ctx->block_budget && ctx->block_bytes + p.elems * 8 > ctx->block_budget) {
        if(int rc = evict_stale(ctx, p.elems * 8)) return -rc;
        if(ctx->block_bytes + p.elems * 8 > ctx->block_budget) {
          pf_pos--;
          return 1;
        }
      }
      BlockRef ref;
      if(int rc = resolve_block(ctx, p.key, p.elems, pf_clock, ref)) return -rc;
      return 0;
    };

    int64_t               claimed = claim_ahead ? next_task() : -2; // -2: not used
    std::vector<SrcPiece> ahead;                                      // missing blocks of the task claimed ahead
    size_t                ahead_pos = 0;
    for(int64_t j = 0;; j++) {
      const int64_t ti = claim_ahead ? claimed : next_task();
      if(ti < 0) break;
      if(claim_ahead) {
        claimed = next_task();
        ahead.clear();
        ahead_pos = 0;
        if(claimed >= 0 && j > 0) missing_blocks(ctx, ctx->tasks[claimed], ahead, scratch);
      }
      mine.push_back(ti);
      StageBuf& b = ctx->buf[j % nbuf];
      // the buffer's previous task must have finished computing before its panels are rebuilt
      if(lookahead > 0)
        while(b.timing_pending && cudaEventQuery(b.done) == cudaErrorNotReady) {
          const int r = prefetch_one(j);
          if(r < 0) return -r;
          if(r > 0) break;
        }
      if(b.timing_pending) {
        if(int rc = collect_timing(ctx, b)) return rc;
      }
      CK(cudaStreamWaitEvent(ctx->s_stage, b.done, 0));
      const double tr0 = now_s() - t0, tf0 = ctx->stats.seconds_fetch;
      if(int rc = stage_task(ctx, b, ctx->tasks[ti], descs, scratch)) return rc;
      if(int rc = launch_task(ctx, b, j)) return rc;
      b.trace_slot = -1;
      if(ctx->opt.verbose >= 2) {
        b.trace_slot = (int64_t) ctx->trace.size();
        ctx->trace.push_back(TraceRow{ti, tr0, now_s() - t0, ctx->stats.seconds_fetch - tf0, 0, 0, 0, 0});
      }
      if(claim_ahead && claimed >= 0) {
        // task j is queued: pull the blocks of the task claimed ahead while the GPU works on j - 1 and j (the first
        // time round its list is made only now, so that the very first task starts as early as possible)
        if(j == 0) missing_blocks(ctx, ctx->tasks[claimed], ahead, scratch);
        StageBuf& nb = ctx->buf[(j + 1) % nbuf];
        while(ahead_pos < ahead.size() && nb.timing_pending && cudaEventQuery(nb.done) == cudaErrorNotReady) {
          const SrcPiece& p = ahead[ahead_pos++];
          if(ctx->blocks.count(p.key)) continue;
          if(ctx->block_budget && ctx->block_bytes + p.elems * 8 > ctx->block_budget) break;
          BlockRef ref;
          if(int rc = resolve_block(ctx, p.key, p.elems, ctx->use_clock + 1, ref)) return rc;
        }
      }
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
    uint32_t flag = 0;
    CK(cudaMemcpy(&flag, ctx->d_error, 4, cudaMemcpyDeviceToHost));
    if(flag) return ctx->fail("device error flag " + std::to_string(flag), 9);
    if(ctx->opt.symmetry && ctx->opt.check_symmetry >= 0) {
      CK(cudaStreamSynchronize(ctx->s_fetch));
      if(int rc = flush_block_symmetry_checks(ctx)) return rc; // blocks prefetched for tasks this run did not reach
      CK(cudaStreamSynchronize(ctx->s_stage));
      CK(cudaMemcpy(&flag, ctx->d_symflag, 4, cudaMemcpyDeviceToHost));
      // (the verdict stays until ccsdt_set_space drops the operands: a block that failed is not looked at again)
      if(flag)
        return ctx->fail("an operand is not antisymmetric (T2 in (a,b) / (i,j), v2ijab in (i,j) / (a,b), v2ijka in (i,j), v2iabc in (b,c)): "
                         "options.symmetry = 1 relies on it -- set symmetry = 0 to evaluate every element as the reference does", 12);
    }
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
  // the host buffers of asynchronous uploads may be released once the run returns -- also when this rank had no task
  CK(cudaStreamSynchronize(ctx->s_copy_a));
  CK(cudaStreamSynchronize(ctx->s_copy_b));
  CK(cudaStreamSynchronize(ctx->s_fetch));
  share_poll(ctx, true); // every block this rank fetched for the node is published, every peer copy accounted for
  for(int tn = 0; tn < 5; tn++) ctx->upload_pending[tn] = false;
  ctx->stats.seconds_total = now_s() - t0;
  if(stats_out) *stats_out = ctx->stats;
  if(ctx->opt.verbose >= 2) {
    fprintf(stderr, "[ccsdt trace] run of %zu tasks, %.3f ms total; per task: id | host stage begin..end ms (fetch ms) | gather ms | kernel k0..k1 ms\n",
            ctx->trace.size(), ctx->stats.seconds_total * 1e3);
    for(const TraceRow& r: ctx->trace) {
      const Task& t = ctx->tasks[r.task];
      fprintf(stderr, "[ccsdt trace] %5lld (%d,%d,%d|%d,%d,%d) | %8.3f..%8.3f (%7.3f) | %8.3f..%8.3f | %8.3f..%8.3f\n", (long long) r.task,
              t.t[0], t.t[1], t.t[2], t.t[3], t.t[4], t.t[5], r.host_begin * 1e3, r.host_staged * 1e3, r.fetch_s * 1e3, r.gather0, r.gather1,
              r.k0, r.k1);
    }
  }
  return 0;
}

// the execution tiling, the task list and the split follow the storage space and the options
static int rebuild_exec(ccsdt_ctx* ctx) {
  if(!ctx->have_space) return 0;
  ctx->sp      = ctx->opt.exec_tilesize ? make_exec_space(ctx->store, ctx->opt.exec_tilesize) : ctx->store;
  ctx->retiled = !same_tiling(ctx->sp, ctx->store);
  const std::string v = ctx->sp.validate();
  if(!v.empty()) return ctx->fail("invalid execution tiling: " + v);
  ctx->tasks       = enumerate_tasks(ctx->sp.noab(), ctx->sp.nvab(), ctx->sp.k_spin.data(), ctx->sp.restricted, &ctx->n_outer);
  ctx->owner       = partition_tasks(ctx->sp, ctx->tasks, ctx->opt.nranks, ctx->opt.symmetry != 0);
  ctx->pools_ready = false; // geometry follows the tiling and the CTA box
  return 0;
}

static void sync_streams(ccsdt_ctx* ctx) {
  for(cudaStream_t s: {ctx->s_compute, ctx->s_compute2, ctx->s_stage, ctx->s_copy_a, ctx->s_copy_b, ctx->s_fetch})
    if(s) cudaStreamSynchronize(s);
}

static void real_destroy(ccsdt_ctx* ctx) {
  cudaSetDevice(ctx->device);
  sync_streams(ctx);
  share_detach(ctx);
  comm_destroy(ctx);
  free_pools(ctx);
  store_destroy(ctx);
  for(auto& b: ctx->buf)
    for(cudaEvent_t ev: {b.staged, b.done, b.k0, b.k1, b.g0, b.g1})
      if(ev) cudaEventDestroy(ev);
  if(ctx->d_evl) cudaFree(ctx->d_evl);
  if(ctx->d_task_energy) cudaFree(ctx->d_task_energy);
  if(ctx->d_error) cudaFree(ctx->d_error);
  if(ctx->d_symflag) cudaFree(ctx->d_symflag);
  for(auto& kv: ctx->box_lists)
    if(kv.second.dev) cudaFree(kv.second.dev);
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
  {
    std::lock_guard<std::mutex> lock(g_park_mutex);
    auto                        it = g_parked.find(device);
    if(it != g_parked.end()) {
      *out = it->second;
      g_parked.erase(it);
      cudaSetDevice(device);
      return 0;
    }
  }
  auto* ctx   = new ccsdt_ctx;
  ctx->device = device;
  ccsdt_default_options(&ctx->opt);
  auto bail = [&](const std::string& m) {
    g_create_error = m;
    real_destroy(ctx);
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
  if((e = cudaMalloc(&ctx->d_error, 16)) != cudaSuccess) return bail(cudaGetErrorString(e));
  cudaMemset(ctx->d_error, 0, 16);
  if((e = cudaMalloc(&ctx->d_symflag, 4)) != cudaSuccess) return bail(cudaGetErrorString(e));
  cudaMemset(ctx->d_symflag, 0, 4);
  if((e = fused_dmma_configure((size_t) ctx->prop.sharedMemPerBlockOptin - 2048)) != cudaSuccess)
    return bail(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
  if(store_create(ctx)) return bail(ctx->err);
  *out = ctx;
  return 0;
}

int ccsdt_destroy(ccsdt_ctx* ctx) {
  if(!ctx) return 0;
  cudaSetDevice(ctx->device);
  sync_streams(ctx);
  const char* env   = std::getenv("CCSDT_B200_CACHE");
  const bool  cache = !(env && std::atoi(env) == 0);
  if(cache) {
    // park: the logical state goes, the device resources (streams, events, panel pools, pinned ring, memory pool,
    // box lists) stay for the next ccsdt_create on this device
    share_detach(ctx);
    comm_destroy(ctx);
    free_operands(ctx);
    ccsdt_default_options(&ctx->opt);
    ctx->store = ctx->sp = Space{};
    ctx->retiled = ctx->have_space = false;
    ctx->tasks.clear();
    ctx->owner.clear();
    ctx->fetch        = nullptr;
    ctx->fetch_user   = nullptr;
    ctx->synthetic    = false;
    ctx->task_counter = nullptr;
    ctx->pending_h2d  = 0;
    ctx->pools_ready  = false;
    ctx->err.clear();
    std::lock_guard<std::mutex> lock(g_park_mutex);
    if(!g_parked.count(ctx->device)) {
      g_parked[ctx->device] = ctx;
      return 0;
    }
  }
  real_destroy(ctx);
  return 0;
}

int ccsdt_release_cached(void) {
  std::lock_guard<std::mutex> lock(g_park_mutex);
  for(auto& kv: g_parked) real_destroy(kv.second);
  g_parked.clear();
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
  if(n.exec_tilesize < -1) return ctx->fail("exec_tilesize must be -1 (auto), 0 (the caller's tiles) or a tile extent");
  if(n.block_budget_bytes < 0) return ctx->fail("block_budget_bytes must be >= 0");
  cudaSetDevice(ctx->device);
  sync_streams(ctx);
  ctx->opt = n;
  return rebuild_exec(ctx);
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

static int export_tiles(const Space& sp, int64_t* k_range, int32_t* k_spin, int32_t counts[4], int cap) {
  const int n = sp.noab() + sp.nvab();
  if(counts) counts[0] = sp.noa, counts[1] = sp.nob, counts[2] = sp.nva, counts[3] = sp.nvb;
  if(n > cap) return -n;
  for(int t = 0; t < n; t++) {
    if(k_range) k_range[t] = sp.k_range[t];
    if(k_spin) k_spin[t] = sp.k_spin[t];
  }
  return n;
}

int ccsdt_exec_tiles(const ccsdt_ctx* ctx, int64_t* k_range, int32_t* k_spin, int32_t counts[4], int cap) {
  if(!ctx || !ctx->have_space) return 0;
  return export_tiles(ctx->sp, k_range, k_spin, counts, cap);
}

int64_t ccsdt_num_tasks(const ccsdt_ctx* ctx) { return ctx ? (int64_t) ctx->tasks.size() : 0; }

int ccsdt_make_exec_tiles(int noa, int nob, int nva, int nvb, const int64_t* k_range, const int32_t* k_spin, int target,
                          int64_t* out_range, int32_t* out_spin, int32_t counts[4], int cap) {
  if(!k_range || !k_spin) return 0;
  const Space store = make_space(noa, nob, nva, nvb, k_range, k_spin, nullptr, true);
  return export_tiles(target ? make_exec_space(store, target) : store, out_range, out_spin, counts, cap);
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

int ccsdt_estimate_memory(int noa, int nob, int nva, int nvb, const int64_t* k_range, const int32_t* k_spin, int target,
                          ccsdt_memory_estimate* out) {
  if(!k_range || !k_spin || !out) return 1;
  const Space store = make_space(noa, nob, nva, nvb, k_range, k_spin, nullptr, true);
  if(!store.validate().empty()) return 1;
  const Space          ex     = target ? make_exec_space(store, target) : store;
  const int            sub[3] = {1, 1, 1};
  const MemoryEstimate e      = estimate_memory(ex, sub, 2);
  out->exec_max_hole_tile     = ex.max_hole_tile();
  out->exec_max_particle_tile = ex.max_particle_tile();
  out->panel_bytes            = e.panel_bytes;
  out->s1_bytes               = e.s1_bytes;
  out->task_block_bytes       = e.task_block_bytes;
  for(int t = 0; t < 5; t++) out->tensor_bytes[t] = e.tensor_bytes[t];
  out->minimum_bytes = e.minimum_bytes;
  return 0;
}

int ccsdt_set_space(ccsdt_ctx* ctx, int noa, int nob, int nva, int nvb, const int64_t* k_range, const int32_t* k_spin,
                    const double* evl, int is_restricted) {
  if(!ctx || !k_range || !k_spin || !evl) return 1;
  Space       sp = make_space(noa, nob, nva, nvb, k_range, k_spin, evl, is_restricted != 0);
  std::string v  = sp.validate();
  if(!v.empty()) return ctx->fail("invalid space: " + v);
  cudaSetDevice(ctx->device);
  sync_streams(ctx);
  free_operands(ctx);
  CK(cudaMemset(ctx->d_symflag, 0, 4)); // the operands of the previous space are gone, and their verdict with them
  ctx->store      = sp;
  ctx->have_space = true;
  if(int rc = rebuild_exec(ctx)) return rc;
  const size_t bytes = sp.evl.size() * 8;
  if(!ctx->d_evl || ctx->evl_cap < bytes) {
    if(ctx->d_evl) CK(cudaFree(ctx->d_evl));
    ctx->d_evl = nullptr;
    CK(cudaMalloc(&ctx->d_evl, bytes));
    ctx->evl_cap = bytes;
  }
  CK(cudaMemcpy(ctx->d_evl, sp.evl.data(), bytes, cudaMemcpyHostToDevice));
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
