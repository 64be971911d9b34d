// Operand store of one context: where T1, T2 and the three V2 tensors live in HBM and how they get there.
//
//   dense tensors   ccsdt_put_dense[_async]: the spin-conserving blocks of a dense host array, strided 3-d copies
//   block store     individual STORAGE blocks (the caller's tiling, what Tensor<T>::get delivers) keyed by block id:
//                   ccsdt_put_block (resident for good) or the fetch callback (on demand).  A fetched block is written
//                   by the callback straight into a pinned ring, copied on its own stream (s_fetch) and the host moves
//                   on: no synchronisation per block -- the panel build of the task that needs it waits for one event
//                   (fetch_fence).  Device memory comes from a private cudaMemPool (the application's default pool is
//                   not touched); least-recently-used blocks that neither the task being staged nor the one in flight
//                   reads are evicted against a budget derived from the memory that is actually free, and again when
//                   an allocation fails.
//   procedural      ccsdt_set_synthetic: nothing is stored, the panel build evaluates the generator
//
// Replaces Tensor<T>::get + the six host-side LRUCache arguments of the reference
// (exachem/cc/ccsd_t/ccsd_t_all_fused_singles.hpp:200,304; ..._doubles1.hpp:208-247,282; ..._doubles2.hpp:215-239,335-343;
//  exachem/cc/ccsd_t/ccsd_t.cpp:236-241) and the overlap contract of ccsd_t_all_fused.hpp:141,180.
#include "ccsdt_ctx.hpp"

#include <algorithm>
#include <chrono>
#include <cstring>

using namespace ccsdt;

namespace ccsdt {

const char* const kKinds[5] = {"vo", "vvoo", "oovv", "ooov", "ovvv"};

int64_t dim_full(const Space& sp, char kind) { return kind == 'o' ? sp.n_occ() : sp.n_virt(); }
int     tile_of(const Space& sp, char kind, uint32_t bid) { return kind == 'o' ? (int) bid : (int) bid + sp.noab(); }
int64_t tile_start(const Space& sp, int tile) {
  return tile < sp.noab() ? sp.k_offset[tile] : sp.k_offset[tile] - sp.k_offset[sp.noab()];
}
size_t block_elems(const Space& sp, int tensor, const uint32_t bid[4]) {
  const char* kinds = kKinds[tensor];
  size_t      n     = 1;
  for(int d = 0; kinds[d]; d++) n *= (size_t) sp.k_range[tile_of(sp, kinds[d], bid[d])];
  return n;
}

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ------------------------------------------------------------------------------------------------
int store_create(ccsdt_ctx* ctx) {
  CK(cudaStreamCreateWithFlags(&ctx->s_fetch, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&ctx->ev_fetched, cudaEventDisableTiming));
  // fetched blocks come from a PRIVATE stream-ordered pool that keeps its memory across synchronisations; the
  // process-wide default pool (which the host application may use with its own settings) is left alone
  cudaMemPoolProps props{};
  props.allocType     = cudaMemAllocationTypePinned;
  props.handleTypes   = cudaMemHandleTypeNone;
  props.location.type = cudaMemLocationTypeDevice;
  props.location.id   = ctx->device;
  CK(cudaMemPoolCreate(&ctx->block_pool, &props));
  uint64_t keep = ~0ull;
  CK(cudaMemPoolSetAttribute(ctx->block_pool, cudaMemPoolAttrReleaseThreshold, &keep));
  return 0;
}

static void ring_drain(ccsdt_ctx* ctx) {
  while(!ctx->ring_inflight.empty()) {
    cudaEventSynchronize(ctx->ring_inflight.front().ev);
    ctx->ring_events.push_back(ctx->ring_inflight.front().ev);
    ctx->ring_inflight.pop_front();
  }
  ctx->ring_head = 0;
}

void store_destroy(ccsdt_ctx* ctx) {
  free_operands(ctx);
  ring_drain(ctx);
  for(cudaEvent_t e: ctx->ring_events) cudaEventDestroy(e);
  ctx->ring_events.clear();
  if(ctx->ring) cudaFreeHost(ctx->ring);
  ctx->ring = nullptr, ctx->ring_cap = 0;
  if(ctx->s_fetch) cudaStreamSynchronize(ctx->s_fetch);
  if(ctx->block_pool) cudaMemPoolDestroy(ctx->block_pool);
  ctx->block_pool = nullptr;
  if(ctx->ev_fetched) cudaEventDestroy(ctx->ev_fetched);
  if(ctx->s_fetch) cudaStreamDestroy(ctx->s_fetch);
  ctx->ev_fetched = nullptr, ctx->s_fetch = nullptr;
}

// false: the block lives in the node-shared store and a peer is still copying it (only with wait = false)
static bool release_block(ccsdt_ctx* ctx, const BlockKey& key, BlockEntry& e, bool wait) {
  // the last reader of a block is a panel build on s_stage: the release is ordered behind it
  if(e.slab >= 0) {
    if(!share_release(ctx, key, e, wait)) return false;
    // the region may be handed out again at once: the next copy into it (on s_fetch) waits for the panel builds queued so far
    cudaEventRecord(ctx->ev_fetched, ctx->s_stage);
    cudaStreamWaitEvent(ctx->s_fetch, ctx->ev_fetched, 0);
  }
  else if(e.pooled) cudaFreeAsync(e.dev, ctx->s_stage);
  else cudaFree(e.dev);
  ctx->block_bytes -= e.bytes;
  return true;
}

int clear_blocks(ccsdt_ctx* ctx, bool keep_pinned) {
  ring_drain(ctx); // copies into blocks about to be released must have landed
  for(auto it = ctx->blocks.begin(); it != ctx->blocks.end();) {
    if(keep_pinned && it->second.pinned) {
      ++it;
      continue;
    }
    release_block(ctx, it->first, it->second, true);
    it = ctx->blocks.erase(it);
  }
  ctx->fetch_dirty = false;
  ctx->sym_check_pending.clear();
  return 0;
}

void free_operands(ccsdt_ctx* ctx) {
  for(int t = 0; t < 5; t++) {
    if(ctx->dense[t]) cudaFree(ctx->dense[t]);
    ctx->dense[t]           = nullptr;
    ctx->dense_elems[t]     = 0;
    ctx->dense_sparse_ok[t] = false;
    ctx->upload_pending[t]  = false;
  }
  clear_blocks(ctx, false);
  share_poll(ctx, true);
  ctx->block_bytes = 0;
}

int update_block_budget(ccsdt_ctx* ctx) {
  if(ctx->opt.block_budget_bytes > 0) {
    ctx->block_budget = (size_t) ctx->opt.block_budget_bytes;
    return 0;
  }
  // what is free now + what our pool holds without using it + what the resident blocks occupy, minus a reserve for
  // the application and for the partial / box-list buffers still to come
  size_t free_b = 0, total_b = 0;
  CK(cudaMemGetInfo(&free_b, &total_b));
  uint64_t reserved = 0, used = 0;
  if(ctx->block_pool) {
    cudaMemPoolGetAttribute(ctx->block_pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
    cudaMemPoolGetAttribute(ctx->block_pool, cudaMemPoolAttrUsedMemCurrent, &used);
  }
  const size_t reserve = std::max<size_t>((size_t) 2 << 30, total_b / 32);
  const size_t avail   = free_b + (size_t) (reserved - used) + share_unused_bytes(ctx) + ctx->block_bytes;
  ctx->block_budget    = avail > reserve ? avail - reserve : avail / 2;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// pinned fetch ring: a FIFO of regions, each released by the event of its host-to-device copy
static int ring_wait_oldest(ccsdt_ctx* ctx) {
  const double t0 = now_s();
  CK(cudaEventSynchronize(ctx->ring_inflight.front().ev));
  ctx->stats.seconds_host_wait += now_s() - t0;
  ctx->ring_events.push_back(ctx->ring_inflight.front().ev);
  ctx->ring_inflight.pop_front();
  return 0;
}

static int ring_alloc(ccsdt_ctx* ctx, size_t bytes, uint8_t** out, size_t* begin) {
  bytes = (bytes + 255) & ~(size_t) 255;
  if(bytes * 2 > ctx->ring_cap) {
    // (re)size: at least two of the largest block, 64 MB to start with
    while(!ctx->ring_inflight.empty())
      if(int rc = ring_wait_oldest(ctx)) return rc;
    if(ctx->ring) CK(cudaFreeHost(ctx->ring));
    ctx->ring     = nullptr;
    size_t cap    = (size_t) 64 << 20;
    while(cap < bytes * 4) cap *= 2;
    CK(cudaHostAlloc((void**) &ctx->ring, cap, cudaHostAllocDefault));
    ctx->ring_cap  = cap;
    ctx->ring_head = 0;
  }
  for(;;) {
    // retire finished copies without blocking
    while(!ctx->ring_inflight.empty() && cudaEventQuery(ctx->ring_inflight.front().ev) == cudaSuccess) {
      ctx->ring_events.push_back(ctx->ring_inflight.front().ev);
      ctx->ring_inflight.pop_front();
    }
    if(ctx->ring_inflight.empty()) ctx->ring_head = 0;
    const size_t head = ctx->ring_head;
    const size_t tail = ctx->ring_inflight.empty() ? head : ctx->ring_inflight.front().begin;
    size_t       at   = (size_t) -1;
    if(ctx->ring_inflight.empty() || head > tail) { // live data in [tail, head): free is [head, cap) and [0, tail)
      if(head + bytes <= ctx->ring_cap) at = head;
      else if(bytes < tail) at = 0;
    }
    else if(head + bytes < tail) at = head;         // wrapped: free is [head, tail)
    if(at != (size_t) -1) {
      *out           = ctx->ring + at;
      *begin         = at;
      ctx->ring_head = at + bytes;
      return 0;
    }
    if(int rc = ring_wait_oldest(ctx)) return rc;
  }
}

// ------------------------------------------------------------------------------------------------
// evicts least-recently-used blocks until `bytes` more fit the budget (or, with force, at least one block goes)
static int evict(ccsdt_ctx* ctx, size_t bytes, bool force, bool* evicted_any) {
  if(evicted_any) *evicted_any = false;
  std::vector<const BlockKey*> busy; // shared blocks a peer is copying right now: not this time
  while(force || (ctx->block_budget && ctx->block_bytes + bytes > ctx->block_budget)) {
    auto victim = ctx->blocks.end();
    for(auto jt = ctx->blocks.begin(); jt != ctx->blocks.end(); ++jt) {
      const BlockEntry& e = jt->second;
      // the task being staged (use_clock) and the one in flight (use_clock - 1) keep their blocks; prefetched blocks
      // carry a future clock
      if(e.pinned || e.last_use + 2 > ctx->use_clock) continue;
      if(std::find(busy.begin(), busy.end(), &jt->first) != busy.end()) continue;
      if(victim == ctx->blocks.end() || e.last_use < victim->second.last_use) victim = jt;
    }
    if(victim == ctx->blocks.end()) break;
    if(!release_block(ctx, victim->first, victim->second, false)) {
      busy.push_back(&victim->first);
      continue;
    }
    ctx->blocks.erase(victim);
    ctx->stats.blocks_evicted++;
    if(evicted_any) *evicted_any = true;
    force = false;
  }
  return 0;
}

int evict_stale(ccsdt_ctx* ctx, size_t bytes) { return evict(ctx, bytes, false, nullptr); }

int resolve_block(ccsdt_ctx* ctx, const BlockKey& key, size_t elems, int64_t for_clock, BlockRef& out) {
  auto it = ctx->blocks.find(key);
  if(it == ctx->blocks.end()) {
    if(!ctx->fetch)
      return ctx->fail("block {" + std::to_string(key.b[0]) + "," + std::to_string(key.b[1]) + "," + std::to_string(key.b[2]) +
                       "," + std::to_string(key.b[3]) + "} of tensor " + std::to_string(key.tensor) +
                       " was never supplied (no dense tensor, block or fetch callback)", 4);
    const size_t bytes = elems * 8;
    if(int rc = evict(ctx, bytes, false, nullptr)) return rc;
    double* dev = nullptr;
    int     slab = -1;
    size_t  slab_off = 0;
    void*   dir_entry = nullptr; // node-shared store: the directory entry to publish once the upload has landed
    bool    from_peer = false;
    for(int attempt = 0;; attempt++) {
      cudaError_t e = cudaSuccess;
      if(ctx->share) {
        const int rc = share_acquire(ctx, key, bytes, &dev, &slab, &slab_off, &dir_entry);
        if(rc > 1) return rc;
        from_peer = rc == 1;
        if(rc >= 0) break;
        e = cudaErrorMemoryAllocation;
      }
      else {
        e = cudaMallocFromPoolAsync((void**) &dev, bytes, ctx->block_pool, ctx->s_fetch);
        if(e == cudaSuccess) break;
        cudaGetLastError();
      }
      if(e != cudaErrorMemoryAllocation || attempt >= 64)
        return ctx->fail(std::string("block store allocation failed: ") + cudaGetErrorString(e), 2);
      // the budget was optimistic (another allocator took the memory): make room and retry
      bool any = false;
      if(int rc = evict(ctx, bytes, true, &any)) return rc;
      if(!any && attempt > 0) return ctx->fail("out of device memory for the block store and nothing left to evict", 2);
      CK(cudaStreamSynchronize(ctx->s_stage)); // pending releases
    }
    if(from_peer) { // the copy from the owner's HBM is on its way on s_fetch
      ctx->block_bytes += bytes;
      BlockEntry be;
      be.dev = dev, be.bytes = bytes, be.last_use = for_clock, be.slab = slab, be.offset = slab_off;
      it = ctx->blocks.emplace(key, be).first;
    }
    else {
    uint8_t* host  = nullptr;
    size_t   begin = 0;
    if(int rc = ring_alloc(ctx, bytes, &host, &begin)) return rc;
    const double t0 = now_s();
    if(ctx->fetch(ctx->fetch_user, key.tensor, key.b, (double*) host, elems) != 0)
      return ctx->fail("fetch callback failed for tensor " + std::to_string(key.tensor), 5);
    ctx->stats.seconds_fetch += now_s() - t0;
    CK(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->s_fetch));
    cudaEvent_t ev = nullptr;
    if(!ctx->ring_events.empty()) {
      ev = ctx->ring_events.back();
      ctx->ring_events.pop_back();
    }
    else CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CK(cudaEventRecord(ev, ctx->s_fetch));
    ctx->ring_inflight.push_back(RingSeg{begin, ctx->ring_head, ev});
    ctx->fetch_dirty = true;
    ctx->stats.h2d_bytes += (int64_t) bytes;
    ctx->stats.blocks_fetched++;
    ctx->block_bytes += bytes;
    BlockEntry be;
    be.dev = dev, be.bytes = bytes, be.last_use = for_clock, be.pooled = slab < 0, be.slab = slab, be.offset = slab_off;
    it = ctx->blocks.emplace(key, be).first;
    if(int rc = share_publish_after(ctx, dir_entry, ctx->s_fetch)) return rc;
    }
    // (checked later, on the staging stream: a kernel on the fetch stream would have to wait for a slot among the fused
    //  CTAs and hold every copy queued behind it back with it)
    if(ctx->opt.symmetry && ctx->opt.check_symmetry >= 0) ctx->sym_check_pending.push_back(key);
  }
  it->second.last_use = std::max(it->second.last_use, for_clock);
  // row-major strides of the storage block
  const char* kinds = kKinds[key.tensor];
  const int   nd    = (int) strlen(kinds);
  int64_t     st    = 1;
  for(int d = nd - 1; d >= 0; d--) {
    out.stride[d] = st;
    st *= ctx->store.k_range[tile_of(ctx->store, kinds[d], key.b[d])];
  }
  for(int d = nd; d < 4; d++) out.stride[d] = 0;
  out.ptr = it->second.dev;
  return 0;
}

int resolve_dense(ccsdt_ctx* ctx, int tensor, const uint32_t exec_bid[4], BlockRef& out) {
  const Space& sp    = ctx->sp;
  const char*  kinds = kKinds[tensor];
  const int    nd    = (int) strlen(kinds);
  int64_t      st = 1, off = 0;
  for(int d = nd - 1; d >= 0; d--) {
    out.stride[d] = st;
    off += tile_start(sp, tile_of(sp, kinds[d], exec_bid[d])) * st;
    st *= dim_full(sp, kinds[d]);
  }
  for(int d = nd; d < 4; d++) out.stride[d] = 0;
  out.ptr = ctx->dense[tensor] + off;
  return 0;
}

int fetch_fence(ccsdt_ctx* ctx, cudaStream_t st) {
  if(!ctx->fetch_dirty) return 0;
  CK(cudaEventRecord(ctx->ev_fetched, ctx->s_fetch));
  CK(cudaStreamWaitEvent(st, ctx->ev_fetched, 0));
  ctx->fetch_dirty = false;
  return 0;
}

// which index pairs of a tensor are antisymmetric: first dimension of each pair, -1 = none
static const int kAntiPairs[5][2] = {{-1, -1}, {0, 2}, {0, 2}, {0, -1}, {2, -1}};

int check_block_symmetry(ccsdt_ctx* ctx, const BlockKey& key, const double* dev, cudaStream_t st) {
  if(!ctx->opt.symmetry || ctx->opt.check_symmetry < 0) return 0;
  const char* kinds = kKinds[key.tensor];
  for(int q = 0; q < 2; q++) {
    const int pair = kAntiPairs[key.tensor][q];
    if(pair < 0 || key.b[pair] != key.b[pair + 1]) continue; // visible inside one block only when the two tiles coincide
    int64_t n[4] = {1, 1, 1, 1}, stv[4] = {0, 0, 0, 0}, acc = 1;
    for(int d = 3; d >= 0; d--) {
      n[d]   = ctx->store.k_range[tile_of(ctx->store, kinds[d], key.b[d])];
      stv[d] = acc;
      acc *= n[d];
    }
    CK(launch_antisym_check(dev, n, stv, pair, ctx->d_symflag, st));
    ctx->stats.kernel_launches++;
  }
  return 0;
}

int flush_block_symmetry_checks(ccsdt_ctx* ctx) {
  for(const BlockKey& key: ctx->sym_check_pending) {
    auto it = ctx->blocks.find(key); // (a prefetched block may have been evicted again)
    if(it == ctx->blocks.end()) continue;
    if(int rc = check_block_symmetry(ctx, key, it->second.dev, ctx->s_stage)) return rc;
  }
  ctx->sym_check_pending.clear();
  return 0;
}

int check_dense_symmetry(ccsdt_ctx* ctx) {
  for(int t = 1; t < 5; t++) {
    if(!ctx->dense_check_pending[t]) continue;
    ctx->dense_check_pending[t] = false;
    if(!ctx->opt.symmetry || ctx->opt.check_symmetry < 0 || !ctx->dense[t]) continue;
    const char* kinds = kKinds[t];
    int64_t     n[4], stv[4], acc = 1;
    for(int d = 3; d >= 0; d--) {
      n[d]   = dim_full(ctx->sp, kinds[d]);
      stv[d] = acc;
      acc *= n[d];
    }
    if(ctx->upload_pending[t]) CK(cudaStreamWaitEvent(ctx->s_stage, ctx->ev_full[t], 0));
    // Same-spin index pairs only: that is what the box skipping relies on (it applies to coinciding TILES, hence equal
    // spins), and what the per-block check of the block store sees.  Mixed-spin blocks of a closed-shell tensor set are
    // related by the spin flip of the solver and agree to its convergence only -- and half of them are never read.
    for(int q = 0; q < 2; q++) {
      const int pair = kAntiPairs[t][q];
      if(pair < 0) continue;
      const bool virt = kinds[pair] == 'v';
      int64_t    lo   = 0;
      for(int spin = 1; spin <= 2; spin++) {
        int     tb, te;
        int64_t cnt = 0;
        ctx->sp.spin_range(virt, spin, tb, te, cnt);
        if(cnt > 0) {
          int64_t m[4] = {n[0], n[1], n[2], n[3]};
          m[pair] = m[pair + 1] = cnt;
          CK(launch_antisym_check(ctx->dense[t] + lo * stv[pair] + lo * stv[pair + 1], m, stv, pair, ctx->d_symflag, ctx->s_stage));
          ctx->stats.kernel_launches++;
        }
        lo += cnt;
      }
    }
  }
  return 0;
}

void source_pieces(const ccsdt_ctx* ctx, const SrcSpec& s, std::vector<SrcPiece>& out) {
  const Space &ex = ctx->sp, &stor = ctx->store;
  const char*  kinds = kKinds[s.tensor];
  const int    nd    = (int) strlen(kinds);
  std::vector<TilePiece> cut[4];
  for(int d = 0; d < nd; d++) {
    const int tile = tile_of(ex, kinds[d], s.bid[d]);
    if(ctx->retiled) cut[d] = split_tile(ex, stor, tile);
    else cut[d].push_back(TilePiece{tile, 0, 0, ex.k_range[tile]});
  }
  size_t idx[4] = {0, 0, 0, 0};
  for(;;) {
    SrcPiece p{};
    p.key.tensor = s.tensor;
    for(int d = 0; d < 4; d++) p.key.b[d] = 0, p.store_off[d] = 0, p.exec_off[d] = 0, p.len[d] = 1;
    for(int d = 0; d < nd; d++) {
      const TilePiece& c = cut[d][idx[d]];
      p.key.b[d]         = (uint32_t) (kinds[d] == 'o' ? c.store_tile : c.store_tile - stor.noab());
      p.store_off[d] = c.store_off, p.exec_off[d] = c.exec_off, p.len[d] = c.len;
    }
    p.sign  = (double) canonical_block(s.tensor, p.key.b, p.perm);
    p.elems = block_elems(stor, s.tensor, p.key.b);
    out.push_back(p);
    int d = nd - 1;
    while(d >= 0 && ++idx[d] == cut[d].size()) idx[d--] = 0;
    if(d < 0) break;
  }
}

} // namespace ccsdt

// =================================================================================================
// C ABI: operand supply
// =================================================================================================
extern "C" {

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
  ctx->dense_check_pending[tensor] = true;   // antisymmetry is verified by the next run (options.symmetry)
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
    const int lim = kKinds[tensor][d] == 'o' ? ctx->store.noab() : ctx->store.nvab();
    if((int) bid[d] >= lim) return ctx->fail("block id out of range");
  }
  const size_t n = block_elems(ctx->store, tensor, bid);
  BlockKey     key{tensor, {bid[0], bid[1], nd > 2 ? bid[2] : 0u, nd > 3 ? bid[3] : 0u}};
  auto         it = ctx->blocks.find(key);
  double*      dev;
  if(it == ctx->blocks.end()) {
    CK(cudaMalloc(&dev, n * 8));
    BlockEntry be;
    be.dev = dev, be.bytes = n * 8, be.pinned = true; // no callback could bring it back: never evicted
    ctx->blocks.emplace(key, be);
    ctx->block_bytes += n * 8;
  }
  else {
    dev = it->second.dev;
    CK(cudaStreamSynchronize(ctx->s_stage)); // a panel build may still read the old contents
  }
  CK(cudaMemcpy(dev, host, n * 8, cudaMemcpyHostToDevice));
  ctx->pending_h2d += (int64_t) n * 8;
  ctx->synthetic = false;
  return check_block_symmetry(ctx, key, dev, ctx->s_stage);
}

int ccsdt_set_fetch(ccsdt_ctx* ctx, ccsdt_fetch_fn fn, void* user) {
  if(!ctx) return 1;
  ctx->fetch      = fn;
  ctx->fetch_user = user;
  if(fn) ctx->synthetic = false;
  return 0;
}

int ccsdt_clear_blocks(ccsdt_ctx* ctx) {
  if(!ctx) return 1;
  cudaSetDevice(ctx->device);
  return clear_blocks(ctx, true);
}

int ccsdt_set_synthetic(ccsdt_ctx* ctx, uint64_t seed) {
  if(!ctx) return 1;
  ctx->synthetic = true;
  ctx->seed      = seed;
  return 0;
}

int64_t ccsdt_split_request(int noa, int nob, int nva, int nvb, const int64_t* store_range, const int32_t* store_spin,
                            const int64_t* exec_range, const int32_t* exec_counts, int tensor, const uint32_t exec_bid[4],
                            int64_t* pieces, int64_t cap) {
  if(tensor < 0 || tensor > 4 || !store_range || !store_spin || !exec_range || !exec_counts || !exec_bid) return -1;
  ccsdt_ctx tmp; // host-only use of the splitting logic: no device state is touched
  tmp.store = make_space(noa, nob, nva, nvb, store_range, store_spin, nullptr, true);
  std::vector<int32_t> es;
  for(int g = 0; g < 4; g++)
    for(int i = 0; i < exec_counts[g]; i++) es.push_back(g % 2 + 1);
  tmp.sp      = make_space(exec_counts[0], exec_counts[1], exec_counts[2], exec_counts[3], exec_range, es.data(), nullptr, true);
  tmp.retiled = !same_tiling(tmp.sp, tmp.store);
  SrcSpec s{};
  s.tensor = tensor;
  for(int d = 0; d < 4; d++) s.bid[d] = exec_bid[d];
  std::vector<SrcPiece> out;
  source_pieces(&tmp, s, out);
  for(int64_t i = 0; i < (int64_t) out.size() && i < cap; i++) {
    int64_t* q = pieces + 22 * i;
    for(int d = 0; d < 4; d++) {
      q[d]      = out[i].key.b[d];
      q[5 + d]  = out[i].perm[d];
      q[9 + d]  = out[i].store_off[d];
      q[13 + d] = out[i].exec_off[d];
      q[17 + d] = out[i].len[d];
    }
    q[4]  = (int64_t) out[i].sign;
    q[21] = (int64_t) out[i].elems;
  }
  return (int64_t) out.size();
}

} // extern "C"
