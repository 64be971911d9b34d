// Internal state of one ccsdt_ctx (one GPU) shared by the translation units of the driver:
//   ccsdt_capi.cu   context life cycle, options, staging of a task, the task loop
//   ccsdt_store.cu  operand store: dense tensors, the HBM block store with its asynchronous fetch ring and LRU
//                   eviction, the cut of execution tiles along storage tiles
//   ccsdt_v2.cu     V2 tensors from Cholesky vectors (hand-written DMMA GEMM)
//   ccsdt_comm.cu   the one collective (NCCL all-reduce of the two energies) and the process-shared task counter
#pragma once
#include "../../include/ccsdt_b200.h"
#include "ccsdt_device.hpp"
#include "ccsdt_host.hpp"

#include <array>
#include <deque>
#include <map>
#include <string>
#include <vector>

namespace ccsdt {

struct BlockRef {
  const double* ptr       = nullptr;
  int64_t       stride[4] = {0, 0, 0, 0};
};

// key of a STORAGE block: tensor id + the block id Tensor::get takes
struct BlockKey {
  int      tensor;
  uint32_t b[4];
  bool     operator<(const BlockKey& o) const {
    if(tensor != o.tensor) return tensor < o.tensor;
    for(int i = 0; i < 4; i++)
      if(b[i] != o.b[i]) return b[i] < o.b[i];
    return false;
  }
};

struct BlockEntry {
  double* dev      = nullptr;
  size_t  bytes    = 0;
  int64_t last_use = 0;     // use_clock of the last task that reads it (a prefetched block carries its FUTURE clock)
  bool    pinned   = false; // supplied by ccsdt_put_block: cannot be fetched again, never evicted
  bool    pooled   = false; // allocated from the context's private memory pool (fetch path)
  int     slab     = -1;    // >= 0: carved from slab `slab` of the node-shared store (ccsdt_share.cu) at `offset`
  size_t  offset   = 0;
};

// which index kind ('o'/'v') each dimension of a tensor has
extern const char* const kKinds[5];

// a region of the pinned fetch ring whose host-to-device copy is still in flight
struct RingSeg {
  size_t      begin, end;
  cudaEvent_t ev;
};

struct ShareState;

struct StageBuf {
  PoolGeom     geom{};
  size_t       hpp_cap[2] = {0, 0}, hhp_cap[2] = {0, 0}, s1a_cap = 0, s1b_cap = 0; // bytes allocated
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
  int64_t      trace_slot = -1;     // options.verbose >= 2: row of the run's trace this buffer's task fills
};

// one row of the per-task trace (options.verbose >= 2, printed to stderr at the end of the run)
struct TraceRow {
  int64_t task;
  double  host_begin, host_staged; // seconds after the start of the run: staging began / kernels were enqueued
  double  fetch_s;                 // host time inside the fetch callback while staging this task (prefetch excluded)
  double  gather0, gather1, k0, k1; // ms after ev_base on the GPU time line
};

// one enabled source of a task's panels, in execution-tile block ids
struct SrcSpec {
  int      tensor;
  uint32_t bid[4];
  int      dim_of[4]; // which tensor dimension feeds dst dim j (dst dims ordered o2, o1, in, k); -1: extent 1
  double   scale;
};

} // namespace ccsdt

struct ccsdt_ctx {
  int            device = 0;
  cudaDeviceProp prop{};
  std::string    err;
  ccsdt_options  opt{};
  ccsdt::Space   store;      // the caller's tiling: block ids of put_block / fetch
  ccsdt::Space   sp;         // execution tiling: task list, panels, kernel (== store unless options.exec_tilesize)
  bool           retiled = false;
  bool           have_space = false;
  std::vector<ccsdt::Task> tasks;
  std::vector<int32_t>     owner;
  int64_t                  n_outer = 0;

  // ---- operands ----
  double*                                      dense[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t                                       dense_elems[5] = {0, 0, 0, 0, 0};
  bool                                         dense_sparse_ok[5] = {false, false, false, false, false};
  std::map<ccsdt::BlockKey, ccsdt::BlockEntry> blocks;
  size_t                                       block_bytes = 0, block_budget = 0;
  int64_t                                      use_clock = 0;
  ccsdt_fetch_fn                               fetch = nullptr;
  void*                                        fetch_user = nullptr;
  bool                                         synthetic = false;
  uint64_t                                     seed = 0;
  // asynchronous fetch: blocks land in a pinned ring and travel on s_fetch; the panel build waits for ev_fetched
  cudaMemPool_t               block_pool = nullptr;
  cudaStream_t                s_fetch = nullptr;
  cudaEvent_t                 ev_fetched = nullptr;
  uint8_t*                    ring = nullptr;
  size_t                      ring_cap = 0, ring_head = 0;
  std::deque<ccsdt::RingSeg>  ring_inflight;
  std::vector<cudaEvent_t>    ring_events; // recycled
  bool                        fetch_dirty = false; // copies were issued since ev_fetched was recorded

  // ---- staging / compute ----
  double*         d_evl = nullptr;
  size_t          evl_cap = 0;
  ccsdt::StageBuf buf[2];
  bool            pools_ready = false;
  double*         d_task_energy = nullptr;
  int64_t         task_energy_cap = 0;
  uint32_t*       d_error = nullptr;   // [0] error word of the fused kernel, [1..2] watchdog limit, [3] unused
  uint32_t*       d_symflag = nullptr; // set by antisym_check_kernel: an operand is not antisymmetric (options.symmetry)
  bool            dense_check_pending[5] = {false, false, false, false, false};
  std::vector<ccsdt::BlockKey> sym_check_pending; // fetched blocks whose antisymmetry check has not been queued yet
  cudaStream_t    s_compute = nullptr, s_compute2 = nullptr, s_stage = nullptr;
  cudaEvent_t     ev_base = nullptr;   // start of the current run: kernel intervals are placed on its time line
  double          kernel_busy_until = 0.0; // end (ms after ev_base) of the union of fused-kernel intervals so far
  void*           encode_fn = nullptr;
  int64_t*        task_counter = nullptr; // process-shared dynamic task counter (NULL = static split)
  ccsdt_stats     stats{};
  std::vector<ccsdt::TraceRow> trace;
  int64_t         pending_h2d = 0; // bytes uploaded by ccsdt_put_* since the last run
  // asynchronous dense uploads (ccsdt_put_dense_async): the all-alpha blocks of every tensor travel on s_copy_a,
  // the other spin patterns on s_copy_b; tasks whose six tiles are all alpha only wait for the first
  cudaStream_t s_copy_a = nullptr, s_copy_b = nullptr;
  cudaEvent_t  ev_alpha[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}, ev_full[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  bool         upload_pending[5] = {false, false, false, false, false};
  // symmetry-reduced box lists (device), keyed by (nbox, brick, nbrick, sym): a handful per job
  struct BoxList {
    int32_t* dev = nullptr;
    int32_t  n   = 0;
    double   units_x = 0, units_y = 0; // sum over the listed boxes of (DMMAs per k-step and warp) x (warps with work), layout X / Y terms
  };
  std::map<std::array<int, 19>, BoxList> box_lists;
  // node-shared block directory (ccsdt_share.cu); NULL = private store
  ccsdt::ShareState* share = nullptr;
  // the one collective (ccsdt_comm.cu)
  void* nccl_comm = nullptr;
  double* d_allreduce = nullptr;

  int fail(const std::string& m, int code = 1) {
    err = m;
    return code;
  }
};

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if(e_ != cudaSuccess)                                                                          \
      return ctx->fail(std::string(#call) + ": " + cudaGetErrorName(e_) + " " + cudaGetErrorString(e_), 2); \
  } while(0)

namespace ccsdt {

// ---- ccsdt_store.cu ----
int64_t dim_full(const Space& sp, char kind);
int     tile_of(const Space& sp, char kind, uint32_t bid);      // block id inside the tensor's sub-space -> global tile id
int64_t tile_start(const Space& sp, int tile);                  // orbital number inside the occupied / virtual range
size_t  block_elems(const Space& sp, int tensor, const uint32_t bid[4]);
int     store_create(ccsdt_ctx* ctx);                           // streams, pool, events of the fetch path
void    store_destroy(ccsdt_ctx* ctx);
void    free_operands(ccsdt_ctx* ctx);                          // dense tensors and every block
int     clear_blocks(ccsdt_ctx* ctx, bool keep_pinned);
int     update_block_budget(ccsdt_ctx* ctx);
int     evict_stale(ccsdt_ctx* ctx, size_t bytes); // LRU: make room for `bytes` within the budget, as far as the rule allows
// Storage pieces of one source: calls fn(key, elems, piece) for every storage block the execution-tile source
// overlaps.  piece.src_off / stride are filled by resolve (fn may ignore them).
struct SrcPiece {
  BlockKey key;          // canonical storage block
  size_t   elems;        // elements of that block
  int      perm[4];      // requested dim d is dim perm[d] of the canonical block
  double   sign;         // antisymmetry sign of the canonicalisation
  int64_t  store_off[4]; // per REQUESTED tensor dim: first element inside the storage tile
  int64_t  exec_off[4];  //                          first element inside the execution tile
  int64_t  len[4];
};
void source_pieces(const ccsdt_ctx* ctx, const SrcSpec& s, std::vector<SrcPiece>& out);
// block -> device pointer + strides (fetching it through the callback if it is not resident); for_clock = use_clock
// of the task that needs it (prefetch passes a future clock)
int  resolve_block(ccsdt_ctx* ctx, const BlockKey& key, size_t elems, int64_t for_clock, BlockRef& out);
int  resolve_dense(ccsdt_ctx* ctx, int tensor, const uint32_t exec_bid[4], BlockRef& out);
// records ev_fetched after the copies issued so far and makes `st` wait for it
int  fetch_fence(ccsdt_ctx* ctx, cudaStream_t st);
// queues the antisymmetry checks (options.symmetry / check_symmetry) of a storage block on `st`, of the dense tensors on s_stage
int  check_block_symmetry(ccsdt_ctx* ctx, const BlockKey& key, const double* dev, cudaStream_t st);
int  check_dense_symmetry(ccsdt_ctx* ctx);
int  flush_block_symmetry_checks(ccsdt_ctx* ctx); // queues the checks of the blocks fetched since the last call on s_stage

// ---- ccsdt_share.cu ----
int    share_acquire(ccsdt_ctx* ctx, const BlockKey& key, size_t bytes, double** dev, int* slab, size_t* offset, void** entry);
int    share_publish_after(ccsdt_ctx* ctx, void* entry, cudaStream_t st);
bool   share_release(ccsdt_ctx* ctx, const BlockKey& key, BlockEntry& be, bool wait);
void   share_free(ccsdt_ctx* ctx, int slab, size_t offset, size_t bytes);
void   share_poll(ccsdt_ctx* ctx, bool wait);
void   share_detach(ccsdt_ctx* ctx);
size_t share_unused_bytes(const ccsdt_ctx* ctx);

// ---- ccsdt_comm.cu ----
void comm_destroy(ccsdt_ctx* ctx);

} // namespace ccsdt
