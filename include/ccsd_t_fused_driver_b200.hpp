// ccsd_t_fused_driver_b200.hpp -- drop-in replacement for ExaChem's
//   exachem/cc/ccsd_t/ccsd_t_fused_driver.hpp
// Same class name, same two member signatures (reference lines 73-81 and 83-87), same return tuple
// (energy1 = E[T] rank-partial, energy2 = E(T) rank-partial, work seconds, total seconds; reference
// line 545), so exachem/cc/ccsd_t/ccsd_t.cpp:253-256 and :298-301 compile and behave unchanged.  The
// body is a thin adapter: it reads the tile space from TAMM, serves tensor blocks to the library
// through Tensor<T>::get, and calls the C ABI of include/ccsdt_b200.h.  All arithmetic happens in
// libccsdt_b200.so on the GPU; there is no CPU path here.
//
// Contract kept from the reference:
//   * one host rank drives one GPU (the rank's current CUDA device, as TAMM bound it).  ONE rank per GPU: the
//     fused kernel sizes its grid for the whole device, and several ranks time-slicing one GPU only add latency;
//   * the returned energies are rank PARTIALS -- the caller sums them over ranks (ccsd_t.cpp:262-263).
//     With CCSDT_B200_INTERNAL_ALLREDUCE defined the library's own ncclAllReduce combines them
//     (ccsdt_comm_*; the unique id travels through ec.pg().broadcast) and the total is returned on rank 0
//     and 0.0 elsewhere, so that the caller's reduce still yields the total;
//   * task hand-out across ranks: on one node (ec.nnodes() == 1) the ranks share an atomic counter in POSIX
//     shared memory and claim tasks longest-first (the role of AtomicCounterGA, reference lines 169-172, 381,
//     456), and a block directory through which a block pulled from Tensor::get by one rank is read from that
//     rank's HBM by the others (ccsdt_share_attach); across nodes, or with CCSDT_B200_DYNAMIC=0, the static
//     cost-balanced split of the library and private block stores;
//   * execution tiling: by default (CCSDT_B200_EXEC_TILESIZE unset or -1) the library re-cuts ragged or tiny
//     tiles (ts 28 -> execution tiles of 40, see ccsdt_options.exec_tilesize); blocks are still requested from
//     Tensor::get in the caller's tiling and only canonically ordered ones, like the reference.  0 keeps the
//     caller's tiles (the reference's task list, task for task);
//   * the six LRUCache arguments are accepted and ignored: the library's HBM block store caches
//     every fetched block for the whole call (the reference caches sorted copies on the host);
//   * hf_ccsd_energy, seq_h3b and tilesize_opt are unused, as in the reference body;
//   * errors are fatal: a non-zero C-ABI status becomes tamm_terminate(message), matching
//     CUDA_SAFE -> exit(100) (ccsd_t_common.hpp:26-31) and ccsd_t.cpp:27-28.
//
// Requirements on the including translation unit (the same the reference header has): the TAMM
// umbrella header (Index, IndexVector, Tensor, LRUCache, TiledIndexSpace, ExecutionContext),
// ChemEnv and exachem::cholesky_2e::V2Tensors must be declared before this header is included.
#pragma once

#include "ccsdt_b200.h"

#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <string>
#include <unistd.h>
#include <tuple>
#include <type_traits>
#include <vector>

#ifndef CCSDT_B200_TERMINATE
// real TAMM provides tamm::tamm_terminate(std::string); a test shim may define this macro instead
#define CCSDT_B200_TERMINATE(msg) tamm_terminate(msg)
#endif
#ifndef CCSDT_B200_SINGLE_NODE
// all ranks of ec.pg() share one node (and its POSIX shared memory): TAMM's ExecutionContext knows
#define CCSDT_B200_SINGLE_NODE(ec) ((ec).nnodes() == 1)
#endif

namespace ccsdt_b200_detail {

template<typename T>
struct FetchUser {
  Tensor<T>*     tensor[5];
  std::vector<T> buf;
};

// ccsdt_fetch_fn trampoline: one unsorted row-major block, exactly what Tensor<T>::get returns
// (the reference's call sites: ccsd_t_all_fused_singles.hpp:200,304; ..._doubles1.hpp:222,237,282;
//  ..._doubles2.hpp:215,230,335)
template<typename T>
int fetch_block(void* user, int tensor, const uint32_t bid[4], double* dst, size_t n) {
  auto*       u  = static_cast<FetchUser<T>*>(user);
  const int   nd = tensor == CCSDT_T1 ? 2 : 4;
  IndexVector id(nd);
  for(int i = 0; i < nd; i++) id[i] = (Index) bid[i];
#if defined(CCSDT_B200_SPAN_TYPE)
  // Tensor<T>::get of this TAMM takes a span: the block is written straight into the library's pinned buffer
  // (define CCSDT_B200_SPAN_TYPE, e.g. to tamm::span, before including this header; T must be double)
  if constexpr(std::is_same<T, double>::value) {
    u->tensor[tensor]->get(id, CCSDT_B200_SPAN_TYPE<T>(dst, n));
    return 0;
  }
#endif
  if(u->buf.size() < n) u->buf.resize(n);
  u->tensor[tensor]->get(id, u->buf);
  for(size_t i = 0; i < n; i++) dst[i] = (double) u->buf[i];
  return 0;
}

struct TileSpace {
  int                  noa, nob, nva, nvb;
  std::vector<int64_t> k_range;
  std::vector<int32_t> k_spin;
};

inline TileSpace read_space(const TiledIndexSpace& MO, const std::vector<int>& k_spin) {
  TileSpace  s;
  const int  noab = (int) MO("occ").num_tiles(), nvab = (int) MO("virt").num_tiles();
  s.noa = (int) MO("occ_alpha").num_tiles();
  s.nva = (int) MO("virt_alpha").num_tiles();
  s.nob = noab - s.noa;
  s.nvb = nvab - s.nva;
  for(auto x: MO.input_tile_sizes()) s.k_range.push_back((int64_t) x);
  for(auto x: k_spin) s.k_spin.push_back((int32_t) x);
  return s;
}

} // namespace ccsdt_b200_detail

template<typename T>
class CCSD_T_Fused_Driver {
public:
  CCSD_T_Fused_Driver() = default;

  virtual ~CCSD_T_Fused_Driver() = default;

  CCSD_T_Fused_Driver(const CCSD_T_Fused_Driver&)            = default;
  CCSD_T_Fused_Driver& operator=(const CCSD_T_Fused_Driver&) = default;
  CCSD_T_Fused_Driver(CCSD_T_Fused_Driver&&)                 = default;
  CCSD_T_Fused_Driver& operator=(CCSD_T_Fused_Driver&&)      = default;

  virtual std::tuple<T, T, double, double>
  execute(ChemEnv& chem_env, ExecutionContext& ec, std::vector<int>& k_spin,
          const TiledIndexSpace& MO, Tensor<T>& d_t1, Tensor<T>& d_t2,
          exachem::cholesky_2e::V2Tensors<T>& d_v2, std::vector<T>& k_evl_sorted, T hf_ccsd_energy,
          bool is_restricted, LRUCache<Index, std::vector<T>>& cache_s1t,
          LRUCache<Index, std::vector<T>>& cache_s1v, LRUCache<Index, std::vector<T>>& cache_d1t,
          LRUCache<Index, std::vector<T>>& cache_d1v, LRUCache<Index, std::vector<T>>& cache_d2t,
          LRUCache<Index, std::vector<T>>& cache_d2v, bool seq_h3b = false,
          bool tilesize_opt = true);

  virtual void calculate_performance_ops(ChemEnv& chem_env, ExecutionContext& ec,
                                         std::vector<int>& k_spin, const TiledIndexSpace& MO,
                                         std::vector<T>& k_evl_sorted, double hf_ccsd_energy,
                                         bool is_restricted, long double& total_num_ops,
                                         bool seq_h3b = false);

  // ---- not part of the reference interface ----
  // statistics of the last execute() on this object
  ccsdt_stats last_stats{};
  // when non-empty, execute() runs only these kernel tasks (indices into the execution task list) instead of the
  // whole job: benchmarks and tests of problems whose whole job takes hours
  std::vector<int64_t> task_subset;
  // library options applied on top of the defaults (rank / nranks are always taken from ec)
  bool          have_options = false;
  ccsdt_options options{};
};

template<typename T>
std::tuple<T, T, double, double> CCSD_T_Fused_Driver<T>::execute(
  ChemEnv& chem_env, ExecutionContext& ec, std::vector<int>& k_spin, const TiledIndexSpace& MO,
  Tensor<T>& d_t1, Tensor<T>& d_t2, exachem::cholesky_2e::V2Tensors<T>& d_v2,
  std::vector<T>& k_evl_sorted, T hf_ccsd_energy, bool is_restricted,
  LRUCache<Index, std::vector<T>>& cache_s1t, LRUCache<Index, std::vector<T>>& cache_s1v,
  LRUCache<Index, std::vector<T>>& cache_d1t, LRUCache<Index, std::vector<T>>& cache_d1v,
  LRUCache<Index, std::vector<T>>& cache_d2t, LRUCache<Index, std::vector<T>>& cache_d2v,
  bool seq_h3b, bool tilesize_opt) {
  (void) chem_env; (void) hf_ccsd_energy; (void) seq_h3b; (void) tilesize_opt;
  (void) cache_s1t; (void) cache_s1v; (void) cache_d1t; (void) cache_d1v; (void) cache_d2t; (void) cache_d2v;
  using namespace ccsdt_b200_detail;

  const TileSpace     s = read_space(MO, k_spin);
  std::vector<double> evl(k_evl_sorted.begin(), k_evl_sorted.end());

  ccsdt_ctx* ctx = nullptr;
  if(ccsdt_create(&ctx, /*device=*/-1 /* the rank's current CUDA device */))
    CCSDT_B200_TERMINATE(std::string("[CCSD(T) B200] ") + ccsdt_last_error(nullptr));
  auto check = [&](int rc) {
    if(rc) {
      std::string msg = std::string("[CCSD(T) B200] ") + ccsdt_last_error(ctx);
      ccsdt_destroy(ctx);
      CCSDT_B200_TERMINATE(msg);
    }
  };

  const int     rank = (int) ec.pg().rank().value(), nranks = (int) ec.pg().size().value();
  ccsdt_options opt;
  ccsdt_default_options(&opt);
  opt.exec_tilesize = -1; // auto: re-cut ragged / tiny tiles; totals do not depend on it
  if(have_options) opt = options;
  opt.rank   = (int32_t) rank;
  opt.nranks = (int32_t) nranks;
  // CCSDT_B200_SYMMETRY=0 evaluates every element of every task as the reference does (see ccsdt_options.symmetry)
  if(const char* e = std::getenv("CCSDT_B200_SYMMETRY")) opt.symmetry = std::atoi(e) != 0;
  if(const char* e = std::getenv("CCSDT_B200_EXEC_TILESIZE")) opt.exec_tilesize = std::atoi(e);
  if(const char* e = std::getenv("CCSDT_B200_VERBOSE")) opt.verbose = std::atoi(e); // 2: per-task trace on stderr
  if(const char* e = std::getenv("CCSDT_B200_BLOCK_BUDGET_MB")) opt.block_budget_bytes = (int64_t) (std::atof(e) * 1048576.0);
  check(ccsdt_set_options(ctx, &opt));
  check(ccsdt_set_space(ctx, s.noa, s.nob, s.nva, s.nvb, s.k_range.data(), s.k_spin.data(), evl.data(),
                        is_restricted ? 1 : 0));

  if(opt.verbose >= 1 && rank == 0) {
    // the counterpart of the reference's memory summary (ccsd_t.cpp:95-152): what the path takes of this GPU's HBM
    ccsdt_memory_estimate m;
    if(ccsdt_estimate_memory(s.noa, s.nob, s.nva, s.nvb, s.k_range.data(), s.k_spin.data(), opt.exec_tilesize, &m) == 0)
      std::fprintf(stderr,
                   "[CCSD(T) B200] execution tiles up to %lld (occ) / %lld (virt); HBM per GPU: operand panels %.2f GiB, s1 %.2f GiB, "
                   "blocks of one task <= %.2f GiB (minimum %.2f GiB); T2 %.1f, v2ijab %.1f, v2ijka %.1f, v2iabc %.1f GiB if fully resident\n",
                   (long long) m.exec_max_hole_tile, (long long) m.exec_max_particle_tile, m.panel_bytes / 1073741824.0,
                   m.s1_bytes / 1073741824.0, m.task_block_bytes / 1073741824.0, m.minimum_bytes / 1073741824.0,
                   m.tensor_bytes[1] / 1073741824.0, m.tensor_bytes[2] / 1073741824.0, m.tensor_bytes[3] / 1073741824.0,
                   m.tensor_bytes[4] / 1073741824.0);
  }

  FetchUser<T> user;
  user.tensor[CCSDT_T1]     = &d_t1;
  user.tensor[CCSDT_T2]     = &d_t2;
  user.tensor[CCSDT_V_IJAB] = &d_v2.v2ijab;
  user.tensor[CCSDT_V_IJKA] = &d_v2.v2ijka;
  user.tensor[CCSDT_V_IABC] = &d_v2.v2iabc;
  check(ccsdt_set_fetch(ctx, &fetch_block<T>, &user));

  // dynamic hand-out among the ranks of one node: rank 0 creates the shared counter, the others attach after a
  // barrier (reference: AtomicCounterGA allocate / fetch_add / deallocate, lines 169-172, 456, 541)
  int64_t*    counter = nullptr;
  std::string counter_name;
  bool        dynamic = nranks > 1 && CCSDT_B200_SINGLE_NODE(ec);
  if(const char* e = std::getenv("CCSDT_B200_DYNAMIC")) dynamic = dynamic && std::atoi(e) != 0;
  if(dynamic) {
    const char* key = std::getenv("CCSDT_B200_COUNTER_KEY");
    counter_name    = "/ccsdt_b200_" + std::to_string((long) getuid()) + "_" + (key ? key : "0");
    if(rank == 0 && ccsdt_task_counter_open(counter_name.c_str(), 1, &counter)) {
      ccsdt_destroy(ctx);
      CCSDT_B200_TERMINATE("[CCSD(T) B200] cannot create the shared task counter " + counter_name);
    }
    ec.pg().barrier();
    if(rank != 0 && ccsdt_task_counter_open(counter_name.c_str(), 0, &counter)) {
      ccsdt_destroy(ctx);
      CCSDT_B200_TERMINATE("[CCSD(T) B200] cannot attach to the shared task counter " + counter_name);
    }
    check(ccsdt_set_task_counter(ctx, counter));
  }
  // The node-shared block store (also with the static split): a block is fetched through Tensor::get by ONE rank of the
  // node and read from its HBM by the others (CCSDT_B200_SHARE=0 keeps every rank's store private)
  bool share = nranks > 1 && CCSDT_B200_SINGLE_NODE(ec);
  if(const char* e = std::getenv("CCSDT_B200_SHARE")) share = share && std::atoi(e) != 0;
  if(share) {
    const char*       key = std::getenv("CCSDT_B200_COUNTER_KEY");
    const std::string dir = "/ccsdt_b200_dir_" + std::to_string((long) getuid()) + "_" + (key ? key : "0");
    if(rank == 0) check(ccsdt_share_attach(ctx, dir.c_str(), rank, nranks, 1));
    ec.pg().barrier();
    if(rank != 0) check(ccsdt_share_attach(ctx, dir.c_str(), rank, nranks, 0));
  }

  double energies[2] = {0.0, 0.0};
  if(task_subset.empty()) check(ccsdt_run(ctx, 0, -1, energies, nullptr, &last_stats));
  else check(ccsdt_run_tasks(ctx, task_subset.data(), (int64_t) task_subset.size(), energies, nullptr, &last_stats));

#if defined(CCSDT_B200_INTERNAL_ALLREDUCE)
  if(nranks > 1) {
    // the one collective, inside the library: ncclAllReduce of {E[T], E(T)}
    unsigned char id[128];
    if(rank == 0 && ccsdt_comm_unique_id(id)) {
      ccsdt_destroy(ctx);
      CCSDT_B200_TERMINATE("[CCSD(T) B200] ncclGetUniqueId failed (is libnccl.so.2 loadable?)");
    }
    ec.pg().broadcast(id, sizeof(id), 0);
    check(ccsdt_comm_init(ctx, id, rank, nranks));
    check(ccsdt_comm_allreduce(ctx, energies));
    if(rank != 0) energies[0] = energies[1] = 0.0; // the caller reduces what execute returns (ccsd_t.cpp:262-263)
  }
#endif
  if(dynamic) {
    check(ccsdt_set_task_counter(ctx, nullptr));
    ec.pg().barrier(); // every rank has drawn its last task
    ccsdt_task_counter_close(counter, counter_name.c_str(), rank == 0 ? 1 : 0);
  }
  ccsdt_destroy(ctx);
  ec.pg().barrier(); // the reference ends its timed region with a barrier (line 536)

  return std::make_tuple((T) energies[0], (T) energies[1], last_stats.seconds_kernel + last_stats.seconds_staging,
                         last_stats.seconds_total);
}

template<typename T>
void CCSD_T_Fused_Driver<T>::calculate_performance_ops(ChemEnv& chem_env, ExecutionContext& ec,
                                                       std::vector<int>& k_spin, const TiledIndexSpace& MO,
                                                       std::vector<T>& k_evl_sorted, double hf_ccsd_energy,
                                                       bool is_restricted, long double& total_num_ops,
                                                       bool seq_h3b) {
  (void) chem_env; (void) ec; (void) k_evl_sorted; (void) hf_ccsd_energy; (void) seq_h3b;
  const auto s = ccsdt_b200_detail::read_space(MO, k_spin);
  if(ccsdt_count_ops(s.noa + s.nob, s.nva + s.nvb, s.k_spin.data(), s.k_range.data(), is_restricted ? 1 : 0,
                     &total_num_ops))
    CCSDT_B200_TERMINATE(std::string("[CCSD(T) B200] ccsdt_count_ops failed"));
}
