// Test harness for include/ccsd_t_fused_driver_b200.hpp: compiles the drop-in header against the
// TAMM stand-in the oracle uses (oracle/shim/tamm/tamm.hpp -- test infrastructure) and exposes the
// same extern "C" entry the reference harness has (oracle/ref_driver.cpp: ref_ccsdt_execute), so a
// parity test can call "reference execute" and "B200 execute" with identical arguments.
// Links against exachem_b200/libccsdt_b200.so; nothing here computes.
#include "tamm/tamm.hpp"

#include <cstring>
#include <array>
#include <map>
#include <memory>
#include <stdexcept>

#define CCSDT_B200_TERMINATE(msg) throw std::runtime_error(msg)
#ifndef ADAPTER_VECTOR_GET // default: the span flavour of Tensor::get (blocks land in the library's pinned buffer)
#define CCSDT_B200_SPAN_TYPE tamm::span
#endif
#include "ccsd_t_fused_driver_b200.hpp"

namespace {

struct Space {
  int                 noab = 0, nvab = 0;
  std::vector<size_t> k_range, k_offset;
  std::vector<int>    k_spin;
  size_t              Ot = 0, Vt = 0;
};

Space make_space(int noab, int nvab, const int64_t* k_range, const int32_t* k_spin) {
  Space s;
  s.noab = noab, s.nvab = nvab;
  size_t sum = 0;
  for(int i = 0; i < noab + nvab; i++) {
    s.k_range.push_back((size_t) k_range[i]);
    s.k_offset.push_back(sum);
    sum += (size_t) k_range[i];
    s.k_spin.push_back(k_spin[i]);
    (i < noab ? s.Ot : s.Vt) += (size_t) k_range[i];
  }
  return s;
}

// a TAMM-like tensor over a dense row-major spin-orbital array; kinds[d] in {'o','v'}
Tensor<double> dense_tensor(const Space& s, const double* data, std::string kinds) {
  auto copy_block = [&s, data, kinds](const IndexVector& bid, std::vector<double>* vec, double* raw) {
    const int d = (int) kinds.size();
    size_t    ext[4], off[4], stride[4], st = 1, n = 1;
    for(int i = d - 1; i >= 0; i--) {
      const bool   occ  = kinds[i] == 'o';
      const size_t tile = occ ? bid[i] : bid[i] + s.noab;
      ext[i]    = s.k_range[tile];
      off[i]    = s.k_offset[tile] - (occ ? 0 : s.Ot);
      stride[i] = st;
      st *= occ ? s.Ot : s.Vt;
      n *= ext[i];
    }
    if(vec && vec->size() < n) vec->resize(n);
    double* buf = vec ? vec->data() : raw;
    // rows of the last index are contiguous in the dense array: one memcpy per row (a local TAMM block read is a
    // plain copy too; an element-by-element stand-in would dominate the end-to-end time of the adapter)
    const size_t        row = ext[d - 1], nrows = n / row;
    std::vector<size_t> idx(d, 0);
    for(size_t r = 0; r < nrows; r++) {
      size_t o = off[d - 1];
      for(int i = 0; i < d - 1; i++) o += (off[i] + idx[i]) * stride[i];
      std::memcpy(&buf[r * row], data + o, row * sizeof(double));
      for(int i = d - 2; i >= 0 && ++idx[i] == ext[i]; i--) idx[i] = 0;
    }
  };
  Tensor<double> t([copy_block](const IndexVector& bid, std::vector<double>& buf) { copy_block(bid, &buf, nullptr); });
  t.set_span_fetch([copy_block](const IndexVector& bid, double* out, size_t) { copy_block(bid, nullptr, out); });
  return t;
}

std::string g_error;

} // namespace

extern "C" {

__attribute__((visibility("default"))) const char* adapter_last_error() { return g_error.c_str(); }

// out[0..3] = the tuple execute returns; gets[0..4] = Tensor::get calls per tensor; stats_out (optional);
// subset (optional): CCSD_T_Fused_Driver::task_subset; opt (optional): CCSD_T_Fused_Driver::options
__attribute__((visibility("default"))) int
adapter_ccsdt_execute_ex(int noa, int nob, int nva, int nvb, const int64_t* k_range, const int32_t* k_spin,
                         const double* evl, const double* t1, const double* t2, const double* v2ijab,
                         const double* v2ijka, const double* v2iabc, int is_restricted, int tilesize, double* out,
                         int64_t* gets, long double* total_num_ops, ccsdt_stats* stats_out, const int64_t* subset,
                         int64_t nsubset, const ccsdt_options* opt) {
  try {
    Space             s = make_space(noa + nob, nva + nvb, k_range, k_spin);
    std::vector<Tile> tiles(s.k_range.begin(), s.k_range.end());
    TiledIndexSpace   MO(tiles, noa, nob, nva, nvb);
    ExecutionContext  ec;
    ChemEnv           chem_env;
    chem_env.ioptions.ccsd_options.ccsdt_tilesize = tilesize;
    Tensor<double>                          d_t1 = dense_tensor(s, t1, "vo");
    Tensor<double>                          d_t2 = dense_tensor(s, t2, "vvoo");
    exachem::cholesky_2e::V2Tensors<double> d_v2;
    d_v2.v2ijab = dense_tensor(s, v2ijab, "oovv");
    d_v2.v2ijka = dense_tensor(s, v2ijka, "ooov");
    d_v2.v2iabc = dense_tensor(s, v2iabc, "ovvv");
    std::vector<double> k_evl(evl, evl + s.Ot + s.Vt);
    // the caller's caches (exachem/cc/ccsd_t/ccsd_t.cpp:236-241); the adapter ignores them
    LRUCache<Index, std::vector<double>> c1{8}, c2{8}, c3{8}, c4{8}, c5{8}, c6{8};

    CCSD_T_Fused_Driver<double> drv;
    if(subset && nsubset > 0) drv.task_subset.assign(subset, subset + nsubset);
    if(opt) drv.have_options = true, drv.options = *opt;
    // same call as exachem/cc/ccsd_t/ccsd_t.cpp:253-256
    auto [e1, e2, tw, tt] = drv.execute(chem_env, ec, s.k_spin, MO, d_t1, d_t2, d_v2, k_evl, 0.0,
                                        is_restricted != 0, c1, c2, c3, c4, c5, c6, true);
    // ... and the caller's reduction over ranks (ccsd_t.cpp:262-263); a no-op on one rank
    out[0] = ec.pg().allreduce_sum(e1), out[1] = ec.pg().allreduce_sum(e2), out[2] = tw, out[3] = tt;
    if(gets) {
      gets[0] = (int64_t) d_t1.num_gets, gets[1] = (int64_t) d_t2.num_gets;
      gets[2] = (int64_t) d_v2.v2ijab.num_gets, gets[3] = (int64_t) d_v2.v2ijka.num_gets;
      gets[4] = (int64_t) d_v2.v2iabc.num_gets;
    }
    if(stats_out) *stats_out = drv.last_stats;
    if(total_num_ops) // same call as ccsd_t.cpp:298-301
      drv.calculate_performance_ops(chem_env, ec, s.k_spin, MO, k_evl, 0.0, is_restricted != 0, *total_num_ops, true);
    return 0;
  } catch(const std::exception& e) {
    g_error = e.what();
    return 1;
  }
}

__attribute__((visibility("default"))) int
adapter_ccsdt_execute(int noa, int nob, int nva, int nvb, const int64_t* k_range, const int32_t* k_spin,
                      const double* evl, const double* t1, const double* t2, const double* v2ijab,
                      const double* v2ijka, const double* v2iabc, int is_restricted, int tilesize, double* out,
                      int64_t* gets, long double* total_num_ops, ccsdt_stats* stats_out) {
  return adapter_ccsdt_execute_ex(noa, nob, nva, nvb, k_range, k_spin, evl, t1, t2, v2ijab, v2ijka, v2iabc, is_restricted,
                                  tilesize, out, gets, total_num_ops, stats_out, nullptr, 0, nullptr);
}

// ---- block-table tensors: problems whose dense tensors do not fit the host ((60,500): v2iabc is 360 GB) -------------
// The "application" holds only the blocks its tasks touch, each as one contiguous host buffer (what a local TAMM block
// is).  A block that is asked for the first time is produced by the library's device generator (ccsdt_synth_block,
// the same counter-based values the procedural device path evaluates) and kept; afterwards Tensor::get is a memcpy.
struct BlockTable {
  Space                 s;
  std::vector<int>      orb; // noa, nob, nva, nvb in ORBITALS
  uint64_t              seed   = 0;
  int                   device = 0;
  std::vector<double>   evl;
  std::map<std::array<uint32_t, 5>, std::unique_ptr<double[]>> blocks;
  size_t                bytes = 0, misses = 0;
  const double* get(int tensor, const IndexVector& bid, size_t& n) {
    static const char* kinds[5] = {"vo", "vvoo", "oovv", "ooov", "ovvv"};
    std::array<uint32_t, 5> key{(uint32_t) tensor, 0, 0, 0, 0};
    for(size_t i = 0; i < bid.size(); i++) key[1 + i] = bid[i];
    int64_t lo[4] = {0, 0, 0, 0}, ext[4] = {1, 1, 1, 1};
    n = 1;
    for(size_t i = 0; i < bid.size(); i++) {
      const bool   occ  = kinds[tensor][i] == 'o';
      const size_t tile = occ ? bid[i] : bid[i] + s.noab;
      lo[i]  = (int64_t) (s.k_offset[tile] - (occ ? 0 : s.Ot));
      ext[i] = (int64_t) s.k_range[tile];
      n *= (size_t) ext[i];
    }
    auto it = blocks.find(key);
    if(it == blocks.end()) {
      std::unique_ptr<double[]> buf(new double[n]);
      if(ccsdt_synth_block(device, seed, tensor, orb[0], orb[1], orb[2], orb[3], lo, ext, buf.get()))
        throw std::runtime_error("ccsdt_synth_block failed");
      bytes += n * 8;
      misses++;
      it = blocks.emplace(key, std::move(buf)).first;
    }
    return it->second.get();
  }
};

Tensor<double> table_tensor(BlockTable* bt, int tensor) {
  Tensor<double> t([bt, tensor](const IndexVector& bid, std::vector<double>& buf) {
    size_t        n   = 0;
    const double* src = bt->get(tensor, bid, n);
    if(buf.size() < n) buf.resize(n);
    std::memcpy(buf.data(), src, n * 8);
  });
  t.set_span_fetch([bt, tensor](const IndexVector& bid, double* out, size_t) {
    size_t        n   = 0;
    const double* src = bt->get(tensor, bid, n);
    std::memcpy(out, src, n * 8);
  });
  return t;
}

__attribute__((visibility("default"))) void*
adapter_table_open(int noa, int nob, int nva, int nvb, const int64_t* k_range, const int32_t* k_spin, const double* evl,
                   uint64_t seed, int device) {
  auto* bt   = new BlockTable;
  bt->s      = make_space(noa + nob, nva + nvb, k_range, k_spin);
  bt->orb    = {0, 0, 0, 0};
  for(int i = 0; i < noa + nob + nva + nvb; i++) {
    const int g = (i < noa + nob ? 0 : 2) + (k_spin[i] == 2 ? 1 : 0);
    bt->orb[g] += (int) k_range[i];
  }
  bt->seed   = seed;
  bt->device = device;
  bt->evl.assign(evl, evl + bt->s.Ot + bt->s.Vt);
  (void) noa; (void) nob; (void) nva; (void) nvb;
  return bt;
}
__attribute__((visibility("default"))) void adapter_table_close(void* h) { delete static_cast<BlockTable*>(h); }
__attribute__((visibility("default"))) void adapter_table_info(void* h, int64_t* out3) {
  auto* bt = static_cast<BlockTable*>(h);
  out3[0] = (int64_t) bt->bytes, out3[1] = (int64_t) bt->blocks.size(), out3[2] = (int64_t) bt->misses;
}

// CCSD_T_Fused_Driver<double>::execute on block-table tensors; tile counts per spin in counts[4]
__attribute__((visibility("default"))) int
adapter_table_execute(void* h, const int32_t counts[4], int is_restricted, int tilesize, const int64_t* subset, int64_t nsubset,
                      const ccsdt_options* opt, double* out, ccsdt_stats* stats_out, int64_t* gets) {
  try {
    auto*             bt = static_cast<BlockTable*>(h);
    std::vector<Tile> tiles(bt->s.k_range.begin(), bt->s.k_range.end());
    TiledIndexSpace   MO(tiles, counts[0], counts[1], counts[2], counts[3]);
    ExecutionContext  ec;
    ChemEnv           chem_env;
    chem_env.ioptions.ccsd_options.ccsdt_tilesize = tilesize;
    Tensor<double>                          d_t1 = table_tensor(bt, CCSDT_T1), d_t2 = table_tensor(bt, CCSDT_T2);
    exachem::cholesky_2e::V2Tensors<double> d_v2;
    d_v2.v2ijab = table_tensor(bt, CCSDT_V_IJAB);
    d_v2.v2ijka = table_tensor(bt, CCSDT_V_IJKA);
    d_v2.v2iabc = table_tensor(bt, CCSDT_V_IABC);
    LRUCache<Index, std::vector<double>> c1{8}, c2{8}, c3{8}, c4{8}, c5{8}, c6{8};
    CCSD_T_Fused_Driver<double>          drv;
    if(subset && nsubset > 0) drv.task_subset.assign(subset, subset + nsubset);
    if(opt) drv.have_options = true, drv.options = *opt;
    auto [e1, e2, tw, tt] = drv.execute(chem_env, ec, bt->s.k_spin, MO, d_t1, d_t2, d_v2, bt->evl, 0.0, is_restricted != 0,
                                        c1, c2, c3, c4, c5, c6, true);
    out[0] = ec.pg().allreduce_sum(e1), out[1] = ec.pg().allreduce_sum(e2), out[2] = tw, out[3] = tt;
    if(stats_out) *stats_out = drv.last_stats;
    if(gets) {
      gets[0] = (int64_t) d_t1.num_gets, gets[1] = (int64_t) d_t2.num_gets;
      gets[2] = (int64_t) d_v2.v2ijab.num_gets, gets[3] = (int64_t) d_v2.v2ijka.num_gets;
      gets[4] = (int64_t) d_v2.v2iabc.num_gets;
    }
    return 0;
  } catch(const std::exception& e) {
    g_error = e.what();
    return 1;
  }
}

// host-only self-test of the multi-rank plumbing the drop-in header relies on (no GPU): the ranks are separate
// processes (TAMM_SHIM_RANK / _SIZE / _KEY).  They meet in ec.pg() (barrier, broadcast), share the library's task counter
// (ccsdt_task_counter_open) exactly as CCSD_T_Fused_Driver::execute does, and claim `n` tickets with atomic fetch-adds.
// out[0] = tickets this rank claimed, out[1] = their sum, out[2] = sum over ranks of out[0], out[3] = of out[1],
// out[4] = the value rank 0 broadcast.
__attribute__((visibility("default"))) int adapter_multirank_selftest(int64_t n, const char* counter_name, int64_t* out) {
  try {
    ExecutionContext ec;
    const int        rank = (int) ec.pg().rank().value();
    int64_t*         counter = nullptr;
    if(rank == 0 && ccsdt_task_counter_open(counter_name, 1, &counter)) throw std::runtime_error("counter create failed");
    ec.pg().barrier();
    if(rank != 0 && ccsdt_task_counter_open(counter_name, 0, &counter)) throw std::runtime_error("counter attach failed");
    int64_t mine = 0, sum = 0;
    for(;;) {
      const int64_t k = __atomic_fetch_add(counter, (int64_t) 1, __ATOMIC_RELAXED);
      if(k >= n) break;
      mine++, sum += k;
    }
    int64_t magic = rank == 0 ? 0x5eed1234 : 0;
    ec.pg().broadcast(&magic, 1, 0);
    out[0] = mine, out[1] = sum, out[4] = magic;
    out[2] = (int64_t) ec.pg().allreduce_sum((double) mine);
    out[3] = (int64_t) ec.pg().allreduce_sum((double) sum);
    ec.pg().barrier();
    ccsdt_task_counter_close(counter, counter_name, rank == 0);
    return 0;
  } catch(const std::exception& e) {
    g_error = e.what();
    return 1;
  }
}

// host-only: just the op counter through the adapter (usable without a GPU)
__attribute__((visibility("default"))) int
adapter_ccsdt_count_ops(int noa, int nob, int nva, int nvb, const int64_t* k_range, const int32_t* k_spin,
                        int is_restricted, long double* total_num_ops) {
  try {
    Space               s = make_space(noa + nob, nva + nvb, k_range, k_spin);
    std::vector<Tile>   tiles(s.k_range.begin(), s.k_range.end());
    TiledIndexSpace     MO(tiles, noa, nob, nva, nvb);
    ExecutionContext    ec;
    ChemEnv             chem_env;
    std::vector<double> k_evl(s.Ot + s.Vt, 0.0);
    CCSD_T_Fused_Driver<double> drv;
    drv.calculate_performance_ops(chem_env, ec, s.k_spin, MO, k_evl, 0.0, is_restricted != 0, *total_num_ops, true);
    return 0;
  } catch(const std::exception& e) {
    g_error = e.what();
    return 1;
  }
}
}
