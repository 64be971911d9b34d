// TEST INFRASTRUCTURE ONLY -- a minimal stand-in for the (un-vendored) TAMM umbrella header.
//
// Purpose: let the UNMODIFIED reference (T) path
//   /root/reference/exachem/cc/ccsd_t/ccsd_t_fused_driver.hpp (+ the headers it includes)
// compile with plain g++ so that oracle/_ref/libccsdt_ref.so runs the reference's own
// `CCSD_T_Fused_Driver<double>::execute`, staging (`ccsd_t_data_{s1,d1,d2}_new`), CPU kernel
// (`total_fused_ccsd_t_cpu`) and op counter.  Nothing in here is reference code: every type is a
// small original re-implementation of the TAMM/HPTT *interface* those headers use, inferred from
// the call sites (cited inline).  TAMM/HPTT only move and permute data on this path; all (T)
// floating-point arithmetic stays in the reference sources.
#pragma once

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <iomanip>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include <cstring>
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

namespace tamm {

using Index       = uint32_t;
using IndexVector = std::vector<Index>;
using Tile        = uint32_t;

// printed at ccsd_t_fused_driver.hpp:133 ("MO Tiles = ")
template<typename T>
std::ostream& operator<<(std::ostream& os, const std::vector<T>& v) {
  os << "[";
  for(size_t i = 0; i < v.size(); i++) os << (i ? "," : "") << v[i];
  return os << "]";
}

// RAII wall-clock accumulator (ccsd_t_all_fused_singles.hpp:198 etc.)
class TimerGuard {
public:
  explicit TimerGuard(double* acc): acc_(acc), t0_(std::chrono::high_resolution_clock::now()) {}
  ~TimerGuard() {
    auto t1 = std::chrono::high_resolution_clock::now();
    *acc_ += std::chrono::duration_cast<std::chrono::duration<double>>(t1 - t0_).count();
  }

private:
  double*                                        acc_;
  std::chrono::high_resolution_clock::time_point t0_;
};

// minimal span (pointer + count), for the get(bid, span) overload below
template<typename T>
class span {
public:
  span(T* p, size_t n): p_(p), n_(n) {}
  T*     data() const { return p_; }
  size_t size() const { return n_; }

private:
  T*     p_;
  size_t n_;
};

// Dense block provider: get(block id, buffer) delivers one row-major block
// (ccsd_t_all_fused_doubles1.hpp:222 `d_t2.get({p4b - noab, p5b - noab, h7b, h1b}, k_a)`).
template<typename T>
class Tensor {
public:
  using Fetch = std::function<void(const IndexVector&, std::vector<T>&)>;
  Tensor()    = default;
  explicit Tensor(Fetch f): fetch_(std::move(f)) {}
  void get(const IndexVector& bid, std::vector<T>& buf) const {
    ++num_gets;
    fetch_(bid, buf);
  }
  // span flavour (used by the adapter test when CCSDT_B200_SPAN_TYPE is defined): the provider writes into foreign memory
  void get(const IndexVector& bid, span<T> out) const {
    ++num_gets;
    if(span_fetch_) span_fetch_(bid, out.data(), out.size());
    else {
      std::vector<T> tmp;
      fetch_(bid, tmp);
      std::copy(tmp.begin(), tmp.begin() + (std::ptrdiff_t) out.size(), out.data());
    }
  }
  using SpanFetch = std::function<void(const IndexVector&, T*, size_t)>;
  void set_span_fetch(SpanFetch f) { span_fetch_ = std::move(f); }
  mutable size_t num_gets = 0;

private:
  Fetch     fetch_;
  SpanFetch span_fetch_;
};

// LRU cache keyed by block-id vectors.  Call sites:
//   auto [hit, value] = cache.log_access(key);  if(hit) use value;  else { ...; value = sorted; }
// so log_access returns a reference that must stay valid until the caller assigns through it.
template<typename K, typename V>
class LRUCache {
public:
  explicit LRUCache(size_t capacity = 0): cap_(capacity) {}
  std::tuple<bool, V&> log_access(const std::vector<K>& key) {
    if(cap_ == 0) {
      scratch_ = V{};
      return {false, scratch_};
    }
    auto it = map_.find(key);
    if(it != map_.end()) {
      order_.splice(order_.begin(), order_, it->second);
      return {true, it->second->second};
    }
    if(order_.size() >= cap_) {
      map_.erase(order_.back().first);
      order_.pop_back();
    }
    order_.emplace_front(key, V{});
    map_[key] = order_.begin();
    return {false, order_.begin()->second};
  }

private:
  using Entry = std::pair<std::vector<K>, V>;
  size_t                                                       cap_;
  std::list<Entry>                                             order_;
  std::map<std::vector<K>, typename std::list<Entry>::iterator> map_;
  V                                                            scratch_;
};

// ec.pg().rank().value(), ec.pg().barrier()  (ccsd_t_fused_driver.hpp:99,536)
// Single rank by default.  With TAMM_SHIM_SIZE > 1 in the environment (and TAMM_SHIM_RANK, TAMM_SHIM_KEY) the ranks
// are separate processes of one node that meet in a POSIX shared-memory segment "/tamm_shim_<key>" (4 KiB, created
// and zeroed by the launcher before the ranks start): a sense-reversing barrier, a broadcast buffer and one slot per
// rank for a sum -- enough for the multi-rank tests of the drop-in header.
struct RankValue {
  int64_t v;
  int64_t value() const { return v; }
};
class ProcGroup {
public:
  ProcGroup() {
    const char *r = std::getenv("TAMM_SHIM_RANK"), *n = std::getenv("TAMM_SHIM_SIZE"), *k = std::getenv("TAMM_SHIM_KEY");
    if(!r || !n || std::atoi(n) <= 1) return;
    rank_ = std::atoi(r), size_ = std::atoi(n);
    const std::string name = std::string("/tamm_shim_") + (k ? k : "0");
    const int         fd   = shm_open(name.c_str(), O_RDWR, 0600);
    if(fd < 0) {
      std::cerr << "shim ProcGroup: cannot open " << name << std::endl;
      std::abort();
    }
    mem_ = (unsigned char*) mmap(nullptr, 4096, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if(mem_ == (unsigned char*) MAP_FAILED) std::abort();
  }
  RankValue rank() const { return {rank_}; }
  RankValue size() const { return {size_}; }
  void      barrier() const {
    if(size_ <= 1) return;
    int64_t *     count = (int64_t*) mem_, *gen = (int64_t*) (mem_ + 64);
    const int64_t g = __atomic_load_n(gen, __ATOMIC_ACQUIRE);
    if(__atomic_add_fetch(count, 1, __ATOMIC_ACQ_REL) == size_) {
      __atomic_store_n(count, 0, __ATOMIC_RELAXED);
      __atomic_add_fetch(gen, 1, __ATOMIC_ACQ_REL);
    }
    else
      while(__atomic_load_n(gen, __ATOMIC_ACQUIRE) == g) usleep(20);
  }
  // TAMM: pg.broadcast(buf, count, root)
  template<typename T>
  void broadcast(T* buf, size_t count, int root) const {
    if(size_ <= 1) return;
    const size_t bytes = count * sizeof(T);
    if(bytes > 2048) std::abort();
    if(rank_ == root) std::memcpy(mem_ + 1024, buf, bytes);
    barrier();
    if(rank_ != root) std::memcpy(buf, mem_ + 1024, bytes);
    barrier();
  }
  // sum over ranks, result on every rank (stands in for the caller's ec.pg().reduce of ccsd_t.cpp:262-263)
  double allreduce_sum(double v) const {
    if(size_ <= 1) return v;
    double* slots = (double*) (mem_ + 3072);
    slots[rank_]  = v;
    barrier();
    double s = 0;
    for(int64_t i = 0; i < size_; i++) s += slots[i];
    barrier();
    return s;
  }

private:
  int64_t        rank_ = 0, size_ = 1;
  unsigned char* mem_  = nullptr;
};
class ExecutionContext {
public:
  ProcGroup&       pg() { return pg_; }
  const ProcGroup& pg() const { return pg_; }
  int              nnodes() const { return 1; } // the stand-in's ranks always share one node
  int              ppn() const { return (int) pg_.size().value(); }

private:
  ProcGroup pg_;
};

// Global task counter (ccsd_t_fused_driver.hpp:169-172,456,541): single-rank stand-in.
class AtomicCounter {
public:
  virtual ~AtomicCounter()                            = default;
  virtual void    allocate(int64_t init)              = 0;
  virtual int64_t fetch_add(int64_t idx, int64_t inc) = 0;
  virtual void    deallocate()                        = 0;
};
class AtomicCounterGA: public AtomicCounter {
public:
  AtomicCounterGA(const ProcGroup&, int64_t n): v_(n, 0) {}
  void    allocate(int64_t init) override { std::fill(v_.begin(), v_.end(), init); }
  int64_t fetch_add(int64_t idx, int64_t inc) override {
    int64_t old = v_[idx];
    v_[idx] += inc;
    return old;
  }
  void deallocate() override {}

private:
  std::vector<int64_t> v_;
};

// MO("occ").num_tiles(), MO.input_tile_sizes()  (ccsd_t_fused_driver.hpp:112-118)
class TiledIndexSpace {
public:
  struct Sub {
    Index n;
    Index num_tiles() const { return n; }
  };
  TiledIndexSpace() = default;
  TiledIndexSpace(std::vector<Tile> tiles, Index noa, Index nob, Index nva, Index nvb):
    tiles_(std::move(tiles)), noa_(noa), nob_(nob), nva_(nva), nvb_(nvb) {}
  Sub operator()(const std::string& name) const {
    if(name == "occ") return {noa_ + nob_};
    if(name == "virt") return {nva_ + nvb_};
    if(name == "occ_alpha") return {noa_};
    if(name == "occ_beta") return {nob_};
    if(name == "virt_alpha") return {nva_};
    if(name == "virt_beta") return {nvb_};
    std::cerr << "shim TiledIndexSpace: unknown subspace " << name << std::endl;
    std::abort();
  }
  const std::vector<Tile>& input_tile_sizes() const { return tiles_; }

private:
  std::vector<Tile> tiles_;
  Index             noa_ = 0, nob_ = 0, nva_ = 0, nvb_ = 0;
};

} // namespace tamm

// HPTT interface used on the path: out-of-place tensor transpose B = alpha * perm(A) + beta * B,
// row-major, B's dimension j is A's dimension perm[j]
// (ccsd_t_all_fused_doubles1.hpp:229 `hptt::create_plan(perm, 4, -1.0, &k_a[0], size, NULL, 0,
//  &k_a_sort[0], NULL, hptt::ESTIMATE, 1, NULL, true)`).
namespace hptt {
enum SelectionMethod { ESTIMATE, MEASURE, PATIENT, CRAZY };

template<typename T>
class Plan {
public:
  Plan(const int* perm, int dim, T alpha, const T* A, const int* sizeA, T beta, T* B):
    perm_(perm, perm + dim), size_(sizeA, sizeA + dim), alpha_(alpha), beta_(beta), A_(A), B_(B) {}
  void execute() {
    const int           d = (int) perm_.size();
    std::vector<size_t> strideA(d), sizeB(d), strideAforB(d);
    size_t              s = 1;
    for(int i = d - 1; i >= 0; i--) {
      strideA[i] = s;
      s *= (size_t) size_[i];
    }
    size_t total = s;
    for(int j = 0; j < d; j++) {
      sizeB[j]       = (size_t) size_[perm_[j]];
      strideAforB[j] = strideA[perm_[j]];
    }
    std::vector<size_t> idx(d, 0);
    for(size_t lin = 0; lin < total; lin++) {
      size_t offA = 0;
      for(int j = 0; j < d; j++) offA += idx[j] * strideAforB[j];
      B_[lin] = alpha_ * A_[offA] + (beta_ == T(0) ? T(0) : beta_ * B_[lin]);
      for(int j = d - 1; j >= 0; j--) {
        if(++idx[j] < sizeB[j]) break;
        idx[j] = 0;
      }
    }
  }

private:
  std::vector<int> perm_, size_;
  T                alpha_, beta_;
  const T*         A_;
  T*               B_;
};

template<typename T, typename AlphaT, typename BetaT>
std::shared_ptr<Plan<T>> create_plan(const int* perm, int dim, AlphaT alpha, const T* A,
                                     const int* sizeA, const int* /*outerSizeA*/, BetaT beta, T* B,
                                     const int* /*outerSizeB*/, SelectionMethod, int /*numThreads*/,
                                     const int* /*threadIds*/, bool useRowMajor) {
  if(!useRowMajor) {
    std::cerr << "shim hptt: only row-major plans are used on the (T) path" << std::endl;
    std::abort();
  }
  return std::make_shared<Plan<T>>(perm, dim, (T) alpha, A, sizeA, (T) beta, B);
}
} // namespace hptt

#if defined(USE_CUDA)
#include "tamm/gpu_streams.hpp" // GPU comparator build only (oracle/_ref/libccsdt_refgpu_*.so)
#endif

using namespace tamm;
using std::cout;
using std::endl;

// Types the driver signature names but whose definitions live elsewhere in ExaChem
// (exachem/cholesky/v2tensors.hpp:136-141, exachem/common/chemenv.hpp,
//  exachem/common/options/input_options.hpp:258-260).
namespace exachem::cholesky_2e {
template<typename T>
struct V2Tensors {
  tamm::Tensor<T> v2ijab, v2ijka, v2iabc;
};
} // namespace exachem::cholesky_2e

struct CCSDOptions {
  int  ccsdt_tilesize = 40;
  int  cache_size     = 8;
  bool skip_ccsd      = false;
};
struct ECOptions {
  CCSDOptions ccsd_options;
};
struct ChemEnv {
  ECOptions ioptions;
};
