// TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// extern "C" harness around the UNMODIFIED reference (T) path, compiled where the sources lie
// under /root/reference against oracle/shim/tamm/tamm.hpp:
//   CCSD_T_Fused_Driver<double>::execute            exachem/cc/ccsd_t/ccsd_t_fused_driver.hpp:91
//   CCSD_T_Fused_Driver<double>::calculate_performance_ops                               :548
//   ccsd_t_data_{s1,d1,d2}_info_only                ccsd_t_all_fused_{singles,doubles1,doubles2}.hpp
//   total_fused_ccsd_t_cpu                          ccsd_t_all_fused_cpu.hpp:21
// Output goes to oracle/_ref/libccsdt_ref.so (git-ignored).  It is used (a) to pin the C restatement
// in oracle/ccsdt_oracle.c, (b) to generate tests/golden/*.json, (c) as the "reference" CPU baseline.
//
// Tensors come in as dense row-major spin-orbital arrays over the tiled ranges
//   T1[Vt][Ot]  T2[Vt][Vt][Ot][Ot]  v2ijab[Ot][Ot][Vt][Vt]  v2ijka[Ot][Ot][Ot][Vt]  v2iabc[Ot][Vt][Vt][Vt]
// (index order of the TAMM tensors declared at exachem/cc/ccsd_t/ccsd_t.cpp:66-69).
#include "tamm/tamm.hpp"

#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif

// the eight globals of exachem/cc/ccsd_t/ccsd_t.hpp:8-15 (that header itself drags in CCSD/Cholesky)
double ccsdt_s1_t1_GetTime  = 0;
double ccsdt_s1_v2_GetTime  = 0;
double ccsdt_d1_t2_GetTime  = 0;
double ccsdt_d1_v2_GetTime  = 0;
double ccsdt_d2_t2_GetTime  = 0;
double ccsdt_d2_v2_GetTime  = 0;
double genTime              = 0;
double ccsd_t_data_per_rank = 0;

#ifndef USE_CUDA
// 1) the reference CPU task function under its own name
#include "exachem/cc/ccsd_t/ccsd_t_all_fused_cpu.hpp"

// 2) a tracing interposer: `execute` (included next) calls total_fused_ccsd_t_cpu<T>(...); the macro
//    below re-routes that call through a recorder that logs the task tuple in call order and then
//    forwards to the real reference function.  The reference sources are not modified.
struct TraceRec {
  int64_t h1, h2, h3, p4, p5, p6, taskid;
  double  factor, e1_after, e2_after;
};
static std::vector<TraceRec> g_trace;
static int                   g_trace_skip_compute = 0;
static int64_t               g_trace_limit        = -1; // stop computing after this many kernel tasks

template<typename T, typename... Rest>
void traced_total_fused_ccsd_t_cpu(
  bool is_restricted, const Index noab, const Index nvab, int64_t rank, std::vector<int>& k_spin,
  std::vector<size_t>& k_range, std::vector<size_t>& k_offset, Tensor<T>& d_t1, Tensor<T>& d_t2,
  exachem::cholesky_2e::V2Tensors<T>& d_v2, std::vector<T>& k_evl_sorted, T* a0, T* a1, T* a2, T* a3,
  T* a4, T* a5, T* host_energies, int* d1h7, int* d2p7, int* s1s, int* d1s, int* d2s, int* s1e,
  int* d1e, int* d2e, size_t t_h1b, size_t t_h2b, size_t t_h3b, size_t t_p4b, size_t t_p5b,
  size_t t_p6b, double factor, size_t taskid, size_t m1, size_t m2, size_t z0, size_t z1, size_t z2,
  size_t z3, size_t z4, size_t z5, std::vector<double>& energy_l, Rest&... caches) {
  bool compute = !g_trace_skip_compute &&
                 (g_trace_limit < 0 || (int64_t) g_trace.size() < g_trace_limit);
  if(compute) {
    total_fused_ccsd_t_cpu<T>(is_restricted, noab, nvab, rank, k_spin, k_range, k_offset, d_t1, d_t2,
                              d_v2, k_evl_sorted, a0, a1, a2, a3, a4, a5, host_energies, d1h7, d2p7,
                              s1s, d1s, d2s, s1e, d1e, d2e, t_h1b, t_h2b, t_h3b, t_p4b, t_p5b, t_p6b,
                              factor, taskid, m1, m2, z0, z1, z2, z3, z4, z5, energy_l, caches...);
  }
  g_trace.push_back({(int64_t) t_h1b, (int64_t) t_h2b, (int64_t) t_h3b, (int64_t) t_p4b,
                     (int64_t) t_p5b, (int64_t) t_p6b, (int64_t) taskid, factor, energy_l[0],
                     energy_l[1]});
}
#define total_fused_ccsd_t_cpu traced_total_fused_ccsd_t_cpu
#include "exachem/cc/ccsd_t/ccsd_t_fused_driver.hpp"
#undef total_fused_ccsd_t_cpu

#else // USE_CUDA: GPU comparator build (nvcc -x cu -DUSE_CUDA [-DUSE_NV_TC]) ------------------------------
// Tracing interposer around the reference's kernel LAUNCHER (the kernel, its constant-memory uploads and
// its launch configuration stay the reference's own): the macro renames the launcher that
// ccsd_t_all_fused.hpp declares and calls; the renamed template is defined here, brackets the real
// launcher with CUDA events on the reference's stream, reads the per-block partial energies back and
// records (extents, blocks, kernel ms, unscaled task energies).
struct TraceRec {
  int64_t h1, h2, h3, p4, p5, p6, nblocks;
  double  kernel_ms, e1_task, e2_task;
};
static std::vector<TraceRec> g_trace;
static int                   g_trace_skip_compute = 0;
static int64_t               g_trace_limit        = -1;
struct StopExecute {}; // thrown out of `execute` once task_limit kernel tasks ran (sampling big shapes)

#include "exachem/cc/ccsd_t/ccsd_t_common.hpp"
#if defined(USE_NV_TC)
// the real launcher, explicit instantiation at ccsd_t_all_fused_gpu.cu:2682
template<typename T>
void ccsd_t_fully_fused_nvidia_tc_fp64(gpuStream_t&, size_t, size_t, size_t, size_t, size_t, size_t, size_t, T*, T*,
                                       T*, T*, T*, T*, int*, int*, int*, int*, int*, size_t, size_t, size_t,
                                       size_t, size_t, size_t, size_t, size_t, T*, T*, T*, T*, T*, T*, T*,
                                       event_ptr_t);
#define REAL_LAUNCHER ccsd_t_fully_fused_nvidia_tc_fp64
#else
// the real launcher, explicit instantiation at ccsd_t_all_fused_nontcCuda_Hip_Sycl.cpp:2922
template<typename T>
void fully_fused_ccsd_t_gpu(gpuStream_t&, size_t, size_t, size_t, size_t, size_t, size_t, size_t, T*, T*, T*, T*,
                            T*, T*, int*, int*, int*, int*, int*, int*, size_t, size_t, size_t, size_t, size_t,
                            size_t, size_t, size_t, T*, T*, T*, T*, T*, T*, T*, event_ptr_t);
#define REAL_LAUNCHER fully_fused_ccsd_t_gpu
#endif

static void trace_before(gpuStream_t& stream, cudaEvent_t* e0, cudaEvent_t* e1) {
  if(g_trace_limit >= 0 && (int64_t) g_trace.size() >= g_trace_limit) {
    cudaStreamSynchronize(stream.first);
    throw StopExecute{};
  }
  cudaEventCreate(e0);
  cudaEventCreate(e1);
  cudaEventRecord(*e0, stream.first);
}
static void trace_after(gpuStream_t& stream, cudaEvent_t e0, cudaEvent_t e1, size_t num_blocks,
                        const double* dev_energies, size_t h1, size_t h2, size_t h3, size_t p4, size_t p5,
                        size_t p6) {
  cudaEventRecord(e1, stream.first);
  cudaError_t err = cudaEventSynchronize(e1);
  if(err != cudaSuccess) {
    std::fprintf(stderr, "reference GPU kernel failed: %s\n", cudaGetErrorString(err));
    std::exit(100);
  }
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  std::vector<double> part(2 * num_blocks);
  cudaMemcpy(part.data(), dev_energies, sizeof(double) * 2 * num_blocks, cudaMemcpyDeviceToHost);
  double a = 0, b = 0;
  for(size_t i = 0; i < num_blocks; i++) a += part[i], b += part[i + num_blocks];
  g_trace.push_back({(int64_t) h1, (int64_t) h2, (int64_t) h3, (int64_t) p4, (int64_t) p5, (int64_t) p6,
                     (int64_t) num_blocks, (double) ms, a, b});
}

#if defined(USE_NV_TC)
#define ccsd_t_fully_fused_nvidia_tc_fp64 traced_gpu_launcher
#else
#define fully_fused_ccsd_t_gpu traced_gpu_launcher
#endif
#include "exachem/cc/ccsd_t/ccsd_t_fused_driver.hpp"
#undef ccsd_t_fully_fused_nvidia_tc_fp64
#undef fully_fused_ccsd_t_gpu

// definition of the renamed launcher template that ccsd_t_all_fused.hpp (:36 / :58) declared above
#if defined(USE_NV_TC)
template<typename T>
void traced_gpu_launcher(gpuStream_t& stream, size_t numBlks, size_t size_h3, size_t size_h2, size_t size_h1,
                         size_t size_p6, size_t size_p5, size_t size_p4, T* a0, T* a1, T* a2, T* a3, T* a4, T* a5,
                         int* i0, int* i1, int* i2, int* i3, int* i4, size_t z0, size_t z1, size_t z2, size_t z3,
                         size_t z4, size_t z5, size_t z6, size_t z7, T* e0p, T* e1p, T* e2p, T* e3p, T* e4p,
                         T* e5p, T* dev_energies, event_ptr_t done_copy) {
  if(g_trace_skip_compute) return;
  cudaEvent_t e0, e1;
  trace_before(stream, &e0, &e1);
  REAL_LAUNCHER<T>(stream, numBlks, size_h3, size_h2, size_h1, size_p6, size_p5, size_p4, a0, a1, a2, a3, a4, a5,
                   i0, i1, i2, i3, i4, z0, z1, z2, z3, z4, z5, z6, z7, e0p, e1p, e2p, e3p, e4p, e5p, dev_energies,
                   done_copy);
  trace_after(stream, e0, e1, numBlks, dev_energies, size_h1, size_h2, size_h3, size_p4, size_p5, size_p6);
}
#else
template<typename T>
void traced_gpu_launcher(gpuStream_t& stream, size_t num_blocks, size_t size_h1, size_t size_h2, size_t size_h3,
                         size_t size_p4, size_t size_p5, size_t size_p6, T* a0, T* a1, T* a2, T* a3, T* a4, T* a5,
                         int* i0, int* i1, int* i2, int* i3, int* i4, int* i5, size_t z0, size_t z1, size_t z2,
                         size_t z3, size_t z4, size_t z5, size_t z6, size_t z7, T* e0p, T* e1p, T* e2p, T* e3p,
                         T* e4p, T* e5p, T* dev_energies, event_ptr_t done_copy) {
  if(g_trace_skip_compute) return;
  cudaEvent_t e0, e1;
  trace_before(stream, &e0, &e1);
  REAL_LAUNCHER<T>(stream, num_blocks, size_h1, size_h2, size_h3, size_p4, size_p5, size_p6, a0, a1, a2, a3, a4,
                   a5, i0, i1, i2, i3, i4, i5, z0, z1, z2, z3, z4, z5, z6, z7, e0p, e1p, e2p, e3p, e4p, e5p,
                   dev_energies, done_copy);
  trace_after(stream, e0, e1, num_blocks, dev_energies, size_h1, size_h2, size_h3, size_p4, size_p5, size_p6);
}
#endif
#endif // USE_CUDA

// trace record -> 10 doubles.  CPU build: h1,h2,h3,p4,p5,p6 tile ids, taskid, factor, energy_l[0..1]
// after the task.  GPU build: h1,h2,h3,p4,p5,p6 tile EXTENTS, thread blocks, kernel ms (CUDA events
// around the reference launcher), unscaled task energies (sum of the kernel's per-block partials).
#ifndef USE_CUDA
static void trace_record_out(const TraceRec& r, double* o) {
  o[0] = (double) r.h1, o[1] = (double) r.h2, o[2] = (double) r.h3, o[3] = (double) r.p4;
  o[4] = (double) r.p5, o[5] = (double) r.p6, o[6] = (double) r.taskid, o[7] = r.factor;
  o[8] = r.e1_after, o[9] = r.e2_after;
}
#define REF_TRY
#define REF_CATCH(e1, e2, tw, tt)
#else
static void trace_record_out(const TraceRec& r, double* o) {
  o[0] = (double) r.h1, o[1] = (double) r.h2, o[2] = (double) r.h3, o[3] = (double) r.p4;
  o[4] = (double) r.p5, o[5] = (double) r.p6, o[6] = (double) r.nblocks, o[7] = r.kernel_ms;
  o[8] = r.e1_task, o[9] = r.e2_task;
}
// a sampled run leaves `execute` by exception: energies are then not available (NaN), the device pool
// and pinned buffers of that call stay allocated until ref_ccsdt_gpu_release()
#define REF_TRY try
#define REF_CATCH(e1, e2, tw, tt)                                   \
  catch(const StopExecute&) {                                        \
    cudaDeviceSynchronize();                                         \
    e1 = e2 = std::nan("");                                         \
    tw = tt = 0.0;                                                   \
  }
#endif

namespace {

struct Space {
  int                 noab, nvab, noa, nva;
  std::vector<size_t> k_range, k_offset;
  std::vector<int>    k_spin;
  size_t              Ot = 0, Vt = 0;
};

Space make_space(int noa, int nob, int nva, int nvb, const int64_t* k_range, const int32_t* k_spin) {
  Space s;
  s.noab = noa + nob;
  s.nvab = nva + nvb;
  s.noa  = noa;
  s.nva  = nva;
  size_t sum = 0;
  for(int i = 0; i < s.noab + s.nvab; i++) {
    s.k_range.push_back((size_t) k_range[i]);
    s.k_offset.push_back(sum);
    sum += (size_t) k_range[i];
    s.k_spin.push_back(k_spin[i]);
    if(i < s.noab) s.Ot += (size_t) k_range[i];
    else s.Vt += (size_t) k_range[i];
  }
  return s;
}

// kinds[d] = 'o' or 'v'; a 'v' block id is relative to the virtual sub-space (tile id - noab)
Tensor<double> dense_tensor(const Space& s, const double* data, std::string kinds) {
  return Tensor<double>([&s, data, kinds](const IndexVector& bid, std::vector<double>& buf) {
    const int d = (int) kinds.size();
    size_t    ext[4], off[4], full[4], stride[4];
    for(int i = 0; i < d; i++) {
      size_t tile = kinds[i] == 'o' ? bid[i] : bid[i] + s.noab;
      ext[i]      = s.k_range[tile];
      off[i]      = s.k_offset[tile] - (kinds[i] == 'o' ? 0 : s.Ot);
      full[i]     = kinds[i] == 'o' ? s.Ot : s.Vt;
    }
    size_t st = 1;
    for(int i = d - 1; i >= 0; i--) {
      stride[i] = st;
      st *= full[i];
    }
    size_t n = 1;
    for(int i = 0; i < d; i++) n *= ext[i];
    if(buf.size() < n) buf.resize(n);
    size_t idx[4] = {0, 0, 0, 0};
    for(size_t lin = 0; lin < n; lin++) {
      size_t o = 0;
      for(int i = 0; i < d; i++) o += (off[i] + idx[i]) * stride[i];
      buf[lin] = data[o];
      for(int i = d - 1; i >= 0; i--) {
        if(++idx[i] < ext[i]) break;
        idx[i] = 0;
      }
    }
  });
}

TiledIndexSpace make_mo(const Space& s, int noa, int nob, int nva, int nvb) {
  std::vector<Tile> tiles;
  for(auto r: s.k_range) tiles.push_back((Tile) r);
  return TiledIndexSpace(tiles, noa, nob, nva, nvb);
}

struct Silence {
  std::streambuf* old;
  std::ostringstream sink;
  Silence(): old(std::cout.rdbuf(sink.rdbuf())) {}
  ~Silence() { std::cout.rdbuf(old); }
};

} // namespace

#include <sstream>

extern "C" {

// Runs the reference `execute`.  out[0]=E[T] partial (energy1), out[1]=E(T) partial (energy2),
// out[2]=work seconds, out[3]=total seconds.  If trace_out != NULL it receives up to trace_cap
// records of 10 doubles (h1,h2,h3,p4,p5,p6,taskid,factor,energy_l[0],energy_l[1] after the task),
// *n_trace the number of kernel tasks executed.  mode: 0 = full run, 1 = enumerate only (no maths),
// task_limit >= 0 computes only the first task_limit kernel tasks (CPU-baseline sampling).
int ref_ccsdt_execute(int noa, int nob, int nva, int nvb, const int64_t* k_range,
                      const int32_t* k_spin, const double* evl, const double* t1, const double* t2,
                      const double* v2ijab, const double* v2ijka, const double* v2iabc,
                      int is_restricted, int tilesize, int cache_size, int mode, int64_t task_limit,
                      double* out, double* trace_out, int64_t trace_cap, int64_t* n_trace) {
  Space            s  = make_space(noa, nob, nva, nvb, k_range, k_spin);
  TiledIndexSpace  MO = make_mo(s, noa, nob, nva, nvb);
  ExecutionContext ec;
  ChemEnv          chem_env;
  chem_env.ioptions.ccsd_options.ccsdt_tilesize = tilesize;
  chem_env.ioptions.ccsd_options.cache_size     = cache_size;

  Tensor<double>                          d_t1 = dense_tensor(s, t1, "vo");
  Tensor<double>                          d_t2 = dense_tensor(s, t2, "vvoo");
  exachem::cholesky_2e::V2Tensors<double> d_v2;
  d_v2.v2ijab = dense_tensor(s, v2ijab, "oovv");
  d_v2.v2ijka = dense_tensor(s, v2ijka, "ooov");
  d_v2.v2iabc = dense_tensor(s, v2iabc, "ovvv");

  std::vector<double> k_evl(evl, evl + s.Ot + s.Vt);
  // cache sizes as exachem/cc/ccsd_t/ccsd_t.cpp:236-241
  LRUCache<Index, std::vector<double>> cache_s1t{(size_t) cache_size};
  LRUCache<Index, std::vector<double>> cache_s1v{(size_t) cache_size};
  LRUCache<Index, std::vector<double>> cache_d1t{(size_t) cache_size * s.noab};
  LRUCache<Index, std::vector<double>> cache_d1v{(size_t) cache_size * s.noab};
  LRUCache<Index, std::vector<double>> cache_d2t{(size_t) cache_size * s.nvab};
  LRUCache<Index, std::vector<double>> cache_d2v{(size_t) cache_size * s.nvab};

  g_trace.clear();
  g_trace_skip_compute = (mode == 1);
  g_trace_limit        = task_limit;

  CCSD_T_Fused_Driver<double> drv;
  double                      e1, e2, tw, tt;
  {
    Silence quiet;
    // same call as exachem/cc/ccsd_t/ccsd_t.cpp:253-256 (seq_h3b = true)
    REF_TRY {
      std::tie(e1, e2, tw, tt) =
        drv.execute(chem_env, ec, s.k_spin, MO, d_t1, d_t2, d_v2, k_evl, 0.0, is_restricted != 0,
                    cache_s1t, cache_s1v, cache_d1t, cache_d1v, cache_d2t, cache_d2v, true);
    }
    REF_CATCH(e1, e2, tw, tt)
  }
  out[0] = e1;
  out[1] = e2;
  out[2] = tw;
  out[3] = tt;
  if(n_trace) *n_trace = (int64_t) g_trace.size();
  if(trace_out) {
    for(int64_t i = 0; i < (int64_t) g_trace.size() && i < trace_cap; i++) {
      const TraceRec& r    = g_trace[i];
      double*         o    = trace_out + 10 * i;
      trace_record_out(r, o);
    }
  }
  return 0;
}

// total_num_ops exactly as the reference counts it (ccsd_t.cpp:298-301 passes seq_h3b = true).
int ref_ccsdt_count_ops(int noa, int nob, int nva, int nvb, const int64_t* k_range,
                        const int32_t* k_spin, int is_restricted, long double* total_num_ops) {
  Space               s  = make_space(noa, nob, nva, nvb, k_range, k_spin);
  TiledIndexSpace     MO = make_mo(s, noa, nob, nva, nvb);
  ExecutionContext    ec;
  ChemEnv             chem_env;
  std::vector<double> k_evl(s.Ot + s.Vt, 0.0);
  CCSD_T_Fused_Driver<double> drv;
  long double                 ops = 0;
  drv.calculate_performance_ops(chem_env, ec, s.k_spin, MO, k_evl, 0.0, is_restricted != 0, ops,
                                true);
  *total_num_ops = ops;
  return 0;
}

// exec tables of one task from the reference's own *_info_only functions:
// s1_exec[9], d1_exec[9*noab], d2_exec[9*nvab]; counts[3] = number of enabled (candidate, k-tile)
int ref_ccsdt_task_info(int noa, int nob, int nva, int nvb, const int64_t* k_range,
                        const int32_t* k_spin, int is_restricted, int h1b, int h2b, int h3b, int p4b,
                        int p5b, int p6b, int* s1_exec, int* d1_exec, int* d2_exec, int* counts) {
  Space               s = make_space(noa, nob, nva, nvb, k_range, k_spin);
  std::vector<double> k_evl(s.Ot + s.Vt, 0.0);
  std::vector<int>    s1_size(6), d1_size(7 * s.noab), d2_size(7 * s.nvab);
  std::fill(s1_exec, s1_exec + 9, -1);
  std::fill(d1_exec, d1_exec + 9 * s.noab, -1);
  std::fill(d2_exec, d2_exec + 9 * s.nvab, -1);
  size_t comm = 0;
  ccsd_t_data_s1_info_only<double>(is_restricted != 0, s.noab, s.nvab, s.k_spin, k_evl, s.k_range,
                                   h1b, h2b, h3b, p4b, p5b, p6b, s1_size.data(), s1_exec, &counts[0],
                                   comm);
  ccsd_t_data_d1_info_only<double>(is_restricted != 0, s.noab, s.nvab, s.k_spin, k_evl, s.k_range,
                                   h1b, h2b, h3b, p4b, p5b, p6b, d1_size.data(), d1_exec, &counts[1],
                                   comm);
  ccsd_t_data_d2_info_only<double>(is_restricted != 0, s.noab, s.nvab, s.k_spin, k_evl, s.k_range,
                                   h1b, h2b, h3b, p4b, p5b, p6b, d2_size.data(), d2_exec, &counts[2],
                                   comm);
  return 0;
}

#ifdef USE_CUDA
// 1 = DMMA kernel K1 (ccsd_t_all_fused_gpu.cu), 0 = FMA kernel K2 (ccsd_t_all_fused_nontcCuda_Hip_Sycl.cpp)
int ref_ccsdt_gpu_kernel_kind() {
#if defined(USE_NV_TC)
  return 1;
#else
  return 0;
#endif
}
void ref_ccsdt_gpu_release() {
  cudaDeviceSynchronize();
  RMMMemoryManager::getInstance().getDeviceMemoryPool().release();
}
#endif

int ref_ccsdt_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

} // extern "C"

// ---- procedural tensors (same counter-based generator as exachem_b200/synthetic.py) so that the
//      reference can be timed on shapes whose dense tensors would not fit in host memory ----------
namespace {
inline uint64_t synth_mix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
inline double synth_raw(uint64_t seed, int tensor, int64_t i0, int64_t i1, int64_t i2, int64_t i3) {
  static const double scale[5] = {0.05, 0.1, 0.1, 0.1, 0.1};
  uint64_t key = ((uint64_t) i0 << 48) | ((uint64_t) i1 << 32) | ((uint64_t) i2 << 16) | (uint64_t) i3;
  uint64_t s   = synth_mix(seed ^ ((uint64_t) (tensor + 1) * 0xD1B54A32D192ED03ull));
  uint64_t h   = synth_mix(s ^ key);
  double   u   = (double) (h >> 11) * (1.0 / 9007199254740992.0);
  return (2.0 * u - 1.0) * scale[tensor];
}
struct SynthOrb {
  uint64_t seed;
  int64_t  noa, nob, nva, nvb;
  int so(int64_t i) const { return i < noa ? 1 : 2; }
  int sv(int64_t a) const { return a < nva ? 1 : 2; }
  double value(int tensor, int64_t a, int64_t b, int64_t c, int64_t d) const {
    double sign = 1.0;
    switch(tensor) {
      case 0: return sv(a) == so(b) ? synth_raw(seed, 0, a, b, 0, 0) : 0.0;
      case 1:
        if(a == b || c == d || sv(a) + sv(b) != so(c) + so(d)) return 0.0;
        if(a > b) std::swap(a, b), sign = -sign;
        if(c > d) std::swap(c, d), sign = -sign;
        return sign * synth_raw(seed, 1, a, b, c, d);
      case 2:
        if(a == b || c == d || so(a) + so(b) != sv(c) + sv(d)) return 0.0;
        if(a > b) std::swap(a, b), sign = -sign;
        if(c > d) std::swap(c, d), sign = -sign;
        return sign * synth_raw(seed, 2, a, b, c, d);
      case 3:
        if(a == b || so(a) + so(b) != so(c) + sv(d)) return 0.0;
        if(a > b) std::swap(a, b), sign = -sign;
        return sign * synth_raw(seed, 3, a, b, c, d);
      default:
        if(c == d || so(a) + sv(b) != sv(c) + sv(d)) return 0.0;
        if(c > d) std::swap(c, d), sign = -sign;
        return sign * synth_raw(seed, 4, a, b, c, d);
    }
  }
};

Tensor<double> synth_tensor(const Space& s, const SynthOrb& orb, int tensor, std::string kinds) {
  return Tensor<double>([&s, orb, tensor, kinds](const IndexVector& bid, std::vector<double>& buf) {
    const int d = (int) kinds.size();
    int64_t   ext[4] = {1, 1, 1, 1}, off[4] = {0, 0, 0, 0};
    for(int i = 0; i < d; i++) {
      size_t tile = kinds[i] == 'o' ? bid[i] : bid[i] + s.noab;
      ext[i]      = (int64_t) s.k_range[tile];
      off[i]      = (int64_t) (s.k_offset[tile] - (kinds[i] == 'o' ? 0 : s.Ot));
    }
    size_t n = (size_t) (ext[0] * ext[1] * ext[2] * ext[3]);
    if(buf.size() < n) buf.resize(n);
    size_t lin = 0;
    for(int64_t a = 0; a < ext[0]; a++)
      for(int64_t b = 0; b < ext[1]; b++)
        for(int64_t c = 0; c < ext[2]; c++)
          for(int64_t e = 0; e < ext[3]; e++)
            buf[lin++] = orb.value(tensor, off[0] + a, off[1] + b, d > 2 ? off[2] + c : 0, d > 3 ? off[3] + e : 0);
  });
}
} // namespace

extern "C" {
// Same as ref_ccsdt_execute but on procedural tensors: n_orb = orbital counts (occ a, occ b, virt a,
// virt b) the generator needs for the spin of an index.  Used by bench.py for the CPU baseline.
int ref_ccsdt_execute_synth(int noa, int nob, int nva, int nvb, const int64_t* k_range,
                            const int32_t* k_spin, const double* evl, const int64_t* n_orb, uint64_t seed,
                            int is_restricted, int tilesize, int cache_size, int64_t task_limit,
                            double* out, int64_t* n_trace, double* trace_out, int64_t trace_cap) {
  Space            s  = make_space(noa, nob, nva, nvb, k_range, k_spin);
  TiledIndexSpace  MO = make_mo(s, noa, nob, nva, nvb);
  ExecutionContext ec;
  ChemEnv          chem_env;
  chem_env.ioptions.ccsd_options.ccsdt_tilesize = tilesize;
  SynthOrb orb{seed, n_orb[0], n_orb[1], n_orb[2], n_orb[3]};
  Tensor<double>                          d_t1 = synth_tensor(s, orb, 0, "vo");
  Tensor<double>                          d_t2 = synth_tensor(s, orb, 1, "vvoo");
  exachem::cholesky_2e::V2Tensors<double> d_v2;
  d_v2.v2ijab = synth_tensor(s, orb, 2, "oovv");
  d_v2.v2ijka = synth_tensor(s, orb, 3, "ooov");
  d_v2.v2iabc = synth_tensor(s, orb, 4, "ovvv");
  std::vector<double>                  k_evl(evl, evl + s.Ot + s.Vt);
  LRUCache<Index, std::vector<double>> cache_s1t{(size_t) cache_size}, cache_s1v{(size_t) cache_size};
  LRUCache<Index, std::vector<double>> cache_d1t{(size_t) cache_size * s.noab}, cache_d1v{(size_t) cache_size * s.noab};
  LRUCache<Index, std::vector<double>> cache_d2t{(size_t) cache_size * s.nvab}, cache_d2v{(size_t) cache_size * s.nvab};
  g_trace.clear();
  g_trace_skip_compute = 0;
  g_trace_limit        = task_limit;
  CCSD_T_Fused_Driver<double> drv;
  double                      e1, e2, tw, tt;
  {
    Silence quiet;
    REF_TRY {
      std::tie(e1, e2, tw, tt) = drv.execute(chem_env, ec, s.k_spin, MO, d_t1, d_t2, d_v2, k_evl, 0.0,
                                             is_restricted != 0, cache_s1t, cache_s1v, cache_d1t, cache_d1v,
                                             cache_d2t, cache_d2v, true);
    }
    REF_CATCH(e1, e2, tw, tt)
  }
  out[0] = e1, out[1] = e2, out[2] = tw, out[3] = tt;
  if(n_trace) *n_trace = (int64_t) g_trace.size();
  if(trace_out)
    for(int64_t i = 0; i < (int64_t) g_trace.size() && i < trace_cap; i++) trace_record_out(g_trace[i], trace_out + 10 * i);
  return 0;
}
}
