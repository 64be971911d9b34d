// TEST INFRASTRUCTURE ONLY -- stand-in for TAMM's GPU runtime wrappers, so that the UNMODIFIED
// reference GPU task function and kernels
//   /root/reference/exachem/cc/ccsd_t/ccsd_t_all_fused.hpp:77-286          (staging, H2D, launch, D2H)
//   /root/reference/exachem/cc/ccsd_t/ccsd_t_all_fused_gpu.cu              (K1: sm_80 DMMA kernel)
//   /root/reference/exachem/cc/ccsd_t/ccsd_t_all_fused_nontcCuda_Hip_Sycl.cpp (K2: FMA kernel)
// compile with nvcc for sm_100a and run on the B200 as the on-box GPU comparator.
// Nothing in here is reference code: each item re-implements the *interface* named at the cited call
// site with plain CUDA runtime calls.
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <utility>
#include <vector>

namespace tamm {

// ccsd_t_all_fused_gpu.cu:2591,2647 use `stream.first` only
using gpuStream_t = std::pair<cudaStream_t, int>;
using gpuEvent_t  = cudaEvent_t;

#define SHIM_CUDA_OK(x)                                                                            \
  do {                                                                                             \
    cudaError_t e__ = (x);                                                                         \
    if(e__ != cudaSuccess) {                                                                       \
      std::fprintf(stderr, "shim gpu_streams: %s failed: %s\n", #x, cudaGetErrorString(e__));      \
      std::exit(100);                                                                              \
    }                                                                                              \
  } while(0)

// tamm::GPUStreamPool::getInstance().getStream()  (ccsd_t_all_fused.hpp:108): one stream per rank
class GPUStreamPool {
public:
  static GPUStreamPool& getInstance() {
    static GPUStreamPool p;
    return p;
  }
  gpuStream_t& getStream() {
    if(!s_.first) SHIM_CUDA_OK(cudaStreamCreateWithFlags(&s_.first, cudaStreamNonBlocking));
    return s_;
  }

private:
  gpuStream_t s_{nullptr, 0};
};

// RMMMemoryManager::getInstance().getDeviceMemoryPool().{allocate,deallocate}
// (ccsd_t_fused_driver.hpp:218-224,247-249,506-512; ccsd_t_all_fused.hpp:180-186,279-284).
// A size-bucketed free list over cudaMalloc: like RMM's pool, repeated allocate/deallocate of the same
// sizes (six orbital-energy slices per task) costs no driver call after the first task.
class DeviceMemoryPool {
public:
  void* allocate(size_t bytes) {
    if(bytes == 0) bytes = 8;
    auto it = free_.find(bytes);
    if(it != free_.end() && !it->second.empty()) {
      void* p = it->second.back();
      it->second.pop_back();
      return p;
    }
    void* p = nullptr;
    SHIM_CUDA_OK(cudaMalloc(&p, bytes));
    return p;
  }
  void deallocate(void* p, size_t bytes) {
    if(bytes == 0) bytes = 8;
    free_[bytes].push_back(p);
  }
  // frees everything parked in the free lists (harness calls it between runs)
  void release() {
    for(auto& kv: free_)
      for(void* p: kv.second) cudaFree(p);
    free_.clear();
  }

private:
  std::map<size_t, std::vector<void*>> free_;
};
class RMMMemoryManager {
public:
  static RMMMemoryManager& getInstance() {
    static RMMMemoryManager m;
    return m;
  }
  DeviceMemoryPool& getDeviceMemoryPool() { return pool_; }

private:
  DeviceMemoryPool pool_;
};

// tamm::getPinnedMem / freePinnedMem (ccsd_t_fused_driver.hpp:226-234,514-520)
inline void* getPinnedMem(size_t bytes) {
  void* p = nullptr;
  SHIM_CUDA_OK(cudaMallocHost(&p, bytes ? bytes : 8));
  return p;
}
inline void freePinnedMem(void* p) { SHIM_CUDA_OK(cudaFreeHost(p)); }

// gpuMemcpyAsync<T>(dst, src, count, kind, stream) (ccsd_t_all_fused.hpp:188-214,261): count in elements
constexpr cudaMemcpyKind gpuMemcpyHostToDevice = cudaMemcpyHostToDevice;
constexpr cudaMemcpyKind gpuMemcpyDeviceToHost = cudaMemcpyDeviceToHost;
template<typename T>
inline void gpuMemcpyAsync(T* dst, const T* src, size_t count, cudaMemcpyKind kind, gpuStream_t& stream) {
  if(count == 0) return;
  SHIM_CUDA_OK(cudaMemcpyAsync(dst, src, count * sizeof(T), kind, stream.first));
}

// gpuEventQuery / gpuEventSynchronize / gpuDeviceSynchronize (ccsd_t_all_fused.hpp:141,180;
// ccsd_t_fused_driver.hpp:489): query is true when the event has completed (or was never recorded)
inline bool gpuEventQuery(gpuEvent_t e) { return cudaEventQuery(e) == cudaSuccess; }
inline void gpuEventSynchronize(gpuEvent_t e) { SHIM_CUDA_OK(cudaEventSynchronize(e)); }
inline void gpuDeviceSynchronize() { SHIM_CUDA_OK(cudaDeviceSynchronize()); }

} // namespace tamm
