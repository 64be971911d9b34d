// The one collective of the (T) path and the process-shared task counter, inside the C ABI.
//
//   ccsdt_comm_*          ncclAllReduce(sum) of {E[T], E(T)} over all ranks on the context's GPU: replaces the two
//                         ec.pg().reduce calls of exachem/cc/ccsd_t/ccsd_t.cpp:262-263.  NCCL is loaded at run time
//                         (libnccl.so.2) so that the library still loads on a box without it (CPU-side tests).
//   ccsdt_task_counter_*  an int64 in POSIX shared memory for ccsdt_set_task_counter: the role of AtomicCounterGA
//                         (exachem/cc/ccsd_t/ccsd_t_fused_driver.hpp:169-172,456,541) among the ranks of one node.
#include "ccsdt_ctx.hpp"

#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstring>

namespace {

// the slice of the NCCL API this file uses (nccl.h is not required at build time)
typedef struct ncclComm* ncclComm_t;
struct NcclUniqueId {
  char internal[128];
};
enum { kNcclFloat64 = 8, kNcclSum = 0 };

struct Nccl {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string open() {
    if(lib) return "";
    // a process that already loaded an NCCL (torch bundles one) gets that copy back
    for(const char* name: {"libnccl.so.2", "libnccl.so"}) {
      lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if(lib) break;
    }
    if(!lib) return std::string("cannot load libnccl.so.2: ") + dlerror();
    GetUniqueId    = (decltype(GetUniqueId)) dlsym(lib, "ncclGetUniqueId");
    CommInitRank   = (decltype(CommInitRank)) dlsym(lib, "ncclCommInitRank");
    AllReduce      = (decltype(AllReduce)) dlsym(lib, "ncclAllReduce");
    CommDestroy    = (decltype(CommDestroy)) dlsym(lib, "ncclCommDestroy");
    GetErrorString = (decltype(GetErrorString)) dlsym(lib, "ncclGetErrorString");
    if(!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy) {
      lib = nullptr;
      return "libnccl lacks ncclGetUniqueId / ncclCommInitRank / ncclAllReduce / ncclCommDestroy";
    }
    return "";
  }
  std::string why(int rc) const { return GetErrorString ? GetErrorString(rc) : ("nccl error " + std::to_string(rc)); }
};
Nccl        g_nccl;
std::string g_comm_error;

} // namespace

namespace ccsdt {
void comm_destroy(ccsdt_ctx* ctx) {
  if(ctx->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t) ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
  if(ctx->d_allreduce) cudaFree(ctx->d_allreduce);
  ctx->d_allreduce = nullptr;
}
} // namespace ccsdt

extern "C" {

int ccsdt_comm_unique_id(void* id128) {
  if(!id128) return 1;
  const std::string e = g_nccl.open();
  if(!e.empty()) return 10;
  NcclUniqueId id;
  const int    rc = g_nccl.GetUniqueId(&id);
  if(rc) return 10;
  memcpy(id128, id.internal, 128);
  return 0;
}

int ccsdt_comm_init(ccsdt_ctx* ctx, const void* id128, int rank, int nranks) {
  if(!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return 1;
  const std::string e = g_nccl.open();
  if(!e.empty()) return ctx->fail(e, 10);
  cudaSetDevice(ctx->device);
  ccsdt::comm_destroy(ctx);
  NcclUniqueId id;
  memcpy(id.internal, id128, 128);
  ncclComm_t comm = nullptr;
  const int  rc   = g_nccl.CommInitRank(&comm, nranks, id, rank);
  if(rc) return ctx->fail("ncclCommInitRank: " + g_nccl.why(rc), 10);
  ctx->nccl_comm = comm;
  CK(cudaMalloc(&ctx->d_allreduce, 16));
  return 0;
}

int ccsdt_comm_allreduce(ccsdt_ctx* ctx, double energies[2]) {
  if(!ctx || !energies) return 1;
  if(!ctx->nccl_comm) return ctx->fail("ccsdt_comm_init has not been called", 10);
  cudaSetDevice(ctx->device);
  CK(cudaMemcpyAsync(ctx->d_allreduce, energies, 16, cudaMemcpyHostToDevice, ctx->s_compute));
  const int rc = g_nccl.AllReduce(ctx->d_allreduce, ctx->d_allreduce, 2, kNcclFloat64, kNcclSum, (ncclComm_t) ctx->nccl_comm,
                                  ctx->s_compute);
  if(rc) return ctx->fail("ncclAllReduce: " + g_nccl.why(rc), 10);
  CK(cudaMemcpyAsync(energies, ctx->d_allreduce, 16, cudaMemcpyDeviceToHost, ctx->s_compute));
  CK(cudaStreamSynchronize(ctx->s_compute));
  return 0;
}

int ccsdt_comm_destroy(ccsdt_ctx* ctx) {
  if(!ctx) return 1;
  cudaSetDevice(ctx->device);
  ccsdt::comm_destroy(ctx);
  return 0;
}

int ccsdt_task_counter_open(const char* name, int create, int64_t** counter) {
  if(!name || !counter) return 1;
  *counter = nullptr;
  if(create) shm_unlink(name); // a stale segment of a crashed run
  const int fd = shm_open(name, create ? (O_CREAT | O_EXCL | O_RDWR) : O_RDWR, 0600);
  if(fd < 0) return 2;
  if(create && ftruncate(fd, 64) != 0) {
    close(fd);
    shm_unlink(name);
    return 2;
  }
  void* p = mmap(nullptr, 64, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if(p == MAP_FAILED) return 2;
  *counter = (int64_t*) p;
  if(create) __atomic_store_n(*counter, (int64_t) 0, __ATOMIC_SEQ_CST);
  return 0;
}

int ccsdt_task_counter_close(int64_t* counter, const char* name, int unlink_it) {
  if(counter) munmap(counter, 64);
  if(unlink_it && name) shm_unlink(name);
  return 0;
}

} // extern "C"
