// Device helpers shared by the fused kernels (ccsdt_kernels.cu).
#pragma once
#include "ccsdt_device.hpp"

namespace ccsdt {

// =================================================================================================
// small PTX wrappers
// =================================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t) __cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
               "selp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok)
               : "r"(bar), "r"(parity)
               : "memory");
  return ok != 0;
}
// bounded wait: a pipeline bug must end in a trap (reported as a CUDA error), never in a hung GPU.
// The slow path is kept out of line: the wait is inlined at every pipeline step and the kernel's hot
// code has to stay small (instruction cache).
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
               "selp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok)
               : "r"(bar), "r"(parity), "r"(ns)
               : "memory");
  return ok != 0;
}
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity, uint32_t* error_flag, int tag) {
  // long suspend hint + back-off: a waiting warp (the producer, almost always) must not eat issue slots of
  // the DMMA warps that share its scheduler.  The watchdog runs on %globaltimer (wall-clock nanoseconds, so a
  // context that is time-sliced out by MPS, a second rank on the GPU or a debugger is not mistaken for a dead
  // pipeline as quickly as with SM cycles) against the limit the host wrote behind the error word
  // (options.watchdog_ms; 0 = never).
  uint64_t limit = 0, t0 = 0;
  if(error_flag) {
    limit = (uint64_t) error_flag[1] | ((uint64_t) error_flag[2] << 32);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  }
  while(!mbar_try_wait_hint(bar, parity, 20000u)) {
    __nanosleep(100);
    if(limit) {
      uint64_t t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if(t1 - t0 > limit) {
        atomicExch(error_flag, 0xDEAD0000u | (uint32_t) tag);
        __threadfence_system();
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t* error_flag, int tag) {
  if(!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity, error_flag, tag);
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
               " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
// D = A(8x4,row) * B(4x8,col) + D.  lane holds a = A[lane>>2][lane&3], b = B[lane&3][lane>>2],
// d0,d1 = D[lane>>2][2*(lane&3) + {0,1}]
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Fragment row r (0..7) of an 8-row operand group is stored in shared-memory row 2*(r&3) + (r>>2).
// With the TMA 128-byte swizzle (16-byte chunk ^= row & 7) this makes every 64-bit fragment load
// conflict-free per half-warp; the natural order has a 2-way conflict (rows 0-3 only reach 4 of the
// 8 chunks).  Consequence: DMMA row r / column n correspond to particle offsets frag_row(r) / frag_row(n).
__device__ __forceinline__ int frag_row(int r) { return 2 * (r & 3) + (r >> 2); }

struct Ring {
  uint32_t stage, phase;
  __device__ __forceinline__ void advance(uint32_t nstages) {
    if(++stage == nstages) {
      stage = 0;
      phase ^= 1u;
    }
  }
};

struct BoxCoord {
  int off[6]; // element offset of the box inside the tile, per index id
  int weight; // permutation multiplicity of the box (TaskParams::sym), 1 without symmetry
};

// id -> box coordinates; false when the id is padding of the brick grid.  Inside a brick and across
// bricks h3 runs fastest, then h2, h1, p6, p5, p4.
__device__ __forceinline__ bool decode_box(const TaskParams& p, int id, BoxCoord& b) {
  const int order[6] = {2, 1, 0, 5, 4, 3};
  int       in[6], bx[6], r = id;
  bool      valid = true;
#pragma unroll
  for(int i = 0; i < 6; i++) {
    const int d = order[i];
    in[d]       = r % p.brick[d];
    r /= p.brick[d];
  }
#pragma unroll
  for(int i = 0; i < 6; i++) {
    const int d  = order[i];
    const int bi = (r % p.nbrick[d]) * p.brick[d] + in[d];
    r /= p.nbrick[d];
    valid &= bi < p.nbox[d];
    bx[d]    = bi;
    b.off[d] = bi * (d < 3 ? p.c[d] : PBOX);
  }
  b.weight = p.sym ? box_weight(p.sym, bx) : 1;
  return valid && b.weight > 0;
}

} // namespace ccsdt
