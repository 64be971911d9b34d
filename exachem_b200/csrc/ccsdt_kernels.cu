// sm_100a kernels of the fused CCSD(T) triples path.
//
//   gather_panels_kernel   : builds the K-major operand panels of one task from tensor blocks
//                            (replaces hptt sort + pinned staging + H2D of the reference,
//                             exachem/cc/ccsd_t/ccsd_t_all_fused_{singles,doubles1,doubles2}.hpp)
//   fused_t_dmma_kernel    : the product kernel.  One CTA = one (2s1,2s2,2s3,8,8,8) box of the
//                            t3 tile; TMA (128B swizzle) + mbarrier ring feeds FP64 DMMA m8n8k4;
//                            all enabled d1/d2 terms accumulate in registers, s1 and the energy
//                            denominators are applied in the epilogue, per-box partials are written
//                            without atomics  (replaces ccsd_t_all_fused_gpu.cu:132-2564)
//   fused_t_simple_kernel  : diagnostic one-thread-per-element FMA kernel over the same panels
//   reduce_partials_kernel : fixed-order sum of the per-box partials (replaces hostEnergyReduce,
//                            ccsd_t_all_fused.hpp:19-32)
//   probes                 : FP64 DFMA / DMMA peak, DMMA fragment layout, TMA swizzle layout
#include "ccsdt_kernel_common.cuh"

#include <algorithm>
#include <cstdio>
#include <type_traits>

namespace ccsdt {

// =================================================================================================
// panel build
// =================================================================================================
// (The staging kernels cannot join an SM that holds three fused CTAs -- tools/probes/coresidency.cu -- so the panel
// build runs in the gap between two fused kernels and is sized for speed, not for co-residency.)
__global__ void __launch_bounds__(256) gather_panels_kernel(const GatherDesc* __restrict__ descs,
                                                                SynthInfo si) {
  const GatherDesc& d  = descs[blockIdx.y];
  const int64_t     n3 = d.n[3], n2 = d.n[2], n1 = d.n[1];
  const int64_t     total = (int64_t) d.n[0] * n1 * n2 * n3;
  for(int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < total;
      e += (int64_t) gridDim.x * blockDim.x) {
    const int64_t k  = e % n3;
    int64_t       r  = e / n3;
    const int64_t in = r % n2;
    r /= n2;
    const int64_t o1 = r % n1, o2 = r / n1;
    double        v;
    if(d.synth_tensor >= 0) {
      int64_t idx[4] = {0, 0, 0, 0};
      idx[d.gpos[0]] = d.gbase[0] + o2;
      idx[d.gpos[1]] = d.gbase[1] + o1;
      idx[d.gpos[2]] = d.gbase[2] + in;
      idx[d.gpos[3]] = d.gbase[3] + k;
      v              = d.scale * synth_value(si, d.synth_tensor, idx);
    }
    else if(d.src) { v = d.scale * __ldg(d.src + o2 * d.ss[0] + o1 * d.ss[1] + in * d.ss[2] + k * d.ss[3]); }
    else { v = 0.0; }
    d.dst[o2 * d.ds[0] + o1 * d.ds[1] + in * d.ds[2] + k * d.ds[3]] = v;
  }
}

cudaError_t launch_gather(const GatherDesc* dev_descs, int ndesc, int64_t max_elems, SynthInfo si,
                          cudaStream_t st) {
  if(ndesc <= 0) return cudaSuccess;
  int64_t bx = (max_elems + 256 * 8 - 1) / (256 * 8);
  if(bx < 1) bx = 1;
  if(bx > 1024) bx = 1024;
  for(int off = 0; off < ndesc; off += 65535) {
    const int n = ndesc - off < 65535 ? ndesc - off : 65535;
    gather_panels_kernel<<<dim3((unsigned) bx, (unsigned) n), 256, 0, st>>>(dev_descs + off, si);
  }
  return cudaGetLastError();
}

// zeroes the per-box partials and the box-scheduler words of a staging buffer (instead of two cudaMemsetAsync,
// whose kernels need not fit next to the resident fused CTAs)
__global__ void __launch_bounds__(128, 16) zero_words_kernel(double* __restrict__ a, int64_t na, uint32_t* __restrict__ b,
                                                             int nb) {
  const int64_t i0 = (int64_t) blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t) gridDim.x * blockDim.x;
  for(int64_t i = i0; i < na; i += stride) a[i] = 0.0;
  for(int64_t i = i0; i < nb; i += stride) b[i] = 0u;
}

// copies the gather descriptors of a task from pinned host memory (read through its device alias) to device memory.
// A cudaMemcpyAsync would queue on the host-to-device copy engine BEHIND any dense upload still in flight
// (ccsdt_put_dense_async) and hold the first panel build back until the last byte has arrived.
__global__ void __launch_bounds__(128, 16) copy_words_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int64_t n) {
  for(int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x) dst[i] = src[i];
}

cudaError_t launch_copy_from_pinned(const void* pinned_host, void* dev, size_t bytes, cudaStream_t st) {
  void*       alias = nullptr;
  cudaError_t e     = cudaHostGetDevicePointer(&alias, const_cast<void*>(pinned_host), 0);
  if(e != cudaSuccess) return e;
  const int64_t n = (int64_t) (bytes / 4);
  copy_words_kernel<<<(unsigned) std::min<int64_t>((n + 127) / 128, 64), 128, 0, st>>>((const uint32_t*) alias, (uint32_t*) dev, n);
  return cudaGetLastError();
}

cudaError_t launch_zero(double* a, int64_t na, uint32_t* b, int nb, cudaStream_t st) {
  int64_t blocks = (na + 128 * 4 - 1) / (128 * 4);
  if(blocks < 1) blocks = 1;
  if(blocks > 592) blocks = 592;
  zero_words_kernel<<<(unsigned) blocks, 128, 0, st>>>(a, na, b, nb);
  return cudaGetLastError();
}

// Antisymmetry of a 4-index array in dimensions (PAIR, PAIR + 1), which must have the same extent: sets *flag when
// A[..x,y..] + A[..y,x..] is not zero to rounding.  Guards options.symmetry (the box skipping of the fused kernel
// assumes t3 antisymmetric in same-spin indices, i.e. antisymmetric T2 / V2); element strides are passed, so the same
// kernel checks a block of the block store and a whole dense tensor.
__global__ void __launch_bounds__(256) antisym_check_kernel(const double* __restrict__ A, int64_t n0, int64_t n1, int64_t n2, int64_t n3,
                                                            int64_t s0, int64_t s1, int64_t s2, int64_t s3, int pair,
                                                            uint32_t* __restrict__ flag) {
  const int64_t total = n0 * n1 * n2 * n3;
  for(int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t) gridDim.x * blockDim.x) {
    int64_t r = e, i[4];
    i[3] = r % n3; r /= n3;
    i[2] = r % n2; r /= n2;
    i[1] = r % n1;
    i[0] = r / n1;
    if(i[pair] > i[pair + 1]) continue;
    const double  x = A[i[0] * s0 + i[1] * s1 + i[2] * s2 + i[3] * s3];
    const int64_t t = i[pair];
    i[pair] = i[pair + 1], i[pair + 1] = t;
    const double y = A[i[0] * s0 + i[1] * s1 + i[2] * s2 + i[3] * s3];
    // (converged amplitudes are antisymmetric to solver noise, not to the last bit: 1e-7 relative / 1e-11 absolute is far
    //  below what moves the energy by 1e-9 Eh and far above that noise)
    if(fabs(x + y) > 1e-7 * (fabs(x) + fabs(y)) + 1e-11) atomicOr(flag, 1u);
  }
}

cudaError_t launch_antisym_check(const double* A, const int64_t n[4], const int64_t st[4], int pair, uint32_t* flag, cudaStream_t stream) {
  const int64_t total = n[0] * n[1] * n[2] * n[3];
  if(total <= 0 || n[pair] != n[pair + 1]) return cudaSuccess;
  const unsigned grid = (unsigned) std::min<int64_t>((total + 255) / 256, 148 * 8);
  antisym_check_kernel<<<grid, 256, 0, stream>>>(A, n[0], n[1], n[2], n[3], st[0], st[1], st[2], st[3], pair, flag);
  return cudaGetLastError();
}

// =================================================================================================
// shared epilogue pieces
// =================================================================================================
__device__ __forceinline__ double s1_sum(const TaskParams& p, const int coord[6]) {
  double s = 0.0;
  for(int k = 0; k < p.ns1; k++) {
    const S1Dev& t  = p.s1[k];
    int          oa = 0, ob = 0;
#pragma unroll
    for(int i = 0; i < 6; i++) {
      oa += coord[i] * t.sa[i];
      ob += coord[i] * t.sb[i];
    }
    s += __ldg(t.a + oa) * __ldg(t.b + ob);
  }
  return s;
}

// d / D for the energy denominators (|D| is an orbital-energy gap, far from 0, inf and the subnormals):
// hardware reciprocal seed (2^-23), two Newton steps, one correction of the quotient.  Within 1 ulp of
// the IEEE quotient; replaces the ~30-instruction generic FP64 division, 32 of which sat in the epilogue.
__device__ __forceinline__ double div_fast(double d, double D) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(D));
  double e = fma(-D, r, 1.0);
  r        = fma(r, e, r);
  e        = fma(-D, r, 1.0);
  r        = fma(r, e, r);
  double q = d * r;
  return fma(fma(-D, q, d), r, q);
}

// One s1 term of a thread's 32 elements: sum_e tq[e] * a[e] * b[e] with a = sign*T1[pa,hx] and
// b = v2ijab[hz,hy,pc,pb].  Element bit <-> accumulator index: bit 4 = i1, 3 = i2, 2 = i3, 1 = ql, 0 = r.
// a carries exactly one hole (HX, compile time) and at most one of the two in-thread particle bits; b
// carries the other two holes and the remaining particle bits.  Absent indices have stride 0, so
// 8 loads of a and 16 of b (some of them repeats that hit L1) serve all 32 elements; only three
// instantiations exist (code size: the epilogue shares the instruction cache with the DMMA loops).
template<int HX>
__device__ __forceinline__ double s1_term(const double (&tq)[32], const double* __restrict__ pa,
                                          const double* __restrict__ pb, const int (&da)[5], const int (&db)[5]) {
  constexpr int HB = 4 - HX;                           // element bit of a's hole
  constexpr int O1 = HX == 0 ? 3 : 4, O2 = HX == 2 ? 3 : 2; // element bits of the other two holes
  double        av[2][4], bv[4][4];
#pragma unroll
  for(int h = 0; h < 2; h++)
#pragma unroll
    for(int q = 0; q < 4; q++) av[h][q] = __ldg(pa + h * da[HB] + (q >> 1) * da[1] + (q & 1) * da[0]);
#pragma unroll
  for(int o = 0; o < 4; o++)
#pragma unroll
    for(int q = 0; q < 4; q++)
      bv[o][q] = __ldg(pb + (o >> 1) * db[O1] + (o & 1) * db[O2] + (q >> 1) * db[1] + (q & 1) * db[0]);
  // sum over (hole of a, particle bits q) of a * (sum over the other two holes of tq * b): 40 FP64 operations
  double sum = 0.0;
#pragma unroll
  for(int h = 0; h < 2; h++)
#pragma unroll
    for(int q = 0; q < 4; q++) {
      double inner = 0.0;
#pragma unroll
      for(int o = 0; o < 4; o++)
        inner = fma(tq[(h << HB) | ((o >> 1) << O1) | ((o & 1) << O2) | q], bv[o][q], inner);
      sum = fma(av[h][q], inner, sum);
    }
  return sum;
}

// =================================================================================================
// diagnostic kernel: one thread per t3 element, plain FMAs over the same panels
// =================================================================================================
__global__ void __launch_bounds__(256) fused_t_simple_kernel(const __grid_constant__ TaskParams p) {
  __shared__ double red[2][8];
  const int64_t total = (int64_t) p.ext[0] * p.ext[1] * p.ext[2] * p.ext[3] * p.ext[4] * p.ext[5];
  const int64_t e     = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  double        e1 = 0.0, e2 = 0.0;
  if(e < total) {
    int     coord[6];
    int64_t r = e;
    // h3 fastest, then h2, h1, p6, p5, p4 (the reference's t3 element order)
    coord[2] = (int) (r % p.ext[2]); r /= p.ext[2];
    coord[1] = (int) (r % p.ext[1]); r /= p.ext[1];
    coord[0] = (int) (r % p.ext[0]); r /= p.ext[0];
    coord[5] = (int) (r % p.ext[5]); r /= p.ext[5];
    coord[4] = (int) (r % p.ext[4]); r /= p.ext[4];
    coord[3] = (int) r;
    double d = 0.0;
    for(int t = 0; t < p.nterms; t++) {
      const TermDev& td = p.term[t];
      const int      pl = td.pool;
      const double*  A  = p.geom.hpp[pl] + td.hpp_panel * p.geom.hpp_stride_panel(pl) +
                        coord[td.hpp_hole] * p.geom.hpp_stride_o2(pl) +
                        coord[td.qt] * p.geom.hpp_stride_o1(pl) + coord[td.inner_hpp] * p.geom.hpp_stride_in(pl);
      const double* B = p.geom.hhp[pl] + td.hhp_panel * p.geom.hhp_stride_panel(pl) +
                        coord[td.hhp_o2] * p.geom.hhp_stride_o2(pl) +
                        coord[td.hhp_o1] * p.geom.hhp_stride_o1(pl) + coord[td.inner_hhp] * p.geom.hhp_stride_in(pl);
      const int nk = td.kslabs * KSLAB;
      for(int k = 0; k < nk; k++) d = fma(A[k], B[k], d);
    }
    const double s = s1_sum(p, coord);
    const double D = __ldg(p.evl[0] + coord[0]) + __ldg(p.evl[1] + coord[1]) + __ldg(p.evl[2] + coord[2]) -
                     __ldg(p.evl[3] + coord[3]) - __ldg(p.evl[4] + coord[4]) - __ldg(p.evl[5] + coord[5]);
    const double tmp = d / D;
    e1 = tmp * d;
    e2 = tmp * (d + s);
  }
#pragma unroll
  for(int o = 16; o > 0; o >>= 1) {
    e1 += __shfl_xor_sync(0xffffffffu, e1, o);
    e2 += __shfl_xor_sync(0xffffffffu, e2, o);
  }
  const int w = threadIdx.x >> 5;
  if((threadIdx.x & 31) == 0) red[0][w] = e1, red[1][w] = e2;
  __syncthreads();
  if(threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for(int i = 0; i < 8; i++) a += red[0][i], b += red[1][i];
    p.partial[2 * (int64_t) blockIdx.x]     = a;
    p.partial[2 * (int64_t) blockIdx.x + 1] = b;
  }
}

cudaError_t launch_fused_simple(const TaskParams& p, cudaStream_t st, int* grid_out) {
  const int64_t total = (int64_t) p.ext[0] * p.ext[1] * p.ext[2] * p.ext[3] * p.ext[4] * p.ext[5];
  const int     grid  = (int) ((total + 255) / 256);
  if(grid_out) *grid_out = grid;
  fused_t_simple_kernel<<<grid, 256, 0, st>>>(p);
  return cudaGetLastError();
}

// =================================================================================================
// the fused DMMA kernel
// =================================================================================================
template<int N>
using Int = std::integral_constant<int, N>;

// All K slabs of one contraction for one consumer warp.
// acc index = (((i1*2 + i2)*2 + i3)*2 + ql)*2 + r  with i* the hole offsets inside the warp group's
// (2,2,2) sub-box, ql the warp's tile-particle slot and r the DMMA column parity.
// SKIP: boundary boxes of an ODD hole extent hold one valid hole index where the sub-box has two slots; the second slot
// of that hole is then left out altogether -- SKIP = 1: it is the HPP hole (hx = 1), 2 / 3: the first / second HHP hole
// (ga = 1 / gb = 1) -- half the fragments, half the DMMAs.  These are separate, smaller instantiations of the loop (a
// predicate per DMMA costs more than it saves: ptxas guards every predicated mma.sync with a convergence barrier); the
// host orders the boxes by class (interior first, then the boundary layers hole by hole), so that the CTAs of an SM run
// the same few instantiations at any time and the instruction cache is not thrashed.
template<int HH, bool A_HPP, int SKIP>
__device__ __forceinline__ void consume_term(double (&acc)[32], const TaskParams& p, const TermDev& td,
                                             Ring& ring, uint32_t ring_base, uint32_t full_bar,
                                             uint32_t empty_bar, const int sub_off[3], int wq, int lane) {
  constexpr int HA = (HH == 0) ? 1 : 0;        // the two other holes, ascending
  constexpr int HB = (HH == 2) ? 1 : 2;
  constexpr int NHX = SKIP == 1 ? 1 : 2, NGA = SKIP == 2 ? 1 : 2, NGB = SKIP == 3 ? 1 : 2;
  const int     q  = frag_row(lane >> 2), l3 = lane & 3;
  // lane part of the swizzled fragment address (see DESIGN.md "shared-memory layout"): fragment row
  // lane>>2 lives in shared-memory row frag_row(lane>>2) of its 8-row group, so that the 16 lanes of
  // each half-warp (fragment rows 0-3 / 4-7) hit 16 distinct 8-byte bank pairs
  const uint32_t lane_const = (uint32_t) (q * ROW_BYTES + ((((l3 >> 1) ^ (q & 1)) << 4) | ((l3 & 1) << 3)));
  const uint32_t jx         = (uint32_t) (q >> 1);

  // row offsets (bytes) of this warp's 4 + 4 fragments inside a stage
  const int hpp_rows = p.c[HH] * 64;
  uint32_t  hpp_off[2][2], hhp_off[2][2];
#pragma unroll
  for(int ih = 0; ih < 2; ih++)
#pragma unroll
    for(int ql = 0; ql < 2; ql++)
      hpp_off[ih][ql] = (uint32_t) ((((sub_off[HH] + ih) * 8 + 2 * wq + ql) * 8) * ROW_BYTES) + lane_const;
  const int c_o1 = p.c[HB]; // HHP is [o2 = HA][o1 = HB][inner]
#pragma unroll
  for(int ia = 0; ia < 2; ia++)
#pragma unroll
    for(int ib = 0; ib < 2; ib++)
      hhp_off[ia][ib] =
        (uint32_t) ((hpp_rows + ((sub_off[HA] + ia) * c_o1 + sub_off[HB] + ib) * 8) * ROW_BYTES) + lane_const;

  // One 4-wide k-step = 16 DMMAs from 4 HPP fragments (fh) and 4 HHP fragments (fg), all single-buffered:
  // the DMMAs run fh-major, each fh is re-requested for the next step right after its fourth and last use,
  // each fg after its last use in the final group -- every reload sits at least three DMMA issue slots
  // (~48 cycles) ahead of its first use, which covers the shared-memory latency.  The k-loop is unrolled by
  // two only (measured: x2 +1.3 %, x4 +2 % on long K but -2 % on short K): the six (HH, A_HPP) instantiations of this loop are the whole hot code of the kernel and must
  // stay resident in the instruction cache while co-resident CTAs run different terms.
  // Loads past the end of the term read valid but unused shared memory (no conditional loads).
  double fh[2][2], fg[2][2];
  mbar_wait(full_bar + 8 * ring.stage, ring.phase, p.error_flag, 1);
  uint32_t base = ring_base + ring.stage * (uint32_t) p.stage_bytes;
#pragma unroll
  for(int x = 0; x < 2; x++)
#pragma unroll
    for(int y = 0; y < 2; y++) {
      if(x < NHX) fh[x][y] = lds_f64(base + (jx << 5) + hpp_off[x][y]);
      if(x < NGA && y < NGB) fg[x][y] = lds_f64(base + (jx << 5) + hhp_off[x][y]);
    }

  for(int s = 0; s < td.kslabs; s++) {
    const bool     last = s == td.kslabs - 1;
    const uint32_t nj   = last ? (uint32_t) td.ksteps_last : 4u; // K tail: 4-wide steps only
    Ring           nxt  = ring;
    nxt.advance((uint32_t) p.stages);
    const uint32_t next_base = last ? base : ring_base + nxt.stage * (uint32_t) p.stage_bytes;
#pragma unroll 2
    for(uint32_t j = 0; j < nj; j++) {
      uint32_t jn; // shared-memory address (without fragment offset) of the next step's fragments
      if(j == 3) {
        if(!last) mbar_wait(full_bar + 8 * nxt.stage, nxt.phase, p.error_flag, 1);
        jn = next_base + (jx << 5);
      }
      else jn = base + (((j + 1) ^ jx) << 5);
#pragma unroll
      for(int hx = 0; hx < NHX; hx++)
#pragma unroll
        for(int ql = 0; ql < 2; ql++) {
#pragma unroll
          for(int ga = 0; ga < NGA; ga++)
#pragma unroll
            for(int gb = 0; gb < NGB; gb++) {
              int hi[3];
              hi[HH] = hx, hi[HA] = ga, hi[HB] = gb;
              const int ai = ((((hi[0] * 2 + hi[1]) * 2 + hi[2]) * 2 + ql) * 2);
              if(A_HPP) dmma884(acc[ai], acc[ai + 1], fh[hx][ql], fg[ga][gb]);
              else dmma884(acc[ai], acc[ai + 1], fg[ga][gb], fh[hx][ql]);
              if(hx == NHX - 1 && ql == 1) fg[ga][gb] = lds_f64(jn + hhp_off[ga][gb]);
            }
          fh[hx][ql] = lds_f64(jn + hpp_off[hx][ql]);
        }
    }
    mbar_arrive(empty_bar + 8 * ring.stage); // every lane arrives: no warp-wide reconvergence in the loop
    ring = nxt;
    base = next_base;
  }
}

// a warp none of whose tile-particle slots lies inside the tile (ragged last box of p4 / p5) has no DMMA to issue for the
// term: it only keeps the ring protocol (wait for every slab, hand it back)
__device__ __forceinline__ void consume_idle(const TaskParams& p, const TermDev& td, Ring& ring, uint32_t full_bar,
                                             uint32_t empty_bar) {
  for(int s = 0; s < td.kslabs; s++) {
    mbar_wait(full_bar + 8 * ring.stage, ring.phase, p.error_flag, 1);
    mbar_arrive(empty_bar + 8 * ring.stage);
    ring.advance((uint32_t) p.stages);
  }
}

// jb: the hole (0..2) whose second slot of this warp's sub-box lies outside the tile, -1 = none
template<int HH, bool A_HPP>
__device__ __forceinline__ void consume_skip(double (&acc)[32], const TaskParams& p, const TermDev& td, Ring& ring,
                                             uint32_t ring_base, uint32_t full_bar, uint32_t empty_bar, const int sub_off[3],
                                             int wq, int lane, int jb) {
  constexpr int HA = (HH == 0) ? 1 : 0;
  if(jb < 0) consume_term<HH, A_HPP, 0>(acc, p, td, ring, ring_base, full_bar, empty_bar, sub_off, wq, lane);
  else if(jb == HH) consume_term<HH, A_HPP, 1>(acc, p, td, ring, ring_base, full_bar, empty_bar, sub_off, wq, lane);
  else if(jb == HA) consume_term<HH, A_HPP, 2>(acc, p, td, ring, ring_base, full_bar, empty_bar, sub_off, wq, lane);
  else consume_term<HH, A_HPP, 3>(acc, p, td, ring, ring_base, full_bar, empty_bar, sub_off, wq, lane);
}

__device__ __forceinline__ void consume_dispatch(double (&acc)[32], const TaskParams& p, const TermDev& td,
                                                 Ring& ring, uint32_t ring_base, uint32_t full_bar,
                                                 uint32_t empty_bar, const int sub_off[3], int wq, int lane, int jb) {
  switch(td.hpp_hole * 2 + td.a_is_hpp) {
    case 0: consume_skip<0, false>(acc, p, td, ring, ring_base, full_bar, empty_bar, sub_off, wq, lane, jb); break;
    case 1: consume_skip<0, true>(acc, p, td, ring, ring_base, full_bar, empty_bar, sub_off, wq, lane, jb); break;
    case 2: consume_skip<1, false>(acc, p, td, ring, ring_base, full_bar, empty_bar, sub_off, wq, lane, jb); break;
    case 3: consume_skip<1, true>(acc, p, td, ring, ring_base, full_bar, empty_bar, sub_off, wq, lane, jb); break;
    case 4: consume_skip<2, false>(acc, p, td, ring, ring_base, full_bar, empty_bar, sub_off, wq, lane, jb); break;
    default: consume_skip<2, true>(acc, p, td, ring, ring_base, full_bar, empty_bar, sub_off, wq, lane, jb); break;
  }
}

// producer: all slabs of one contraction for one box
__device__ __forceinline__ void produce_term(const TaskParams& p, const TermDev& td, const BoxCoord& bc,
                                             Ring& ring, uint32_t ring_base, uint32_t full_bar,
                                             uint32_t empty_bar) {
  const CUtensorMap* map_hpp = &p.tmap[td.pool * 2 + 0];
  const CUtensorMap* map_hhp = &p.tmap[td.pool * 2 + 1];
  const int          c_h     = p.c[td.hpp_hole];
  const int          c_o2 = p.c[td.hhp_o2], c_o1 = p.c[td.hhp_o1];
  const uint32_t     hpp_bytes = (uint32_t) (c_h * 64 * ROW_BYTES);
  const uint32_t     tx        = hpp_bytes + (uint32_t) (c_o2 * c_o1 * 8 * ROW_BYTES);
  const int          hpp_in = bc.off[td.inner_hpp], hpp_o1 = bc.off[td.qt];
  const int          hpp_o2 = td.hpp_panel * p.geom.THp + bc.off[td.hpp_hole];
  const int          hhp_in = bc.off[td.inner_hhp], hhp_o1 = bc.off[td.hhp_o1];
  const int          hhp_o2 = td.hhp_panel * p.geom.THp + bc.off[td.hhp_o2];
  for(int s = 0; s < td.kslabs; s++) {
    mbar_wait(empty_bar + 8 * ring.stage, ring.phase ^ 1u, p.error_flag, 2);
    const uint32_t bar = full_bar + 8 * ring.stage;
    const uint32_t dst = ring_base + ring.stage * (uint32_t) p.stage_bytes;
    mbar_arrive_expect_tx(bar, tx);
    const int k0 = s * KSLAB;
    for(int j = 0; j < c_h; j++)
      tma_load_4d(dst + (uint32_t) (j * 64 * ROW_BYTES), map_hpp, bar, k0, hpp_in, hpp_o1, hpp_o2 + j);
    for(int j2 = 0; j2 < c_o2; j2++)
      for(int j1 = 0; j1 < c_o1; j1 += 2)
        tma_load_4d(dst + hpp_bytes + (uint32_t) ((j2 * c_o1 + j1) * 8 * ROW_BYTES), map_hhp, bar, k0, hhp_in,
                    hhp_o1 + j1, hhp_o2 + j2);
    ring.advance((uint32_t) p.stages);
  }
}

// Box queue: the producer warp owns the dynamic box scheduler (one atomicAdd per box on a global
// counter, ids in brick-major order) and tells the consumers which box the slabs it is about to issue
// belong to.  Every box has the same number of slabs, so the consumers only need the coordinates for
// the epilogue.  Entry seq % BOXQ is written before the first slab of box number seq is issued and is
// published by that slab's mbarrier completion; the producer is never more than `stages` slabs (hence
// boxes) ahead, so MAX_STAGES + 2 entries cannot be overrun.  id < 0 ends the CTA.
constexpr int BOXQ = MAX_STAGES + 2;
struct BoxQueueEntry {
  int      id;
  int      slot; // index of the box's partial: the id, or the position in the symmetry-reduced list
  BoxCoord bc;
};

// blockDim.x = 32 * (consumer_warps + 1); the last warp is the TMA producer.
// Two instantiations: <160,3> (4 consumer warps, three CTAs per SM) and <416,1> (8 or 12 consumer
// warps, one CTA per SM); both get the full 128 registers per thread, which the 64 accumulator
// registers + 16 fragment registers need to stay spill-free.
template<int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) fused_t_dmma_kernel(const __grid_constant__ TaskParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * MAX_STAGES];
  __shared__ double        red[2][16];
  __shared__ BoxQueueEntry boxq[BOXQ];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ncw = (int) (blockDim.x >> 5) - 1; // consumer warps

  // 1024-byte aligned ring (the swizzle pattern is anchored on absolute smem address bits)
  const uint32_t ring_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t full_bar  = smem_u32(&bars[0]);
  const uint32_t empty_bar = smem_u32(&bars[MAX_STAGES]);

  if(tid == 0) {
    for(int s = 0; s < p.stages; s++) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, (uint32_t) (ncw * 32));
    }
    mbar_fence_init();
  }
  if(warp == ncw && lane == 0) {
    for(int i = 0; i < 4; i++) tma_prefetch_desc(&p.tmap[i]);
  }
  __syncthreads();

  const bool relayout = p.nterms_x > 0 && p.nterms_x < p.nterms;
  Ring       ring{0u, 0u};

  if(warp == ncw) {
    // ================= producer warp =================
    // The CTAs that share an SM all take the same time per box; started together they would all reach
    // their (tensor-pipe-idle) relayouts and epilogues at the same moment.  The k-th CTA to start on an SM
    // therefore delays its first load by k/ctas_per_sm of a box time, so that one CTA's epilogue runs under
    // the other CTAs' DMMAs.
    if(lane == 0 && p.stagger_cycles > 0) {
      uint32_t smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      const uint32_t  slot = atomicAdd(p.box_counter + 1 + (smid & 255u), 1u) % (uint32_t) p.ctas_per_sm;
      const long long t0   = clock64();
      while(clock64() - t0 < (long long) slot * p.stagger_cycles) {}
    }
    __syncwarp();
    for(int seq = 0;; seq++) {
      int      id = -1, slot = 0;
      BoxCoord bc;
      if(lane == 0) {
        for(;;) {
          id = slot = (int) atomicAdd(p.box_counter, 1u);
          if(p.box_list) { // symmetry-reduced task: dense list of the surviving ids
            id = id < p.nlist ? __ldg(p.box_list + id) : -1;
            if(id >= 0) decode_box(p, id, bc);
            break;
          }
          if(id >= p.nboxes_padded) {
            id = -1;
            break;
          }
          if(decode_box(p, id, bc)) break;
        }
        BoxQueueEntry& e = boxq[seq % BOXQ];
        e.id             = id;
        e.slot           = slot;
        e.bc             = bc;
        if(id < 0) {
          // end marker: complete one more phase of the next full barrier without data
          mbar_wait(empty_bar + 8 * ring.stage, ring.phase ^ 1u, p.error_flag, 2);
          mbar_arrive(full_bar + 8 * ring.stage);
        }
      }
      id = __shfl_sync(0xffffffffu, id, 0);
      if(id < 0) break;
      for(int t = 0; t < p.nterms; t++) {
        if(relayout && t == p.nterms_x) {
          __syncwarp();    // reconverge before the CTA-wide (aligned) barriers
          __syncthreads(); // consumers are done with every X slab: the ring is idle
          __syncthreads(); // accumulators parked in the ring
          __syncthreads(); // accumulators reloaded: the ring may be refilled
        }
        if(lane == 0) produce_term(p, p.term[t], bc, ring, ring_base, full_bar, empty_bar);
      }
      __syncwarp();
    }
    return;
  }

  // ================= consumer warps =================
  const int grp = warp >> 2, wq = warp & 3;
  int       sub_off[3];
  {
    const int g2 = grp % p.sub[2], g1 = (grp / p.sub[2]) % p.sub[1], g0 = grp / (p.sub[2] * p.sub[1]);
    sub_off[0] = 2 * g0, sub_off[1] = 2 * g1, sub_off[2] = 2 * g2;
  }
  const int l3  = lane & 3;
  const int q6  = frag_row(lane >> 2);            // particle offset of this lane's DMMA row
  const int qc0 = frag_row(2 * l3);               // particle offset of DMMA column 2*l3 (+2 for column 2*l3+1)

  for(int seq = 0;; seq++) {
    // the first slab of the box publishes its queue entry
    mbar_wait(full_bar + 8 * ring.stage, ring.phase, p.error_flag, 4);
    const int box = boxq[seq % BOXQ].id;
    if(box < 0) break;
    const BoxCoord bc = boxq[seq % BOXQ].bc;
    double         acc[32];
#pragma unroll
    for(int i = 0; i < 32; i++) acc[i] = 0.0;
    // boundary layers of the tile (see consume_term): the hole whose second slot is outside, and whether this warp's
    // two tile-particle slots are both outside in layout X (p4) / Y (p5)
    int jb = -1;
#pragma unroll
    for(int j = 2; j >= 0; j--)
      if(bc.off[j] + sub_off[j] + 1 >= p.ext[j]) jb = j;
#ifdef CCSDT_NO_BOUNDARY_VARIANTS // A/B builds (tools/ab_build.sh)
    jb = -1;
    const bool idle_x = false, idle_y = false;
#else
    const bool idle_x = bc.off[3] + 2 * wq >= p.ext[3], idle_y = bc.off[4] + 2 * wq >= p.ext[4];
#endif

    for(int t = 0; t < p.nterms; t++) {
      if(relayout && t == p.nterms_x) {
        // X -> Y: swap the roles of p4 and p5 (tile particle <-> DMMA column) through shared memory.
        // scratch index = ((((h1*c2 + h2)*c3 + h3)*8 + p4)*8 + p5)*8 + p6  (box-local coordinates)
        __syncthreads();
        double* scratch = reinterpret_cast<double*>(smem_raw + (ring_base - smem_u32(smem_raw)));
#pragma unroll
        for(int i1 = 0; i1 < 2; i1++)
#pragma unroll
          for(int i2 = 0; i2 < 2; i2++)
#pragma unroll
            for(int i3 = 0; i3 < 2; i3++)
#pragma unroll
              for(int ql = 0; ql < 2; ql++)
#pragma unroll
                for(int r = 0; r < 2; r++) {
                  const int hl = ((sub_off[0] + i1) * p.c[1] + sub_off[1] + i2) * p.c[2] + sub_off[2] + i3;
                  const int p4 = 2 * wq + ql, p5 = qc0 + 2 * r;
                  scratch[((hl * 8 + p4) * 8 + p5) * 8 + q6] = acc[((((i1 * 2 + i2) * 2 + i3) * 2 + ql) * 2) + r];
                }
        __syncthreads();
#pragma unroll
        for(int i1 = 0; i1 < 2; i1++)
#pragma unroll
          for(int i2 = 0; i2 < 2; i2++)
#pragma unroll
            for(int i3 = 0; i3 < 2; i3++)
#pragma unroll
              for(int ql = 0; ql < 2; ql++)
#pragma unroll
                for(int r = 0; r < 2; r++) {
                  const int hl = ((sub_off[0] + i1) * p.c[1] + sub_off[1] + i2) * p.c[2] + sub_off[2] + i3;
                  const int p5 = 2 * wq + ql, p4 = qc0 + 2 * r;
                  acc[((((i1 * 2 + i2) * 2 + i3) * 2 + ql) * 2) + r] = scratch[((hl * 8 + p4) * 8 + p5) * 8 + q6];
                }
        // generic-proxy accesses to the ring must be ordered before the TMA (async proxy) refills it
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
      }
      if(t < p.nterms_x ? idle_x : idle_y) consume_idle(p, p.term[t], ring, full_bar, empty_bar);
      else consume_dispatch(acc, p, p.term[t], ring, ring_base, full_bar, empty_bar, sub_off, wq, lane, jb);
    }

    // ---------------- epilogue: denominators, E[T], then the s1 part of E(T) ----------------
    // E[T] += d*d/D ; E(T) += d*(d+s)/D = E[T] part + (d/D)*s.  Pass 1 turns every accumulator into
    // t = d/D (0 for padding elements); pass 2 adds t * (a_k * b_k) term by term, walking the staged
    // s1 operands with per-term strides (no per-element index arithmetic).
    const bool layout_y = p.nterms > p.nterms_x; // layout of the accumulators now
    const int  id_qt = layout_y ? 4 : 3, id_qc = layout_y ? 3 : 4;
    int        hc[3], tcq, ccq;                  // coordinates of element (0,0,0,0,0) of this thread
    hc[0] = bc.off[0] + sub_off[0], hc[1] = bc.off[1] + sub_off[1], hc[2] = bc.off[2] + sub_off[2];
    tcq = (layout_y ? bc.off[4] : bc.off[3]) + 2 * wq;
    ccq = (layout_y ? bc.off[3] : bc.off[4]) + qc0;
    const int  c6 = bc.off[5] + q6;
    double     eh[3][2], et[2], ec[2];
    unsigned   vmask = 0; // bit layout: [h1:2][h2:2][h3:2][qt:2][qc:2]
#pragma unroll
    for(int j = 0; j < 3; j++)
#pragma unroll
      for(int i = 0; i < 2; i++) {
        const bool ok = hc[j] + i < p.ext[j];
        eh[j][i]      = ok ? __ldg(p.evl[j] + hc[j] + i) : 0.0;
        vmask |= (unsigned) ok << (2 * j + i);
      }
#pragma unroll
    for(int i = 0; i < 2; i++) {
      const int     ext_t = layout_y ? p.ext[4] : p.ext[3], ext_c = layout_y ? p.ext[3] : p.ext[4];
      const double* ev_t  = layout_y ? p.evl[4] : p.evl[3];
      const double* ev_c  = layout_y ? p.evl[3] : p.evl[4];
      const bool    okt = tcq + i < ext_t, okc = ccq + 2 * i < ext_c;
      et[i]             = okt ? __ldg(ev_t + tcq + i) : 0.0;
      ec[i]             = okc ? __ldg(ev_c + ccq + 2 * i) : 0.0;
      vmask |= ((unsigned) okt << (6 + i)) | ((unsigned) okc << (8 + i));
    }
    const bool   ok6 = c6 < p.ext[5];
    const double e6  = ok6 ? __ldg(p.evl[5] + c6) : 0.0;
    double       e1 = 0.0, e2 = 0.0;
#pragma unroll
    for(int i1 = 0; i1 < 2; i1++)
#pragma unroll
      for(int i2 = 0; i2 < 2; i2++)
#pragma unroll
        for(int i3 = 0; i3 < 2; i3++)
#pragma unroll
          for(int ql = 0; ql < 2; ql++)
#pragma unroll
            for(int r = 0; r < 2; r++) {
              const int      ai   = ((((i1 * 2 + i2) * 2 + i3) * 2 + ql) * 2) + r;
              const unsigned need = (1u << i1) | (1u << (2 + i2)) | (1u << (4 + i3)) | (1u << (6 + ql)) | (1u << (8 + r));
              const bool     ok   = ok6 && (vmask & need) == need;
              const double   d    = acc[ai];
              const double   D    = eh[0][i1] + eh[1][i2] + eh[2][i3] - et[ql] - ec[r] - e6;
              const double   tq   = ok ? div_fast(d, D) : 0.0;
              e1 += tq * d;
              acc[ai] = tq;
            }
    e2 = e1;
    for(int k = 0; k < p.ns1; k++) {
      const S1Dev&  sd = p.s1[k];
      const double* pa = sd.a + (hc[0] * sd.sa[0] + hc[1] * sd.sa[1] + hc[2] * sd.sa[2] + tcq * sd.sa[id_qt] +
                                 ccq * sd.sa[id_qc] + c6 * sd.sa[5]);
      const double* pb = sd.b + (hc[0] * sd.sb[0] + hc[1] * sd.sb[1] + hc[2] * sd.sb[2] + tcq * sd.sb[id_qt] +
                                 ccq * sd.sb[id_qc] + c6 * sd.sb[5]);
      // strides per in-thread element bit: bit 4 = i1, 3 = i2, 2 = i3, 1 = ql, 0 = r (= accumulator index)
      const int da[5] = {2 * sd.sa[id_qc], sd.sa[id_qt], sd.sa[2], sd.sa[1], sd.sa[0]};
      const int db[5] = {2 * sd.sb[id_qc], sd.sb[id_qt], sd.sb[2], sd.sb[1], sd.sb[0]};
      switch(sd.hx) {
        case 0: e2 += s1_term<0>(acc, pa, pb, da, db); break;
        case 1: e2 += s1_term<1>(acc, pa, pb, da, db); break;
        default: e2 += s1_term<2>(acc, pa, pb, da, db); break;
      }
    }
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) {
      e1 += __shfl_xor_sync(0xffffffffu, e1, o);
      e2 += __shfl_xor_sync(0xffffffffu, e2, o);
    }
    if(lane == 0) red[0][warp] = e1, red[1][warp] = e2;
    named_bar_sync(1, ncw * 32);
    if(tid == 0) {
      double a = 0.0, b = 0.0;
      for(int i = 0; i < ncw; i++) a += red[0][i], b += red[1][i];
      const BoxQueueEntry& e = boxq[seq % BOXQ]; // the entry outlives the box (see BOXQ)
      const double         w = (double) e.bc.weight;
      p.partial[2 * (int64_t) e.slot]     = w * a;
      p.partial[2 * (int64_t) e.slot + 1] = w * b;
    }
    named_bar_sync(1, ncw * 32); // red[] may be overwritten by the next box
  }
}

// <160,3>: 4+1 warps, three CTAs per SM at 128 registers; <288,1>: 8+1 warps, one CTA per SM with up to
// 224 registers (no spills); <416,1>: 12+1 warps, one CTA per SM at 152 registers.
__global__ void reduce_partials_kernel(const double* __restrict__ partial, int n, double* __restrict__ out2);

cudaError_t fused_dmma_configure(size_t smem_bytes) {
  cudaError_t e;
  // the fused kernel wants all of the L1/shared-memory array as shared memory (the staging kernels keep the default
  // split: they run between fused kernels, never next to them, and their strided reads like the larger L1)
  const int carve = (int) cudaSharedmemCarveoutMaxShared;
  if((e = cudaFuncSetAttribute(fused_t_dmma_kernel<160, 3>, cudaFuncAttributePreferredSharedMemoryCarveout, carve)) != cudaSuccess) return e;
  if((e = cudaFuncSetAttribute(fused_t_dmma_kernel<288, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, carve)) != cudaSuccess) return e;
  if((e = cudaFuncSetAttribute(fused_t_dmma_kernel<416, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, carve)) != cudaSuccess) return e;
  if((e = cudaFuncSetAttribute(fused_t_dmma_kernel<160, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int) smem_bytes)) != cudaSuccess)
    return e;
  if((e = cudaFuncSetAttribute(fused_t_dmma_kernel<288, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int) smem_bytes)) != cudaSuccess)
    return e;
  return cudaFuncSetAttribute(fused_t_dmma_kernel<416, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int) smem_bytes);
}

int fused_dmma_max_ctas_per_sm(int threads, size_t smem_bytes) {
  int         n = 0;
  cudaError_t e =
    threads <= 160   ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fused_t_dmma_kernel<160, 3>, threads, smem_bytes)
    : threads <= 288 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fused_t_dmma_kernel<288, 1>, threads, smem_bytes)
                     : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fused_t_dmma_kernel<416, 1>, threads, smem_bytes);
  return e == cudaSuccess ? n : 0;
}

cudaError_t launch_fused_dmma(const TaskParams& p, int grid, int consumer_warps, size_t smem_bytes,
                              cudaStream_t st) {
  const int threads = 32 * (consumer_warps + 1);
  if(threads <= 160) fused_t_dmma_kernel<160, 3><<<grid, threads, smem_bytes, st>>>(p);
  else if(threads <= 288) fused_t_dmma_kernel<288, 1><<<grid, threads, smem_bytes, st>>>(p);
  else if(threads <= 416) fused_t_dmma_kernel<416, 1><<<grid, threads, smem_bytes, st>>>(p);
  else return cudaErrorInvalidConfiguration;
  return cudaGetLastError();
}

// =================================================================================================
// fixed-order reduction of per-box partials: out2[0..1] = sum(partial[:,0]), sum(partial[:,1])
// =================================================================================================
__global__ void __launch_bounds__(128, 16) reduce_partials_kernel(const double* __restrict__ partial, int n,
                                                                  double* __restrict__ out2) {
  __shared__ double sh[2][128];
  double            a = 0.0, b = 0.0;
  for(int i = threadIdx.x; i < n; i += 128) {
    a += partial[2 * (int64_t) i];
    b += partial[2 * (int64_t) i + 1];
  }
  sh[0][threadIdx.x] = a;
  sh[1][threadIdx.x] = b;
  __syncthreads();
  for(int s = 64; s > 0; s >>= 1) {
    if((int) threadIdx.x < s) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + s];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + s];
    }
    __syncthreads();
  }
  if(threadIdx.x == 0) out2[0] = sh[0][0], out2[1] = sh[1][0];
}

cudaError_t launch_reduce_partials(const double* partial, int n, double* out2, cudaStream_t st) {
  reduce_partials_kernel<<<1, 128, 0, st>>>(partial, n, out2);
  return cudaGetLastError();
}

// =================================================================================================
// probes
// =================================================================================================
__global__ void __launch_bounds__(256) fp64_fma_peak_kernel(double* out, int iters, double x) {
  double a[16];
#pragma unroll
  for(int i = 0; i < 16; i++) a[i] = threadIdx.x * 1e-9 + i;
  const double b = x, c = 1.0 - x * 1e-3;
  for(int it = 0; it < iters; it++) {
#pragma unroll
    for(int i = 0; i < 16; i++) a[i] = fma(a[i], c, b);
  }
  double s = 0;
#pragma unroll
  for(int i = 0; i < 16; i++) s += a[i];
  out[(int64_t) blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) fp64_dmma_peak_kernel(double* out, int iters, double x) {
  double acc[32];
#pragma unroll
  for(int i = 0; i < 32; i++) acc[i] = 0.0;
  const double a = x + threadIdx.x * 1e-9, b = 1.0 - x * 1e-3;
  for(int it = 0; it < iters; it++) {
#pragma unroll
    for(int i = 0; i < 16; i++) dmma884(acc[2 * i], acc[2 * i + 1], a, b);
  }
  double s = 0;
#pragma unroll
  for(int i = 0; i < 32; i++) s += acc[i];
  out[(int64_t) blockIdx.x * blockDim.x + threadIdx.x] = s;
}

cudaError_t probe_fp64_peak(int use_dmma, int iters, double* tflops, double* ms_out) {
  cudaDeviceProp prop;
  int            dev = 0;
  cudaGetDevice(&dev);
  cudaError_t err = cudaGetDeviceProperties(&prop, dev);
  if(err != cudaSuccess) return err;
  const int grid = prop.multiProcessorCount * 8, block = 256;
  double*   out  = nullptr;
  if((err = cudaMalloc(&out, sizeof(double) * grid * block)) != cudaSuccess) return err;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for(int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    if(use_dmma) fp64_dmma_peak_kernel<<<grid, block>>>(out, iters, 0.5);
    else fp64_fma_peak_kernel<<<grid, block>>>(out, iters, 0.5);
    cudaEventRecord(e1);
    if((err = cudaEventSynchronize(e1)) != cudaSuccess) break;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if(rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if(err != cudaSuccess) return err;
  const double flops = use_dmma ? (double) grid * (block / 32) * (double) iters * 16.0 * 512.0
                                : (double) grid * block * (double) iters * 16.0 * 2.0;
  *ms_out            = best;
  *tflops            = flops / (best * 1e-3) / 1e12;
  return cudaGetLastError();
}


// Mainloop emulation: every warp repeats { TA + TB fragment loads from shared memory (swizzled,
// conflict-free), TA*TB DMMAs } with no TMA, barriers or epilogue.  Shows how close a given
// warps-per-SM / tiles-per-warp configuration can get to the DMMA peak.
template<int TA, int TB>
__global__ void __launch_bounds__(512) mainloop_probe_kernel(double* out, int iters) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  double*        sm   = reinterpret_cast<double*>(smem_raw + (base - smem_u32(smem_raw)));
  for(int i = threadIdx.x; i < 41 * 1024 / 8; i += blockDim.x) sm[i] = 1e-3 * (i % 7);
  __syncthreads();
  const int      lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int      q = lane >> 2, l3 = lane & 3;
  const int      pr = 2 * (q & 3) + (q >> 2);
  const uint32_t lane_const = (uint32_t) (pr * ROW_BYTES + ((((l3 >> 1) ^ (pr & 1)) << 4) | ((l3 & 1) << 3)));
  const uint32_t jx = (uint32_t) (pr >> 1);
  double         acc[TA * TB * 2];
#pragma unroll
  for(int i = 0; i < TA * TB * 2; i++) acc[i] = 0.0;
  uint32_t offa[TA], offb[TB];
#pragma unroll
  for(int i = 0; i < TA; i++) offa[i] = (uint32_t) (((warp * 3 + i) & 31) * 8 * ROW_BYTES) + lane_const;
#pragma unroll
  for(int i = 0; i < TB; i++) offb[i] = (uint32_t) (((warp * 5 + i + 16) & 31) * 8 * ROW_BYTES) + lane_const;
  for(int it = 0; it < iters; it++) {
#pragma unroll
    for(uint32_t j = 0; j < 4; j++) {
      const uint32_t jo = base + ((j ^ jx) << 5) + (uint32_t) ((it & 1) * 64 * ROW_BYTES);
      double         fa[TA], fb[TB];
#pragma unroll
      for(int i = 0; i < TA; i++) fa[i] = lds_f64(jo + offa[i]);
#pragma unroll
      for(int i = 0; i < TB; i++) fb[i] = lds_f64(jo + offb[i]);
#pragma unroll
      for(int x = 0; x < TA; x++)
#pragma unroll
        for(int y = 0; y < TB; y++) dmma884(acc[2 * (x * TB + y)], acc[2 * (x * TB + y) + 1], fa[x], fb[y]);
    }
  }
  double sacc = 0;
#pragma unroll
  for(int i = 0; i < TA * TB * 2; i++) sacc += acc[i];
  out[(int64_t) blockIdx.x * blockDim.x + threadIdx.x] = sacc;
}

// MODE 1: the product kernel's k-step loop shape (rolled, 16 DMMAs per trip, fragments reloaded right after
// their last use) without any barrier; MODE 2: plus the per-slab mbarrier traffic (every lane arrives on an
// "empty" barrier, try_wait on a completed "full" one).  ta/tb select MODE here: (1,0) and (2,0).
template<int MODE>
__global__ void __launch_bounds__(128) mainloop_probe2_kernel(double* out, int iters) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  double*        sm   = reinterpret_cast<double*>(smem_raw + (base - smem_u32(smem_raw)));
  for(int i = threadIdx.x; i < 41 * 1024 / 8; i += blockDim.x) sm[i] = 1e-3 * (i % 7);
  if(threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), blockDim.x);
    mbar_fence_init();
  }
  __syncthreads();
  if(threadIdx.x == 0) mbar_arrive(smem_u32(&bars[0])); // phase 0 of the "full" barrier is complete for good
  __syncthreads();
  const int      lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int      q = frag_row(lane >> 2), l3 = lane & 3;
  const uint32_t lane_const = (uint32_t) (q * ROW_BYTES + ((((l3 >> 1) ^ (q & 1)) << 4) | ((l3 & 1) << 3)));
  const uint32_t jx = (uint32_t) (q >> 1);
  double         acc[32];
#pragma unroll
  for(int i = 0; i < 32; i++) acc[i] = 0.0;
  uint32_t hpp_off[2][2], hhp_off[2][2];
#pragma unroll
  for(int x = 0; x < 2; x++)
#pragma unroll
    for(int y = 0; y < 2; y++) {
      hpp_off[x][y] = (uint32_t) (((warp * 3 + 2 * x + y) & 31) * 8 * ROW_BYTES) + lane_const;
      hhp_off[x][y] = (uint32_t) (((warp * 5 + 2 * x + y + 16) & 31) * 8 * ROW_BYTES) + lane_const;
    }
  double fh[2][2], fg[2][2];
#pragma unroll
  for(int x = 0; x < 2; x++)
#pragma unroll
    for(int y = 0; y < 2; y++) {
      fh[x][y] = lds_f64(base + (jx << 5) + hpp_off[x][y]);
      fg[x][y] = lds_f64(base + (jx << 5) + hhp_off[x][y]);
    }
  uint32_t phase = 0;
  for(int s = 0; s < iters; s++) {
#pragma unroll(MODE == 3 ? 2 : (MODE == 4 ? 4 : 1))
    for(uint32_t j = 0; j < 4; j++) {
      uint32_t jn;
      if(j == 3) {
        if(MODE == 2) mbar_wait(smem_u32(&bars[0]), 0, nullptr, 9);
        jn = base + (jx << 5) + (uint32_t) (((s + 1) & 1) * 64 * ROW_BYTES);
      }
      else jn = base + (((j + 1) ^ jx) << 5) + (uint32_t) ((s & 1) * 64 * ROW_BYTES);
#pragma unroll
      for(int hx = 0; hx < 2; hx++)
#pragma unroll
        for(int ql = 0; ql < 2; ql++) {
#pragma unroll
          for(int ga = 0; ga < 2; ga++)
#pragma unroll
            for(int gb = 0; gb < 2; gb++) {
              const int ai = ((((hx * 2 + ga) * 2 + gb) * 2 + ql) * 2);
              dmma884(acc[ai], acc[ai + 1], fh[hx][ql], fg[ga][gb]);
              if(hx == 1 && ql == 1) fg[ga][gb] = lds_f64(jn + hhp_off[ga][gb]);
            }
          fh[hx][ql] = lds_f64(jn + hpp_off[hx][ql]);
        }
    }
    if(MODE == 2) {
      mbar_arrive(smem_u32(&bars[1]));
      phase ^= 1u;
    }
  }
  double sacc = (double) phase;
#pragma unroll
  for(int i = 0; i < 32; i++) sacc += acc[i];
  out[(int64_t) blockIdx.x * blockDim.x + threadIdx.x] = sacc;
}

cudaError_t probe_mainloop(int ta, int tb, int warps_per_cta, int ctas_per_sm, int iters, double* tflops) {
  cudaDeviceProp prop;
  int            dev = 0;
  cudaGetDevice(&dev);
  cudaError_t err = cudaGetDeviceProperties(&prop, dev);
  if(err != cudaSuccess) return err;
  cudaGetLastError();
  void (*k)(double*, int) = nullptr;
  if(ta == 4 && tb == 4) k = mainloop_probe_kernel<4, 4>;
  else if(ta == 1 && tb == 0) k = mainloop_probe2_kernel<1>;
  else if(ta == 2 && tb == 0) k = mainloop_probe2_kernel<2>;
  else if(ta == 3 && tb == 0) k = mainloop_probe2_kernel<3>;
  else if(ta == 5 && tb == 0) k = mainloop_probe2_kernel<4>;
  else return cudaErrorInvalidValue;
  if(tb == 0 && warps_per_cta != 4) return cudaErrorInvalidValue;
  // shared memory sized so that exactly ctas_per_sm CTAs fit
  size_t smem = (size_t) prop.sharedMemPerMultiprocessor / ctas_per_sm - 4096;
  if(smem > prop.sharedMemPerBlockOptin) smem = prop.sharedMemPerBlockOptin;
  if(smem < 43 * 1024) return cudaErrorInvalidValue;
  if((err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)) != cudaSuccess) return err;
  const int grid = prop.multiProcessorCount * ctas_per_sm, block = 32 * warps_per_cta;
  int       occ  = 0;
  if((err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, block, smem)) != cudaSuccess) return err;
  if(occ < ctas_per_sm) return cudaErrorLaunchOutOfResources;
  double* out = nullptr;
  if((err = cudaMalloc(&out, sizeof(double) * grid * block)) != cudaSuccess) return err;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for(int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    k<<<grid, block, smem>>>(out, iters);
    cudaEventRecord(e1);
    if((err = cudaEventSynchronize(e1)) != cudaSuccess) break;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if(rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if(err != cudaSuccess) return err;
  const double tiles = tb == 0 ? 16.0 : (double) ta * tb;
  *tflops = (double) grid * warps_per_cta * (double) iters * 4.0 * tiles * 512.0 / (best * 1e-3) / 1e12;
  return cudaGetLastError();
}

__global__ void dmma_layout_kernel(double* c_out, const double* a, const double* b) {
  const int lane = threadIdx.x;
  double    d0 = 0.0, d1 = 0.0;
  dmma884(d0, d1, a[(lane >> 2) * 4 + (lane & 3)], b[(lane & 3) * 8 + (lane >> 2)]);
  c_out[(lane >> 2) * 8 + 2 * (lane & 3)]     = d0;
  c_out[(lane >> 2) * 8 + 2 * (lane & 3) + 1] = d1;
}

cudaError_t probe_dmma_layout(double* c_host, const double* a_host, const double* b_host) {
  double *    a, *b, *c;
  cudaError_t err;
  if((err = cudaMalloc(&a, 32 * 8)) != cudaSuccess) return err;
  cudaMalloc(&b, 32 * 8);
  cudaMalloc(&c, 64 * 8);
  cudaMemcpy(a, a_host, 32 * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(b, b_host, 32 * 8, cudaMemcpyHostToDevice);
  dmma_layout_kernel<<<1, 32>>>(c, a, b);
  err = cudaMemcpy(c_host, c, 64 * 8, cudaMemcpyDeviceToHost);
  cudaFree(a), cudaFree(b), cudaFree(c);
  return err;
}

// loads `rows` rows x 16 doubles of a [rows/8][8][16] tensor holding value = row*16 + k through a
// (16,8,2,1)-box / 128B-swizzle tensor map and dumps shared memory verbatim
__global__ void tma_swizzle_kernel(const __grid_constant__ CUtensorMap map, double* dump, int rows,
                                   uint32_t* error_flag) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b    = smem_u32(&bar);
  if(threadIdx.x == 0) {
    mbar_init(b, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if(threadIdx.x == 0) {
    mbar_arrive_expect_tx(b, (uint32_t) (rows * ROW_BYTES));
    for(int r = 0; r < rows; r += 16) tma_load_4d(base + (uint32_t) (r * ROW_BYTES), &map, b, 0, 0, r / 8, 0);
  }
  mbar_wait(b, 0, error_flag, 3);
  const double* s = reinterpret_cast<const double*>(smem_raw + (base - smem_u32(smem_raw)));
  for(int i = threadIdx.x; i < rows * 16; i += blockDim.x) dump[i] = s[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

cudaError_t probe_tma_swizzle(double* dump_host, int rows) {
  if(rows % 16 || rows <= 0 || rows > 512) return cudaErrorInvalidValue;
  void*                           fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t err = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if(err != cudaSuccess || !fn) return err != cudaSuccess ? err : cudaErrorUnknown;
  double *src, *dump;
  uint32_t* flag;
  cudaMalloc(&src, (size_t) rows * 16 * 8);
  cudaMalloc(&dump, (size_t) rows * 16 * 8);
  cudaMalloc(&flag, 16);
  const uint32_t flag_init[4] = {0u, 2000000000u, 0u, 0u}; // error word, then the watchdog limit: 2 s of %globaltimer
  cudaMemcpy(flag, flag_init, 16, cudaMemcpyHostToDevice);
  double* h = (double*) malloc((size_t) rows * 16 * 8);
  for(int i = 0; i < rows * 16; i++) h[i] = (double) i;
  cudaMemcpy(src, h, (size_t) rows * 16 * 8, cudaMemcpyHostToDevice);
  free(h);
  CUtensorMap map;
  cuuint64_t  dims[4]    = {16, 8, (cuuint64_t) (rows / 8), 1};
  cuuint64_t  strides[3] = {128, 1024, (cuuint64_t) rows * 128};
  cuuint32_t  box[4]     = {16, 8, 2, 1};
  cuuint32_t  es[4]      = {1, 1, 1, 1};
  CUresult    r = ((EncodeTiledFn) fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, src, dims, strides, box, es,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if(r != CUDA_SUCCESS) return cudaErrorInvalidValue;
  const size_t smem = (size_t) rows * ROW_BYTES + 1024;
  cudaFuncSetAttribute(tma_swizzle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  tma_swizzle_kernel<<<1, 128, smem>>>(map, dump, rows, flag);
  err = cudaMemcpy(dump_host, dump, (size_t) rows * 16 * 8, cudaMemcpyDeviceToHost);
  cudaFree(src), cudaFree(dump), cudaFree(flag);
  return err;
}

__global__ void synth_block_kernel(SynthInfo si, int tensor, int64_t l0, int64_t l1, int64_t l2, int64_t l3,
                                   int64_t n0, int64_t n1, int64_t n2, int64_t n3, double* out) {
  const int64_t total = n0 * n1 * n2 * n3;
  for(int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t) gridDim.x * blockDim.x) {
    int64_t r = e;
    int64_t idx[4];
    idx[3] = l3 + r % n3; r /= n3;
    idx[2] = l2 + r % n2; r /= n2;
    idx[1] = l1 + r % n1; r /= n1;
    idx[0] = l0 + r;
    out[e] = synth_value(si, tensor, idx);
  }
}

cudaError_t synth_block_device(SynthInfo si, int tensor, const int64_t lo[4], const int64_t n[4], double* host_out) {
  const int64_t total = n[0] * n[1] * n[2] * n[3];
  double*       d;
  cudaError_t   err = cudaMalloc(&d, (size_t) total * 8);
  if(err != cudaSuccess) return err;
  synth_block_kernel<<<(unsigned) ((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256), 256>>>(
    si, tensor, lo[0], lo[1], lo[2], lo[3], n[0], n[1], n[2], n[3], d);
  err = cudaMemcpy(host_out, d, (size_t) total * 8, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return err;
}

} // namespace ccsdt
