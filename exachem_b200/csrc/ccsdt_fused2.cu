// fused_t_dmma2_kernel -- the product kernel, second generation.
//
// Same math, operand layout and accumulator mapping as fused_t_dmma_kernel (ccsdt_kernels.cu), but
// without a dedicated producer warp: every warp of the CTA issues DMMAs, and lane 0 of warp 0 also
// "pumps" the TMA ring (it issues the loads of future slabs whenever a ring stage has been released).
// That keeps the whole CTA at a multiple of four warps, so two 8-warp CTAs (or four 4-warp CTAs) fit
// one SM at the full 128 registers per thread.  Two independent CTAs per SM hide each other's
// epilogues, relayouts and slab-boundary bubbles; measured effect in DESIGN.md.
#include "ccsdt_kernel_common.cuh"

namespace ccsdt {

struct PumpState { // shared memory; only lane 0 of warp 0 touches it
  int      box, term, slab;
  uint32_t issued;
  BoxCoord bc;
};

// Issue as many pending slabs as the ring allows.  `consumed` = slabs warp 0 has finished (so slab
// number `consumed` is the one it is about to read); the cursor never runs more than `stages` ahead.
// Never blocks unless the slab warp 0 needs next has not been issued yet.
__device__ __forceinline__ void pump(const TaskParams& p, PumpState& ps, uint32_t consumed, int cons_box,
                                     bool cons_in_y, bool relayout, uint32_t ring_base, uint32_t full_bar,
                                     uint32_t empty_bar) {
  uint32_t       issued = ps.issued;
  const uint32_t stages = (uint32_t) p.stages;
  if(issued >= consumed + stages || ps.box >= p.nboxes) return;
  int box = ps.box, term = ps.term, slab = ps.slab;
  while(issued < consumed + stages && box < p.nboxes) {
    // the Y terms of a box may only be fetched once the consumers have re-laid-out its accumulators
    // (the relayout parks them in the ring)
    if(relayout && term >= p.nterms_x && !(box == cons_box && cons_in_y)) break;
    const uint32_t stage = issued % stages, use = issued / stages;
    if(use > 0) {
      const uint32_t par = (use - 1u) & 1u;
      if(!mbar_try_wait(empty_bar + 8 * stage, par)) {
        if(issued != consumed) break;                              // not urgent: try again at the next slab
        mbar_wait(empty_bar + 8 * stage, par, p.error_flag, 2);    // warp 0 itself needs this slab now
      }
    }
    const TermDev&     td      = p.term[term];
    const CUtensorMap* map_hpp = &p.tmap[td.pool * 2 + 0];
    const CUtensorMap* map_hhp = &p.tmap[td.pool * 2 + 1];
    const int          c_h = p.c[td.hpp_hole], c_o2 = p.c[td.hhp_o2], c_o1 = p.c[td.hhp_o1];
    const uint32_t     hpp_bytes = (uint32_t) (c_h * 64 * ROW_BYTES);
    const uint32_t     bar       = full_bar + 8 * stage;
    const uint32_t     dst       = ring_base + stage * (uint32_t) p.stage_bytes;
    mbar_arrive_expect_tx(bar, hpp_bytes + (uint32_t) (c_o2 * c_o1 * 8 * ROW_BYTES));
    const int k0     = slab * KSLAB;
    const int hpp_o2 = td.hpp_panel * p.geom.THp + ps.bc.off[td.hpp_hole];
    for(int j = 0; j < c_h; j++)
      tma_load_4d(dst + (uint32_t) (j * 64 * ROW_BYTES), map_hpp, bar, k0, ps.bc.off[td.inner_hpp], ps.bc.off[td.qt],
                  hpp_o2 + j);
    const int hhp_o2 = td.hhp_panel * p.geom.THp + ps.bc.off[td.hhp_o2];
    for(int j2 = 0; j2 < c_o2; j2++)
      for(int j1 = 0; j1 < c_o1; j1 += 2)
        tma_load_4d(dst + hpp_bytes + (uint32_t) ((j2 * c_o1 + j1) * 8 * ROW_BYTES), map_hhp, bar, k0,
                    ps.bc.off[td.inner_hhp], ps.bc.off[td.hhp_o1] + j1, hhp_o2 + j2);
    issued++;
    if(++slab == td.kslabs) {
      slab = 0;
      if(++term == p.nterms) {
        term = 0;
        box += gridDim.x;
        if(box < p.nboxes) ps.bc = decode_box(p, box);
      }
    }
  }
  ps.box = box, ps.term = term, ps.slab = slab, ps.issued = issued;
}

struct ConsumerCtx {
  uint32_t ring_base, full_bar, empty_bar;
  int      wq, lane, warp;
  int      sub_off[3];
  uint32_t consumed; // slabs this warp has finished
  int      box;      // box being processed
  bool     in_y, relayout;
};

template<int HH, bool A_HPP>
__device__ __forceinline__ void consume_term2(double (&acc)[32], const TaskParams& p, const TermDev& td, Ring& ring,
                                              ConsumerCtx& cc, PumpState& ps) {
  constexpr int  HA = (HH == 0) ? 1 : 0; // the two other holes, ascending
  constexpr int  HB = (HH == 2) ? 1 : 2;
  const int      q = frag_row(cc.lane >> 2), l3 = cc.lane & 3;
  const uint32_t lane_const = (uint32_t) (q * ROW_BYTES + ((((l3 >> 1) ^ (q & 1)) << 4) | ((l3 & 1) << 3)));
  const uint32_t jx         = (uint32_t) (q >> 1);
  // fragment (ih, ql) of HPP sits at hpp0 + ih*8192 + ql*1024; fragment (ia, ib) of HHP at
  // hhp0 + ia*hhp_sa + ib*1024 (bytes inside a stage): three registers, the rest are immediates
  const uint32_t hpp0   = (uint32_t) (((cc.sub_off[HH] * 8 + 2 * cc.wq) * 8) * ROW_BYTES) + lane_const;
  const uint32_t hhp_sa = (uint32_t) (p.c[HB] * 8 * ROW_BYTES);
  const uint32_t hhp0 =
    (uint32_t) ((p.c[HH] * 64 + (cc.sub_off[HA] * p.c[HB] + cc.sub_off[HB]) * 8) * ROW_BYTES) + lane_const;

  for(int s = 0; s < td.kslabs; s++) {
    if(cc.warp == 0) {
      if(cc.lane == 0) pump(p, ps, cc.consumed, cc.box, cc.in_y, cc.relayout, cc.ring_base, cc.full_bar, cc.empty_bar);
      __syncwarp();
    }
    mbar_wait(cc.full_bar + 8 * ring.stage, ring.phase, p.error_flag, 1);
    const uint32_t base = cc.ring_base + ring.stage * (uint32_t) p.stage_bytes;
    const uint32_t nj   = (s == td.kslabs - 1) ? (uint32_t) td.ksteps_last : 4u; // K tail: 4-wide steps only
#pragma unroll
    for(uint32_t j = 0; j < 4; j++) {
      if(j >= nj) break;
      const uint32_t jo = base + ((j ^ jx) << 5);
      double         fh[2][2], fg[2][2];
#pragma unroll
      for(int x = 0; x < 2; x++)
#pragma unroll
        for(int y = 0; y < 2; y++) {
          fh[x][y] = lds_f64(jo + hpp0 + (uint32_t) (x * 64 * ROW_BYTES + y * 8 * ROW_BYTES));
          fg[x][y] = lds_f64(jo + hhp0 + (uint32_t) x * hhp_sa + (uint32_t) (y * 8 * ROW_BYTES));
        }
#pragma unroll
      for(int i1 = 0; i1 < 2; i1++)
#pragma unroll
        for(int i2 = 0; i2 < 2; i2++)
#pragma unroll
          for(int i3 = 0; i3 < 2; i3++)
#pragma unroll
            for(int ql = 0; ql < 2; ql++) {
              const int    hi[3] = {i1, i2, i3};
              const double vh    = fh[hi[HH]][ql];
              const double vg    = fg[hi[HA]][hi[HB]];
              const int    ai    = ((((i1 * 2 + i2) * 2 + i3) * 2 + ql) * 2);
              if(A_HPP) dmma884(acc[ai], acc[ai + 1], vh, vg);
              else dmma884(acc[ai], acc[ai + 1], vg, vh);
            }
    }
    __syncwarp();
    if(cc.lane == 0) mbar_arrive(cc.empty_bar + 8 * ring.stage);
    ring.advance((uint32_t) p.stages);
    cc.consumed++;
  }
}

// blockDim.x = 32 * (4 * sub0*sub1*sub2); no producer warp.
template<int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) fused_t_dmma2_kernel(const __grid_constant__ TaskParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * MAX_STAGES];
  __shared__ double    red[2][16];
  __shared__ PumpState ps;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ncw = (int) (blockDim.x >> 5);

  ConsumerCtx cc;
  cc.ring_base = (smem_u32(smem_raw) + 1023u) & ~1023u; // the swizzle is anchored on absolute smem address bits
  cc.full_bar  = smem_u32(&bars[0]);
  cc.empty_bar = smem_u32(&bars[MAX_STAGES]);
  cc.lane = lane, cc.warp = warp, cc.wq = warp & 3;
  cc.consumed = 0;
  cc.relayout = p.nterms_x > 0 && p.nterms_x < p.nterms;

  if(tid == 0) {
    for(int s = 0; s < p.stages; s++) {
      mbar_init(cc.full_bar + 8 * s, 1);
      mbar_init(cc.empty_bar + 8 * s, (uint32_t) ncw);
    }
    mbar_fence_init();
    for(int i = 0; i < 4; i++) tma_prefetch_desc(&p.tmap[i]);
    ps.box = blockIdx.x, ps.term = 0, ps.slab = 0, ps.issued = 0;
    if(ps.box < p.nboxes) ps.bc = decode_box(p, ps.box);
  }
  __syncthreads();

  {
    const int grp = warp >> 2;
    const int g2 = grp % p.sub[2], g1 = (grp / p.sub[2]) % p.sub[1], g0 = grp / (p.sub[2] * p.sub[1]);
    cc.sub_off[0] = 2 * g0, cc.sub_off[1] = 2 * g1, cc.sub_off[2] = 2 * g2;
  }
  const int l3  = lane & 3;
  const int q6  = frag_row(lane >> 2);  // particle offset of this lane's DMMA row
  const int qc0 = frag_row(2 * l3);     // particle offset of DMMA column 2*l3 (+2 for column 2*l3+1)
  Ring      ring{0u, 0u};

  for(int box = blockIdx.x; box < p.nboxes; box += gridDim.x) {
    const BoxCoord bc = decode_box(p, box);
    cc.box  = box;
    cc.in_y = p.nterms_x == 0;
    double acc[32];
#pragma unroll
    for(int i = 0; i < 32; i++) acc[i] = 0.0;

    for(int t = 0; t < p.nterms; t++) {
      if(cc.relayout && t == p.nterms_x) {
        // X -> Y: swap the roles of p4 and p5 (tile particle <-> DMMA column) through shared memory.
        // All X slabs of this box are consumed and the pump cannot have issued anything later, so the
        // ring is idle.  scratch index = ((((h1*c2 + h2)*c3 + h3)*8 + p4)*8 + p5)*8 + p6 (box-local)
        __syncthreads();
        double* scratch = reinterpret_cast<double*>(smem_raw + (cc.ring_base - smem_u32(smem_raw)));
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
                  const int hl = ((cc.sub_off[0] + i1) * p.c[1] + cc.sub_off[1] + i2) * p.c[2] + cc.sub_off[2] + i3;
                  const int p4 = 2 * cc.wq + ql, p5 = qc0 + 2 * r;
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
                  const int hl = ((cc.sub_off[0] + i1) * p.c[1] + cc.sub_off[1] + i2) * p.c[2] + cc.sub_off[2] + i3;
                  const int p5 = 2 * cc.wq + ql, p4 = qc0 + 2 * r;
                  acc[((((i1 * 2 + i2) * 2 + i3) * 2 + ql) * 2) + r] = scratch[((hl * 8 + p4) * 8 + p5) * 8 + q6];
                }
        // generic-proxy accesses to the ring must be ordered before the TMA (async proxy) refills it
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        cc.in_y = true;
      }
      const TermDev& td = p.term[t];
      switch(td.hpp_hole * 2 + td.a_is_hpp) {
        case 0: consume_term2<0, false>(acc, p, td, ring, cc, ps); break;
        case 1: consume_term2<0, true>(acc, p, td, ring, cc, ps); break;
        case 2: consume_term2<1, false>(acc, p, td, ring, cc, ps); break;
        case 3: consume_term2<1, true>(acc, p, td, ring, cc, ps); break;
        case 4: consume_term2<2, false>(acc, p, td, ring, cc, ps); break;
        default: consume_term2<2, true>(acc, p, td, ring, cc, ps); break;
      }
    }
    // keep the ring filling while this CTA is busy with its epilogue
    if(warp == 0) {
      if(lane == 0) pump(p, ps, cc.consumed, cc.box, cc.in_y, cc.relayout, cc.ring_base, cc.full_bar, cc.empty_bar);
      __syncwarp();
    }

    // ---------------- epilogue: denominators, E[T], then the s1 part of E(T) ----------------
    const bool layout_y = p.nterms > p.nterms_x;
    const int  id_qt = layout_y ? 4 : 3, id_qc = layout_y ? 3 : 4;
    int        hc[3], tcq, ccq;
    hc[0] = bc.off[0] + cc.sub_off[0], hc[1] = bc.off[1] + cc.sub_off[1], hc[2] = bc.off[2] + cc.sub_off[2];
    tcq = (layout_y ? bc.off[4] : bc.off[3]) + 2 * cc.wq;
    ccq = (layout_y ? bc.off[3] : bc.off[4]) + qc0;
    const int c6 = bc.off[5] + q6;
    double    eh[3][2], et[2], ec[2];
    unsigned  vmask = 0; // bit layout: [h1:2][h2:2][h3:2][qt:2][qc:2]
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
              const double   tq   = ok ? d / D : 0.0;
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
      const int da[5] = {sd.sa[0], sd.sa[1], sd.sa[2], sd.sa[id_qt], 2 * sd.sa[id_qc]};
      const int db[5] = {sd.sb[0], sd.sb[1], sd.sb[2], sd.sb[id_qt], 2 * sd.sb[id_qc]};
      double    acc_s = 0.0;
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
                const int ai = ((((i1 * 2 + i2) * 2 + i3) * 2 + ql) * 2) + r;
                const int oa = i1 * da[0] + i2 * da[1] + i3 * da[2] + ql * da[3] + r * da[4];
                const int ob = i1 * db[0] + i2 * db[1] + i3 * db[2] + ql * db[3] + r * db[4];
                acc_s += acc[ai] * (__ldg(pa + oa) * __ldg(pb + ob));
              }
      e2 += acc_s;
    }
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) {
      e1 += __shfl_xor_sync(0xffffffffu, e1, o);
      e2 += __shfl_xor_sync(0xffffffffu, e2, o);
    }
    if(lane == 0) red[0][warp] = e1, red[1][warp] = e2;
    __syncthreads();
    if(tid == 0) {
      double a = 0.0, b = 0.0;
      for(int i = 0; i < ncw; i++) a += red[0][i], b += red[1][i];
      p.partial[2 * (int64_t) box]     = a;
      p.partial[2 * (int64_t) box + 1] = b;
    }
    __syncthreads(); // red[] may be overwritten by the next box
  }
}

cudaError_t fused_dmma2_configure(size_t smem_bytes) {
  cudaError_t e;
  if((e = cudaFuncSetAttribute(fused_t_dmma2_kernel<128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int) smem_bytes)) != cudaSuccess)
    return e;
  if((e = cudaFuncSetAttribute(fused_t_dmma2_kernel<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int) smem_bytes)) != cudaSuccess)
    return e;
  return cudaFuncSetAttribute(fused_t_dmma2_kernel<384, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int) smem_bytes);
}

cudaError_t launch_fused_dmma2(const TaskParams& p, int grid, int warps, size_t smem_bytes, cudaStream_t st) {
  const int threads = 32 * warps;
  if(warps == 4) fused_t_dmma2_kernel<128, 4><<<grid, threads, smem_bytes, st>>>(p);
  else if(warps == 8) fused_t_dmma2_kernel<256, 2><<<grid, threads, smem_bytes, st>>>(p);
  else if(warps == 12) fused_t_dmma2_kernel<384, 1><<<grid, threads, smem_bytes, st>>>(p);
  else return cudaErrorInvalidConfiguration;
  return cudaGetLastError();
}

} // namespace ccsdt
