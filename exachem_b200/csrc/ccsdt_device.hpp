// Device-facing data structures shared by the host driver (ccsdt_capi.cu) and the kernels
// (ccsdt_kernels.cu).  See DESIGN.md for the HBM layout these describe.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace ccsdt {

// Index ids: 0=h1 1=h2 2=h3 3=p4 4=p5 5=p6.

// Operand panels.  Every d1/d2 contraction reads two K-major panels:
//   HPP[hole][particle qt][particle inner][K]      (one hole, two particles)
//   HHP[hole o2][hole o1][particle inner][K]       (two holes o2<o1, one particle)
// stored with pool-wide padded strides so that four TMA tensor maps (d1/d2 x HPP/HHP) cover
// every task.  K is padded with zeros to a multiple of KSLAB.
constexpr int KSLAB      = 16;  // doubles per K slab = one 128-byte swizzle row
constexpr int ROW_BYTES  = 128; // KSLAB * 8
constexpr int PBOX       = 8;   // particle extent of a CTA box (= DMMA m and n)
constexpr int MAX_TERMS  = 18;
constexpr int MAX_STAGES = 8;
constexpr int COUNTER_WORDS = 1 + 256; // box counter + per-SM start counters

struct PoolGeom {
  int     THp, TPp; // padded max hole / particle tile extent (multiples of 4 / 8)
  int     Kp[2];    // padded contraction length of the d1 (occupied) and d2 (virtual) pools
  double* hpp[2];   // [9][THp][TPp][TPp][Kp]
  double* hhp[2];   // [9][THp][THp][TPp][Kp]
  __host__ __device__ int64_t hpp_stride_in(int pool) const { return Kp[pool]; }
  __host__ __device__ int64_t hpp_stride_o1(int pool) const { return (int64_t) TPp * Kp[pool]; }
  __host__ __device__ int64_t hpp_stride_o2(int pool) const { return (int64_t) TPp * TPp * Kp[pool]; }
  __host__ __device__ int64_t hpp_stride_panel(int pool) const { return (int64_t) THp * TPp * TPp * Kp[pool]; }
  __host__ __device__ int64_t hhp_stride_in(int pool) const { return Kp[pool]; }
  __host__ __device__ int64_t hhp_stride_o1(int pool) const { return (int64_t) TPp * Kp[pool]; }
  __host__ __device__ int64_t hhp_stride_o2(int pool) const { return (int64_t) THp * TPp * Kp[pool]; }
  __host__ __device__ int64_t hhp_stride_panel(int pool) const { return (int64_t) THp * THp * TPp * Kp[pool]; }
};

// One enabled d1/d2 contraction of a task, in the form the fused kernel consumes.
// Accumulator layout X: DMMA rows = p6, cols = p5, tile particle qt = p4.
//                    Y: DMMA rows = p6, cols = p4, tile particle qt = p5.
// The a-side panel (DMMA "A", rows) is the one whose inner particle is p6.
struct TermDev {
  int32_t pool;      // 0 = d1, 1 = d2
  int32_t kslabs;    // number of KSLAB-wide slabs
  int32_t ksteps_last; // 4-wide DMMA k-steps of the last slab (1..4): K is padded to 4, not 16, in compute
  int32_t layout_y;  // 0 = X, 1 = Y
  int32_t a_is_hpp;  // 1: HPP is the a-side, 0: HHP is
  int32_t hpp_hole;  // hole index (0..2) of HPP dim o2
  int32_t hhp_o2, hhp_o1;       // hole indices of HHP dims, hhp_o2 < hhp_o1
  int32_t inner_hpp, inner_hhp; // particle index ids (3..5)
  int32_t qt;                   // particle index id of HPP dim o1
  int32_t hpp_panel, hhp_panel; // panel slot inside the pool
};

// One enabled s1 term: s += a[sum coord*sa] * b[sum coord*sb]  (sign folded into a)
struct S1Dev {
  const double* a;
  const double* b;
  int32_t       sa[6], sb[6]; // element strides per index id (0 where the operand lacks the index)
  int32_t       hx, pa;       // index ids of the two indices a carries (hole 0..2, particle 3..5)
};

struct alignas(64) TaskParams {
  CUtensorMap tmap[4]; // [pool*2 + 0] = HPP box (KSLAB,8,8,1), [pool*2 + 1] = HHP box (KSLAB,8,2,1)
  PoolGeom    geom;
  int32_t     ext[6];  // tile extents of h1,h2,h3,p4,p5,p6
  int32_t     nbox[6]; // boxes per index
  int32_t     c[3];    // hole extents of a CTA box = 2*sub
  int32_t     sub[3];
  int32_t     nterms, nterms_x; // terms [0,nterms_x) use layout X, the rest Y
  TermDev     term[MAX_TERMS];
  int32_t     ns1;
  S1Dev       s1[9];
  const double* evl[6]; // orbital-energy slices of the six tiles
  double*     partial;  // [nboxes_padded][2] per-box energy partials (zero for ids that are not boxes);
                        // [nlist][2] in list order for a symmetry-reduced task
  int32_t     nboxes;   // boxes of the tile
  // Box ids are handed out in brick-major order: a brick is a (brick[0..5])-shaped group of boxes that
  // is as close to a 6-d cube (in elements) as the tile allows, so that the boxes in flight at any
  // time share operand rows under all 18 index groupings (L2 reuse).  ids run over the padded grid
  // nbrick*brick; ids whose coordinates fall outside nbox are skipped by the scheduler.
  int32_t     brick[6], nbrick[6];
  int32_t     nboxes_padded;
  uint32_t*   box_counter; // dynamic box scheduler: [0] = next id, [1 + smid] = CTAs that started on that SM;
                           // zeroed before every launch
  int32_t     stagger_cycles; // start delay per co-resident CTA slot (de-phases the epilogues of the CTAs of an SM)
  int32_t     ctas_per_sm;
  int32_t     stages, stage_bytes;
  uint32_t*   error_flag;
  // Permutational symmetry inside a task (options.symmetry): when two hole (particle) TILES of the task
  // coincide, the summand d*d/D and d*(d+s)/D is symmetric under exchange of the two element indices
  // (t3 is antisymmetric in same-spin indices), so only boxes with ascending box coordinates along
  // coinciding indices are evaluated and their partials are weighted by the number of distinct
  // permutations (1, 2, 3 or 6 per index family).  sym bits: 0 = h1~h2, 1 = h2~h3, 2 = p4~p5, 3 = p5~p6,
  // 4 = hole boxes are 2 wide (see box_weight).
  // box_list (device, brick-major order of the surviving ids of the padded grid) replaces the id range.
  int32_t        sym;
  int32_t        nlist;
  const int32_t* box_list;
};

// weight of a box under the task's symmetry bits; 0 = the box is a mirror image and is skipped.
// sym bit 4: the hole boxes are 2 wide, so a box on the triple diagonal of three coinciding hole tiles only
// holds elements with a repeated hole index, where the antisymmetric t3 vanishes: skipped as well.
__host__ __device__ inline int box_weight(int sym, const int bi[6]) {
  int w = 1;
  for(int f = 0; f < 2; f++) {
    const int  a = bi[3 * f], b = bi[3 * f + 1], c = bi[3 * f + 2];
    const bool s01 = (sym >> (2 * f)) & 1, s12 = (sym >> (2 * f + 1)) & 1;
    if(s01 && a > b) return 0;
    if(s12 && b > c) return 0;
    if(f == 0 && (sym & 16) && s01 && s12 && a == b && b == c) return 0;
    if(s01 && s12) w *= (a < b && b < c) ? 6 : ((a < b || b < c) ? 3 : 1);
    else if(s01) w *= a < b ? 2 : 1;
    else if(s12) w *= b < c ? 2 : 1;
  }
  return w;
}

// Panel-build work item: dst[o2][o1][in][k] = scale * src[...]  (or a procedural value)
struct GatherDesc {
  double*       dst;
  const double* src;         // NULL with synth_tensor < 0 writes zeros
  int64_t       ds[4], ss[4]; // element strides, dims ordered (o2, o1, in, k)
  int32_t       n[4];
  double        scale;
  int32_t       synth_tensor; // >= 0: value = synth(tensor, indices); src unused
  int32_t       gpos[4];      // tensor index position fed by each dst dim
  int64_t       gbase[4];     // global orbital index of element 0 of each dst dim
  int32_t       pad_;
};

struct SynthInfo {
  uint64_t seed;
  int32_t  noa, nob, nva, nvb; // ORBITAL counts per spin
};

// ---- launchers (ccsdt_kernels.cu) ----
cudaError_t launch_gather(const GatherDesc* dev_descs, int ndesc, int64_t max_elems, SynthInfo si,
                          cudaStream_t st);
cudaError_t launch_fused_dmma(const TaskParams& p, int grid, int consumer_warps, size_t smem_bytes,
                              cudaStream_t st);
cudaError_t launch_fused_simple(const TaskParams& p, cudaStream_t st, int* grid_out);
cudaError_t launch_reduce_partials(const double* partial, int n, double* out2, cudaStream_t st);
cudaError_t launch_zero(double* a, int64_t na, uint32_t* b, int nb, cudaStream_t st);
cudaError_t launch_copy_from_pinned(const void* pinned_host, void* dev, size_t bytes, cudaStream_t st);
cudaError_t launch_antisym_check(const double* A, const int64_t n[4], const int64_t st[4], int pair, uint32_t* flag, cudaStream_t stream);
cudaError_t fused_dmma_configure(size_t smem_bytes);
int         fused_dmma_max_ctas_per_sm(int threads, size_t smem_bytes);

cudaError_t probe_fp64_peak(int use_dmma, int iters, double* tflops, double* ms);
cudaError_t probe_mainloop(int ta, int tb, int warps_per_cta, int ctas_per_sm, int iters, double* tflops);
cudaError_t probe_dmma_layout(double* c_out, const double* a, const double* b);
cudaError_t probe_tma_swizzle(double* smem_dump, int rows);
cudaError_t synth_block_device(SynthInfo si, int tensor, const int64_t lo[4], const int64_t n[4],
                               double* host_out);

// counter-based synthetic tensors (must match exachem_b200/synthetic.py bit for bit)
__host__ __device__ inline uint64_t synth_mix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__host__ __device__ inline double synth_raw(uint64_t seed, int tensor, int64_t i0, int64_t i1,
                                            int64_t i2, int64_t i3) {
  const double   scale[5] = {0.05, 0.1, 0.1, 0.1, 0.1};
  const uint64_t key = ((uint64_t) i0 << 48) | ((uint64_t) i1 << 32) | ((uint64_t) i2 << 16) | (uint64_t) i3;
  const uint64_t s   = synth_mix(seed ^ ((uint64_t) (tensor + 1) * 0xD1B54A32D192ED03ull));
  const uint64_t h   = synth_mix(s ^ key);
  const double   u   = (double) (h >> 11) * (1.0 / 9007199254740992.0);
  return (2.0 * u - 1.0) * scale[tensor];
}
// idx in the tensor's own index order; occupied/virtual spin-orbital numbers (alpha first)
__host__ __device__ inline double synth_value(const SynthInfo& si, int tensor, const int64_t idx[4]) {
  auto so = [&](int64_t i) { return i < si.noa ? 1 : 2; };
  auto sv = [&](int64_t a) { return a < si.nva ? 1 : 2; };
  int64_t a = idx[0], b = idx[1], c = idx[2], d = idx[3];
  double  sign = 1.0;
  switch(tensor) {
    case 0: // T1[a,i]
      if(sv(a) != so(b)) return 0.0;
      return synth_raw(si.seed, 0, a, b, 0, 0);
    case 1: // T2[a,b,i,j]
      if(a == b || c == d || sv(a) + sv(b) != so(c) + so(d)) return 0.0;
      if(a > b) { int64_t t = a; a = b; b = t; sign = -sign; }
      if(c > d) { int64_t t = c; c = d; d = t; sign = -sign; }
      return sign * synth_raw(si.seed, 1, a, b, c, d);
    case 2: // v2ijab[i,j,a,b]
      if(a == b || c == d || so(a) + so(b) != sv(c) + sv(d)) return 0.0;
      if(a > b) { int64_t t = a; a = b; b = t; sign = -sign; }
      if(c > d) { int64_t t = c; c = d; d = t; sign = -sign; }
      return sign * synth_raw(si.seed, 2, a, b, c, d);
    case 3: // v2ijka[i,j,k,a]
      if(a == b || so(a) + so(b) != so(c) + sv(d)) return 0.0;
      if(a > b) { int64_t t = a; a = b; b = t; sign = -sign; }
      return sign * synth_raw(si.seed, 3, a, b, c, d);
    default: // v2iabc[i,a,b,c]
      if(c == d || so(a) + sv(b) != sv(c) + sv(d)) return 0.0;
      if(c > d) { int64_t t = c; c = d; d = t; sign = -sign; }
      return sign * synth_raw(si.seed, 4, a, b, c, d);
  }
}

} // namespace ccsdt
