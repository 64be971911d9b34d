// Host-side model of the (T) task space: tiles, spins, the canonical task list, which of the
// 9 s1 + 9 d1 + 9 d2 contractions a task executes, the reference's flop count, and the static
// cost-balanced split over ranks.  Pure C++17 (no CUDA) so that it is testable on a CPU-only box.
//
// Index ids everywhere: 0=h1 1=h2 2=h3 3=p4 4=p5 5=p6.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace ccsdt {

struct Space {
  int                  noa = 0, nob = 0, nva = 0, nvb = 0;
  std::vector<int64_t> k_range, k_offset; // per tile, all tiles (occ a | occ b | virt a | virt b)
  std::vector<int32_t> k_spin;            // 1 = alpha, 2 = beta
  std::vector<double>  evl;               // orbital energies, spin-orbital order of the tiles
  bool                 restricted = true;

  int     noab() const { return noa + nob; }
  int     nvab() const { return nva + nvb; }
  int64_t n_occ() const { return k_offset[noab()]; }
  int64_t n_virt() const { return k_offset.back() - k_offset[noab()]; }
  int64_t max_hole_tile() const;
  int64_t max_particle_tile() const;
  // orbitals of one spin inside the occupied (particle=false) or virtual range: contiguous tiles
  void    spin_range(bool particle, int spin, int& tile_begin, int& tile_end, int64_t& n_orb) const;
  std::string validate() const; // empty = ok
};

Space make_space(int noa, int nob, int nva, int nvb, const int64_t* k_range, const int32_t* k_spin,
                 const double* evl, bool restricted);

// setupMOIS(triples=true) tiling + the half/half spin labels (see include/ccsdt_b200.h)
int make_tiles(int64_t n_occ_alpha, int64_t n_occ_beta, int64_t n_vir_alpha, int64_t n_vir_beta,
               int64_t tilesize, std::vector<int64_t>& k_range, std::vector<int32_t>& k_spin,
               int32_t counts[4]);

struct Task {
  int32_t t[6];  // tile ids h1b,h2b,h3b,p4b,p5b,p6b
  int64_t outer; // id of the (h1b,p4b,h2b,p5b,p6b) 5-tuple: the reference's unit of distribution
  double  factor;
};

std::vector<Task> enumerate_tasks(int noab, int nvab, const int32_t* k_spin, bool restricted,
                                  int64_t* n_outer);

// ---- the 27 contractions ---------------------------------------------------------------------
// d1_k:  t3 += sign * sum_l  T2[pa,pb,hx,l] * v2ijka[hy,hz,l,pc]     (hy<hz the other holes)
// d2_k:  t3 += sign * sum_d  T2[pa,d,hx,hy] * v2iabc[hz,d,pb,pc]     (hx<hy, pb<pc)
// s1_k:  t3s+= sign * T1[pa,hx] * v2ijab[hz,hy,pc,pb]                (hy<hz, pb<pc)
struct D1Term {
  int pa, pb, hx, hy, hz, pc, sign;
};
struct D2Term {
  int pa, hx, hy, hz, pb, pc, sign;
};
struct S1Term {
  int pa, hx, hy, hz, pb, pc, sign;
};
extern const D1Term kD1[9];
extern const D2Term kD2[9];
extern const S1Term kS1[9];

// spin of the contracted index required by a d1/d2 term of this task (0 = term disabled)
int d1_contracted_spin(const Space& sp, const Task& t, int k);
int d2_contracted_spin(const Space& sp, const Task& t, int k);
bool s1_enabled(const Space& sp, const Task& t, int k);
bool task_nonempty(const Space& sp, const Task& t);

// enabled masks in the reference's exec-table layout: s1[9], d1[k + 9*h7b], d2[k + 9*(p7b-noab)]
void task_terms(const Space& sp, const Task& t, uint8_t* s1_on, uint8_t* d1_on, uint8_t* d2_on);

// flops of one task / of the whole list, as the reference counts them
long double task_ops(const Space& sp, const Task& t);
long double count_ops(const Space& sp);

// Fraction of a task's CTA boxes the fused kernel evaluates when it exploits the permutational symmetry of
// coinciding tiles (boxes of hbox[i] elements per hole index and pbox per particle index): with n boxes
// along two coinciding indices n(n+1)/2 of n^2 survive, along three n(n+1)(n+2)/6 of n^3.
double symmetry_fraction(const Space& sp, const Task& t, const int hbox[3], int pbox);
// cost model of one task: task_ops, times symmetry_fraction for the default boxes (2,2,2,8,8,8) if asked
long double task_cost(const Space& sp, const Task& t, bool symmetry);

// static split: longest-processing-time greedy on task_cost; owner[i] in [0,nranks)
std::vector<int32_t> partition_tasks(const Space& sp, const std::vector<Task>& tasks, int nranks,
                                     bool symmetry = true);

// ---- execution tiling --------------------------------------------------------------------------
// The STORAGE tiling is the caller's (setupMOIS: the block ids Tensor::get understands).  The EXECUTION tiling is the
// one the task list, the panels and the kernel work on.  They coincide by default (per-task parity with the
// reference); otherwise every spin block is re-cut into near-equal tiles whose extents are multiples of the CTA box
// (8 for particles, 2 for holes) so that no task pads a ragged tile up to the box: ts 28 (caffeine, H2O) executes
// 32^3 / 28^3 = x1.49 DMMAs per particle triple on its own tiles and x1.08 on execution tiles of 40.
// target > 0: requested execution tile extent; target < 0 (auto): max(40, storage tile rounded up to 8), and the
// storage tiling is kept as it is when all its particle tiles but the last of each spin block are multiples of 8.
Space make_exec_space(const Space& store, int target);
// true when a and b describe the same tiles
bool  same_tiling(const Space& a, const Space& b);

// One tensor dimension of an execution tile cut along the storage tiles it overlaps.
struct TilePiece {
  int32_t store_tile; // tile id in the storage space (global id: particle tiles include the noab offset)
  int64_t store_off;  // first element inside the storage tile
  int64_t exec_off;   // first element inside the execution tile
  int64_t len;
};
// particle = the dimension is virtual; exec_tile / store_tile are global tile ids of their spaces
std::vector<TilePiece> split_tile(const Space& exec, const Space& store, int exec_tile);

// ---- device memory the path needs (host-only arithmetic; mirrors ensure_pools of ccsdt_capi.cu) -----------------
// What the reference prints before the (T) loop is the HOST memory of its staging buffers and block caches
// (exachem/cc/ccsd_t/ccsd_t.cpp:95-152); here the staging lives in HBM.
struct MemoryEstimate {
  int64_t thp, tpp, kp_occ, kp_virt; // padded panel extents: hole / particle tile, contraction lengths (d1 / d2)
  int64_t panel_bytes;               // K-major operand panels (d1 + d2, HPP + HHP, 9 slots each), all staging buffers
  int64_t s1_bytes;                  // staged s1 operands, all staging buffers
  int64_t task_block_bytes;          // upper bound of the tensor blocks ONE task reads (what must be resident to stage it)
  int64_t tensor_bytes[5];           // spin-conserving blocks of T1, T2, v2ijab, v2ijka, v2iabc: a fully resident block store
  int64_t minimum_bytes;             // panels + s1 + blocks of the task in flight and of the task being staged
};
MemoryEstimate estimate_memory(const Space& exec, const int sub[3], int staging_buffers);

// The blocks the reference requests are canonically ordered: T2{p_lo,p_hi,h_lo,h_hi},
// v2ijka{h_lo,h_hi,h7,p}, v2iabc{h,p7,p_lo,p_hi}, v2ijab{h_hi,h_lo,p_hi,p_lo}.  A storage sub-block of a diagonal
// execution block may come out in the other order; canonical_block swaps the offending pair(s) of `bid`, records the
// dimension permutation (requested dim d is dim perm[d] of the canonical block) and returns the sign (-1)^swaps.
// tensor ids as in include/ccsdt_b200.h.
int canonical_block(int tensor, uint32_t bid[4], int perm[4]);

} // namespace ccsdt
