/*
 * TEST INFRASTRUCTURE ONLY.  CPU restatement (the "oracle") of ExaChem's fused CCSD(T) triples path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this file's library.  The product (exachem_b200/csrc) never links, imports or calls it.
 *
 * Parity is PINNED: tests/test_oracle.py compares every function here against the reference's own
 * code compiled unmodified (oracle/_ref/libccsdt_ref.so, see oracle/ref_driver.cpp) -- task lists and
 * exec tables bit-exact, energies bit-exact (same operation order, same compiler flags) -- and checks
 * the six integer known-answer `total_num_ops` values held by the reference's CI goldens
 * (ci/reference_output/ *.ccsd_t.json) plus the fixtures in tests/golden/ generated from the reference
 * library (tests/golden/make_golden.py); tests/test_molecules.py and tests/test_provider.py add the
 * reference's published [T] / (T) energies (butanol2, CH4 UHF) on amplitudes from tools/provider.
 *
 * All paths below are relative to /root/reference/exachem/.  Written from scratch in plain C99; the
 * reference is C++ templates over TAMM tensors.
 *
 * Index ids used in the tables: 0=h1 1=h2 2=h3 3=p4 4=p5 5=p6.
 * Dense inputs (row-major, spin-orbital, tile order | occ a | occ b | virt a | virt b |):
 *   T1[Vt][Ot]  T2[Vt][Vt][Ot][Ot]  v2ijab[Ot][Ot][Vt][Vt]  v2ijka[Ot][Ot][Ot][Vt]  v2iabc[Ot][Vt][Vt][Vt]
 * (declaration order of the TAMM tensors, cc/ccsd_t/ccsd_t.cpp:66-69).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* Tiling: cholesky/cholesky_2e.cpp:186-190,230-279 (triples => balance_tiles=false, tile =
 * ccsdt_tilesize, no active space) and the half/half spin labelling of cc/ccsd_t/ccsd_t.cpp:245-249. */
static int push_tiles(int64_t n, int64_t ts, int64_t* out) {
  int     cnt  = 0;
  int64_t full = n / ts, last = n % ts;
  for(int64_t x = 0; x < full; x++) out[cnt++] = ts;
  if(last > 0) out[cnt++] = last;
  return cnt;
}

ORC_API int orc_tiles(int64_t n_occ_alpha, int64_t n_occ_beta, int64_t n_vir_alpha,
                      int64_t n_vir_beta, int64_t tilesize, int64_t* k_range, int32_t* k_spin,
                      int32_t* counts /* noa, nob, nva, nvb */) {
  int n      = 0;
  counts[0]  = push_tiles(n_occ_alpha, tilesize, k_range + n);
  n += counts[0];
  counts[1] = push_tiles(n_occ_beta, tilesize, k_range + n);
  n += counts[1];
  counts[2] = push_tiles(n_vir_alpha, tilesize, k_range + n);
  n += counts[2];
  counts[3] = push_tiles(n_vir_beta, tilesize, k_range + n);
  n += counts[3];
  int noab = counts[0] + counts[1], nvab = counts[2] + counts[3];
  int k    = 0;
  /* The reference labels the first noab/2 (nvab/2) tiles alpha and the rest beta (ccsd_t.cpp:245-249), which is the
   * actual tile layout only when both spins have the same number of tiles (or beta has one more).  Here the labels
   * follow the tile counts, so open-shell spaces with different numbers of alpha and beta tiles are labelled
   * correctly (row f5 of SURVEY.md 8f); whenever the reference's split is valid the two agree. */
  (void) noab; (void) nvab;
  for(int g = 0; g < 4; g++)
    for(int x = 0; x < counts[g]; x++) k_spin[k++] = (g % 2) ? 2 : 1;
  return n;
}

/* ------------------------------------------------------------------------------------------ */
/* Task enumeration: cc/ccsd_t/ccsd_t_fused_driver.hpp:368-483 (loop order h1b,p4b,h2b,p5b,p6b then
 * h3b; spin filters :383-386; factor :387-395).  tasks = 7 int64 per kernel task:
 * h1b,h2b,h3b,p4b,p5b,p6b,outer_task_id(taskcount).  Returns the number of kernel tasks; writes at
 * most cap of them.  *n_outer = number of 5-tuples (the unit the reference distributes). */
static double task_factor(int restricted, int64_t h1, int64_t h2, int64_t h3, int64_t p4, int64_t p5,
                          int64_t p6) {
  double factor = restricted ? 2.0 : 1.0;
  if(p4 == p5 && p5 == p6) factor /= 6.0;
  else if(p4 == p5 || p5 == p6) factor /= 2.0;
  if(h1 == h2 && h2 == h3) factor /= 6.0;
  else if(h1 == h2 || h2 == h3) factor /= 2.0;
  return factor;
}

ORC_API int64_t orc_enumerate(int noab, int nvab, const int32_t* k_spin, int is_restricted,
                              int64_t* tasks, double* factors, int64_t cap, int64_t* n_outer) {
  int64_t n = 0, taskcount = 0;
  for(int64_t h1 = 0; h1 < noab; h1++)
    for(int64_t p4 = noab; p4 < noab + nvab; p4++)
      for(int64_t h2 = h1; h2 < noab; h2++)
        for(int64_t p5 = p4; p5 < noab + nvab; p5++)
          for(int64_t p6 = p5; p6 < noab + nvab; p6++) {
            for(int64_t h3 = h2; h3 < noab; h3++) {
              int sp = k_spin[p4] + k_spin[p5] + k_spin[p6];
              int sh = k_spin[h1] + k_spin[h2] + k_spin[h3];
              if(sp != sh) continue;
              if(is_restricted && sp + sh > 8) continue;
              if(n < cap) {
                if(tasks) {
                  int64_t* t = tasks + 7 * n;
                  t[0] = h1, t[1] = h2, t[2] = h3, t[3] = p4, t[4] = p5, t[5] = p6, t[6] = taskcount;
                }
                if(factors) factors[n] = task_factor(is_restricted, h1, h2, h3, p4, p5, p6);
              }
              n++;
            }
            taskcount++;
          }
  if(n_outer) *n_outer = taskcount;
  return n;
}

/* ------------------------------------------------------------------------------------------ */
/* Enabled terms / slot tables.
 * Candidate role permutations and the alias ("cur_*") tuples:
 *   s1: cc/ccsd_t/ccsd_t_all_fused_singles.hpp:55-72 (candidates+dedup), :100-119 (enable), :125-177 (alias)
 *   d1: cc/ccsd_t/ccsd_t_all_fused_doubles1.hpp:57-74, :94-117, :121-188
 *   d2: cc/ccsd_t/ccsd_t_all_fused_doubles2.hpp:58-75, :98-121, :125-191
 * A candidate is ref permuted: roles[j] = ref[cand[c][j]] with ref order (p4,p5,p6,h1,h2,h3). The
 * alias tuple k of a candidate is roles permuted by alias[k]; exec[k] is written when it equals ref. */
typedef struct {
  int cand[9][6];
  int alias[9][6];
} Family;

/* positions: 0=p4 1=p5 2=p6 3=h1 4=h2 5=h3 (tuple order of the reference) */
static const Family FAM_S1 = {
  {{0, 1, 2, 3, 4, 5}, {0, 1, 2, 4, 3, 5}, {0, 1, 2, 5, 3, 4}, {1, 0, 2, 3, 4, 5}, {1, 0, 2, 4, 3, 5},
   {1, 0, 2, 5, 3, 4}, {2, 0, 1, 3, 4, 5}, {2, 0, 1, 4, 3, 5}, {2, 0, 1, 5, 3, 4}},
  /* cur_p456_h123, p456_h213, p456_h231, p546_h123, p546_h213, p546_h231, p564_h123, ... */
  {{0, 1, 2, 3, 4, 5}, {0, 1, 2, 4, 3, 5}, {0, 1, 2, 4, 5, 3}, {1, 0, 2, 3, 4, 5}, {1, 0, 2, 4, 3, 5},
   {1, 0, 2, 4, 5, 3}, {1, 2, 0, 3, 4, 5}, {1, 2, 0, 4, 3, 5}, {1, 2, 0, 4, 5, 3}}};
static const Family FAM_D1 = {
  {{0, 1, 2, 3, 4, 5}, {0, 1, 2, 4, 3, 5}, {0, 1, 2, 5, 3, 4}, {1, 2, 0, 3, 4, 5}, {1, 2, 0, 4, 3, 5},
   {1, 2, 0, 5, 3, 4}, {0, 2, 1, 3, 4, 5}, {0, 2, 1, 4, 3, 5}, {0, 2, 1, 5, 3, 4}},
  /* cur_p456_h123, h213, h231, p645_*, p465_* */
  {{0, 1, 2, 3, 4, 5}, {0, 1, 2, 4, 3, 5}, {0, 1, 2, 4, 5, 3}, {2, 0, 1, 3, 4, 5}, {2, 0, 1, 4, 3, 5},
   {2, 0, 1, 4, 5, 3}, {0, 2, 1, 3, 4, 5}, {0, 2, 1, 4, 3, 5}, {0, 2, 1, 4, 5, 3}}};
static const Family FAM_D2 = {
  {{0, 1, 2, 3, 4, 5}, {0, 1, 2, 4, 5, 3}, {0, 1, 2, 3, 5, 4}, {1, 0, 2, 3, 4, 5}, {1, 0, 2, 4, 5, 3},
   {1, 0, 2, 3, 5, 4}, {2, 0, 1, 3, 4, 5}, {2, 0, 1, 4, 5, 3}, {2, 0, 1, 3, 5, 4}},
  /* cur_p456_h123, h312, h132, p546_*, p564_* */
  {{0, 1, 2, 3, 4, 5}, {0, 1, 2, 5, 3, 4}, {0, 1, 2, 3, 5, 4}, {1, 0, 2, 3, 4, 5}, {1, 0, 2, 5, 3, 4},
   {1, 0, 2, 3, 5, 4}, {1, 2, 0, 3, 4, 5}, {1, 2, 0, 5, 3, 4}, {1, 2, 0, 3, 5, 4}}};

typedef struct {
  int64_t roles[9][6]; /* p4b,p5b,p6b,h1b,h2b,h3b of each candidate; all 0 = removed duplicate */
} Cands;

static void build_cands(const Family* f, const int64_t ref[6], Cands* c) {
  for(int i = 0; i < 9; i++)
    for(int j = 0; j < 6; j++) c->roles[i][j] = ref[f->cand[i][j]];
  for(int i = 0; i < 8; i++) {
    if(c->roles[i][0] == 0) continue;
    for(int j = i + 1; j < 9; j++)
      if(memcmp(c->roles[i], c->roles[j], sizeof(c->roles[i])) == 0)
        memset(c->roles[j], 0, sizeof(c->roles[j]));
  }
}

static int alias_hits(const Family* f, const int64_t roles[6], const int64_t ref[6], int k) {
  for(int j = 0; j < 6; j++)
    if(roles[f->alias[k][j]] != ref[j]) return 0;
  return 1;
}

/* common gate of one candidate (ordering condition differs per family: which two pairs must be sorted) */
static int cand_gate(int fam, const int64_t r[6], const int32_t* k_spin, const int64_t* k_range,
                     int is_restricted) {
  int64_t p4b = r[0], p5b = r[1], p6b = r[2], h1b = r[3], h2b = r[4], h3b = r[5];
  if(p4b == 0) return 0;
  if(fam == 0 && !(p5b <= p6b && h2b <= h3b)) return 0; /* singles.hpp:103 */
  if(fam == 1 && !(p4b <= p5b && h2b <= h3b)) return 0; /* doubles1.hpp:97 */
  if(fam == 2 && !(p5b <= p6b && h1b <= h2b)) return 0; /* doubles2.hpp:101 */
  int sp = k_spin[p4b] + k_spin[p5b] + k_spin[p6b], sh = k_spin[h1b] + k_spin[h2b] + k_spin[h3b];
  if(is_restricted && sp + sh == 12) return 0;
  if(sp != sh) return 0;
  if(!(k_range[p4b] > 0 && k_range[p5b] > 0 && k_range[p6b] > 0 && k_range[h1b] > 0 &&
       k_range[h2b] > 0 && k_range[h3b] > 0))
    return 0;
  return 1;
}

/* task = h1b,h2b,h3b,p4b,p5b,p6b.  s1_exec[9], d1_exec[9*noab] (index k + 9*h7b),
 * d2_exec[9*nvab] (index k + 9*(p7b-noab)); entries are slot numbers or -1.
 * slot_cand_*[slot] / slot_k_*[slot] (optional, may be NULL) give the candidate and the h7b/p7b of
 * each slot, in slot order.  counts[3] = number of slots (enabled kernels) of s1,d1,d2. */
ORC_API void orc_task_exec(int noab, int nvab, const int32_t* k_spin, const int64_t* k_range,
                           int is_restricted, const int64_t* task, int32_t* s1_exec,
                           int32_t* d1_exec, int32_t* d2_exec, int32_t* counts, int32_t* slot_cand_s1,
                           int32_t* slot_cand_d1, int32_t* slot_k_d1, int32_t* slot_cand_d2,
                           int32_t* slot_k_d2) {
  const int64_t ref[6] = {task[3], task[4], task[5], task[0], task[1], task[2]};
  Cands         c;
  for(int k = 0; k < 9; k++) s1_exec[k] = -1;
  for(int k = 0; k < 9 * noab; k++) d1_exec[k] = -1;
  for(int k = 0; k < 9 * nvab; k++) d2_exec[k] = -1;

  /* s1 */
  build_cands(&FAM_S1, ref, &c);
  int slot = 0;
  for(int ia = 0; ia < 9; ia++) {
    const int64_t* r = c.roles[ia];
    if(!cand_gate(0, r, k_spin, k_range, is_restricted)) continue;
    if(k_spin[r[0]] != k_spin[r[3]]) continue; /* singles.hpp:113 */
    for(int k = 0; k < 9; k++)
      if(alias_hits(&FAM_S1, r, ref, k)) s1_exec[k] = slot;
    if(slot_cand_s1) slot_cand_s1[slot] = ia;
    slot++;
  }
  counts[0] = slot;

  /* d1 */
  build_cands(&FAM_D1, ref, &c);
  slot = 0;
  for(int ia = 0; ia < 9; ia++) {
    const int64_t* r = c.roles[ia];
    if(!cand_gate(1, r, k_spin, k_range, is_restricted)) continue;
    for(int64_t h7b = 0; h7b < noab; h7b++) {
      if(k_spin[r[0]] + k_spin[r[1]] != k_spin[r[3]] + k_spin[h7b]) continue; /* doubles1.hpp:108 */
      if(!(h7b <= r[2])) continue;                                            /* :113 */
      for(int k = 0; k < 9; k++)
        if(alias_hits(&FAM_D1, r, ref, k)) d1_exec[k + 9 * h7b] = slot;
      if(slot_cand_d1) slot_cand_d1[slot] = ia, slot_k_d1[slot] = (int32_t) h7b;
      slot++;
    }
  }
  counts[1] = slot;

  /* d2 */
  build_cands(&FAM_D2, ref, &c);
  slot = 0;
  for(int ia = 0; ia < 9; ia++) {
    const int64_t* r = c.roles[ia];
    if(!cand_gate(2, r, k_spin, k_range, is_restricted)) continue;
    for(int64_t p7b = noab; p7b < noab + nvab; p7b++) {
      if(k_spin[r[0]] + k_spin[p7b] != k_spin[r[3]] + k_spin[r[4]]) continue; /* doubles2.hpp:112 */
      if(!(r[5] <= p7b)) continue;                                            /* :117 */
      for(int k = 0; k < 9; k++)
        if(alias_hits(&FAM_D2, r, ref, k)) d2_exec[k + 9 * (p7b - noab)] = slot;
      if(slot_cand_d2) slot_cand_d2[slot] = ia, slot_k_d2[slot] = (int32_t) p7b;
      slot++;
    }
  }
  counts[2] = slot;
}

/* ------------------------------------------------------------------------------------------ */
/* Op counter: cc/ccsd_t/fused_common.hpp:131-261 over the task list of
 * cc/ccsd_t/ccsd_t_fused_driver.hpp:604-634. */
ORC_API long double orc_count_ops(int noab, int nvab, const int32_t* k_spin, const int64_t* k_range,
                                  int is_restricted) {
  long double total_s1 = 0, total_d1 = 0, total_d2 = 0;
  int32_t*    d1e = (int32_t*) malloc(sizeof(int32_t) * 9 * (size_t) noab);
  int32_t*    d2e = (int32_t*) malloc(sizeof(int32_t) * 9 * (size_t) nvab);
  int32_t     s1e[9], counts[3];
  for(int64_t p4 = noab; p4 < noab + nvab; p4++)
    for(int64_t p5 = p4; p5 < noab + nvab; p5++)
      for(int64_t p6 = p5; p6 < noab + nvab; p6++)
        for(int64_t h1 = 0; h1 < noab; h1++)
          for(int64_t h2 = h1; h2 < noab; h2++)
            for(int64_t h3 = h2; h3 < noab; h3++) {
              int sp = k_spin[p4] + k_spin[p5] + k_spin[p6];
              int sh = k_spin[h1] + k_spin[h2] + k_spin[h3];
              if(sp != sh) continue;
              if(is_restricted && sp + sh > 8) continue;
              int64_t task[6] = {h1, h2, h3, p4, p5, p6};
              orc_task_exec(noab, nvab, k_spin, k_range, is_restricted, task, s1e, d1e, d2e, counts,
                            NULL, NULL, NULL, NULL, NULL);
              long double base = (long double) k_range[h1] * (long double) k_range[h2] *
                                 (long double) k_range[h3] * (long double) k_range[p4] *
                                 (long double) k_range[p5] * (long double) k_range[p6];
              long double s1 = 0, d1 = 0, d2 = 0;
              for(int k = 0; k < 9; k++)
                if(s1e[k] >= 0) s1 += base * 2;
              for(int h7 = 0; h7 < noab; h7++)
                for(int k = 0; k < 9; k++)
                  if(d1e[k + 9 * h7] >= 0) d1 += base * (long double) k_range[h7] * 2;
              for(int p7 = 0; p7 < nvab; p7++)
                for(int k = 0; k < 9; k++)
                  if(d2e[k + 9 * p7] >= 0) d2 += base * (long double) k_range[noab + p7] * 2;
              total_s1 += s1;
              total_d1 += d1;
              total_d2 += d2;
            }
  free(d1e);
  free(d2e);
  return total_s1 + total_d1 + total_d2;
}

/* ------------------------------------------------------------------------------------------ */
/* The maths of one task: staging (cc/ccsd_t/ccsd_t_all_fused_{singles,doubles1,doubles2}.hpp fetch +
 * hptt permutations, cited per block below) and the CPU kernel
 * cc/ccsd_t/ccsd_t_all_fused_cpu.hpp:118-273 (d1), :276-431 (d2), :434-580 (s1), :586-623 (energy),
 * in the reference's operation order so that results are bit-identical when compiled alike. */
typedef struct {
  int             noab, nvab;
  const int64_t*  k_range;
  const int64_t*  k_offset; /* cumulative over all tiles */
  const int32_t*  k_spin;
  int64_t         Ot, Vt;
  const double *  t1, *t2, *v2ijab, *v2ijka, *v2iabc, *evl;
} Problem;

/* term tables: index ids 0=h1 1=h2 2=h3 3=p4 4=p5 5=p6 */
static const int D1_SIGN[9] = {-1, +1, -1, -1, +1, -1, +1, -1, +1};
static const int D1_A[9][3] = {/* pa, pb, hx : T2slot[hx][pb][pa][h7] */
                               {3, 4, 0}, {3, 4, 1}, {3, 4, 2}, {4, 5, 0}, {4, 5, 1},
                               {4, 5, 2}, {3, 5, 0}, {3, 5, 1}, {3, 5, 2}};
static const int D1_B[9][3] = {/* hz, hy, pc : Vslot[h7][pc][hy][hz] */
                               {2, 1, 5}, {2, 0, 5}, {1, 0, 5}, {2, 1, 3}, {2, 0, 3},
                               {1, 0, 3}, {2, 1, 4}, {2, 0, 4}, {1, 0, 4}};
static const int D2_SIGN[9] = {-1, -1, +1, +1, +1, -1, -1, -1, +1};
static const int D2_A[9][3] = {/* pa, hx, hy : T2slot[hy][hx][pa][p7] */
                               {3, 0, 1}, {3, 1, 2}, {3, 0, 2}, {4, 0, 1}, {4, 1, 2},
                               {4, 0, 2}, {5, 0, 1}, {5, 1, 2}, {5, 0, 2}};
static const int D2_B[9][3] = {/* hz, pc, pb : Vslot[pb][pc][hz][p7] */
                               {2, 5, 4}, {0, 5, 4}, {1, 5, 4}, {2, 5, 3}, {0, 5, 3},
                               {1, 5, 3}, {2, 4, 3}, {0, 4, 3}, {1, 4, 3}};
static const int S1_SIGN[9] = {+1, -1, +1, -1, +1, -1, +1, -1, +1};
static const int S1_A[9][2] = {/* pa, hx : T1slot[hx][pa] */
                               {3, 0}, {3, 1}, {3, 2}, {4, 0}, {4, 1}, {4, 2}, {5, 0}, {5, 1}, {5, 2}};
static const int S1_B[9][4] = {/* hz, hy, pc, pb : Vslot[pb][pc][hy][hz] */
                               {2, 1, 5, 4}, {2, 0, 5, 4}, {1, 0, 5, 4}, {2, 1, 5, 3}, {2, 0, 5, 3},
                               {1, 0, 5, 3}, {2, 1, 4, 3}, {2, 0, 4, 3}, {1, 0, 4, 3}};

static inline int64_t oo(const Problem* P, int64_t tile) { return P->k_offset[tile]; }
static inline int64_t vo(const Problem* P, int64_t tile) { return P->k_offset[tile] - P->Ot; }

/* one task; returns its two energy contributions (already times factor, as the reference) */
static void task_energy(const Problem* P, int is_restricted, const int64_t* task, double factor,
                        double* e_out) {
  const int      noab = P->noab, nvab = P->nvab;
  const int64_t* kr = P->k_range;
  const int64_t  Ot = P->Ot, Vt = P->Vt;
  int32_t        s1e[9], counts[3];
  int32_t*       d1e  = (int32_t*) malloc(sizeof(int32_t) * 9 * (size_t) noab);
  int32_t*       d2e  = (int32_t*) malloc(sizeof(int32_t) * 9 * (size_t) nvab);
  int32_t *      sc1 = (int32_t*) malloc(sizeof(int32_t) * 9), *sc_d1 = (int32_t*) malloc(sizeof(int32_t) * 9 * (size_t) noab),
          *sk_d1 = (int32_t*) malloc(sizeof(int32_t) * 9 * (size_t) noab),
          *sc_d2 = (int32_t*) malloc(sizeof(int32_t) * 9 * (size_t) nvab),
          *sk_d2 = (int32_t*) malloc(sizeof(int32_t) * 9 * (size_t) nvab);
  orc_task_exec(noab, nvab, P->k_spin, kr, is_restricted, task, s1e, d1e, d2e, counts, sc1, sc_d1,
                sk_d1, sc_d2, sk_d2);

  const int64_t ref[6] = {task[3], task[4], task[5], task[0], task[1], task[2]};
  /* extents by index id (0=h1..5=p6) */
  const int64_t ext[6] = {kr[task[0]], kr[task[1]], kr[task[2]], kr[task[3]], kr[task[4]], kr[task[5]]};
  int64_t       max_p = 0, max_h = 0;
  for(int t = 0; t < noab; t++)
    if(kr[t] > max_h) max_h = kr[t];
  for(int t = noab; t < noab + nvab; t++)
    if(kr[t] > max_p) max_p = kr[t];
  /* slot strides: cc/ccsd_t/ccsd_t_fused_driver.hpp:189-195 divided by slots per pool */
  const int64_t dim_s1_t1 = max_p * max_h, dim_s1_v2 = max_p * max_p * max_h * max_h;
  const int64_t dim_d1_t2 = max_p * max_p * max_h * max_h, dim_d1_v2 = max_p * max_h * max_h * max_h;
  const int64_t dim_d2_t2 = max_p * max_p * max_h * max_h, dim_d2_v2 = max_p * max_p * max_p * max_h;

  double* s1_t1 = (double*) malloc(sizeof(double) * (size_t) (dim_s1_t1 * (counts[0] + 1)));
  double* s1_v2 = (double*) malloc(sizeof(double) * (size_t) (dim_s1_v2 * (counts[0] + 1)));
  double* d1_t2 = (double*) malloc(sizeof(double) * (size_t) (dim_d1_t2 * (counts[1] + 1)));
  double* d1_v2 = (double*) malloc(sizeof(double) * (size_t) (dim_d1_v2 * (counts[1] + 1)));
  double* d2_t2 = (double*) malloc(sizeof(double) * (size_t) (dim_d2_t2 * (counts[2] + 1)));
  double* d2_v2 = (double*) malloc(sizeof(double) * (size_t) (dim_d2_v2 * (counts[2] + 1)));

  Cands c;
  /* ---- s1 staging: singles.hpp:185-212 (t1{p4b,h1b} -> [h1][p4]) and :283-316
   *      (v2ijab{h3b,h2b,p6b,p5b} -> [p5][p6][h2][h3]) */
  build_cands(&FAM_S1, ref, &c);
  for(int s = 0; s < counts[0]; s++) {
    const int64_t* r   = c.roles[sc1[s]];
    int64_t        p4b = r[0], p5b = r[1], p6b = r[2], h1b = r[3], h2b = r[4], h3b = r[5];
    int64_t        P4 = kr[p4b], P5 = kr[p5b], P6 = kr[p6b], H1 = kr[h1b], H2 = kr[h2b], H3 = kr[h3b];
    double*        a = s1_t1 + dim_s1_t1 * s;
    for(int64_t h1 = 0; h1 < H1; h1++)
      for(int64_t p4 = 0; p4 < P4; p4++)
        a[p4 + P4 * h1] = 1 * P->t1[(vo(P, p4b) + p4) * Ot + (oo(P, h1b) + h1)];
    double* b = s1_v2 + dim_s1_v2 * s;
    for(int64_t p5 = 0; p5 < P5; p5++)
      for(int64_t p6 = 0; p6 < P6; p6++)
        for(int64_t h2 = 0; h2 < H2; h2++)
          for(int64_t h3 = 0; h3 < H3; h3++)
            b[h3 + H3 * (h2 + H2 * (p6 + P6 * p5))] =
              1.0 * P->v2ijab[(((oo(P, h3b) + h3) * Ot + (oo(P, h2b) + h2)) * Vt + (vo(P, p6b) + p6)) * Vt +
                              (vo(P, p5b) + p5)];
  }
  /* ---- d1 staging: doubles1.hpp:194-259 (t2 -> [h1][p5][p4][h7], sign -1 when h7b<h1b) and
   *      :261-303 (v2ijka{h2b,h3b,h7b,p6b} -> [h7][p6][h2][h3]) */
  build_cands(&FAM_D1, ref, &c);
  for(int s = 0; s < counts[1]; s++) {
    const int64_t* r   = c.roles[sc_d1[s]];
    int64_t        h7b = sk_d1[s];
    int64_t        p4b = r[0], p5b = r[1], p6b = r[2], h1b = r[3], h2b = r[4], h3b = r[5];
    int64_t        P4 = kr[p4b], P5 = kr[p5b], P6 = kr[p6b], H1 = kr[h1b], H2 = kr[h2b], H3 = kr[h3b],
            H7 = kr[h7b];
    double* a  = d1_t2 + dim_d1_t2 * s;
    for(int64_t h1 = 0; h1 < H1; h1++)
      for(int64_t p5 = 0; p5 < P5; p5++)
        for(int64_t p4 = 0; p4 < P4; p4++)
          for(int64_t h7 = 0; h7 < H7; h7++) {
            int64_t base = ((vo(P, p4b) + p4) * Vt + (vo(P, p5b) + p5)) * Ot;
            double  v;
            if(h7b < h1b) v = -1.0 * P->t2[(base + (oo(P, h7b) + h7)) * Ot + (oo(P, h1b) + h1)];
            else v = 1.0 * P->t2[(base + (oo(P, h1b) + h1)) * Ot + (oo(P, h7b) + h7)];
            a[h7 + H7 * (p4 + P4 * (p5 + P5 * h1))] = v;
          }
    double* b = d1_v2 + dim_d1_v2 * s;
    for(int64_t h7 = 0; h7 < H7; h7++)
      for(int64_t p6 = 0; p6 < P6; p6++)
        for(int64_t h2 = 0; h2 < H2; h2++)
          for(int64_t h3 = 0; h3 < H3; h3++)
            b[h3 + H3 * (h2 + H2 * (p6 + P6 * h7))] =
              1.0 * P->v2ijka[(((oo(P, h2b) + h2) * Ot + (oo(P, h3b) + h3)) * Ot + (oo(P, h7b) + h7)) * Vt +
                              (vo(P, p6b) + p6)];
  }
  /* ---- d2 staging: doubles2.hpp:196-252 (t2 -> [h2][h1][p4][p7], sign -1 when p7b<p4b) and
   *      :318-352 (v2iabc{h3b,p7b,p5b,p6b} -> [p5][p6][h3][p7]) */
  build_cands(&FAM_D2, ref, &c);
  for(int s = 0; s < counts[2]; s++) {
    const int64_t* r   = c.roles[sc_d2[s]];
    int64_t        p7b = sk_d2[s];
    int64_t        p4b = r[0], p5b = r[1], p6b = r[2], h1b = r[3], h2b = r[4], h3b = r[5];
    int64_t        P4 = kr[p4b], P5 = kr[p5b], P6 = kr[p6b], H1 = kr[h1b], H2 = kr[h2b], H3 = kr[h3b],
            P7 = kr[p7b];
    double* a  = d2_t2 + dim_d2_t2 * s;
    for(int64_t h2 = 0; h2 < H2; h2++)
      for(int64_t h1 = 0; h1 < H1; h1++)
        for(int64_t p4 = 0; p4 < P4; p4++)
          for(int64_t p7 = 0; p7 < P7; p7++) {
            int64_t hh = (oo(P, h1b) + h1) * Ot + (oo(P, h2b) + h2);
            double  v;
            if(p7b < p4b) v = -1.0 * P->t2[((vo(P, p7b) + p7) * Vt + (vo(P, p4b) + p4)) * Ot * Ot + hh];
            else v = 1.0 * P->t2[((vo(P, p4b) + p4) * Vt + (vo(P, p7b) + p7)) * Ot * Ot + hh];
            a[p7 + P7 * (p4 + P4 * (h1 + H1 * h2))] = v;
          }
    double* b = d2_v2 + dim_d2_v2 * s;
    for(int64_t p5 = 0; p5 < P5; p5++)
      for(int64_t p6 = 0; p6 < P6; p6++)
        for(int64_t h3 = 0; h3 < H3; h3++)
          for(int64_t p7 = 0; p7 < P7; p7++)
            b[p7 + P7 * (h3 + H3 * (p6 + P6 * p5))] =
              1.0 * P->v2iabc[(((oo(P, h3b) + h3) * Vt + (vo(P, p7b) + p7)) * Vt + (vo(P, p5b) + p5)) * Vt +
                              (vo(P, p6b) + p6)];
  }

  /* ---- t3 buffers: ccsd_t_all_fused_cpu.hpp:103-113 */
  const int64_t H1 = ext[0], H2 = ext[1], H3 = ext[2], P4 = ext[3], P5 = ext[4], P6 = ext[5];
  const int64_t n3 = H1 * H2 * H3 * P4 * P5 * P6;
  double*       t3d = (double*) calloc((size_t) n3, sizeof(double));
  double*       t3s = (double*) calloc((size_t) n3, sizeof(double));

  /* ---- d1: ccsd_t_all_fused_cpu.hpp:118-273.  Per h7 tile, per element, h7 innermost, terms 1..9 */
  for(int h7b = 0; h7b < noab; h7b++) {
    const int64_t H7 = kr[h7b];
    const int32_t* fl = d1e + 9 * h7b;
    int any = 0;
    for(int k = 0; k < 9; k++) any |= fl[k] >= 0;
    if(!any) continue;
#pragma omp parallel for collapse(3)
    for(int64_t i3 = 0; i3 < H3; i3++)
      for(int64_t i2 = 0; i2 < H2; i2++)
        for(int64_t i1 = 0; i1 < H1; i1++)
          for(int64_t i6 = 0; i6 < P6; i6++)
            for(int64_t i5 = 0; i5 < P5; i5++)
              for(int64_t i4 = 0; i4 < P4; i4++) {
                const int64_t ix[6] = {i1, i2, i3, i4, i5, i6};
                int64_t       t3i   = i3 + (i2 + (i1 + (i6 + (i5 + i4 * P5) * P6) * H1) * H2) * H3;
                double        acc   = t3d[t3i];
                for(int64_t h7 = 0; h7 < H7; h7++)
                  for(int k = 0; k < 9; k++) {
                    if(fl[k] < 0) continue;
                    const int pa = D1_A[k][0], pb = D1_A[k][1], hx = D1_A[k][2];
                    const int hz = D1_B[k][0], hy = D1_B[k][1], pc = D1_B[k][2];
                    double    a = d1_t2[dim_d1_t2 * fl[k] + h7 + H7 * (ix[pa] + ext[pa] * (ix[pb] + ext[pb] * ix[hx]))];
                    double    b = d1_v2[dim_d1_v2 * fl[k] + ix[hz] + ext[hz] * (ix[hy] + ext[hy] * (ix[pc] + ext[pc] * h7))];
                    double    prod = a * b;
                    if(D1_SIGN[k] < 0) acc -= prod;
                    else acc += prod;
                  }
                t3d[t3i] = acc;
              }
  }
  /* ---- d2: ccsd_t_all_fused_cpu.hpp:276-431 */
  for(int p7i = 0; p7i < nvab; p7i++) {
    const int64_t  P7 = kr[noab + p7i];
    const int32_t* fl = d2e + 9 * p7i;
    int any = 0;
    for(int k = 0; k < 9; k++) any |= fl[k] >= 0;
    if(!any) continue;
#pragma omp parallel for collapse(3)
    for(int64_t i3 = 0; i3 < H3; i3++)
      for(int64_t i2 = 0; i2 < H2; i2++)
        for(int64_t i1 = 0; i1 < H1; i1++)
          for(int64_t i6 = 0; i6 < P6; i6++)
            for(int64_t i5 = 0; i5 < P5; i5++)
              for(int64_t i4 = 0; i4 < P4; i4++) {
                const int64_t ix[6] = {i1, i2, i3, i4, i5, i6};
                int64_t       t3i   = i3 + (i2 + (i1 + (i6 + (i5 + i4 * P5) * P6) * H1) * H2) * H3;
                double        acc   = t3d[t3i];
                for(int64_t p7 = 0; p7 < P7; p7++)
                  for(int k = 0; k < 9; k++) {
                    if(fl[k] < 0) continue;
                    const int pa = D2_A[k][0], hx = D2_A[k][1], hy = D2_A[k][2];
                    const int hz = D2_B[k][0], pc = D2_B[k][1], pb = D2_B[k][2];
                    double    a = d2_t2[dim_d2_t2 * fl[k] + p7 + P7 * (ix[pa] + ext[pa] * (ix[hx] + ext[hx] * ix[hy]))];
                    double    b = d2_v2[dim_d2_v2 * fl[k] + p7 + P7 * (ix[hz] + ext[hz] * (ix[pc] + ext[pc] * ix[pb]))];
                    double    prod = a * b;
                    if(D2_SIGN[k] < 0) acc -= prod;
                    else acc += prod;
                  }
                t3d[t3i] = acc;
              }
  }
  /* ---- s1: ccsd_t_all_fused_cpu.hpp:434-580 (signs as in the CODE, not its comments) */
#pragma omp parallel for collapse(3)
  for(int64_t i3 = 0; i3 < H3; i3++)
    for(int64_t i2 = 0; i2 < H2; i2++)
      for(int64_t i1 = 0; i1 < H1; i1++)
        for(int64_t i6 = 0; i6 < P6; i6++)
          for(int64_t i5 = 0; i5 < P5; i5++)
            for(int64_t i4 = 0; i4 < P4; i4++) {
              const int64_t ix[6] = {i1, i2, i3, i4, i5, i6};
              int64_t       t3i   = i3 + (i2 + (i1 + (i6 + (i5 + i4 * P5) * P6) * H1) * H2) * H3;
              double        acc   = t3s[t3i];
              for(int k = 0; k < 9; k++) {
                if(s1e[k] < 0) continue;
                const int pa = S1_A[k][0], hx = S1_A[k][1];
                const int hz = S1_B[k][0], hy = S1_B[k][1], pc = S1_B[k][2], pb = S1_B[k][3];
                double    a    = s1_t1[dim_s1_t1 * s1e[k] + ix[pa] + ext[pa] * ix[hx]];
                double    b    = s1_v2[dim_s1_v2 * s1e[k] + ix[hz] + ext[hz] * (ix[hy] + ext[hy] * (ix[pc] + ext[pc] * ix[pb]))];
                double    prod = a * b;
                if(S1_SIGN[k] < 0) acc -= prod;
                else acc += prod;
              }
              t3s[t3i] = acc;
            }

  /* ---- energy: ccsd_t_all_fused_cpu.hpp:586-623, sequential in p4,p5,p6,h1,h2,h3 order */
  const double* e_h1 = P->evl + oo(P, task[0]);
  const double* e_h2 = P->evl + oo(P, task[1]);
  const double* e_h3 = P->evl + oo(P, task[2]);
  const double* e_p4 = P->evl + oo(P, task[3]);
  const double* e_p5 = P->evl + oo(P, task[4]);
  const double* e_p6 = P->evl + oo(P, task[5]);
  double        fe1 = 0.0, fe2 = 0.0;
  for(int64_t i4 = 0; i4 < P4; i4++)
    for(int64_t i5 = 0; i5 < P5; i5++)
      for(int64_t i6 = 0; i6 < P6; i6++)
        for(int64_t i1 = 0; i1 < H1; i1++)
          for(int64_t i2 = 0; i2 < H2; i2++)
            for(int64_t i3 = 0; i3 < H3; i3++) {
              int64_t t3i   = i3 + (i2 + (i1 + (i6 + (i5 + i4 * P5) * P6) * H1) * H2) * H3;
              double  inner = (e_h3[i3] + e_h2[i2] + e_h1[i1] - e_p6[i6] - e_p5[i5] - e_p4[i4]);
              fe1 += factor * t3d[t3i] * (t3d[t3i]) / inner;
              fe2 += factor * t3d[t3i] * (t3d[t3i] + t3s[t3i]) / inner;
            }
  e_out[0] = fe1;
  e_out[1] = fe2;

  free(t3d), free(t3s);
  free(s1_t1), free(s1_v2), free(d1_t2), free(d1_v2), free(d2_t2), free(d2_v2);
  free(d1e), free(d2e), free(sc1), free(sc_d1), free(sk_d1), free(sc_d2), free(sk_d2);
}

/* Runs kernel tasks [task_begin, task_end) of the canonical list (task_end < 0 = all) and
 * accumulates energy_l[0..1] += per-task energy in task order (cpu.hpp:621-623).
 * per_task (optional): 2 doubles per executed task. */
ORC_API int orc_run(int noab, int nvab, const int64_t* k_range, const int32_t* k_spin,
                    const double* evl, const double* t1, const double* t2, const double* v2ijab,
                    const double* v2ijka, const double* v2iabc, int is_restricted, int64_t task_begin,
                    int64_t task_end, double* energies, double* per_task) {
  int      ntile = noab + nvab;
  int64_t* off   = (int64_t*) malloc(sizeof(int64_t) * (size_t) (ntile + 1));
  int64_t  sum = 0, Ot = 0;
  for(int i = 0; i < ntile; i++) {
    off[i] = sum;
    sum += k_range[i];
    if(i < noab) Ot += k_range[i];
  }
  Problem P = {noab, nvab, k_range, off, k_spin, Ot, sum - Ot, t1, t2, v2ijab, v2ijka, v2iabc, evl};

  int64_t  n_outer;
  int64_t  n  = orc_enumerate(noab, nvab, k_spin, is_restricted, NULL, NULL, 0, &n_outer);
  int64_t* tk = (int64_t*) malloc(sizeof(int64_t) * 7 * (size_t) (n + 1));
  double*  fc = (double*) malloc(sizeof(double) * (size_t) (n + 1));
  orc_enumerate(noab, nvab, k_spin, is_restricted, tk, fc, n, &n_outer);
  if(task_end < 0 || task_end > n) task_end = n;
  double e1 = 0.0, e2 = 0.0;
  for(int64_t i = task_begin; i < task_end; i++) {
    double e[2];
    task_energy(&P, is_restricted, tk + 7 * i, fc[i], e);
    e1 += e[0];
    e2 += e[1];
    if(per_task) per_task[2 * (i - task_begin)] = e[0], per_task[2 * (i - task_begin) + 1] = e[1];
  }
  energies[0] = e1;
  energies[1] = e2;
  free(tk), free(fc), free(off);
  return 0;
}
