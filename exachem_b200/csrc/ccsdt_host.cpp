#include "ccsdt_host.hpp"

#include <algorithm>
#include <numeric>

namespace ccsdt {

// ------------------------------------------------------------------------------------------------
// term tables (signs and index roles of the 27 equations; SURVEY.md §8 a9, reference statement
// exachem/cc/ccsd_t/ccsd_t_all_fused_cpu.hpp:172-270, 330-428, 472-578)
//                       pa pb hx hy hz pc sign
const D1Term kD1[9] = {{3, 4, 0, 1, 2, 5, -1}, {3, 4, 1, 0, 2, 5, +1}, {3, 4, 2, 0, 1, 5, -1},
                       {4, 5, 0, 1, 2, 3, -1}, {4, 5, 1, 0, 2, 3, +1}, {4, 5, 2, 0, 1, 3, -1},
                       {3, 5, 0, 1, 2, 4, +1}, {3, 5, 1, 0, 2, 4, -1}, {3, 5, 2, 0, 1, 4, +1}};
//                       pa hx hy hz pb pc sign
const D2Term kD2[9] = {{3, 0, 1, 2, 4, 5, -1}, {3, 1, 2, 0, 4, 5, -1}, {3, 0, 2, 1, 4, 5, +1},
                       {4, 0, 1, 2, 3, 5, +1}, {4, 1, 2, 0, 3, 5, +1}, {4, 0, 2, 1, 3, 5, -1},
                       {5, 0, 1, 2, 3, 4, -1}, {5, 1, 2, 0, 3, 4, -1}, {5, 0, 2, 1, 3, 4, +1}};
//                       pa hx hy hz pb pc sign
const S1Term kS1[9] = {{3, 0, 1, 2, 4, 5, +1}, {3, 1, 0, 2, 4, 5, -1}, {3, 2, 0, 1, 4, 5, +1},
                       {4, 0, 1, 2, 3, 5, -1}, {4, 1, 0, 2, 3, 5, +1}, {4, 2, 0, 1, 3, 5, -1},
                       {5, 0, 1, 2, 3, 4, +1}, {5, 1, 0, 2, 3, 4, -1}, {5, 2, 0, 1, 3, 4, +1}};

// ------------------------------------------------------------------------------------------------
int64_t Space::max_hole_tile() const {
  int64_t m = 0;
  for(int i = 0; i < noab(); i++) m = std::max(m, k_range[i]);
  return m;
}
int64_t Space::max_particle_tile() const {
  int64_t m = 0;
  for(int i = noab(); i < noab() + nvab(); i++) m = std::max(m, k_range[i]);
  return m;
}

void Space::spin_range(bool particle, int spin, int& tile_begin, int& tile_end, int64_t& n_orb) const {
  const int lo = particle ? noab() : 0, hi = particle ? noab() + nvab() : noab();
  tile_begin = hi;
  tile_end   = hi;
  n_orb      = 0;
  for(int t = lo; t < hi; t++) {
    if(k_spin[t] != spin) continue;
    if(tile_begin == hi) tile_begin = t;
    tile_end = t + 1;
    n_orb += k_range[t];
  }
  if(tile_begin == hi) tile_begin = tile_end = lo;
}

std::string Space::validate() const {
  const int n = noab() + nvab();
  if(noab() <= 0 || nvab() <= 0) return "empty occupied or virtual tile space";
  if((int) k_range.size() != n || (int) k_spin.size() != n) return "k_range/k_spin length mismatch";
  for(int t = 0; t < n; t++) {
    if(k_range[t] <= 0) return "tile with non-positive extent";
    if(k_spin[t] != 1 && k_spin[t] != 2) return "k_spin entries must be 1 (alpha) or 2 (beta)";
  }
  // tiles of one spin must be contiguous inside occ and inside virt (| a | b | ordering)
  for(int part = 0; part < 2; part++) {
    const int lo = part ? noab() : 0, hi = part ? n : noab();
    int       changes = 0;
    for(int t = lo + 1; t < hi; t++) changes += k_spin[t] != k_spin[t - 1];
    if(changes > 1) return "tiles of equal spin are not contiguous";
    if(changes == 1 && k_spin[lo] != 1) return "alpha tiles must precede beta tiles";
  }
  if((int64_t) evl.size() != k_offset.back()) return "orbital-energy vector length mismatch";
  return "";
}

Space make_space(int noa, int nob, int nva, int nvb, const int64_t* k_range, const int32_t* k_spin,
                 const double* evl, bool restricted) {
  Space s;
  s.noa = noa, s.nob = nob, s.nva = nva, s.nvb = nvb, s.restricted = restricted;
  const int n   = noa + nob + nva + nvb;
  int64_t   sum = 0;
  for(int t = 0; t < n; t++) {
    s.k_range.push_back(k_range[t]);
    s.k_offset.push_back(sum);
    s.k_spin.push_back(k_spin[t]);
    sum += k_range[t];
  }
  s.k_offset.push_back(sum);
  if(evl) s.evl.assign(evl, evl + sum);
  else s.evl.assign((size_t) sum, 0.0);
  return s;
}

int make_tiles(int64_t n_occ_alpha, int64_t n_occ_beta, int64_t n_vir_alpha, int64_t n_vir_beta,
               int64_t tilesize, std::vector<int64_t>& k_range, std::vector<int32_t>& k_spin,
               int32_t counts[4]) {
  k_range.clear();
  k_spin.clear();
  const int64_t n[4] = {n_occ_alpha, n_occ_beta, n_vir_alpha, n_vir_beta};
  for(int g = 0; g < 4; g++) {
    int c = 0;
    for(int64_t left = n[g]; left > 0; left -= tilesize, c++) k_range.push_back(std::min(left, tilesize));
    counts[g] = c;
  }
  // spins by tile count.  The reference labels the first noab/2 (nvab/2) tiles alpha and the rest beta
  // (ccsd_t.cpp:245-249), which is the actual layout only when both spins have the same number of tiles (or beta
  // one more); the labels here agree with it in those cases and stay correct for open-shell spaces with
  // different alpha / beta tile counts (row f5 of SURVEY.md 8f).
  for(int g = 0; g < 4; g++)
    for(int x = 0; x < counts[g]; x++) k_spin.push_back(g % 2 ? 2 : 1);
  return (int) k_range.size();
}

// ------------------------------------------------------------------------------------------------
static double symmetry_factor(bool restricted, const int32_t t[6]) {
  double f = restricted ? 2.0 : 1.0;
  if(t[3] == t[4] && t[4] == t[5]) f /= 6.0;
  else if(t[3] == t[4] || t[4] == t[5]) f /= 2.0;
  if(t[0] == t[1] && t[1] == t[2]) f /= 6.0;
  else if(t[0] == t[1] || t[1] == t[2]) f /= 2.0;
  return f;
}

static bool spin_allowed(const int32_t* k_spin, bool restricted, const int32_t t[6]) {
  const int sh = k_spin[t[0]] + k_spin[t[1]] + k_spin[t[2]];
  const int sp = k_spin[t[3]] + k_spin[t[4]] + k_spin[t[5]];
  return sh == sp && (!restricted || sh + sp <= 8);
}

std::vector<Task> enumerate_tasks(int noab, int nvab, const int32_t* k_spin, bool restricted,
                                  int64_t* n_outer) {
  std::vector<Task> out;
  int64_t           outer = 0;
  const int         pend  = noab + nvab;
  // distribution order of the reference: (h1b, p4b, h2b>=h1b, p5b>=p4b, p6b>=p5b) outer 5-tuples,
  // h3b>=h2b innermost
  for(int h1 = 0; h1 < noab; h1++)
    for(int p4 = noab; p4 < pend; p4++)
      for(int h2 = h1; h2 < noab; h2++)
        for(int p5 = p4; p5 < pend; p5++)
          for(int p6 = p5; p6 < pend; p6++, outer++)
            for(int h3 = h2; h3 < noab; h3++) {
              Task t{{h1, h2, h3, p4, p5, p6}, outer, 0.0};
              if(!spin_allowed(k_spin, restricted, t.t)) continue;
              t.factor = symmetry_factor(restricted, t.t);
              out.push_back(t);
            }
  if(n_outer) *n_outer = outer;
  return out;
}

// ------------------------------------------------------------------------------------------------
static inline int spin_of(const Space& sp, const Task& t, int idx) { return sp.k_spin[t.t[idx]]; }

bool task_nonempty(const Space& sp, const Task& t) {
  for(int i = 0; i < 6; i++)
    if(sp.k_range[t.t[i]] <= 0) return false;
  int sum = 0;
  for(int i = 0; i < 6; i++) sum += spin_of(sp, t, i);
  // the reference skips all-beta sextuples in restricted mode inside its staging functions too
  return !(sp.restricted && sum == 12);
}

int d1_contracted_spin(const Space& sp, const Task& t, int k) {
  if(!task_nonempty(sp, t)) return 0;
  const int s = spin_of(sp, t, kD1[k].pa) + spin_of(sp, t, kD1[k].pb) - spin_of(sp, t, kD1[k].hx);
  return (s == 1 || s == 2) ? s : 0;
}
int d2_contracted_spin(const Space& sp, const Task& t, int k) {
  if(!task_nonempty(sp, t)) return 0;
  const int s = spin_of(sp, t, kD2[k].hx) + spin_of(sp, t, kD2[k].hy) - spin_of(sp, t, kD2[k].pa);
  return (s == 1 || s == 2) ? s : 0;
}
bool s1_enabled(const Space& sp, const Task& t, int k) {
  return task_nonempty(sp, t) && spin_of(sp, t, kS1[k].pa) == spin_of(sp, t, kS1[k].hx);
}

void task_terms(const Space& sp, const Task& t, uint8_t* s1_on, uint8_t* d1_on, uint8_t* d2_on) {
  const int noab = sp.noab(), nvab = sp.nvab();
  for(int k = 0; k < 9; k++) s1_on[k] = s1_enabled(sp, t, k);
  for(int k = 0; k < 9; k++) {
    const int s = d1_contracted_spin(sp, t, k);
    for(int h7 = 0; h7 < noab; h7++) d1_on[k + 9 * h7] = (s != 0 && sp.k_spin[h7] == s);
  }
  for(int k = 0; k < 9; k++) {
    const int s = d2_contracted_spin(sp, t, k);
    for(int p7 = 0; p7 < nvab; p7++) d2_on[k + 9 * p7] = (s != 0 && sp.k_spin[noab + p7] == s);
  }
}

static void task_ops_split(const Space& sp, const Task& t, long double& s1, long double& d1,
                           long double& d2) {
  long double base = 2;
  for(int i = 0; i < 6; i++) base *= (long double) sp.k_range[t.t[i]];
  s1 = d1 = d2 = 0;
  for(int k = 0; k < 9; k++)
    if(s1_enabled(sp, t, k)) s1 += base;
  int     tb, te;
  int64_t n;
  for(int k = 0; k < 9; k++) {
    const int s = d1_contracted_spin(sp, t, k);
    if(!s) continue;
    sp.spin_range(false, s, tb, te, n);
    for(int h7 = tb; h7 < te; h7++) d1 += base * (long double) sp.k_range[h7];
  }
  for(int k = 0; k < 9; k++) {
    const int s = d2_contracted_spin(sp, t, k);
    if(!s) continue;
    sp.spin_range(true, s, tb, te, n);
    for(int p7 = tb; p7 < te; p7++) d2 += base * (long double) sp.k_range[p7];
  }
}

long double task_ops(const Space& sp, const Task& t) {
  long double a, b, c;
  task_ops_split(sp, t, a, b, c);
  return a + b + c;
}

long double count_ops(const Space& sp) {
  // same accumulation structure as the reference's counter (three running totals, task order
  // p4b,p5b,p6b,h1b,h2b,h3b) so the long-double result is identical even beyond 2^64
  long double tot_s1 = 0, tot_d1 = 0, tot_d2 = 0;
  const int   noab = sp.noab(), pend = sp.noab() + sp.nvab();
  for(int p4 = noab; p4 < pend; p4++)
    for(int p5 = p4; p5 < pend; p5++)
      for(int p6 = p5; p6 < pend; p6++)
        for(int h1 = 0; h1 < noab; h1++)
          for(int h2 = h1; h2 < noab; h2++)
            for(int h3 = h2; h3 < noab; h3++) {
              Task t{{h1, h2, h3, p4, p5, p6}, 0, 0.0};
              if(!spin_allowed(sp.k_spin.data(), sp.restricted, t.t)) continue;
              long double a, b, c;
              task_ops_split(sp, t, a, b, c);
              tot_s1 += a, tot_d1 += b, tot_d2 += c;
            }
  return tot_s1 + tot_d1 + tot_d2;
}

double symmetry_fraction(const Space& sp, const Task& t, const int hbox[3], int pbox) {
  double frac = 1.0;
  for(int f = 0; f < 2; f++) {
    int64_t n[3];
    for(int i = 0; i < 3; i++) {
      const int64_t box = f == 0 ? hbox[i] : pbox;
      n[i]              = (sp.k_range[t.t[3 * f + i]] + box - 1) / box;
    }
    const bool s01 = t.t[3 * f] == t.t[3 * f + 1] && (f == 1 || hbox[0] == hbox[1]);
    const bool s12 = t.t[3 * f + 1] == t.t[3 * f + 2] && (f == 1 || hbox[1] == hbox[2]);
    const double all = (double) n[0] * n[1] * n[2];
    if(all <= 0) continue;
    double keep = all;
    if(s01 && s12) keep = (double) n[0] * (n[0] + 1) * (n[0] + 2) / 6.0;
    else if(s01) keep = (double) n[0] * (n[0] + 1) / 2.0 * n[2];
    else if(s12) keep = (double) n[0] * n[1] * (n[1] + 1) / 2.0;
    frac *= keep / all;
  }
  return frac;
}

long double task_cost(const Space& sp, const Task& t, bool symmetry) {
  const int hbox[3] = {2, 2, 2};
  return task_ops(sp, t) * (symmetry ? (long double) symmetry_fraction(sp, t, hbox, 8) : 1.0L);
}

std::vector<int32_t> partition_tasks(const Space& sp, const std::vector<Task>& tasks, int nranks, bool symmetry) {
  std::vector<int32_t> owner(tasks.size(), 0);
  if(nranks <= 1) return owner;
  std::vector<long double> cost(tasks.size());
  for(size_t i = 0; i < tasks.size(); i++) cost[i] = task_cost(sp, tasks[i], symmetry);
  std::vector<size_t> order(tasks.size());
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return cost[a] > cost[b]; });
  std::vector<long double> load(nranks, 0);
  for(size_t i: order) {
    int best = 0;
    for(int r = 1; r < nranks; r++)
      if(load[r] < load[best]) best = r;
    owner[i] = best;
    load[best] += cost[i];
  }
  return owner;
}

// ------------------------------------------------------------------------------------------------
// execution tiling
static void cut_block(int64_t n, int64_t target, int64_t gran, std::vector<int64_t>& out) {
  if(n <= 0) return;
  // ceil(n / target) tiles; the extent is dealt out in units of `gran` as evenly as possible (larger tiles first) and
  // the last tile gives back the padding of the last unit: 195 at target 40 -> 40,40,40,40,35; 93 at 32 -> 32,32,29
  const int64_t nt = (n + target - 1) / target, units = (n + gran - 1) / gran;
  const int64_t base = units / nt, extra = units % nt;
  int64_t       left = n;
  for(int64_t i = 0; i < nt && left > 0; i++) {
    const int64_t e = std::min(left, (base + (i < extra ? 1 : 0)) * gran);
    if(e <= 0) continue; // fewer units than tiles
    out.push_back(e);
    left -= e;
  }
}

Space make_exec_space(const Space& store, int target) {
  const int nt = store.noab() + store.nvab();
  int64_t   tmax = 0;
  for(int t = 0; t < nt; t++) tmax = std::max(tmax, store.k_range[t]);
  if(target < 0) {
    // auto: keep a storage tiling the kernel already likes
    bool keep = tmax >= 24;
    for(int t = store.noab(); t < nt && keep; t++) {
      const bool last_of_spin = t == nt - 1 || store.k_spin[t + 1] != store.k_spin[t];
      if(!last_of_spin && store.k_range[t] % 8) keep = false;
    }
    if(keep) return store;
    target = (int) std::max<int64_t>(40, (tmax + 7) / 8 * 8);
  }
  if(target == 0) return store;
  const int64_t        tp = std::max(8, (target + 7) / 8 * 8), th = std::max(2, (target + 1) / 2 * 2);
  std::vector<int64_t> kr;
  std::vector<int32_t> ks;
  int                  cnt[4] = {0, 0, 0, 0};
  for(int g = 0; g < 4; g++) {
    const bool particle = g >= 2;
    int        tb, te;
    int64_t    n;
    store.spin_range(particle, g % 2 + 1, tb, te, n);
    const size_t before = kr.size();
    cut_block(n, particle ? tp : th, particle ? 8 : 2, kr);
    cnt[g] = (int) (kr.size() - before);
    for(int i = 0; i < cnt[g]; i++) ks.push_back(g % 2 + 1);
  }
  return make_space(cnt[0], cnt[1], cnt[2], cnt[3], kr.data(), ks.data(), store.evl.empty() ? nullptr : store.evl.data(),
                    store.restricted);
}

bool same_tiling(const Space& a, const Space& b) {
  return a.noa == b.noa && a.nob == b.nob && a.nva == b.nva && a.nvb == b.nvb && a.k_range == b.k_range;
}

std::vector<TilePiece> split_tile(const Space& exec, const Space& store, int exec_tile) {
  std::vector<TilePiece> out;
  const bool    particle = exec_tile >= exec.noab();
  const int64_t lo = exec.k_offset[exec_tile], hi = lo + exec.k_range[exec_tile]; // spin-orbital numbers: both spaces
  const int     sb = particle ? store.noab() : 0, se = particle ? store.noab() + store.nvab() : store.noab(); // order alike
  // exec and store number the orbitals identically (occ a | occ b | virt a | virt b), but the occupied/virtual split
  // is at k_offset[noab] in each: the same value, since both hold the same orbitals
  for(int t = sb; t < se; t++) {
    const int64_t a = std::max(lo, store.k_offset[t]), b = std::min(hi, store.k_offset[t] + store.k_range[t]);
    if(a < b) out.push_back(TilePiece{t, a - store.k_offset[t], a - lo, b - a});
  }
  return out;
}

int canonical_block(int tensor, uint32_t bid[4], int perm[4]) {
  for(int d = 0; d < 4; d++) perm[d] = d;
  int  sign = 1;
  auto order = [&](int i, int j, bool ascending) {
    if(ascending ? bid[i] > bid[j] : bid[i] < bid[j]) {
      std::swap(bid[i], bid[j]);
      std::swap(perm[i], perm[j]);
      sign = -sign;
    }
  };
  switch(tensor) {
    case 1: order(0, 1, true), order(2, 3, true); break;   // T2[a,b,i,j]
    case 2: order(0, 1, false), order(2, 3, false); break; // v2ijab[i,j,a,b]: the reference asks {h_hi,h_lo,p_hi,p_lo}
    case 3: order(0, 1, true); break;                      // v2ijka[i,j,k,a]
    case 4: order(2, 3, true); break;                      // v2iabc[i,a,b,c]
    default: break;                                        // T1
  }
  return sign;
}

// ------------------------------------------------------------------------------------------------
MemoryEstimate estimate_memory(const Space& sp, const int sub[3], int staging_buffers) {
  auto up = [](int64_t x, int64_t m) { return (x + m - 1) / m * m; };
  MemoryEstimate e{};
  int64_t        thp = 0;
  for(int i = 0; i < 3; i++) thp = std::max(thp, up(sp.max_hole_tile(), 2 * std::max(sub[i], 1)));
  e.thp = up(thp, 2);
  e.tpp = up(sp.max_particle_tile(), 8);
  int     tb, te;
  int64_t na, nb;
  sp.spin_range(false, 1, tb, te, na);
  sp.spin_range(false, 2, tb, te, nb);
  e.kp_occ = up(std::max<int64_t>(std::max(na, nb), 1), 16);
  const int64_t occ_spin = std::max(na, nb);
  sp.spin_range(true, 1, tb, te, na);
  sp.spin_range(true, 2, tb, te, nb);
  e.kp_virt = up(std::max<int64_t>(std::max(na, nb), 1), 16);
  const int64_t virt_spin = std::max(na, nb);
  // per staging buffer: 9 slots x (HPP [THp][TPp][TPp][Kp] + HHP [THp][THp][TPp][Kp]) for the d1 (Kp_occ) and d2 (Kp_virt) pools
  const int64_t hpp = 9 * e.thp * e.tpp * e.tpp, hhp = 9 * e.thp * e.thp * e.tpp;
  e.panel_bytes     = (int64_t) staging_buffers * 8 * (hpp + hhp) * (e.kp_occ + e.kp_virt);
  e.s1_bytes        = (int64_t) staging_buffers * 8 * (9 * e.thp * e.tpp + 9 * e.tpp * e.tpp * e.thp * e.thp);
  // one task reads at most 9 (T2 + V) block pairs per contracted tile; summed over the contracted tiles of one spin that
  // is 9 x 2 x T^3 x (orbitals of the spin) elements for d1 and for d2, plus the 9 s1 pairs
  const int64_t th = sp.max_hole_tile(), tp = sp.max_particle_tile();
  e.task_block_bytes = 8 * (9 * (th * tp * tp + th * th * tp) * occ_spin + 9 * (th * th * tp + th * tp * tp) * virt_spin +
                            9 * (th * tp + th * th * tp * tp));
  // spin-conserving blocks: T1 s_a = s_i; four-index tensors s0 + s1 = s2 + s3 (6 of the 16 spin patterns)
  int64_t n[2][2]; // [occ, virt][alpha, beta]
  for(int part = 0; part < 2; part++)
    for(int spin = 1; spin <= 2; spin++) {
      int64_t c = 0;
      sp.spin_range(part == 1, spin, tb, te, c);
      n[part][spin - 1] = c;
    }
  auto four = [&](int k0, int k1, int k2, int k3) { // kinds: 0 = occupied, 1 = virtual
    int64_t tot = 0;
    for(int pat = 0; pat < 16; pat++) {
      const int s0 = (pat >> 3) & 1, s1 = (pat >> 2) & 1, s2 = (pat >> 1) & 1, s3 = pat & 1;
      if(s0 + s1 != s2 + s3) continue;
      tot += n[k0][s0] * n[k1][s1] * n[k2][s2] * n[k3][s3];
    }
    return 8 * tot;
  };
  e.tensor_bytes[0] = 8 * (n[1][0] * n[0][0] + n[1][1] * n[0][1]);
  e.tensor_bytes[1] = four(1, 1, 0, 0);
  e.tensor_bytes[2] = four(0, 0, 1, 1);
  e.tensor_bytes[3] = four(0, 0, 0, 1);
  e.tensor_bytes[4] = four(0, 1, 1, 1);
  e.minimum_bytes   = e.panel_bytes + e.s1_bytes + 2 * e.task_block_bytes;
  return e;
}

} // namespace ccsdt
