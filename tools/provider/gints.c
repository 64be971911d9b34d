/* TEST-INPUT PROVIDER (not product code, not on the hot path): one- and two-electron integrals over
 * contracted cartesian Gaussians by the McMurchie-Davidson scheme (Hermite expansion coefficients E,
 * Hermite Coulomb integrals R from the Boys function).  Written from the published algorithm
 * (McMurchie & Davidson, J. Comput. Phys. 26, 218 (1978); Helgaker, Jorgensen, Olsen ch. 9); the
 * reference gets these numbers from libint, which is absent here.
 *
 * Used by tools/provider/provider.py to produce converged RHF/CCSD amplitudes and MO integrals for
 * small molecules, so the (T) path can be checked on REAL amplitudes against the reference's own CI
 * goldens.  Primitive coefficients arrive already multiplied by a^((2l+3)/4) (radial normalisation up
 * to a constant); every basis function is normalised by its self-overlap on the Python side.
 *
 * gcc -O2 -fopenmp -shared -fPIC gints.c -o _build/libgints.so -lm
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define LMAX 3               /* up to f shells */
#define NC(l) (((l) + 1) * ((l) + 2) / 2)
#define LSUM (4 * LMAX)      /* highest Hermite order of an ERI */
#define EDIM (LMAX + 3)      /* kinetic integrals need j + 2 */

/* ------------------------------------------------------------------ Boys function F_0..F_n(x) */
static void boys(int nmax, double x, double* F) {
  if(x < 35.0) {
    /* series for the highest order, downward recursion for the rest */
    double term = 1.0 / (2 * nmax + 1), sum = term;
    for(int k = 1; k < 300; k++) {
      term *= 2.0 * x / (2 * nmax + 2 * k + 1);
      sum += term;
      if(term < 1e-17 * sum) break;
    }
    const double ex = exp(-x);
    F[nmax]         = ex * sum;
    for(int n = nmax; n > 0; n--) F[n - 1] = (2.0 * x * F[n] + ex) / (2 * n - 1);
  }
  else {
    /* asymptotic F_0 and (stable for large x) upward recursion */
    const double ex = exp(-x);
    F[0]            = 0.5 * sqrt(M_PI / x);
    for(int n = 0; n < nmax; n++) F[n + 1] = ((2 * n + 1) * F[n] - ex) / (2.0 * x);
  }
}

/* ------------------------------------------------------------------ Hermite expansion, one dimension
 * E[i][j][t], 0<=i<=imax, 0<=j<=jmax, 0<=t<=i+j, for exponents a (centre A) and b (centre B) */
typedef double Etab[EDIM][EDIM][2 * EDIM];
static void hermite_E(int imax, int jmax, double a, double b, double A, double B, Etab E) {
  const double p = a + b, mu = a * b / p, P = (a * A + b * B) / p;
  const double XPA = P - A, XPB = P - B, XAB = A - B;
  memset(E, 0, sizeof(Etab));
  E[0][0][0] = exp(-mu * XAB * XAB);
  for(int i = 0; i <= imax; i++) {
    if(i > 0)
      for(int t = 0; t <= i; t++) {
        double v = XPA * E[i - 1][0][t];
        if(t > 0) v += E[i - 1][0][t - 1] / (2 * p);
        if(t + 1 <= i - 1) v += (t + 1) * E[i - 1][0][t + 1];
        E[i][0][t] = v;
      }
    for(int j = 1; j <= jmax; j++)
      for(int t = 0; t <= i + j; t++) {
        double v = XPB * E[i][j - 1][t];
        if(t > 0) v += E[i][j - 1][t - 1] / (2 * p);
        if(t + 1 <= i + j - 1) v += (t + 1) * E[i][j - 1][t + 1];
        E[i][j][t] = v;
      }
  }
}

/* ------------------------------------------------------------------ Hermite Coulomb integrals
 * R[t][u][v] = R^0_{tuv}(alpha, PC), t+u+v <= L */
#define RD (LSUM + 1)
static void hermite_R(int L, double alpha, double X, double Y, double Z, double R[RD][RD][RD]) {
  static __thread double W[RD + 1][RD][RD][RD]; /* W[n][t][u][v] */
  double                 F[RD + 1];
  boys(L, alpha * (X * X + Y * Y + Z * Z), F);
  double m2a = 1.0;
  for(int n = 0; n <= L; n++) {
    W[n][0][0][0] = m2a * F[n];
    m2a *= -2.0 * alpha;
  }
  for(int n = L - 1; n >= 0; n--) {
    const int M = L - n; /* total order available at this level */
    for(int t = 0; t <= M; t++)
      for(int u = 0; u + t <= M; u++)
        for(int v = 0; v + u + t <= M; v++) {
          if(t + u + v == 0) continue;
          double val;
          if(t > 0) val = (t > 1 ? (t - 1) * W[n + 1][t - 2][u][v] : 0.0) + X * W[n + 1][t - 1][u][v];
          else if(u > 0) val = (u > 1 ? (u - 1) * W[n + 1][t][u - 2][v] : 0.0) + Y * W[n + 1][t][u - 1][v];
          else val = (v > 1 ? (v - 1) * W[n + 1][t][u][v - 2] : 0.0) + Z * W[n + 1][t][u][v - 1];
          W[n][t][u][v] = val;
        }
  }
  for(int t = 0; t <= L; t++)
    for(int u = 0; u + t <= L; u++)
      for(int v = 0; v + u + t <= L; v++) R[t][u][v] = W[0][t][u][v];
}

static void cart_powers(int l, int pw[][3]) {
  int n = 0;
  for(int lx = l; lx >= 0; lx--)
    for(int ly = l - lx; ly >= 0; ly--) {
      pw[n][0] = lx, pw[n][1] = ly, pw[n][2] = l - lx - ly;
      n++;
    }
}

#if defined(__GNUC__)
#define EXPORT __attribute__((visibility("default")))
#else
#define EXPORT
#endif

/* shells: centre[3*s], l[s], nprim[s], poff[s] (offset into exps/coefs), coff[s] (first cartesian
 * function); atoms: Z[n], xyz[3*n].  Outputs ncart x ncart row-major. */
EXPORT int gints_one_electron(int nshell, const double* centre, const int* l, const int* nprim, const int* poff,
                              const int* coff, const double* exps, const double* coefs, int natom,
                              const double* Z, const double* xyz, int ncart, double* S, double* T, double* V) {
  memset(S, 0, sizeof(double) * ncart * ncart);
  memset(T, 0, sizeof(double) * ncart * ncart);
  memset(V, 0, sizeof(double) * ncart * ncart);
#pragma omp parallel for schedule(dynamic)
  for(int sa = 0; sa < nshell; sa++)
    for(int sb = 0; sb < nshell; sb++) {
      const int     la = l[sa], lb = l[sb];
      int           pa[NC(LMAX)][3], pb[NC(LMAX)][3];
      const double *A = centre + 3 * sa, *B = centre + 3 * sb;
      cart_powers(la, pa);
      cart_powers(lb, pb);
      static __thread double R[RD][RD][RD];
      for(int ka = 0; ka < nprim[sa]; ka++)
        for(int kb = 0; kb < nprim[sb]; kb++) {
          const double a = exps[poff[sa] + ka], b = exps[poff[sb] + kb];
          const double c = coefs[poff[sa] + ka] * coefs[poff[sb] + kb];
          const double p = a + b;
          Etab         E[3];
          for(int d = 0; d < 3; d++) hermite_E(la, lb + 2, a, b, A[d], B[d], E[d]);
          const double s3 = pow(M_PI / p, 1.5);
          for(int ia = 0; ia < NC(la); ia++)
            for(int ib = 0; ib < NC(lb); ib++) {
              double s1[3], k1[3];
              for(int d = 0; d < 3; d++) {
                const int i = pa[ia][d], j = pb[ib][d];
                s1[d]       = E[d][i][j][0];
                k1[d]       = -2.0 * b * b * E[d][i][j + 2][0] + b * (2 * j + 1) * E[d][i][j][0] -
                        (j >= 2 ? 0.5 * j * (j - 1) * E[d][i][j - 2][0] : 0.0);
              }
              const int row = coff[sa] + ia, col = coff[sb] + ib;
              S[row * ncart + col] += c * s3 * s1[0] * s1[1] * s1[2];
              T[row * ncart + col] += c * s3 * (k1[0] * s1[1] * s1[2] + s1[0] * k1[1] * s1[2] + s1[0] * s1[1] * k1[2]);
            }
          /* nuclear attraction */
          const double Px = (a * A[0] + b * B[0]) / p, Py = (a * A[1] + b * B[1]) / p, Pz = (a * A[2] + b * B[2]) / p;
          for(int n = 0; n < natom; n++) {
            hermite_R(la + lb, p, Px - xyz[3 * n], Py - xyz[3 * n + 1], Pz - xyz[3 * n + 2], R);
            for(int ia = 0; ia < NC(la); ia++)
              for(int ib = 0; ib < NC(lb); ib++) {
                const int ix = pa[ia][0], iy = pa[ia][1], iz = pa[ia][2];
                const int jx = pb[ib][0], jy = pb[ib][1], jz = pb[ib][2];
                double    sum = 0.0;
                for(int t = 0; t <= ix + jx; t++)
                  for(int u = 0; u <= iy + jy; u++)
                    for(int v = 0; v <= iz + jz; v++)
                      sum += E[0][ix][jx][t] * E[1][iy][jy][u] * E[2][iz][jz][v] * R[t][u][v];
                V[(coff[sa] + ia) * ncart + coff[sb] + ib] += -Z[n] * c * 2.0 * M_PI / p * sum;
              }
          }
        }
    }
  return 0;
}

/* (ab|cd) over cartesian functions, chemists' notation, full ncart^4 tensor (8-fold symmetry used
 * at the shell-quartet level, every permutation written). */
EXPORT int gints_eri(int nshell, const double* centre, const int* l, const int* nprim, const int* poff,
                     const int* coff, const double* exps, const double* coefs, int ncart, double* out) {
  const long npair = (long) nshell * (nshell + 1) / 2;
  const long N     = ncart;
#pragma omp parallel for schedule(dynamic)
  for(long pq = 0; pq < npair; pq++) {
    int sa = (int) ((sqrt(8.0 * pq + 1.0) - 1.0) / 2.0);
    while((long) (sa + 1) * (sa + 2) / 2 <= pq) sa++;
    while((long) sa * (sa + 1) / 2 > pq) sa--;
    const int sb = (int) (pq - (long) sa * (sa + 1) / 2);
    static __thread double R[RD][RD][RD];
    for(long rs = 0; rs <= pq; rs++) {
      int sc = (int) ((sqrt(8.0 * rs + 1.0) - 1.0) / 2.0);
      while((long) (sc + 1) * (sc + 2) / 2 <= rs) sc++;
      while((long) sc * (sc + 1) / 2 > rs) sc--;
      const int sd = (int) (rs - (long) sc * (sc + 1) / 2);
      const int la = l[sa], lb = l[sb], lc = l[sc], ld = l[sd];
      const int na = NC(la), nb = NC(lb), nc = NC(lc), nd = NC(ld);
      int       pa[NC(LMAX)][3], pb[NC(LMAX)][3], pc[NC(LMAX)][3], pd[NC(LMAX)][3];
      cart_powers(la, pa);
      cart_powers(lb, pb);
      cart_powers(lc, pc);
      cart_powers(ld, pd);
      const double *A = centre + 3 * sa, *B = centre + 3 * sb, *Cc = centre + 3 * sc, *D = centre + 3 * sd;
      double*       buf = (double*) calloc((size_t) na * nb * nc * nd, sizeof(double));
      const int     L   = la + lb + lc + ld;
      for(int ka = 0; ka < nprim[sa]; ka++)
        for(int kb = 0; kb < nprim[sb]; kb++) {
          const double a = exps[poff[sa] + ka], b = exps[poff[sb] + kb], p = a + b;
          const double cab = coefs[poff[sa] + ka] * coefs[poff[sb] + kb];
          Etab         Eb[3];
          for(int d = 0; d < 3; d++) hermite_E(la, lb, a, b, A[d], B[d], Eb[d]);
          const double Px = (a * A[0] + b * B[0]) / p, Py = (a * A[1] + b * B[1]) / p, Pz = (a * A[2] + b * B[2]) / p;
          for(int kc = 0; kc < nprim[sc]; kc++)
            for(int kd = 0; kd < nprim[sd]; kd++) {
              const double c = exps[poff[sc] + kc], d_ = exps[poff[sd] + kd], q = c + d_;
              const double ccd = coefs[poff[sc] + kc] * coefs[poff[sd] + kd];
              Etab         Ek[3];
              for(int d = 0; d < 3; d++) hermite_E(lc, ld, c, d_, Cc[d], D[d], Ek[d]);
              const double Qx = (c * Cc[0] + d_ * D[0]) / q, Qy = (c * Cc[1] + d_ * D[1]) / q,
                           Qz = (c * Cc[2] + d_ * D[2]) / q;
              const double alpha = p * q / (p + q);
              hermite_R(L, alpha, Px - Qx, Py - Qy, Pz - Qz, R);
              const double pref = cab * ccd * 2.0 * pow(M_PI, 2.5) / (p * q * sqrt(p + q));
              for(int ia = 0; ia < na; ia++)
                for(int ib = 0; ib < nb; ib++) {
                  const int tx = pa[ia][0] + pb[ib][0], ty = pa[ia][1] + pb[ib][1], tz = pa[ia][2] + pb[ib][2];
                  for(int ic = 0; ic < nc; ic++)
                    for(int id = 0; id < nd; id++) {
                      const int kx = pc[ic][0] + pd[id][0], ky = pc[ic][1] + pd[id][1], kz = pc[ic][2] + pd[id][2];
                      double    sum = 0.0;
                      for(int t = 0; t <= tx; t++) {
                        const double e1 = Eb[0][pa[ia][0]][pb[ib][0]][t];
                        for(int u = 0; u <= ty; u++) {
                          const double e2 = e1 * Eb[1][pa[ia][1]][pb[ib][1]][u];
                          for(int v = 0; v <= tz; v++) {
                            const double e3 = e2 * Eb[2][pa[ia][2]][pb[ib][2]][v];
                            double       ks = 0.0;
                            for(int tt = 0; tt <= kx; tt++) {
                              const double f1 = Ek[0][pc[ic][0]][pd[id][0]][tt];
                              for(int uu = 0; uu <= ky; uu++) {
                                const double f2 = f1 * Ek[1][pc[ic][1]][pd[id][1]][uu];
                                for(int vv = 0; vv <= kz; vv++) {
                                  const double sg = ((tt + uu + vv) & 1) ? -1.0 : 1.0;
                                  ks += sg * f2 * Ek[2][pc[ic][2]][pd[id][2]][vv] * R[t + tt][u + uu][v + vv];
                                }
                              }
                            }
                            sum += e3 * ks;
                          }
                        }
                      }
                      buf[((ia * nb + ib) * nc + ic) * nd + id] += pref * sum;
                    }
                }
            }
        }
      for(int ia = 0; ia < na; ia++)
        for(int ib = 0; ib < nb; ib++)
          for(int ic = 0; ic < nc; ic++)
            for(int id = 0; id < nd; id++) {
              const double v = buf[((ia * nb + ib) * nc + ic) * nd + id];
              const long   i = coff[sa] + ia, j = coff[sb] + ib, k = coff[sc] + ic, m = coff[sd] + id;
              out[((i * N + j) * N + k) * N + m] = v;
              out[((j * N + i) * N + k) * N + m] = v;
              out[((i * N + j) * N + m) * N + k] = v;
              out[((j * N + i) * N + m) * N + k] = v;
              out[((k * N + m) * N + i) * N + j] = v;
              out[((m * N + k) * N + i) * N + j] = v;
              out[((k * N + m) * N + j) * N + i] = v;
              out[((m * N + k) * N + j) * N + i] = v;
            }
      free(buf);
    }
  }
  return 0;
}
