// V2 tensors from Cholesky vectors on the device (row f2 of SURVEY.md 8f): the contraction over the Cholesky index is a
// hand-written FP64 DMMA GEMM (dgemm_abt_kernel below), no library call.
#include "ccsdt_ctx.hpp"

#include <algorithm>

using namespace ccsdt;

extern "C" {
// =================================================================================================
// V2 from Cholesky vectors (row f2 of SURVEY.md 8f): replaces setupV2Tensors
// (exachem/cholesky/v2tensors.cpp:52-90, called at exachem/cc/ccsd_t/ccsd_t.cpp:168-193)
//   v2ijab(h1,h2,p1,p2) = L(h1,p1,c) L(h2,p2,c) - L(h1,p2,c) L(h2,p1,c)      v2tensors.cpp:68-69
//   v2ijka(h1,h2,h3,p1) = L(h1,h3,c) L(h2,p1,c) - L(h2,h3,c) L(h1,p1,c)      v2tensors.cpp:77-78
//   v2iabc(h1,p1,p2,p3) = L(h1,p2,c) L(p1,p3,c) - L(h1,p3,c) L(p1,p2,c)      v2tensors.cpp:85-86
// Each is one FP64 GEMM over the Cholesky index (dgemm_abt_kernel: C = A B^T with both operands K-major, DMMA m8n8k4)
// followed by an antisymmetrising gather; v2iabc is formed one occupied row at a time, so the GEMM scratch is V^3,
// not O V^3.
// =================================================================================================
} // extern "C"

namespace {

// rows (p in [p0,p0+np), q in [q0,q0+nq)) of chol[N][N][ncv] packed as out[(p,q)][ncv]
__global__ void __launch_bounds__(256) pack_pairs_kernel(const double* __restrict__ chol, int64_t N, int64_t ncv, int64_t p0,
                                                         int64_t np, int64_t q0, int64_t nq, double* __restrict__ out) {
  const int64_t total = np * nq * ncv;
  for(int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t) gridDim.x * blockDim.x) {
    const int64_t c = e % ncv, pq = e / ncv, q = pq % nq, pp = pq / nq;
    out[e]          = chol[((p0 + pp) * N + (q0 + q)) * ncv + c];
  }
}
// v2ijab[h1,h2,p1,p2] from G[(h1,p1),(h2,p2)]  (G is OV x OV)
__global__ void __launch_bounds__(256) v2ijab_kernel(const double* __restrict__ G, int64_t O, int64_t V, double* __restrict__ out) {
  const int64_t total = O * O * V * V, OV = O * V;
  for(int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t) gridDim.x * blockDim.x) {
    int64_t       r  = e;
    const int64_t p2 = r % V; r /= V;
    const int64_t p1 = r % V; r /= V;
    const int64_t h2 = r % O, h1 = r / O;
    out[e] = G[(h1 * V + p1) * OV + h2 * V + p2] - G[(h1 * V + p2) * OV + h2 * V + p1];
  }
}
// v2ijka[h1,h2,h3,p1] from G[(h1,h3),(h2,p1)]  (G is OO x OV)
__global__ void __launch_bounds__(256) v2ijka_kernel(const double* __restrict__ G, int64_t O, int64_t V, double* __restrict__ out) {
  const int64_t total = O * O * O * V, OV = O * V;
  for(int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t) gridDim.x * blockDim.x) {
    int64_t       r  = e;
    const int64_t p1 = r % V; r /= V;
    const int64_t h3 = r % O; r /= O;
    const int64_t h2 = r % O, h1 = r / O;
    out[e] = G[(h1 * O + h3) * OV + h2 * V + p1] - G[(h2 * O + h3) * OV + h1 * V + p1];
  }
}
// one occupied row: v2iabc[h1,p1,p2,p3] from G[(p2),(p1,p3)]  (G is V x VV for this h1)
__global__ void __launch_bounds__(256) v2iabc_row_kernel(const double* __restrict__ G, int64_t V, double* __restrict__ out) {
  const int64_t total = V * V * V, VV = V * V;
  for(int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t) gridDim.x * blockDim.x) {
    int64_t       r  = e;
    const int64_t p3 = r % V; r /= V;
    const int64_t p2 = r % V, p1 = r / V;
    out[e] = G[p2 * VV + p1 * V + p3] - G[p3 * VV + p1 * V + p2];
  }
}

// C[M][N] = A[M][K] * B[N][K]^T, row-major, K contiguous in both operands (the Cholesky index is the fastest index of
// the packed vectors).  CTA tile 64 x 64 x 16, four warps of 32 x 32 (16 DMMA m8n8k4 tiles, 64 accumulator registers),
// operands double-buffered in shared memory through registers.  Row stride 20 doubles: the 16 lanes of a half-warp
// (fragment rows 0-3, k 0-3) then hit 16 distinct 8-byte banks.  Edges are zero-filled on load and masked on store.
constexpr int GEMM_T = 64, GEMM_K = 16, GEMM_S = 20;
// D = A(8x4,row) * B(4x8,col) + D: lane holds a = A[lane>>2][lane&3], b = B[lane&3][lane>>2], d0,d1 = D[lane>>2][2*(lane&3) + {0,1}]
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(128) dgemm_abt_kernel(const double* __restrict__ A, const double* __restrict__ B,
                                                        double* __restrict__ C, int M, int N, int K, int64_t lda, int64_t ldb,
                                                        int64_t ldc) {
  __shared__ double As[2][GEMM_T][GEMM_S], Bs[2][GEMM_T][GEMM_S];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wm = warp >> 1, wn = warp & 1;
  const int m0 = blockIdx.y * GEMM_T, n0 = blockIdx.x * GEMM_T;
  const int lr = tid >> 1, lc = (tid & 1) * 8; // loader: row of the tile, first of 8 consecutive k
  double    acc[4][4][2];
#pragma unroll
  for(int i = 0; i < 4; i++)
#pragma unroll
    for(int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
  double ra[8], rb[8];
  auto   load = [&](int kt) {
    const int     k0 = kt * GEMM_K + lc;
    const double* pa = A + (int64_t) (m0 + lr) * lda + k0;
    const double* pb = B + (int64_t) (n0 + lr) * ldb + k0;
#pragma unroll
    for(int x = 0; x < 8; x++) {
      ra[x] = (m0 + lr < M && k0 + x < K) ? __ldg(pa + x) : 0.0;
      rb[x] = (n0 + lr < N && k0 + x < K) ? __ldg(pb + x) : 0.0;
    }
  };
  auto store = [&](int buf) {
#pragma unroll
    for(int x = 0; x < 8; x++) As[buf][lr][lc + x] = ra[x], Bs[buf][lr][lc + x] = rb[x];
  };
  const int nk = (K + GEMM_K - 1) / GEMM_K;
  load(0);
  store(0);
  __syncthreads();
  for(int kt = 0; kt < nk; kt++) {
    const int cur = kt & 1;
    if(kt + 1 < nk) load(kt + 1);
#pragma unroll
    for(int k4 = 0; k4 < 4; k4++) {
      double a[4], b[4];
#pragma unroll
      for(int i = 0; i < 4; i++) {
        a[i] = As[cur][wm * 32 + i * 8 + (lane >> 2)][k4 * 4 + (lane & 3)];
        b[i] = Bs[cur][wn * 32 + i * 8 + (lane >> 2)][k4 * 4 + (lane & 3)];
      }
#pragma unroll
      for(int i = 0; i < 4; i++)
#pragma unroll
        for(int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    if(kt + 1 < nk) store(cur ^ 1);
    __syncthreads();
  }
#pragma unroll
  for(int i = 0; i < 4; i++)
#pragma unroll
    for(int j = 0; j < 4; j++) {
      const int r = m0 + wm * 32 + i * 8 + (lane >> 2), c = n0 + wn * 32 + j * 8 + 2 * (lane & 3);
      if(r < M) {
        if(c < N) C[(int64_t) r * ldc + c] = acc[i][j][0];
        if(c + 1 < N) C[(int64_t) r * ldc + c + 1] = acc[i][j][1];
      }
    }
}

// row-major C[m][n] = A[m][k] * B[n][k]^T
inline void gemm_abt(cudaStream_t st, int m, int n, int k, const double* A, const double* B, double* C) {
  const dim3 grid((unsigned) ((n + GEMM_T - 1) / GEMM_T), (unsigned) ((m + GEMM_T - 1) / GEMM_T));
  dgemm_abt_kernel<<<grid, 128, 0, st>>>(A, B, C, m, n, k, k, k, n);
}

inline unsigned grid_for(int64_t n) { return (unsigned) std::min<int64_t>((n + 255) / 256, 148 * 16); }

int dense_alloc(ccsdt_ctx* ctx, int tensor, size_t n) {
  if(ctx->dense[tensor] && ctx->dense_elems[tensor] != n) {
    CK(cudaFree(ctx->dense[tensor]));
    ctx->dense[tensor] = nullptr;
  }
  if(!ctx->dense[tensor]) {
    CK(cudaMalloc(&ctx->dense[tensor], n * 8));
    ctx->dense_elems[tensor] = n;
  }
  ctx->dense_sparse_ok[tensor] = false; // written in full by the caller (ccsdt_put_cholesky)
  return 0;
}

} // namespace

extern "C" {

int ccsdt_put_cholesky(ccsdt_ctx* ctx, const double* host_chol, int64_t ncv) {
  if(!ctx || !host_chol || ncv <= 0) return 1;
  if(!ctx->have_space) return ctx->fail("ccsdt_set_space must be called first");
  cudaSetDevice(ctx->device);
  const int64_t O = ctx->sp.n_occ(), V = ctx->sp.n_virt(), N = O + V;
  if(O * V * ncv > 0x7fffffffll || V * V > 0x7fffffffll) return ctx->fail("ccsdt_put_cholesky: GEMM dimension exceeds int32", 10);
  CK(cudaStreamSynchronize(ctx->s_compute));
  CK(cudaStreamSynchronize(ctx->s_compute2));
  cudaStream_t st = ctx->s_stage;
  double *d_chol = nullptr, *Loo = nullptr, *Lov = nullptr, *Lvv = nullptr, *G = nullptr;
  struct Scratch { // the temporaries are released on every way out (the CK macro returns on failure)
    double** p[5];
    ~Scratch() {
      for(double** q: p)
        if(*q) cudaFree(*q);
    }
  } scratch{{&d_chol, &Loo, &Lov, &Lvv, &G}};
  const size_t nchol = (size_t) N * N * ncv;
  CK(cudaMalloc(&d_chol, nchol * 8));
  CK(cudaMemcpyAsync(d_chol, host_chol, nchol * 8, cudaMemcpyHostToDevice, st));
  ctx->pending_h2d += (int64_t) nchol * 8;
  CK(cudaMalloc(&Loo, (size_t) O * O * ncv * 8));
  CK(cudaMalloc(&Lov, (size_t) O * V * ncv * 8));
  CK(cudaMalloc(&Lvv, (size_t) V * V * ncv * 8));
  pack_pairs_kernel<<<grid_for(O * O * ncv), 256, 0, st>>>(d_chol, N, ncv, 0, O, 0, O, Loo);
  pack_pairs_kernel<<<grid_for(O * V * ncv), 256, 0, st>>>(d_chol, N, ncv, 0, O, O, V, Lov);
  pack_pairs_kernel<<<grid_for(V * V * ncv), 256, 0, st>>>(d_chol, N, ncv, O, V, O, V, Lvv);
  CK(cudaGetLastError());
  // scratch for the largest product: (OV x OV), (OO x OV) or one row of (V x VV)
  const size_t gmax = std::max({(size_t) O * V * O * V, (size_t) O * O * O * V, (size_t) V * V * V});
  CK(cudaMalloc(&G, gmax * 8));
  if(int rc = dense_alloc(ctx, CCSDT_V_IJAB, (size_t) O * O * V * V)) return rc;
  if(int rc = dense_alloc(ctx, CCSDT_V_IJKA, (size_t) O * O * O * V)) return rc;
  if(int rc = dense_alloc(ctx, CCSDT_V_IABC, (size_t) O * V * V * V)) return rc;
  gemm_abt(st, (int) (O * V), (int) (O * V), (int) ncv, Lov, Lov, G);
  v2ijab_kernel<<<grid_for(O * O * V * V), 256, 0, st>>>(G, O, V, ctx->dense[CCSDT_V_IJAB]);
  gemm_abt(st, (int) (O * O), (int) (O * V), (int) ncv, Loo, Lov, G);
  v2ijka_kernel<<<grid_for(O * O * O * V), 256, 0, st>>>(G, O, V, ctx->dense[CCSDT_V_IJKA]);
  for(int64_t h1 = 0; h1 < O; h1++) {
    gemm_abt(st, (int) V, (int) (V * V), (int) ncv, Lov + h1 * V * ncv, Lvv, G);
    v2iabc_row_kernel<<<grid_for(V * V * V), 256, 0, st>>>(G, V, ctx->dense[CCSDT_V_IABC] + h1 * V * V * V);
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  ctx->stats.kernel_launches += 7 + 2 * O;
  ctx->synthetic = false;
  return 0;
}


} // extern "C"
