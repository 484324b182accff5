// Eigen-loss kernels: Rayleigh quotient, residual, Gram / orthonormality (forward partials,
// finalisation, analytic backward).  Reference arithmetic: src/multigrid_model.py:291-348.
//
// Data layout: U, KU, MU are row-major n x k fp32 in HBM.  Everything that is summed over
// vertices is accumulated in fp32 for at most one 32-row tile and then folded into fp64, so the
// one-pass residual expansion sum(KU^2) - 2 lam sum(KU MU) + lam^2 sum(MU^2) does not cancel.
// Algorithmic traffic of the partials kernel: 12 n k bytes read, flops 2 n k^2 + 8 n k.
#include "ep_common.cuh"

namespace {

constexpr int kPartialThreads = 256;

__host__ __device__ inline int partials_len(int k) { return k * k + 4 * k; }

inline int partial_blocks(int n, int rows_per_tile) {
  int tiles = ep::ceil_div(n, rows_per_tile);
  int cap = ep::sm_count() * 4;
  int g = tiles < cap ? tiles : cap;
  return g < 1 ? 1 : g;
}

// TG x TG register block of the Gram matrix per thread, 16 x 16 threads -> KP = 16 TG columns.
template <int TG>
__global__ void __launch_bounds__(kPartialThreads)
eigen_partials_kernel(int n, int k, const float* __restrict__ U, int ldu, const float* __restrict__ KU,
                      const float* __restrict__ MU, int ld, double* __restrict__ block_out) {
  constexpr int KP = 16 * TG;
  constexpr int R = (TG == 8) ? 16 : 32;        // rows per tile (3 tiles must fit 48 KB static smem)
  constexpr int CPL = (KP + 31) / 32;           // columns per lane for the column sums
  __shared__ __align__(16) float Us[R][KP];
  __shared__ __align__(16) float KUs[R][KP];
  __shared__ __align__(16) float MUs[R][KP];

  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int warp = tid >> 5, lane = tid & 31;

  double g64[TG][TG];
  double c64[4][CPL];
#pragma unroll
  for (int i = 0; i < TG; ++i)
#pragma unroll
    for (int j = 0; j < TG; ++j) g64[i][j] = 0.0;
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int c = 0; c < CPL; ++c) c64[q][c] = 0.0;

  const int n_tiles = (n + R - 1) / R;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int row0 = tile * R;
    __syncthreads();
    for (int e = tid; e < R * KP; e += kPartialThreads) {
      const int r = e / KP, c = e - r * KP;
      const int row = row0 + r;
      const bool ok = (row < n) && (c < k);
      Us[r][c] = ok ? __ldg(U + (size_t)row * ldu + c) : 0.f;
      KUs[r][c] = ok ? __ldg(KU + (size_t)row * ld + c) : 0.f;
      MUs[r][c] = ok ? __ldg(MU + (size_t)row * ld + c) : 0.f;
    }
    __syncthreads();
    // Gram block: G[ty*TG+i][tx*TG+j] += U[r][ty*TG+i] * MU[r][tx*TG+j]
    float g32[TG][TG];
#pragma unroll
    for (int i = 0; i < TG; ++i)
#pragma unroll
      for (int j = 0; j < TG; ++j) g32[i][j] = 0.f;
#pragma unroll 4
    for (int r = 0; r < R; ++r) {
      float a[TG], b[TG];
#pragma unroll
      for (int i = 0; i < TG; ++i) a[i] = Us[r][ty * TG + i];
#pragma unroll
      for (int j = 0; j < TG; ++j) b[j] = MUs[r][tx * TG + j];
#pragma unroll
      for (int i = 0; i < TG; ++i)
#pragma unroll
        for (int j = 0; j < TG; ++j) g32[i][j] = fmaf(a[i], b[j], g32[i][j]);
    }
#pragma unroll
    for (int i = 0; i < TG; ++i)
#pragma unroll
      for (int j = 0; j < TG; ++j) g64[i][j] += (double)g32[i][j];
    // column sums: warp w owns rows w, w+8, ...; lane owns columns lane, lane+32, ...
#pragma unroll
    for (int cc = 0; cc < CPL; ++cc) {
      const int c = lane + 32 * cc;
      if (c < KP) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        for (int r = warp; r < R; r += kPartialThreads / 32) {
          const float u = Us[r][c], ku = KUs[r][c], mu = MUs[r][c];
          s0 = fmaf(u, ku, s0);
          s1 = fmaf(ku, ku, s1);
          s2 = fmaf(ku, mu, s2);
          s3 = fmaf(mu, mu, s3);
        }
        c64[0][cc] += (double)s0; c64[1][cc] += (double)s1;
        c64[2][cc] += (double)s2; c64[3][cc] += (double)s3;
      }
    }
  }
  // ---- write this block's partials: Gram directly, column sums reduced over the 8 warps
  double* out = block_out + (size_t)blockIdx.x * partials_len(k);
#pragma unroll
  for (int i = 0; i < TG; ++i)
#pragma unroll
    for (int j = 0; j < TG; ++j) {
      const int a = ty * TG + i, b = tx * TG + j;
      if (a < k && b < k) out[a * k + b] = g64[i][j];
    }
  __syncthreads();
  double* red = reinterpret_cast<double*>(&Us[0][0]);       // 8 warps x KP doubles fit in the U tile
  static_assert(sizeof(double) * 8 * KP <= sizeof(float) * R * KP, "reduction scratch too small");
  for (int q = 0; q < 4; ++q) {
#pragma unroll
    for (int cc = 0; cc < CPL; ++cc) {
      const int c = lane + 32 * cc;
      if (c < KP) red[warp * KP + c] = c64[q][cc];
    }
    __syncthreads();
    if (tid < k) {
      double s = 0.0;
      for (int w = 0; w < kPartialThreads / 32; ++w) s += red[w * KP + tid];
      out[k * k + q * k + tid] = s;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
reduce_partials_kernel(int n_blocks, int len, const double* __restrict__ block_out, double* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= len) return;
  double s = 0.0;
  for (int b = 0; b < n_blocks; ++b) s += block_out[(size_t)b * len + e];
  out[e] = s;
}

__device__ double block_sum_256(double v, double* scratch) {
  v = ep::warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < 8; ++w) s += scratch[w];
  return s;
}

__global__ void __launch_bounds__(256)
eigen_finalize_kernel(int k, double n_global, const double* __restrict__ P, float w_res, float w_orth,
                      int level0, const float* __restrict__ lam_target, float w_trace, float w_order,
                      float w_eigen, const float* __restrict__ lam_bar_extra, float* __restrict__ lam_out,
                      float* __restrict__ coef, double* __restrict__ loss_acc) {
  __shared__ double s_lam[128];
  __shared__ double scratch[8];
  const int tid = threadIdx.x;
  const double* G = P;
  const double* num = P + (size_t)k * k;
  const double* sKK = num + k;
  const double* sKM = sKK + k;
  const double* sMM = sKM + k;
  double res_j = 0.0, lam_j = 0.0, den_j = 1.0;
  if (tid < k) {
    den_j = G[(size_t)tid * k + tid] + 1e-12;
    lam_j = num[tid] / den_j;
    s_lam[tid] = lam_j;
    res_j = sKK[tid] - 2.0 * lam_j * sKM[tid] + lam_j * lam_j * sMM[tid];
    if (res_j < 0.0) res_j = 0.0;
  }
  const double res_sum = block_sum_256(res_j, scratch);
  double orth = 0.0;
  for (int e = tid; e < k * k; e += 256) {
    const int a = e / k, b = e - a * k;
    const double d = G[e] - (a == b ? 1.0 : 0.0);
    orth += d * d;
  }
  const double orth_sum = block_sum_256(orth, scratch);
  const double L_res = res_sum / (n_global * (double)k);
  const double L_orth = orth_sum / (double)k;
  // eigenvalue terms on this level's lambda (reference uses level 0 only)
  double tr = 0.0, ordr = 0.0, eig = 0.0, extra_bar = 0.0;
  if (level0 && tid < k) {
    tr = lam_j / (double)k;
    extra_bar += (double)w_trace / (double)k;
    if (tid + 1 < k) {                       // pair (tid, tid+1): relu(lam_tid - lam_{tid+1})
      const double d = s_lam[tid] - s_lam[tid + 1];
      if (d > 0.0) { ordr += d; extra_bar += (double)w_order; }
    }
    if (tid > 0) {                           // pair (tid-1, tid) contributes -w_order to this lam
      const double d = s_lam[tid - 1] - s_lam[tid];
      if (d > 0.0) extra_bar -= (double)w_order;
    }
    if (lam_target != nullptr) {
      const double d = lam_j - (double)lam_target[tid];
      eig = d * d / (double)k;
      extra_bar += (double)w_eigen * 2.0 * d / (double)k;
    }
  }
  const double tr_sum = block_sum_256(tr, scratch);
  const double ord_sum = block_sum_256(ordr, scratch);
  const double eig_sum = block_sum_256(eig, scratch);
  const double c_res = 2.0 * (double)w_res / (n_global * (double)k);
  float* c_lam = coef + 1;
  float* c_num = c_lam + k;
  float* c_den = c_num + k;
  float* c_G = c_den + k;
  double den_bar_j = 0.0;
  if (tid < k) {
    double lam_bar = -c_res * (sKM[tid] - lam_j * sMM[tid]) + extra_bar;
    if (lam_bar_extra != nullptr) lam_bar += (double)lam_bar_extra[tid];
    const double num_bar = lam_bar / den_j;
    den_bar_j = -lam_bar * lam_j / den_j;
    c_lam[tid] = (float)lam_j;
    c_num[tid] = (float)num_bar;
    c_den[tid] = (float)den_bar_j;
    if (lam_out) lam_out[tid] = (float)lam_j;
  }
  __syncthreads();
  const double go = 2.0 * (double)w_orth / (double)k;
  for (int e = tid; e < k * k; e += 256) {
    const int a = e / k, b = e - a * k;
    c_G[e] = (float)(go * (G[e] - (a == b ? 1.0 : 0.0)));   // den_bar is applied separately (coef)
  }
  if (tid == 0) {
    coef[0] = (float)c_res;
    const double t0 = (double)w_res * L_res, t1 = (double)w_orth * L_orth;
    const double t2 = (double)w_trace * tr_sum, t3 = (double)w_order * ord_sum, t4 = (double)w_eigen * eig_sum;
    loss_acc[0] += t0; loss_acc[1] += t1; loss_acc[2] += t2; loss_acc[3] += t3; loss_acc[4] += t4;
    loss_acc[5] += t0 + t1 + t2 + t3 + t4;
  }
}

// Backward preparation.  Tile = R rows x KP columns, 4 x 4 outputs per thread.
//   KU_bar = R_bar + a U;   MU_bar = -lam R_bar + U Gp;   D = a KU + MU Gp^T
//   with R_bar = c (KU - lam MU),  Gp = G_bar + diag(den_bar).
template <int KP>
__global__ void __launch_bounds__(256)
eigen_bwd_prepare_kernel(int n, int k, const float* __restrict__ U, int ldu, const float* __restrict__ KU,
                         const float* __restrict__ MU, int ld, const float* __restrict__ coef,
                         float* __restrict__ KU_bar, float* __restrict__ MU_bar, float* __restrict__ D) {
  constexpr int R = 4096 / KP;
  constexpr int GP = KP + 4;                    // pitch of the two Gram copies (float4 aligned)
  constexpr int TP = KP + 1;                    // pitch of the row tiles (conflict-free scalar reads)
  extern __shared__ __align__(16) float smem[];
  float* Gs = smem;                             // Gs[m][j]  = Gp[m][j]
  float* GTs = Gs + KP * GP;                    // GTs[m][j] = Gp[j][m]
  float* Us = GTs + KP * GP;                    // R x TP
  float* MUs = Us + R * TP;                     // R x TP
  float* s_lam = MUs + R * TP;                  // KP
  float* s_a = s_lam + KP;                      // KP   (num_bar)
  const int tid = threadIdx.x;
  const float c_res = coef[0];
  const float* c_lam = coef + 1;
  const float* c_num = c_lam + k;
  const float* c_den = c_num + k;
  const float* c_G = c_den + k;
  for (int e = tid; e < KP * KP; e += 256) {
    const int a = e / KP, b = e - a * KP;
    float g = 0.f;
    if (a < k && b < k) {
      g = c_G[a * k + b];
      if (a == b) g += c_den[a];
    }
    Gs[a * GP + b] = g;
    GTs[b * GP + a] = g;
  }
  for (int e = tid; e < KP; e += 256) {
    s_lam[e] = e < k ? c_lam[e] : 0.f;
    s_a[e] = e < k ? c_num[e] : 0.f;
  }
  constexpr int CG = KP / 4;                    // column groups
  const int tc = tid % CG, tr = tid / CG;       // tr in [0, R/4)
  const int n_tiles = (n + R - 1) / R;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int row0 = tile * R;
    __syncthreads();
    for (int e = tid; e < R * KP; e += 256) {
      const int r = e / KP, c = e - r * KP;
      const int row = row0 + r;
      const bool ok = (row < n) && (c < k);
      Us[r * TP + c] = ok ? __ldg(U + (size_t)row * ldu + c) : 0.f;
      MUs[r * TP + c] = ok ? __ldg(MU + (size_t)row * ld + c) : 0.f;
    }
    __syncthreads();
    float acc1[4][4], acc2[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc1[i][j] = 0.f; acc2[i][j] = 0.f; }
#pragma unroll 4
    for (int m = 0; m < KP; ++m) {
      const float4 g = *reinterpret_cast<const float4*>(Gs + m * GP + 4 * tc);
      const float4 gt = *reinterpret_cast<const float4*>(GTs + m * GP + 4 * tc);
      float u[4], mu[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { u[i] = Us[(4 * tr + i) * TP + m]; mu[i] = MUs[(4 * tr + i) * TP + m]; }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc1[i][0] = fmaf(u[i], g.x, acc1[i][0]); acc1[i][1] = fmaf(u[i], g.y, acc1[i][1]);
        acc1[i][2] = fmaf(u[i], g.z, acc1[i][2]); acc1[i][3] = fmaf(u[i], g.w, acc1[i][3]);
        acc2[i][0] = fmaf(mu[i], gt.x, acc2[i][0]); acc2[i][1] = fmaf(mu[i], gt.y, acc2[i][1]);
        acc2[i][2] = fmaf(mu[i], gt.z, acc2[i][2]); acc2[i][3] = fmaf(mu[i], gt.w, acc2[i][3]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = 4 * tr + i, row = row0 + r;
      if (row >= n) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = 4 * tc + j;
        if (c >= k) continue;
        const float ku = __ldg(KU + (size_t)row * ld + c);
        const float mu = MUs[r * TP + c], u = Us[r * TP + c];
        const float lam = s_lam[c], a = s_a[c];
        const float rbar = c_res * (ku - lam * mu);
        KU_bar[(size_t)row * ld + c] = rbar + a * u;
        MU_bar[(size_t)row * ld + c] = acc1[i][j] - lam * rbar;
        D[(size_t)row * ld + c] = fmaf(a, ku, acc2[i][j]);
      }
    }
  }
}

template <int KP>
int launch_bwd_prepare(int n, int k, const float* U, int ldu, const float* KU, const float* MU, int ld,
                       const float* coef, float* KU_bar, float* MU_bar, float* D, cudaStream_t st) {
  constexpr int R = 4096 / KP;
  const size_t smem = sizeof(float) * (2 * KP * (KP + 4) + 2 * R * (KP + 1) + 2 * KP);
  static bool configured = false;
  if (!configured) {
    EP_CUDA_CHECK(cudaFuncSetAttribute(eigen_bwd_prepare_kernel<KP>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int grid = partial_blocks(n, R);
  eigen_bwd_prepare_kernel<KP><<<grid, 256, smem, st>>>(n, k, U, ldu, KU, MU, ld, coef, KU_bar, MU_bar, D);
  EP_LAUNCH_CHECK("eigen_bwd_prepare_kernel");
  return EP_OK;
}

__global__ void __launch_bounds__(256)
scale_columns_kernel(long long n, int k, const float* __restrict__ U, int ldu, const double* __restrict__ G,
                     int ldg, double eps, float* __restrict__ out, int ldo) {
  const long long total = n * k;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / k;
    const int c = (int)(i - r * k);
    const float nrm = (float)sqrt(G[(size_t)c * ldg + c] + eps);
    out[r * ldo + c] = U[r * ldu + c] / nrm;
  }
}

__global__ void __launch_bounds__(256)
axpy_out_kernel(size_t n, float alpha, const float* __restrict__ alpha_dev, const float* __restrict__ a,
                const float* __restrict__ b, float* __restrict__ out) {
  const float al = alpha_dev ? *alpha_dev : alpha;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = __fadd_rn(a[i], __fmul_rn(al, b[i]));     // corr = scale*raw (rounded), U = base + corr
}

int stream_grid(size_t total) {
  size_t g = (total + 255) / 256;
  const size_t cap = (size_t)ep::sm_count() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

extern "C" {

size_t ep_eigen_partials_len(int k) { return k > 0 ? (size_t)partials_len(k) : 0; }

size_t ep_eigen_partials_workspace_bytes(int k) {
  if (k <= 0) return 0;
  return sizeof(double) * (size_t)partials_len(k) * (size_t)(ep::sm_count() * 4);
}

size_t ep_eigen_coef_len(int k) { return k > 0 ? (size_t)(1 + 3 * k + k * k) : 0; }

int ep_eigen_partials_f32(int n, int k, const float* U, int ldu, const float* KU, const float* MU, int ld,
                          double* out, void* workspace, size_t workspace_bytes, ep_stream_t stream) {
  EP_REQUIRE(n >= 0 && k > 0, "bad size");
  EP_REQUIRE(k <= 128, "k > 128 not instantiated");
  EP_REQUIRE(out && workspace && (n == 0 || (U && KU && MU)), "null pointer");
  EP_REQUIRE(ldu >= k && ld >= k, "leading dimension < k");
  cudaStream_t st = ep::as_stream(stream);
  const int len = partials_len(k);
  const int tg = k <= 16 ? 1 : k <= 32 ? 2 : k <= 64 ? 4 : 8;
  const int rows = tg == 8 ? 16 : 32;
  const int grid = partial_blocks(n, rows);
  if (workspace_bytes < sizeof(double) * (size_t)len * grid) {
    ep::set_error("ep_eigen_partials_f32: workspace too small");
    return EP_ERR_WORKSPACE;
  }
  double* blocks = static_cast<double*>(workspace);
  switch (tg) {
    case 1: eigen_partials_kernel<1><<<grid, kPartialThreads, 0, st>>>(n, k, U, ldu, KU, MU, ld, blocks); break;
    case 2: eigen_partials_kernel<2><<<grid, kPartialThreads, 0, st>>>(n, k, U, ldu, KU, MU, ld, blocks); break;
    case 4: eigen_partials_kernel<4><<<grid, kPartialThreads, 0, st>>>(n, k, U, ldu, KU, MU, ld, blocks); break;
    default: eigen_partials_kernel<8><<<grid, kPartialThreads, 0, st>>>(n, k, U, ldu, KU, MU, ld, blocks); break;
  }
  EP_LAUNCH_CHECK("eigen_partials_kernel");
  reduce_partials_kernel<<<ep::ceil_div(len, 256), 256, 0, st>>>(grid, len, blocks, out);
  EP_LAUNCH_CHECK("reduce_partials_kernel");
  return EP_OK;
}

int ep_eigen_finalize_f32(int k, double n_global, const double* partials, float w_res, float w_orth,
                          int level0, const float* lam_target, float w_trace, float w_order, float w_eigen,
                          const float* lam_bar_extra, float* lam_out, float* coef, double* loss_acc,
                          ep_stream_t stream) {
  EP_REQUIRE(k > 0 && k <= 128, "k out of range");
  EP_REQUIRE(n_global > 0, "n_global must be positive");
  EP_REQUIRE(partials && coef && loss_acc, "null pointer");
  eigen_finalize_kernel<<<1, 256, 0, ep::as_stream(stream)>>>(k, n_global, partials, w_res, w_orth, level0,
                                                               lam_target, w_trace, w_order, w_eigen,
                                                               lam_bar_extra, lam_out, coef, loss_acc);
  EP_LAUNCH_CHECK("eigen_finalize_kernel");
  return EP_OK;
}

int ep_eigen_bwd_prepare_f32(int n, int k, const float* U, int ldu, const float* KU, const float* MU, int ld,
                             const float* coef, float* KU_bar, float* MU_bar, float* D, ep_stream_t stream) {
  EP_REQUIRE(n >= 0 && k > 0 && k <= 128, "bad size");
  if (n == 0) return EP_OK;
  EP_REQUIRE(U && KU && MU && coef && KU_bar && MU_bar && D, "null pointer");
  EP_REQUIRE(ldu >= k && ld >= k, "leading dimension < k");
  cudaStream_t st = ep::as_stream(stream);
  if (k <= 16) return launch_bwd_prepare<16>(n, k, U, ldu, KU, MU, ld, coef, KU_bar, MU_bar, D, st);
  if (k <= 32) return launch_bwd_prepare<32>(n, k, U, ldu, KU, MU, ld, coef, KU_bar, MU_bar, D, st);
  if (k <= 64) return launch_bwd_prepare<64>(n, k, U, ldu, KU, MU, ld, coef, KU_bar, MU_bar, D, st);
  return launch_bwd_prepare<128>(n, k, U, ldu, KU, MU, ld, coef, KU_bar, MU_bar, D, st);
}

int ep_scale_columns_rsqrt_f32(int n, int k, const float* U, int ldu, const double* G, int ldg, double eps,
                               float* out, int ldo, ep_stream_t stream) {
  EP_REQUIRE(n >= 0 && k > 0, "bad size");
  if (n == 0) return EP_OK;
  EP_REQUIRE(U && G && out, "null pointer");
  scale_columns_kernel<<<stream_grid((size_t)n * k), 256, 0, ep::as_stream(stream)>>>(n, k, U, ldu, G, ldg, eps,
                                                                                      out, ldo);
  EP_LAUNCH_CHECK("scale_columns_kernel");
  return EP_OK;
}

int ep_axpy_out_f32(size_t n, float alpha, const float* alpha_dev, const float* a, const float* b, float* out,
                    ep_stream_t stream) {
  if (n == 0) return EP_OK;
  EP_REQUIRE(a && b && out, "null pointer");
  axpy_out_kernel<<<stream_grid(n), 256, 0, ep::as_stream(stream)>>>(n, alpha, alpha_dev, a, b, out);
  EP_LAUNCH_CHECK("axpy_out_kernel");
  return EP_OK;
}

}  // extern "C"
