// Sparse linear-FEM assembly on the device: per-triangle stiffness / mass blocks in fp64
// and an ordered segmented sum into CSR.  Reference arithmetic: src/Mesh.py:180-198 (Bmatrix), :228-234
// (StiffnessMatrix = B^T B / (2J), MassMatrix = J/12 [[2,1,1],[1,2,1],[1,1,2]]), :348-364 (computeLaplacian
// accumulates K[tri[a], tri[b]] += k[a][b] in triangle order).
//
// The (row, col) keys of the 9 T element entries are sorted with a STABLE sort on the host side of the C ABI
// (any stable radix sort; the Python layer uses torch.sort), so equal keys keep triangle order and the
// segmented sum below adds them in exactly the order of the reference loop - deterministic, no atomics.
#include "ep_common.cuh"

namespace {

__device__ __forceinline__ double dot3(const double* a, const double* b) {
  return __dadd_rn(__dadd_rn(__dmul_rn(a[0], b[0]), __dmul_rn(a[1], b[1])), __dmul_rn(a[2], b[2]));
}

__global__ void __launch_bounds__(256)
fem_element_kernel(long long T, const double* __restrict__ verts, const int32_t* __restrict__ tris, long long n_verts,
                   double* __restrict__ k_el, double* __restrict__ m_el, long long* __restrict__ keys) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += (long long)gridDim.x * blockDim.x) {
    const int v[3] = {tris[t * 3], tris[t * 3 + 1], tris[t * 3 + 2]};
    double p[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int c = 0; c < 3; ++c) p[a][c] = verts[(long long)v[a] * 3 + c];
    double d10[3], d20[3], d02[3], d21[3], d12[3], d01[3], e1[3], e2[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      d10[c] = p[1][c] - p[0][c]; d20[c] = p[2][c] - p[0][c]; d02[c] = p[0][c] - p[2][c];
      d21[c] = p[2][c] - p[1][c]; d12[c] = p[1][c] - p[2][c]; d01[c] = p[0][c] - p[1][c];
    }
    const double n1 = __dsqrt_rn(dot3(d10, d10));
#pragma unroll
    for (int c = 0; c < 3; ++c) e1[c] = __ddiv_rn(d10[c], n1);
    const double pr = dot3(d20, e1);
#pragma unroll
    for (int c = 0; c < 3; ++c) e2[c] = __dsub_rn(d20[c], __dmul_rn(pr, e1[c]));
    const double n2 = __dsqrt_rn(dot3(e2, e2));
#pragma unroll
    for (int c = 0; c < 3; ++c) e2[c] = __ddiv_rn(e2[c], n2);
    const double x21 = dot3(d10, e1), x13 = dot3(d02, e1), x32 = dot3(d21, e1);
    const double y23 = dot3(d12, e2), y31 = dot3(d20, e2), y12 = dot3(d01, e2);
    const double J = __dsub_rn(__dmul_rn(x13, y23), __dmul_rn(y31, x32));
    const double B0[3] = {y23, y31, y12}, B1[3] = {x32, x13, x21};
    const double twoJ = __dmul_rn(2.0, J);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const double kk = __ddiv_rn(__dadd_rn(__dmul_rn(B0[a], B0[b]), __dmul_rn(B1[a], B1[b])), twoJ);
        const double mm = __ddiv_rn(__dmul_rn(a == b ? 2.0 : 1.0, J), 12.0);
        const long long e = t * 9 + a * 3 + b;
        k_el[e] = kk;
        m_el[e] = mm;
        keys[e] = (long long)v[a] * n_verts + v[b];
      }
  }
}

// one thread per CSR entry: sum its run of sorted element entries in order
__global__ void __launch_bounds__(256)
fem_segment_sum_kernel(long long nnz, const long long* __restrict__ seg_start, const long long* __restrict__ seg_count,
                       const long long* __restrict__ perm, const double* __restrict__ k_el,
                       const double* __restrict__ m_el, const long long* __restrict__ uniq_keys, long long n_verts,
                       int32_t* __restrict__ col, float* __restrict__ valK, float* __restrict__ valM,
                       double* __restrict__ valK64, double* __restrict__ valM64) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (long long)gridDim.x * blockDim.x) {
    const long long s = seg_start[e], c = seg_count[e];
    double ak = 0.0, am = 0.0;
    for (long long j = 0; j < c; ++j) {
      const long long src = perm[s + j];
      ak = __dadd_rn(ak, k_el[src]);
      am = __dadd_rn(am, m_el[src]);
    }
    col[e] = (int32_t)(uniq_keys[e] % n_verts);
    valK[e] = (float)ak;
    valM[e] = (float)am;
    if (valK64) valK64[e] = ak;
    if (valM64) valM64[e] = am;
  }
}

int grid_for(long long total) {
  long long g = (total + 255) / 256;
  const long long cap = (long long)ep::sm_count() * 16;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace

extern "C" {

int ep_fem_elements_f64(int64_t n_tris, const double* verts, const int32_t* tris, int64_t n_verts, double* k_el,
                        double* m_el, int64_t* keys, ep_stream_t stream) {
  EP_REQUIRE(n_tris > 0 && n_verts > 0 && verts && tris && k_el && m_el && keys, "bad argument");
  fem_element_kernel<<<grid_for(n_tris), 256, 0, ep::as_stream(stream)>>>(n_tris, verts, tris, n_verts, k_el, m_el,
                                                                          reinterpret_cast<long long*>(keys));
  EP_LAUNCH_CHECK("fem_element_kernel");
  return EP_OK;
}

int ep_fem_segment_sum_f64(int64_t nnz, const int64_t* seg_start, const int64_t* seg_count, const int64_t* perm,
                           const double* k_el, const double* m_el, const int64_t* uniq_keys, int64_t n_verts,
                           int32_t* col, float* valK, float* valM, double* valK64, double* valM64,
                           ep_stream_t stream) {
  EP_REQUIRE(nnz > 0 && seg_start && seg_count && perm && k_el && m_el && uniq_keys && col && valK && valM, "bad argument");
  fem_segment_sum_kernel<<<grid_for(nnz), 256, 0, ep::as_stream(stream)>>>(
      nnz, reinterpret_cast<const long long*>(seg_start), reinterpret_cast<const long long*>(seg_count),
      reinterpret_cast<const long long*>(perm), k_el, m_el, reinterpret_cast<const long long*>(uniq_keys), n_verts, col,
      valK, valM, valK64, valM64);
  EP_LAUNCH_CHECK("fem_segment_sum_kernel");
  return EP_OK;
}

}  // extern "C"
