// k nearest neighbours on a uniform grid hash (fp64 distances): the graph / prolongation pre-processing the reference
// does with scikit-learn (src/utils.py:39-75: NearestNeighbors(n_neighbors=k).fit(X_ref).kneighbors(X_query)).
//
// Reference points are binned into cubic cells (host side: cell ids -> stable sort -> cell_start), the kernel scans,
// for every query point, the cells of growing Chebyshev rings around the query's cell and keeps the k best
// (distance^2, index) pairs in an insertion-sorted list - ties are broken by the smaller reference index, so the result
// is a deterministic function of the point sets.  A ring r can be skipped as soon as k candidates are known whose
// k-th distance is below the distance to that ring ((r - 1) * cell, measured from the query's own cell walls).
// One thread per query; lists live in local memory (k <= 64).  Output rows are sorted by (distance, index).
#include <math.h>
#include "ep_common.cuh"

namespace {

constexpr int KNN_MAX_K = 64;

__global__ void __launch_bounds__(128)
knn_grid_kernel(long long n_query, const double* __restrict__ Q, long long n_ref, const double* __restrict__ R,
                const long long* __restrict__ order, const long long* __restrict__ cell_start, double lo_x, double lo_y,
                double lo_z, double cell, int gx, int gy, int gz, int k, long long* __restrict__ out_idx,
                double* __restrict__ out_dist) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_query) return;
  const double px = Q[3 * q], py = Q[3 * q + 1], pz = Q[3 * q + 2];
  double bd[KNN_MAX_K];
  long long bi[KNN_MAX_K];
  int cnt = 0;
  const double inv = 1.0 / cell;
  const int cx = min(max((int)floor((px - lo_x) * inv), 0), gx - 1);
  const int cy = min(max((int)floor((py - lo_y) * inv), 0), gy - 1);
  const int cz = min(max((int)floor((pz - lo_z) * inv), 0), gz - 1);
  // distance from the query to the walls of its own cell (lower bound of what ring r >= 1 can contain: see below)
  const double fx = (px - lo_x) - cx * cell, fy = (py - lo_y) - cy * cell, fz = (pz - lo_z) - cz * cell;
  const double wall = fmax(0.0, fmin(fmin(fmin(fx, cell - fx), fmin(fy, cell - fy)), fmin(fz, cell - fz)));
  const int rmax = max(max(gx, gy), gz);
  for (int r = 0; r <= rmax; ++r) {
    if (cnt == k && r >= 1) {
      const double reach = wall + (double)(r - 1) * cell;       // every point of ring r is at least this far away
      if (bd[k - 1] < reach * reach) break;
    }
    const int x0 = cx - r, x1 = cx + r, y0 = cy - r, y1 = cy + r, z0 = cz - r, z1 = cz + r;
    for (int x = max(x0, 0); x <= min(x1, gx - 1); ++x) {
      for (int y = max(y0, 0); y <= min(y1, gy - 1); ++y) {
        const bool shell_xy = (x == x0 || x == x1 || y == y0 || y == y1);
        for (int z = max(z0, 0); z <= min(z1, gz - 1); ++z) {
          if (!shell_xy && z != z0 && z != z1) { z = max(z, z1 - 1); continue; }      // interior of the cube: already visited
          const long long c = ((long long)x * gy + y) * gz + z;
          const long long s = cell_start[c], e = cell_start[c + 1];
          for (long long t = s; t < e; ++t) {
            const long long j = order[t];
            const double dx = R[3 * j] - px, dy = R[3 * j + 1] - py, dz = R[3 * j + 2] - pz;
            const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            if (cnt == k && !(d2 < bd[k - 1] || (d2 == bd[k - 1] && j < bi[k - 1]))) continue;
            int pos = cnt < k ? cnt : k - 1;                    // insertion sort by (d2, index)
            while (pos > 0 && (bd[pos - 1] > d2 || (bd[pos - 1] == d2 && bi[pos - 1] > j))) {
              bd[pos] = bd[pos - 1];
              bi[pos] = bi[pos - 1];
              --pos;
            }
            bd[pos] = d2;
            bi[pos] = j;
            if (cnt < k) ++cnt;
          }
        }
      }
    }
  }
  for (int t = 0; t < k; ++t) {
    out_idx[q * k + t] = t < cnt ? bi[t] : -1;
    if (out_dist) out_dist[q * k + t] = t < cnt ? sqrt(bd[t]) : INFINITY;
  }
}

}  // namespace

extern "C" {

int ep_knn_grid_f64(int64_t n_query, const double* query, int64_t n_ref, const double* ref, const int64_t* order,
                    const int64_t* cell_start, const double* lo, double cell, const int64_t* dims, int k,
                    int64_t* out_idx, double* out_dist, ep_stream_t stream) {
  EP_REQUIRE(n_query >= 0 && n_ref > 0 && k > 0 && k <= KNN_MAX_K, "bad size (1 <= k <= 64)");
  if (n_query == 0) return EP_OK;
  EP_REQUIRE(query && ref && order && cell_start && lo && dims && out_idx && cell > 0.0, "bad argument");
  EP_REQUIRE(dims[0] > 0 && dims[1] > 0 && dims[2] > 0 && dims[0] * dims[1] * dims[2] < (1ll << 40), "bad grid");
  const long long blocks = (n_query + 127) / 128;
  knn_grid_kernel<<<(unsigned)blocks, 128, 0, ep::as_stream(stream)>>>(
      n_query, query, n_ref, ref, reinterpret_cast<const long long*>(order), reinterpret_cast<const long long*>(cell_start),
      lo[0], lo[1], lo[2], cell, (int)dims[0], (int)dims[1], (int)dims[2], k, reinterpret_cast<long long*>(out_idx), out_dist);
  EP_LAUNCH_CHECK("knn_grid_kernel");
  return EP_OK;
}

}  // extern "C"
