// clip_grad_norm_ + Adam (coupled L2 weight decay) on one flat parameter buffer.
// Reference: src/multigrid_model.py:218-220 (Adam(lr, weight_decay)), :259-260 (clip 10.0, step).
// Update order follows torch.optim.Adam's single-tensor path:
//   g += wd * p;  m.lerp_(g, 1-b1);  v = b2 v + (1-b2) g g;
//   p -= (lr / (1-b1^t)) * m / (sqrt(v) / sqrt(1-b2^t) + eps)
#include "ep_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
sqnorm_partial_kernel(size_t n, const float* __restrict__ g, double* __restrict__ parts) {
  __shared__ double sh[8];
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double v = (double)g[i];
    s += v * v;
  }
  s = ep::warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    parts[blockIdx.x] = t;
  }
}

constexpr int kNormBlocks = 64;
__device__ double g_norm_parts[kNormBlocks];

__global__ void sqnorm_final_kernel(int n_parts, double* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < n_parts; ++i) t += g_norm_parts[i];
    *out = t;
  }
}

__global__ void __launch_bounds__(256)
adam_clip_kernel(size_t n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                 float* __restrict__ v, float lr, const float* __restrict__ hyper_dev, float beta1, float beta2,
                 float eps, float wd, int step_v, float max_norm, const double* __restrict__ sq_norm) {
  float coef = 1.0f;
  if (sq_norm != nullptr && max_norm > 0.f) {
    const float total = (float)sqrt(*sq_norm);
    coef = fminf(max_norm / (total + 1e-6f), 1.0f);
  }
  // hyper_dev = {float lr, int32 step} in device memory (CUDA-graph replay), else by value.  The bias corrections
  // are evaluated HERE, in double, from the integer step in both cases: replay and eager launches agree bit for bit.
  const float lr_now = hyper_dev ? hyper_dev[0] : lr;
  const int t = hyper_dev ? __float_as_int(hyper_dev[1]) : step_v;
  const double bc1 = 1.0 - pow((double)beta1, (double)t);
  const float step_size = (float)((double)lr_now / bc1);
  const float bc2s = (float)sqrt(1.0 - pow((double)beta2, (double)t));
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float pi = p[i];
    float gi = g[i] * coef;
    gi = fmaf(wd, pi, gi);
    const float mi = m[i] + (gi - m[i]) * (1.0f - beta1);
    const float vi = fmaf(1.0f - beta2, gi * gi, v[i] * beta2);
    const float denom = sqrtf(vi) / bc2s + eps;
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - step_size * (mi / denom);
  }
}

}  // namespace

extern "C" {

int ep_grad_sqnorm_f32(size_t n, const float* g, double* sq_out, ep_stream_t stream) {
  EP_REQUIRE(sq_out && (n == 0 || g), "null pointer");
  cudaStream_t st = ep::as_stream(stream);
  double* parts = nullptr;
  EP_CUDA_CHECK(cudaGetSymbolAddress(reinterpret_cast<void**>(&parts), g_norm_parts));
  sqnorm_partial_kernel<<<kNormBlocks, 256, 0, st>>>(n, g, parts);
  EP_LAUNCH_CHECK("sqnorm_partial_kernel");
  sqnorm_final_kernel<<<1, 32, 0, st>>>(kNormBlocks, sq_out);
  EP_LAUNCH_CHECK("sqnorm_final_kernel");
  return EP_OK;
}

int ep_adam_clip_step_f32(size_t n, float* p, const float* g, float* m, float* v, float lr, const float* hyper_dev,
                          float beta1, float beta2, float eps, float weight_decay, int step, float max_norm,
                          const double* sq_norm, ep_stream_t stream) {
  if (n == 0) return EP_OK;
  EP_REQUIRE(p && g && m && v, "null pointer");
  EP_REQUIRE(step >= 1 || hyper_dev, "step counts from 1");
  size_t grid = (n + 255) / 256;
  const size_t cap = (size_t)ep::sm_count() * 8;
  if (grid > cap) grid = cap;
  adam_clip_kernel<<<(unsigned)grid, 256, 0, ep::as_stream(stream)>>>(n, p, g, m, v, lr, hyper_dev, beta1, beta2, eps,
                                                                     weight_decay, step >= 1 ? step : 1, max_norm,
                                                                     sq_norm);
  EP_LAUNCH_CHECK("adam_clip_kernel");
  return EP_OK;
}

}  // extern "C"
