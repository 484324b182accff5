// How fast does one SM pull L2-resident data into shared memory?  cp.async.bulk (1-D) with several copies in flight
// versus plain 16-byte loads + st.shared, all SMs reading the SAME 1 MB or each its own.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bulk_bw bulk_bw.cu && ./bulk_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int RING = 128 * 1024;

// one thread issues; `depth` copies of `chunk` bytes in flight; total bytes per CTA = total
__global__ void __launch_bounds__(128, 1) bulk_kernel(const uint8_t* src, size_t per_cta_stride, int region, int chunk, int depth,
                                                     int total, long long* clk, int issuers = 1) {
  extern __shared__ __align__(128) uint8_t sm_all[];
  __shared__ uint64_t bar_all[64];
  if (threadIdx.x == 0) { for (int i = 0; i < 64; ++i) mbar_init(&bar_all[i], 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && w < issuers) {
    uint64_t* bar = bar_all + 16 * w;
    uint8_t* sm = sm_all + (size_t)w * (RING / issuers);
    const uint8_t* base = src + blockIdx.x * per_cta_stride + (size_t)w * 65536;
    const int n = total / chunk / issuers;
    long long t0 = clock64();
    for (int i = 0; i < n + depth; ++i) {
      if (i >= depth) { const int j = i - depth; mbar_wait(&bar[j % depth], (j / depth) & 1); }
      if (i < n) {
        const int s = i % depth;
        mbar_expect_tx(&bar[s], chunk);
        bulk_g2s(sm + (size_t)s * chunk, base + ((size_t)i * chunk) % region, chunk, &bar[s]);
      }
    }
    long long t1 = clock64();
    if (w == 0) clk[blockIdx.x] = t1 - t0;
  }
}

// 256 threads: 16-byte loads (L2 -> registers) + st.shared, `unroll` loads in flight per thread
__global__ void __launch_bounds__(256, 1) ldg_kernel(const uint8_t* src, size_t per_cta_stride, int region, int total, long long* clk) {
  extern __shared__ __align__(128) uint8_t sm[];
  const uint4* base = reinterpret_cast<const uint4*>(src + blockIdx.x * per_cta_stride);
  const int n16 = total / 16, r16 = region / 16;
  __syncthreads();
  long long t0 = clock64();
  for (int i = threadIdx.x; i < n16; i += 256 * 8) {
    uint4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcg(base + (i + u * 256) % r16);
#pragma unroll
    for (int u = 0; u < 8; ++u) reinterpret_cast<uint4*>(sm)[(i + u * 256) % (RING / 16)] = v[u];
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t MB = 1 << 20;
  uint8_t* buf; cudaMalloc(&buf, 160 * MB); cudaMemset(buf, 1, 160 * MB);
  long long* clk; cudaMallocManaged(&clk, sms * sizeof(long long));
  cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RING);
  cudaFuncSetAttribute(ldg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RING);
  const int total = 16 * MB;
  const int region = 1 * MB;
  for (int shared_src = 1; shared_src >= 0; --shared_src) {
    const size_t stride = shared_src ? 0 : MB;
    printf("== %s, %d SMs, %d MB per SM from a 1 MB region\n", shared_src ? "all SMs read the SAME 1 MB" : "each SM reads its OWN 1 MB", sms, total >> 20);
    const int chunks[] = {2048, 8192, 16384, 32768};
    const int depths[] = {1, 2, 4, 8};
    for (int c : chunks)
      for (int d : depths) {
        if ((size_t)c * d > RING) continue;
        for (int rep = 0; rep < 2; ++rep) {
          bulk_kernel<<<sms, 128, RING>>>(buf, stride, region, c, d, total, clk);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        }
        double mean = 0; long long mx = 0;
        for (int i = 0; i < sms; ++i) { mean += clk[i]; if (clk[i] > mx) mx = clk[i]; }
        mean /= sms;
        printf("bulk chunk %6d B x depth %d: %6.1f B/clk/SM mean (%5.1f slowest SM)\n", c, d, total / mean, (double)total / mx);
      }
    for (int rep = 0; rep < 2; ++rep) { ldg_kernel<<<sms, 256, RING>>>(buf, stride, region, total, clk); cudaDeviceSynchronize(); }
    double mean = 0; for (int i = 0; i < sms; ++i) mean += clk[i]; mean /= sms;
    printf("ld.global.cg 16 B x 8 in flight x 256 threads + st.shared: %6.1f B/clk/SM\n", total / mean);
  }
  for (int iss : {2, 4})
    for (int c : {2048, 8192}) {
      for (int rep = 0; rep < 2; ++rep) { bulk_kernel<<<sms, 128, RING>>>(buf, 0, region, c, 4, total, clk, iss); cudaDeviceSynchronize(); }
      double mean = 0; for (int i = 0; i < sms; ++i) mean += clk[i]; mean /= sms;
      printf("%d issuing warps, bulk chunk %d B x depth 4 each (same 1 MB): %6.1f B/clk/SM\n", iss, c, total / mean);
    }
  // one SM alone
  bulk_kernel<<<1, 128, RING>>>(buf, 0, region, 16384, 4, total, clk); cudaDeviceSynchronize();
  bulk_kernel<<<1, 128, RING>>>(buf, 0, region, 16384, 4, total, clk); cudaDeviceSynchronize();
  printf("ONE SM alone, bulk 16 KB x 4: %6.1f B/clk\n", (double)total / clk[0]);
  ldg_kernel<<<1, 256, RING>>>(buf, 0, region, total, clk); cudaDeviceSynchronize();
  printf("ONE SM alone, ld.global path: %6.1f B/clk\n", (double)total / clk[0]);
  return 0;
}
