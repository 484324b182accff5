// Micro-benchmark: TMEM -> register read throughput of tcgen05.ld for several shapes / warp counts (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm_bw ldtm_bw.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int SHAPE>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t& sink) {
  uint32_t v[32];
  if (SHAPE == 0) {          // 32x32b.x32 : 4 KB per warp instruction (thread = lane, 32 columns)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),
        "=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31]) : "r"(taddr));
  } else if (SHAPE == 1) {   // 16x256b.x8 : 16 lanes x 64 columns = 4 KB (32 regs per thread)
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),
        "=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31]) : "r"(taddr));
  } else if (SHAPE == 2) {   // 16x128b.x16 : 16 lanes x 64 columns = 4 KB
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),
        "=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31]) : "r"(taddr));
  } else {                   // 32x32b.x8 : 1 KB per instruction
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]) : "r"(taddr));
#pragma unroll
    for (int i = 8; i < 32; ++i) v[i] = 0;
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s ^= v[i];
  sink ^= s;
}

template <int SHAPE>
__global__ void bench(int n_warps_active, int reps, long long* out, uint32_t* sink_out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  uint32_t sink = 0;
  const int q = warp & 3;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < n_warps_active) {
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int cb = 0; cb < 8; ++cb) {
        const uint32_t col = (uint32_t)((cb * 32 + (warp >> 2) * 64) & 511);
        ld<SHAPE>(base + ((uint32_t)(q * 32) << 16) + col, sink);
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (sink == 0x12345678u) sink_out[0] = sink;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512));
}

int main() {
  long long* out; uint32_t* sink;
  cudaMalloc(&out, 8 * 148); cudaMalloc(&sink, 4);
  const char* names[4] = {"32x32b.x32 (4 KB)", "16x256b.x8 (4 KB)", "16x128b.x16 (4 KB)", "32x32b.x8 (1 KB)"};
  const int bytes[4] = {4096, 4096, 4096, 1024};
  for (int shape = 0; shape < 4; ++shape) {
    for (int nw : {1, 4, 8, 16}) {
      const int reps = 64;
      for (int it = 0; it < 2; ++it) {
        if (shape == 0) bench<0><<<148, 512>>>(nw, reps, out, sink);
        if (shape == 1) bench<1><<<148, 512>>>(nw, reps, out, sink);
        if (shape == 2) bench<2><<<148, 512>>>(nw, reps, out, sink);
        if (shape == 3) bench<3><<<148, 512>>>(nw, reps, out, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", names[shape], cudaGetErrorString(e)); return 1; }
      }
      long long h[148];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      const double total = (double)nw * reps * 8 * bytes[shape];
      printf("%-20s warps %2d : %8lld clk  -> %.1f B/clk/SM\n", names[shape], nw, h[0], total / (double)h[0]);
    }
  }
  return 0;
}
