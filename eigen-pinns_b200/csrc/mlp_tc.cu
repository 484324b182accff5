// Corrector MLP on the 5th-generation tensor cores (bf16 operands, fp32 accumulation in TMEM).
// Reference arithmetic: src/corrector_model.py:12-21,31 and the autograd backward of
// src/multigrid_model.py:258 (here analytic, layer by layer).
//
// ---- packed activation layout -------------------------------------------------------------------
// Activations (and their gradients) live in HBM as bf16 in 128-vertex tiles of "core-matrix-major"
// order:   packed[tile][c][r][8]   c = feature / 8 (16-byte chunk), r = vertex in tile (0..127),
// so one tile of a 256-wide layer is 64 KB contiguous.  This is exactly the tcgen05 canonical
// NO-SWIZZLE shared-memory layout, both ways round:
//   * as a K-major operand  (rows = vertices, K = features):  8 x 16 B core matrices, SBO = 128 B
//     between 8-vertex groups, LBO = 2048 B between feature chunks           -> forward, dX GEMMs
//   * as an MN-major operand (MN = features, K = vertices):   LBO(K-group) = 128 B, SBO(MN chunk)
//     = 2048 B                                                               -> dW = dZ^T H GEMM
// Hence a tile goes HBM -> SMEM with plain 1-D TMA bulk copies (cp.async.bulk, no tensor map), the
// epilogue's 16-byte stores of thread r / chunk c are perfectly coalesced (consecutive threads ->
// consecutive 16 B), and no transposition is ever materialised.
// Weights are re-packed (fp32 -> bf16) once per step into [K/8][N][8] (K-major B operand), both W
// and W^T.
//
// ---- kernels ------------------------------------------------------------------------------------
//  tc_linear_kernel   C = A W^T with W resident in SMEM for the whole (persistent) CTA, A tiles
//                     streamed through an 8-stage TMA ring, two TMEM accumulators so the epilogue of
//                     tile i overlaps the MMAs of tile i+1.  Epilogues: bias+ReLU -> packed bf16 |
//                     ReLU-mask -> packed bf16 (dX) | bias -> fp32 rows + U_pred = U_base + s*corr.
//  tc_dw_kernel       dW = X^T Y accumulated in TMEM over all tiles of the CTA (256 x 256 fp32 = all
//                     512 TMEM columns), both operands MN-major; the otherwise idle warps sum the
//                     columns of dZ from SMEM for db.  Per-CTA partials are reduced in fixed order.
// Warp roles: warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issuer (one lane),
// warps 2..5 = epilogue (TMEM lane quadrant = warp % 4).
//
// Every layer is HBM-bound in this form (AI = 128 flop/B at width 256 vs. a ridge of ~250);
// DESIGN.md derives the byte counts the roofline numbers use.
#include <cuda_bf16.h>
#include "ep_common.cuh"

namespace tc {

constexpr int TILE_M = 128;
constexpr int CHUNK_BYTES = TILE_M * 16;          // one 8-feature chunk of a tile: 2048 B
constexpr int STAGE_CHUNKS = 4;                   // K = 32 per ring stage
constexpr int STAGE_BYTES = STAGE_CHUNKS * CHUNK_BYTES;   // 8 KB
constexpr int N_STAGES = 8;
constexpr int LINEAR_THREADS = 192;
constexpr uint32_t TMEM_COLS = 512;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0; spin < (1u << 26); ++spin)
    if (mbar_try_wait(bar, parity)) return;
  printf("eigenpinns_b200: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
  __trap();
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor, SWIZZLE_NONE, Blackwell version field = 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: bf16 x bf16 -> fp32
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(cols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

enum LinearMode { MODE_HIDDEN = 0, MODE_DX = 1, MODE_FINAL = 2 };

struct LinearArgs {
  const uint8_t* A;          // packed input tiles, KC chunks each
  const uint8_t* B;          // packed weights [KC][N][8] bf16
  int n_tiles, KC, N;
  const float* bias; int n_bias;   // bias[0..n_bias) (columns beyond are padding) or NULL
  int relu;
  uint8_t* out_packed;       // MODE_HIDDEN / MODE_DX: [tile][N/8][128][8]
  uint32_t* mask_out;        // MODE_HIDDEN: ReLU bit mask [row][N/32] (bit j of word w: activation 32w+j > 0), may be NULL
  const uint32_t* mask_in;   // MODE_DX: the bit mask written by the forward pass of the previous layer
  float* corr; int ldc;      // MODE_FINAL: fp32 rows
  const float* U_base; float* U_pred; int ldu; float scale; const float* scale_dev;
  int n_rows, n_out;
};

template <int MODE>
__global__ void __launch_bounds__(LINEAR_THREADS, 1) tc_linear_kernel(LinearArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int KC = a.KC, N = a.N;
  uint8_t* Bs = smem;                                           // KC * N * 16 bytes
  uint8_t* As = Bs + (size_t)KC * N * 16;                       // N_STAGES * STAGE_BYTES
  uint64_t* bars = reinterpret_cast<uint64_t*>(As + N_STAGES * STAGE_BYTES);
  uint64_t* full = bars;                 // [N_STAGES]
  uint64_t* empty = bars + N_STAGES;     // [N_STAGES]
  uint64_t* tfull = bars + 2 * N_STAGES; // [2]
  uint64_t* tempty = tfull + 2;          // [2]
  uint64_t* bready = tempty + 2;         // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bready + 1);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 2);      // N floats
  float* stage_f = bias_s + N;                                  // MODE_FINAL only: 4 x 32 x 33 floats

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < N_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
    mbar_init(&tempty[0], 128); mbar_init(&tempty[1], 128);
    mbar_init(bready, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < N; i += blockDim.x) bias_s[i] = (a.bias && i < a.n_bias) ? a.bias[i] : 0.f;
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int k_stages = KC / STAGE_CHUNKS;

  if (warp == 0) {
    if (lane == 0) {
      // resident weights: one barrier, copies of <= 64 KB
      const uint32_t b_bytes = (uint32_t)KC * N * 16;
      mbar_expect_tx(bready, b_bytes);
      for (uint32_t off = 0; off < b_bytes; off += 65536u) {
        const uint32_t len = min(65536u, b_bytes - off);
        bulk_g2s(Bs + off, a.B + off, len, bready);
      }
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const uint8_t* src = a.A + (size_t)tile * KC * CHUNK_BYTES;
        for (int ks = 0; ks < k_stages; ++ks) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], STAGE_BYTES);
          bulk_g2s(As + stage * STAGE_BYTES, src + (size_t)ks * STAGE_BYTES, STAGE_BYTES, &full[stage]);
          if (++stage == N_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(TILE_M, N, false, false);
      mbar_wait(bready, 0);
      tc_fence_after();
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      const uint32_t b_lbo = (uint32_t)N * 16;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_addr = tmem_base + (uint32_t)acc * 256u;
        for (int ks = 0; ks < k_stages; ++ks) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(As + stage * STAGE_BYTES);
#pragma unroll
          for (int j = 0; j < STAGE_CHUNKS / 2; ++j) {
            const uint64_t adesc = make_desc(a_base + j * 2 * CHUNK_BYTES, CHUNK_BYTES, 128);
            const uint64_t bdesc = make_desc(smem_u32(Bs) + (uint32_t)(ks * STAGE_CHUNKS + j * 2) * b_lbo, b_lbo, 128);
            umma_bf16(d_addr, adesc, bdesc, idesc, (ks | j) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == N_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // epilogue: thread = one vertex row of the tile; TMEM lane quadrant = warp % 4
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int ncb = N / 32;                              // 32-column blocks (<= 8)
    int acc = 0; uint32_t acc_phase = 0;
    const float scale = (MODE == MODE_FINAL && a.scale_dev) ? *a.scale_dev : a.scale;
    float* stg = stage_f + (size_t)q * 32 * 33;           // MODE_FINAL: per-warp 32 x 33 transpose buffer
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      const long long row = (long long)tile * TILE_M + r;
      uint32_t mw[8];
      if (MODE == MODE_DX) {                              // ReLU mask of this row, fetched before the MMAs finish
#pragma unroll
        for (int w = 0; w < 8; ++w) mw[w] = (w < ncb) ? __ldg(a.mask_in + (size_t)row * ncb + w) : 0u;
      }
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 256u;
      const size_t tile_off = (size_t)tile * (N / 8) * CHUNK_BYTES + (size_t)r * 16;
      const int cb_end = (MODE == MODE_FINAL) ? ncb : 8;     // FINAL: runtime bound, not unrolled (register budget)
#pragma unroll
      for (int cb = 0; cb < cb_end; ++cb) {
        if (cb < ncb) {
          float ub[32];
          if (MODE == MODE_FINAL) {
            // U_base rows of this warp's 32 x 32 block, fetched (coalesced, all in flight) before the accumulator is read
            const int col = cb * 32 + lane;
            const long long row0 = (long long)tile * TILE_M + q * 32;
#pragma unroll
            for (int rr = 0; rr < 32; ++rr)
              ub[rr] = (a.U_base && col < a.n_out && row0 + rr < a.n_rows)
                           ? __ldg(a.U_base + (size_t)(row0 + rr) * a.ldu + col) : 0.f;
          }
          uint32_t v[32];
          tmem_ld32(t_addr + cb * 32, v);
          if (MODE == MODE_FINAL) {
            // transpose through shared memory so that every store instruction writes 128 contiguous bytes
#pragma unroll
            for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = __uint_as_float(v[j]) + bias_s[cb * 32 + j];
            __syncwarp();
            const int col = cb * 32 + lane;
            const long long row0 = (long long)tile * TILE_M + q * 32;
            if (col < a.n_out) {
#pragma unroll
              for (int rr = 0; rr < 32; ++rr) {
                const long long grow = row0 + rr;
                if (grow < a.n_rows) {
                  const float c = stg[rr * 33 + lane];
                  a.corr[(size_t)grow * a.ldc + col] = c;
                  if (a.U_pred) a.U_pred[(size_t)grow * a.ldu + col] = __fadd_rn(ub[rr], __fmul_rn(scale, c));
                }
              }
            }
            __syncwarp();
          } else {
            uint32_t bits = 0;
#pragma unroll
            for (int g = 0; g < 4; ++g) {                     // 4 chunks of 8 columns
              const int c = cb * 4 + g;
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[g * 8 + j]);
              if (MODE == MODE_HIDDEN) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  f[j] += bias_s[c * 8 + j];
                  if (a.relu) f[j] = fmaxf(f[j], 0.f);
                  bits |= (f[j] > 0.f ? 1u : 0u) << (g * 8 + j);
                }
              } else {                                        // MODE_DX: keep where the forward activation was > 0
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = ((mw[cb] >> (g * 8 + j)) & 1u) ? f[j] : 0.f;
              }
              uint4 o;
              o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
              o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
              *reinterpret_cast<uint4*>(a.out_packed + tile_off + (size_t)c * CHUNK_BYTES) = o;
            }
            if (MODE == MODE_HIDDEN && a.mask_out) a.mask_out[(size_t)row * ncb + cb] = bits;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ------------------------------------------------------------------------------------------- dW
constexpr int DW_THREADS = 192;
constexpr int DW_XSTAGES = 3;          // half tiles of X: 16 chunks = 32 KB
constexpr int DW_YSTAGES = 2;          // whole Y tiles: NCy chunks
constexpr int XHALF_CHUNKS = 16;
constexpr int XHALF_BYTES = XHALF_CHUNKS * CHUNK_BYTES;

struct DwArgs {
  const uint8_t* X;   // packed, Mdim features (128 or 256)
  const uint8_t* Y;   // packed, Ndim features (multiple of 32, <= 256)
  int n_tiles, Mdim, Ndim;
  int db_from_y;      // 0: column sums of X, 1: column sums of Y
  float* partial;     // [grid][Mdim * Ndim]
  float* db_partial;  // [grid][256]
};

__global__ void __launch_bounds__(DW_THREADS, 1) tc_dw_kernel(DwArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int NCy = a.Ndim / 8;
  const int y_bytes = NCy * CHUNK_BYTES;
  const int m_halves = a.Mdim / 128;
  uint8_t* Xs = smem;                                         // DW_XSTAGES * XHALF_BYTES
  uint8_t* Ys = Xs + DW_XSTAGES * XHALF_BYTES;                // DW_YSTAGES * y_bytes
  uint64_t* bars = reinterpret_cast<uint64_t*>(Ys + (size_t)DW_YSTAGES * y_bytes);
  uint64_t* xfull = bars;                         // [3]
  uint64_t* xempty = bars + DW_XSTAGES;           // [3]
  uint64_t* yfull = xempty + DW_XSTAGES;          // [2]
  uint64_t* yempty = yfull + DW_YSTAGES;          // [2]
  uint64_t* done = yempty + DW_YSTAGES;           // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < DW_XSTAGES; ++s) { mbar_init(&xfull[s], 1); mbar_init(&xempty[s], a.db_from_y ? 1 : 129); }
    for (int s = 0; s < DW_YSTAGES; ++s) { mbar_init(&yfull[s], 1); mbar_init(&yempty[s], a.db_from_y ? 129 : 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int xs = 0, ys = 0; uint32_t xph = 0, yph = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        mbar_wait(&yempty[ys], yph ^ 1);
        mbar_expect_tx(&yfull[ys], (uint32_t)y_bytes);
        bulk_g2s(Ys + (size_t)ys * y_bytes, a.Y + (size_t)tile * y_bytes, (uint32_t)y_bytes, &yfull[ys]);
        if (++ys == DW_YSTAGES) { ys = 0; yph ^= 1; }
        for (int mh = 0; mh < m_halves; ++mh) {
          mbar_wait(&xempty[xs], xph ^ 1);
          mbar_expect_tx(&xfull[xs], XHALF_BYTES);
          bulk_g2s(Xs + xs * XHALF_BYTES, a.X + ((size_t)tile * m_halves + mh) * XHALF_BYTES, XHALF_BYTES, &xfull[xs]);
          if (++xs == DW_XSTAGES) { xs = 0; xph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, a.Ndim, true, true);
      int xs = 0, ys = 0; uint32_t xph = 0, yph = 0;
      bool first = true;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        mbar_wait(&yfull[ys], yph);
        tc_fence_after();
        const uint32_t y_base = smem_u32(Ys + (size_t)ys * y_bytes);
        for (int mh = 0; mh < m_halves; ++mh) {
          mbar_wait(&xfull[xs], xph);
          tc_fence_after();
          const uint32_t x_base = smem_u32(Xs + xs * XHALF_BYTES);
#pragma unroll
          for (int s = 0; s < TILE_M / 16; ++s) {              // 16 vertices (two 8-row K groups) per MMA
            const uint64_t adesc = make_desc(x_base + s * 256, 128, CHUNK_BYTES);
            const uint64_t bdesc = make_desc(y_base + s * 256, 128, CHUNK_BYTES);
            umma_bf16(tmem_base + (uint32_t)mh * 256u, adesc, bdesc, idesc, (!first || s > 0) ? 1u : 0u);
          }
          umma_commit(&xempty[xs]);
          if (++xs == DW_XSTAGES) { xs = 0; xph ^= 1; }
        }
        umma_commit(&yempty[ys]);
        if (++ys == DW_YSTAGES) { ys = 0; yph ^= 1; }
        first = false;
      }
      umma_commit(done);
    }
  } else {
    // ---- db: column sums of the dZ operand straight from shared memory (conflict-free 16 B reads)
    const int we = warp - 2;                     // 0..3
    float acc[2][4][8];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[h][c][j] = 0.f;
    int xs = 0, ys = 0; uint32_t xph = 0, yph = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      if (a.db_from_y) {
        mbar_wait(&yfull[ys], yph);
        const uint8_t* yb = Ys + (size_t)ys * y_bytes;
#pragma unroll
        for (int slot = 0; slot < 8; ++slot) {          // chunk = we + 4 * slot (NCy <= 32)
          const int c = we + 4 * slot;
          if (c < NCy) {
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
              const uint4 w = *reinterpret_cast<const uint4*>(yb + (size_t)c * CHUNK_BYTES + (rr * 32 + lane) * 16);
              const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
              for (int j = 0; j < 8; ++j)
                acc[slot >> 2][slot & 3][j] += __uint_as_float(((ww[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu) << 16);
            }
          }
        }
        mbar_arrive(&yempty[ys]);
        if (++ys == DW_YSTAGES) { ys = 0; yph ^= 1; }
      } else {
        for (int mh = 0; mh < m_halves; ++mh) {
          mbar_wait(&xfull[xs], xph);
          const uint8_t* xb = Xs + xs * XHALF_BYTES;
#pragma unroll
          for (int cl = 0; cl < 4; ++cl) {
            const int c = we * 4 + cl;
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
              const uint4 w = *reinterpret_cast<const uint4*>(xb + (size_t)c * CHUNK_BYTES + (rr * 32 + lane) * 16);
              const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float val = __uint_as_float(((ww[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu) << 16);
                if (mh == 0) acc[0][cl][j] += val; else acc[1][cl][j] += val;
              }
            }
          }
          mbar_arrive(&xempty[xs]);
          if (++xs == DW_XSTAGES) { xs = 0; xph ^= 1; }
        }
      }
    }
    float* dbp = a.db_partial + (size_t)blockIdx.x * 256;
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float s = ep::warp_sum(acc[h][c][j]);
          if (lane == 0) {
            // feature index of this accumulator
            const int feat = a.db_from_y ? ((we + 4 * (h * 4 + c)) * 8 + j) : (h * 128 + (we * 4 + c) * 8 + j);
            if (feat < 256) dbp[feat] = s;
          }
        }
    // ---- final: TMEM -> per-CTA fp32 partial
    mbar_wait(done, 0);
    tc_fence_after();
    const int q = warp & 3;
    float* part = a.partial + (size_t)blockIdx.x * a.Mdim * a.Ndim;
    for (int mh = 0; mh < m_halves; ++mh) {
      const int m = mh * 128 + q * 32 + lane;
      for (int cb = 0; cb < a.Ndim / 32; ++cb) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)mh * 256u + cb * 32, v);
        float4* dst = reinterpret_cast<float4*>(part + (size_t)m * a.Ndim + cb * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                               __uint_as_float(v[4 * j + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// out[o][i] = sum_cta partial[cta][m][n], (m, n) = (o, i) or (i, o) when transposed.  64 outputs per block,
// four thread groups split the CTA partials and are combined in a fixed order (deterministic).
__global__ void __launch_bounds__(256)
dw_reduce_kernel(int n_cta, int Mdim, int Ndim, const float* __restrict__ partial, int out_rows, int out_cols,
                 int transposed, float* __restrict__ dW, const float* __restrict__ db_partial, int db_len,
                 float* __restrict__ db) {
  __shared__ float sh[4][64];
  const int total = out_rows * out_cols;
  const int el = threadIdx.x & 63, g = threadIdx.x >> 6;
  const int e = blockIdx.x * 64 + el;
  float s = 0.f;
  if (e < total) {
    const int o = e / out_cols, i = e - o * out_cols;
    const size_t idx = transposed ? (size_t)i * Ndim + o : (size_t)o * Ndim + i;
    const size_t stride = (size_t)Mdim * Ndim;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int c = g;
    for (; c + 12 < n_cta; c += 16) {
      s0 += partial[(size_t)c * stride + idx];
      s1 += partial[(size_t)(c + 4) * stride + idx];
      s2 += partial[(size_t)(c + 8) * stride + idx];
      s3 += partial[(size_t)(c + 12) * stride + idx];
    }
    for (; c < n_cta; c += 4) s0 += partial[(size_t)c * stride + idx];
    s = (s0 + s1) + (s2 + s3);
  }
  sh[g][el] = s;
  __syncthreads();
  if (g == 0 && e < total) dW[e] = (sh[0][el] + sh[1][el]) + (sh[2][el] + sh[3][el]);
  if (blockIdx.x == 0) {
    for (int f = threadIdx.x; f < db_len; f += blockDim.x) {
      float t = 0.f;
      for (int c = 0; c < n_cta; ++c) t += db_partial[(size_t)c * 256 + f];
      db[f] = t;
    }
  }
}

// ------------------------------------------------------------------------------------------- packing
// fp32 rows [n x d] -> packed bf16 tiles [tile][dp/8][128][8], zero padded (rows >= n, cols >= d)
__global__ void __launch_bounds__(256)
pack_rows_kernel(int n, int d, int dp, const float* __restrict__ X, int ldx, uint8_t* __restrict__ out, int n_tiles) {
  // item = one 16-byte chunk.  A warp covers 8 rows x 4 chunks: reads are 128 contiguous bytes per row,
  // writes 128 contiguous bytes per chunk column (8 rows x 16 B).
  const int nc = dp / 8, ncg = nc / 4;                                // dp is a multiple of 32
  const long long total = (long long)n_tiles * nc * TILE_M;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < total;
       it += (long long)gridDim.x * blockDim.x) {
    const int lane = (int)(it & 31);
    const long long w = it >> 5;
    const int rg = (int)(w & 15);
    const long long w2 = w >> 4;
    const int cg = (int)(w2 % ncg);
    const long long tile = w2 / ncg;
    const int r = rg * 8 + (lane >> 2);
    const int c = cg * 4 + (lane & 3);
    const long long row = tile * TILE_M + r;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = c * 8 + j;
      f[j] = (row < n && col < d) ? __ldg(X + (size_t)row * ldx + col) : 0.f;
    }
    uint4 o;
    o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
    o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(out + (((size_t)tile * nc + c) * TILE_M + r) * 16) = o;
  }
}

// W fp32 [out x in] -> Wp [inp/8][outp][8] (B operand of the forward GEMM) and WTp [outp/8][inp][8] (dX GEMM)
__global__ void __launch_bounds__(256)
pack_weight_kernel(int out, int in, int outp, int inp, const float* __restrict__ W, __nv_bfloat16* __restrict__ Wp,
                   __nv_bfloat16* __restrict__ WTp) {
  const int total = outp * inp;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int o = e / inp, i = e - o * inp;
    const float w = (o < out && i < in) ? __ldg(W + (size_t)o * in + i) : 0.f;
    const __nv_bfloat16 b = __float2bfloat16_rn(w);
    Wp[(size_t)(i >> 3) * outp * 8 + (size_t)o * 8 + (i & 7)] = b;
    if (WTp) WTp[(size_t)(o >> 3) * inp * 8 + (size_t)i * 8 + (o & 7)] = b;
  }
}

inline int pad_to(int v, int m) { return (v + m - 1) / m * m; }
inline int n_tiles_for(int n) { return (n + TILE_M - 1) / TILE_M; }

}  // namespace tc

using namespace tc;

namespace {

template <int MODE>
int launch_linear(const LinearArgs& a, cudaStream_t st, int max_ctas = 0) {
  const size_t smem = (size_t)a.KC * a.N * 16 + (size_t)N_STAGES * STAGE_BYTES + 8 * (2 * N_STAGES + 5) + 16 +
                      sizeof(float) * a.N + (MODE == MODE_FINAL ? sizeof(float) * 4 * 32 * 33 : 0) + 128;
  static size_t configured = 0;
  if (smem > configured) {
    EP_CUDA_CHECK(cudaFuncSetAttribute(tc_linear_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  int grid = ep::sm_count();
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  if (grid > a.n_tiles) grid = a.n_tiles;
  tc_linear_kernel<MODE><<<grid, LINEAR_THREADS, smem, st>>>(a);
  EP_LAUNCH_CHECK("tc_linear_kernel");
  return EP_OK;
}

bool dims_ok(int kp, int np) { return kp % 32 == 0 && np % 32 == 0 && kp >= 32 && np >= 32 && kp <= 256 && np <= 256; }

}  // namespace

extern "C" {

int ep_tc_pad_features(int d, int wide) { return wide ? pad_to(d, 128) : pad_to(d, 32); }

size_t ep_tc_packed_rows_bytes(int n, int d_padded) {
  if (n <= 0 || d_padded <= 0) return 0;
  return (size_t)n_tiles_for(n) * TILE_M * d_padded * 2;
}

size_t ep_tc_packed_weight_bytes(int out_padded, int in_padded) { return (size_t)out_padded * in_padded * 2; }

size_t ep_tc_relu_mask_bytes(int n, int d_padded) {
  if (n <= 0 || d_padded <= 0) return 0;
  return (size_t)n_tiles_for(n) * TILE_M * (d_padded / 32) * 4;
}

int ep_tc_pack_rows_bf16(int n, int d, int d_padded, const float* X, int ldx, void* packed, ep_stream_t stream) {
  EP_REQUIRE(n > 0 && d > 0 && d_padded >= d && d_padded % 32 == 0, "bad size (d_padded must be a multiple of 32)");
  EP_REQUIRE(X && packed && ldx >= d, "bad argument");
  const int nt = n_tiles_for(n);
  const long long total = (long long)nt * (d_padded / 8) * TILE_M;
  long long grid = (total + 255) / 256;
  if (grid > (long long)ep::sm_count() * 16) grid = (long long)ep::sm_count() * 16;
  pack_rows_kernel<<<(unsigned)grid, 256, 0, ep::as_stream(stream)>>>(n, d, d_padded, X, ldx, static_cast<uint8_t*>(packed), nt);
  EP_LAUNCH_CHECK("pack_rows_kernel");
  return EP_OK;
}

int ep_tc_pack_weight_bf16(int out, int in, int out_padded, int in_padded, const float* W, void* Wp, void* WTp,
                           ep_stream_t stream) {
  EP_REQUIRE(out > 0 && in > 0 && out_padded >= out && in_padded >= in, "bad size");
  EP_REQUIRE(out_padded % 8 == 0 && in_padded % 8 == 0 && W && Wp, "bad argument");
  const int total = out_padded * in_padded;
  pack_weight_kernel<<<ep::ceil_div(total, 256), 256, 0, ep::as_stream(stream)>>>(
      out, in, out_padded, in_padded, W, static_cast<__nv_bfloat16*>(Wp), static_cast<__nv_bfloat16*>(WTp));
  EP_LAUNCH_CHECK("pack_weight_kernel");
  return EP_OK;
}

int ep_tc_linear_fwd_bf16(int n, int in_padded, int out, int out_padded, const void* A_packed, const void* Wp,
                          const float* bias, int relu, void* out_packed, void* relu_mask_out, ep_stream_t stream) {
  EP_REQUIRE(n > 0 && A_packed && Wp && out_packed, "bad argument");
  if (!dims_ok(in_padded, out_padded)) { ep::set_error("ep_tc_linear_fwd_bf16: padded dims must be multiples of 32 in [32, 256]"); return EP_ERR_UNSUPPORTED; }
  LinearArgs a{};
  a.A = static_cast<const uint8_t*>(A_packed); a.B = static_cast<const uint8_t*>(Wp);
  a.n_tiles = n_tiles_for(n); a.KC = in_padded / 8; a.N = out_padded; a.bias = bias; a.n_bias = out; a.relu = relu;
  a.out_packed = static_cast<uint8_t*>(out_packed); a.mask_out = static_cast<uint32_t*>(relu_mask_out);
  a.n_rows = n; a.n_out = out_padded;
  return launch_linear<MODE_HIDDEN>(a, ep::as_stream(stream));
}

int ep_tc_linear_final_bf16(int n, int in_padded, int out, int out_padded, const void* A_packed, const void* Wp,
                            const float* bias, float* corr, int ldc, const float* U_base, float scale,
                            const float* scale_dev, float* U_pred, int ldu, ep_stream_t stream) {
  EP_REQUIRE(n > 0 && A_packed && Wp && corr && out > 0 && out <= out_padded && ldc >= out, "bad argument");
  EP_REQUIRE((U_base == nullptr) == (U_pred == nullptr) && (!U_pred || ldu >= out), "U_base / U_pred mismatch");
  if (!dims_ok(in_padded, out_padded)) { ep::set_error("ep_tc_linear_final_bf16: unsupported dims"); return EP_ERR_UNSUPPORTED; }
  LinearArgs a{};
  a.A = static_cast<const uint8_t*>(A_packed); a.B = static_cast<const uint8_t*>(Wp);
  a.n_tiles = n_tiles_for(n); a.KC = in_padded / 8; a.N = out_padded; a.bias = bias; a.n_bias = out;
  a.corr = corr; a.ldc = ldc; a.U_base = U_base; a.U_pred = U_pred; a.ldu = ldu; a.scale = scale; a.scale_dev = scale_dev;
  a.n_rows = n; a.n_out = out;
  return launch_linear<MODE_FINAL>(a, ep::as_stream(stream));
}

int ep_tc_linear_dx_bf16(int n, int out_padded, int in_padded, const void* dZ_packed, const void* WTp,
                         const void* relu_mask, void* dZprev_packed, int max_ctas, ep_stream_t stream) {
  EP_REQUIRE(n > 0 && dZ_packed && WTp && relu_mask && dZprev_packed, "bad argument");
  if (!dims_ok(out_padded, in_padded)) { ep::set_error("ep_tc_linear_dx_bf16: unsupported dims"); return EP_ERR_UNSUPPORTED; }
  LinearArgs a{};
  a.A = static_cast<const uint8_t*>(dZ_packed); a.B = static_cast<const uint8_t*>(WTp);
  a.n_tiles = n_tiles_for(n); a.KC = out_padded / 8; a.N = in_padded;
  a.out_packed = static_cast<uint8_t*>(dZprev_packed); a.mask_in = static_cast<const uint32_t*>(relu_mask);
  a.n_rows = n; a.n_out = in_padded;
  return launch_linear<MODE_DX>(a, ep::as_stream(stream), max_ctas);
}

size_t ep_tc_dw_workspace_bytes(void) { return sizeof(float) * (size_t)ep::sm_count() * (256 * 256 + 256); }

int ep_tc_linear_dw_bf16(int n, int out, int in, int out_padded, int in_padded, const void* dZ_packed,
                         const void* act_packed, float* dW, float* db, void* workspace, size_t workspace_bytes,
                         int max_ctas, ep_stream_t stream) {
  EP_REQUIRE(n > 0 && dZ_packed && act_packed && dW && db && workspace, "bad argument");
  EP_REQUIRE(out <= out_padded && in <= in_padded, "bad padding");
  if (workspace_bytes < ep_tc_dw_workspace_bytes()) { ep::set_error("ep_tc_linear_dw_bf16: workspace too small"); return EP_ERR_WORKSPACE; }
  DwArgs a{};
  int transposed;
  if (out_padded % 128 == 0 && out_padded <= 256 && in_padded % 32 == 0 && in_padded <= 256) {
    a.X = static_cast<const uint8_t*>(dZ_packed); a.Y = static_cast<const uint8_t*>(act_packed);
    a.Mdim = out_padded; a.Ndim = in_padded; a.db_from_y = 0; transposed = 0;
  } else if (in_padded % 128 == 0 && in_padded <= 256 && out_padded % 32 == 0 && out_padded <= 256) {
    a.X = static_cast<const uint8_t*>(act_packed); a.Y = static_cast<const uint8_t*>(dZ_packed);
    a.Mdim = in_padded; a.Ndim = out_padded; a.db_from_y = 1; transposed = 1;
  } else {
    ep::set_error("ep_tc_linear_dw_bf16: one padded dimension must be 128 or 256, the other a multiple of 32 <= 256");
    return EP_ERR_UNSUPPORTED;
  }
  a.n_tiles = n_tiles_for(n);
  int grid = ep::sm_count();
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  if (grid > a.n_tiles) grid = a.n_tiles;
  a.partial = static_cast<float*>(workspace);
  a.db_partial = a.partial + (size_t)ep::sm_count() * 256 * 256;
  const size_t smem = (size_t)DW_XSTAGES * XHALF_BYTES + (size_t)DW_YSTAGES * (a.Ndim / 8) * CHUNK_BYTES +
                      8 * (2 * DW_XSTAGES + 2 * DW_YSTAGES + 1) + 16 + 128;
  static size_t configured = 0;
  if (smem > configured) {
    EP_CUDA_CHECK(cudaFuncSetAttribute(tc_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  cudaStream_t st = ep::as_stream(stream);
  tc_dw_kernel<<<grid, DW_THREADS, smem, st>>>(a);
  EP_LAUNCH_CHECK("tc_dw_kernel");
  const int total = out * in;
  dw_reduce_kernel<<<ep::ceil_div(total, 64), 256, 0, st>>>(
      grid, a.Mdim, a.Ndim, a.partial, out, in, transposed, dW, a.db_partial, out, db);
  EP_LAUNCH_CHECK("dw_reduce_kernel");
  return EP_OK;
}

}  // extern "C"
