// Persistent warp-specialised tcgen05 GEMM for sm_100a:  out = epilogue(A[M,K] * W[N,K]^T).
//
//   warp 0      : TMA producer (one elected lane) -- A and W tiles, 128B-swizzled, mbarrier ring
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (128 x BN x 16 per instruction)
//   warps 2..9  : epilogue -- tcgen05.ld the fp32 accumulator (one row per thread, half the columns per warp), bias / RoPE / GELU /
//                 residual / image2-embed row bias, vectorised global stores
// The accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the MMAs of
// tile i+1.  Tiles are walked m-fastest so that co-resident CTAs share the weight tile in L2.
#include "ptx.cuh"
#include "m3r_internal.h"

namespace m3r {

constexpr int BM = 128;
constexpr int BK = 64;        // 64 x 16-bit = 128 B = one swizzle atom
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 320;      // TMA warp, MMA warp, 8 epilogue warps (2 per TMEM lane quarter, half the columns each)
constexpr int SMEM_BUDGET = 196608;  // bytes of operand ring

template <int BN>
struct GemmCfg {
  static constexpr int STAGE_BYTES = (BM + BN) * BK * 2;
  static constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES;   // 256:4  128:6  64:8
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

struct GemmParams {
  int M, N, K;
  int is_bf16;
  const float* bias;
  int act;
  const float* residual;
  long long ldr;
  const float* rowbias;
  int rb_period, rb_first;
  const float* rope_tab;
  int rope_cols, rope_period;
  void* out;
  long long ldc;
  int out_dtype;
  int rows_per_batch;
  long long batch_stride_rows;
  int n_peer_out;
  void* peer_out[M3R_MAX_PEERS];
  unsigned long long* trace;   // debug only (m3r_debug_trace): 16 words per CTA; normally null
  int w_static;                // weights may be requested before the programmatic-dependency wait
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// Epilogue of one 32-column chunk of one accumulator row: bias, RoPE, row bias, GELU, residual, store.
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, const uint32_t (&raw)[32], int row, bool row_ok, long long orow,
                                               bool rb_on, const float* rope_row, int col0) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
  if (p.bias) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i));
      v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
    }
  }
  if (rope_row != nullptr && col0 < p.rope_cols && row_ok) {
    // chunk = half a head: even chunks are the y half (u: 0..15, v: 16..31), odd chunks the x half
    const float* cs = rope_row + ((col0 >> 5) & 1) * 32;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 c4 = __ldg(reinterpret_cast<const float4*>(cs + i));
      const float4 s4 = __ldg(reinterpret_cast<const float4*>(cs + 16 + i));
      const float cc[4] = {c4.x, c4.y, c4.z, c4.w};
      const float ss[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float u = v[i + j], w = v[i + j + 16];
        v[i + j] = u * cc[j] - w * ss[j];
        v[i + j + 16] = w * cc[j] + u * ss[j];
      }
    }
  }
  if (rb_on) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.rowbias + col0 + i));
      v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
    }
  }
  if (p.act == 1) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
  }
  if (row_ok) {
    if (p.residual) {
      const float4* r4 = reinterpret_cast<const float4*>(p.residual + (long long)row * p.ldr + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 r = r4[i];
        v[4 * i] += r.x; v[4 * i + 1] += r.y; v[4 * i + 2] += r.z; v[4 * i + 3] += r.w;
      }
    }
    if (p.out_dtype == 0) {
      float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + orow * p.ldc + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i) o4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
      uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out) + orow * p.ldc + col0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 w;
        w.x = pack16(v[8 * i], v[8 * i + 1], p.is_bf16);
        w.y = pack16(v[8 * i + 2], v[8 * i + 3], p.is_bf16);
        w.z = pack16(v[8 * i + 4], v[8 * i + 5], p.is_bf16);
        w.w = pack16(v[8 * i + 6], v[8 * i + 7], p.is_bf16);
        o4[i] = w;
        // fused all-gather: the same 16 bytes go to the other GPUs' copies of the buffer (NVLink peer stores)
        for (int pr = 0; pr < p.n_peer_out; ++pr)
          reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.peer_out[pr]) + orow * p.ldc + col0)[i] = w;
      }
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                               // STAGES x [128 x 64] 16-bit
  uint8_t* sB = smem + STAGES * BM * BK * 2;        // STAGES x [BN x 64] 16-bit
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;                  // [STAGES]
  uint64_t* empty = bars + STAGES;        // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES;    // [2]
  uint64_t* tempty = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  M3R_TR(unsigned long long* tr = p.trace ? p.trace + 16ull * blockIdx.x : nullptr;
         if (tr && threadIdx.x == 0) { tr[0] = gtime_ns(); tr[8] = smid(); })
  const int warp = threadIdx.x >> 5;
  const int tiles_m = (p.M + BM - 1) / BM;
  const int tiles_n = p.N / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = p.K / BK;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 256); }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  M3R_TR(if (tr && threadIdx.x == 0) tr[1] = gtime_ns();)

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      // The weights never depend on the kernel before this one, so the W halves of the first ring of stages are requested
      // BEFORE the programmatic-dependency wait: their (cold, HBM) latency overlaps the predecessor's tail.  The
      // activations follow after the wait; each stage's barrier expects both halves.
      int pre = 0;
      if (p.w_static && blockIdx.x < num_tiles) {
        const int n0 = (blockIdx.x / tiles_m) * BN;
        pre = num_kb < STAGES ? num_kb : STAGES;
        for (int kb = 0; kb < pre; ++kb) {
          mbar_arrive_expect_tx(&full[kb], Cfg::STAGE_BYTES);
          tma_load_2d(sB + kb * BN * BK * 2, &tmW, &full[kb], kb * BK, n0);
        }
      }
      griddep_wait();
      griddep_launch();
      M3R_TR(if (tr) tr[2] = gtime_ns();)
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m0 = (t % tiles_m) * BM;
        const int n0 = (t / tiles_m) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          if (pre > 0) {                      // stage already armed, W half in flight
            --pre;
            tma_load_2d(sA + stage * BM * BK * 2, &tmA, &full[stage], kb * BK, m0);
          } else {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES);
            tma_load_2d(sA + stage * BM * BK * 2, &tmA, &full[stage], kb * BK, m0);
            tma_load_2d(sB + stage * BN * BK * 2, &tmW, &full[stage], kb * BK, n0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      M3R_TR(if (tr) tr[3] = gtime_ns();)
    } else {
      griddep_wait();
      griddep_launch();
    }
  } else if (warp == 1) {
    griddep_wait();        // everything above overlapped the previous kernel's tail
    griddep_launch();
    // ------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc(BM, BN, p.is_bf16 ? 1u : 0u, 0, 0);
    int stage = 0; uint32_t phase = 0;
    int as = 0; uint32_t aphase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      mbar_wait(&tempty[as], aphase ^ 1);      // epilogue has drained this accumulator buffer
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        M3R_TR(if (tr && kb == 0 && t == (int)blockIdx.x && (threadIdx.x & 31) == 0) tr[4] = gtime_ns();)
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = smem_desc_sw128(smem_u32(sA + stage * BM * BK * 2));
          const uint64_t bdesc = smem_desc_sw128(smem_u32(sB + stage * BN * BK * 2));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 elements (32 bytes) along K inside the 128B swizzle atom: +2 in the >>4 address field
            umma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);                       // frees the smem stage when the MMAs retire
          if (kb == num_kb - 1) umma_commit(&tfull[as]);    // accumulator complete
          M3R_TR(if (tr && kb == num_kb - 1) tr[5] = gtime_ns();)
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (2..9)
    griddep_wait();
    griddep_launch();
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int chalf = (warp - 2) >> 2;            // which half of the tile's columns this warp drains
    const int lane = threadIdx.x & 31;
    int as = 0; uint32_t aphase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m0 = (t % tiles_m) * BM;
      const int n0 = (t / tiles_m) * BN;
      const int row = m0 + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      long long orow = row;
      if (p.rows_per_batch > 0) orow = (long long)(row / p.rows_per_batch) * p.batch_stride_rows + row % p.rows_per_batch;
      const bool rb_on = p.rowbias != nullptr && (row % p.rb_period) >= p.rb_first;
      const float* rope_row = p.rope_tab ? p.rope_tab + (long long)(row % p.rope_period) * 64 : nullptr;

      mbar_wait(&tfull[as], aphase);
      M3R_TR(if (tr && threadIdx.x == 64) tr[6] = gtime_ns();)
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t(quarter * 32) << 16) + as * BN;
#pragma unroll 1
      for (int c = chalf * (BN / 64); c < (chalf + 1) * (BN / 64); ++c) {
        uint32_t raw[32];
        tmem_ld32(t_addr + c * 32, raw);
        tmem_wait_ld();
        if (c == (chalf + 1) * (BN / 64) - 1) {  // this warp's share of the accumulator is read: hand it back early
          tc_fence_before();
          mbar_arrive(&tempty[as]);
        }
        epilogue_chunk(p, raw, row, row_ok, orow, rb_on, rope_row, n0 + c * 32);
      }
      M3R_TR(if (tr && threadIdx.x == 64) tr[7] = gtime_ns();)
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  M3R_TR(if (tr && threadIdx.x == 0) tr[9] = gtime_ns();)
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN>
static int launch_gemm(const m3r_gemm_args* a, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap tmA, tmW;
  if (make_tmap_2d(&tmA, a->A, a->is_bf16, (uint64_t)a->K, (uint64_t)a->M, (uint64_t)a->lda, BK, BM)) return 1;
  if (make_tmap_2d(&tmW, a->W, a->is_bf16, (uint64_t)a->K, (uint64_t)a->N, (uint64_t)a->ldw, BK, BN)) return 1;
  GemmParams p;
  p.M = a->M; p.N = a->N; p.K = a->K; p.is_bf16 = a->is_bf16;
  p.bias = a->bias; p.act = a->act; p.residual = a->residual; p.ldr = a->ldr;
  p.rowbias = a->rowbias; p.rb_period = a->rb_period > 0 ? a->rb_period : 1; p.rb_first = a->rb_first;
  p.rope_tab = a->rope_tab; p.rope_cols = a->rope_cols; p.rope_period = a->rope_period > 0 ? a->rope_period : 1;
  p.out = a->out; p.ldc = a->ldc; p.out_dtype = a->out_dtype;
  p.rows_per_batch = a->rows_per_batch; p.batch_stride_rows = a->batch_stride_rows;
  p.n_peer_out = a->n_peer_out;
  p.trace = trace_buffer();
  p.w_static = a->w_static;
  for (int i = 0; i < M3R_MAX_PEERS; ++i) p.peer_out[i] = i < a->n_peer_out ? a->peer_out[i] : nullptr;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error("gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int tiles = ((a->M + BM - 1) / BM) * (a->N / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  {
    ProfScope prof(BN == 256 ? PROF_GEMM256 : (BN == 128 ? PROF_GEMM128 : PROF_GEMM64), 2.0 * a->M * (double)a->N * a->K, 2.0 * ((double)a->M * a->K + (double)a->N * a->K) + (double)a->M * a->N * (a->out_dtype ? 2 : 4), stream);
    cudaError_t le = launch_pdl(gemm_kernel<BN>, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, tmA, tmW, p);
    if (le != cudaSuccess) return set_error("gemm launch: %s", cudaGetErrorString(le));
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("gemm launch: %s", cudaGetErrorString(e));
  count_launch();
  return 0;
}


// ------------------------------------------------------------------------------------------------ CTA-pair GEMM
// Two CTAs of a cluster (the two SMs of a TPC) compute one 256 x 256 output tile with tcgen05.mma.cta_group::2: each
// CTA stages its own 128 rows of A and HALF of the weight tile (128 of the 256 W rows), so per CTA a k-block costs
// 32 KB of L2->smem traffic instead of 48 KB and the 192 KB ring holds 6 stages instead of 4 (the 1-CTA kernel's MMA
// warp spends half its time waiting for operands, profiles/r01_ncu_gemm_v1_summary.txt).  The leader CTA issues the
// MMAs; completion is multicast to both CTAs' barriers; each CTA's epilogue drains its own 128 accumulator rows.
constexpr int P_STAGES = 6;
constexpr int P_STAGE_BYTES = (128 + 128) * BK * 2;          // A rows + half of W per CTA
constexpr int P_SMEM_BYTES = P_STAGES * P_STAGE_BYTES + 1024 + 256;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                                  // P_STAGES x [128 x 64]
  uint8_t* sB = smem + P_STAGES * 128 * BK * 2;        // P_STAGES x [128 x 64] (this CTA's half of the W tile)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P_STAGES * P_STAGE_BYTES);
  uint64_t* full = bars;                       // [P_STAGES] used in the leader only
  uint64_t* empty = bars + P_STAGES;           // [P_STAGES] both CTAs (multicast commit)
  uint64_t* tfull = bars + 2 * P_STAGES;       // [2] both CTAs (multicast commit)
  uint64_t* tempty = bars + 2 * P_STAGES + 2;  // [2] leader only: 512 arrivals (both CTAs' epilogue threads)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * P_STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int tiles_m = (p.M + 255) / 256;
  const int tiles_n = p.N / 256;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = p.K / BK;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < P_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 512); }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc2(tmem_slot, 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();            // barriers of both CTAs initialised before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();
  griddep_launch();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int t = pair; t < num_tiles; t += n_pairs) {
        const int m0 = (t % tiles_m) * 256 + rank * 128;
        const int n0 = (t / tiles_m) * 256 + rank * 128;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (leader) mbar_arrive_expect_tx(&full[stage], 2 * P_STAGE_BYTES);       // bytes of BOTH CTAs
          const uint32_t lbar = mapa_u32(&full[stage], 0);
          tma_load_2d_pair(sA + stage * 128 * BK * 2, &tmA, lbar, kb * BK, m0);
          tma_load_2d_pair(sB + stage * 128 * BK * 2, &tmW, lbar, kb * BK, n0);
          if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      const uint32_t idesc = make_idesc(256, 256, p.is_bf16 ? 1u : 0u, 0, 0);
      int stage = 0; uint32_t phase = 0;
      int as = 0; uint32_t aphase = 0;
      for (int t = pair; t < num_tiles; t += n_pairs) {
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t adesc = smem_desc_sw128(smem_u32(sA + stage * 128 * BK * 2));
            const uint64_t bdesc = smem_desc_sw128(smem_u32(sB + stage * 128 * BK * 2));
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) umma_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
            umma_commit_pair(&empty[stage]);
            if (kb == num_kb - 1) umma_commit_pair(&tfull[as]);
          }
          __syncwarp();
          if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
        }
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (both CTAs, own 128 rows)
    const int quarter = warp & 3;
    const int chalf = (warp - 2) >> 2;
    const int lane = threadIdx.x & 31;
    int as = 0; uint32_t aphase = 0;
    const uint32_t tempty_leader[2] = {mapa_u32(&tempty[0], 0), mapa_u32(&tempty[1], 0)};
    for (int t = pair; t < num_tiles; t += n_pairs) {
      const int m0 = (t % tiles_m) * 256 + rank * 128;
      const int n0 = (t / tiles_m) * 256;
      const int row = m0 + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      long long orow = row;
      if (p.rows_per_batch > 0) orow = (long long)(row / p.rows_per_batch) * p.batch_stride_rows + row % p.rows_per_batch;
      const bool rb_on = p.rowbias != nullptr && (row % p.rb_period) >= p.rb_first;
      const float* rope_row = p.rope_tab ? p.rope_tab + (long long)(row % p.rope_period) * 64 : nullptr;
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t(quarter * 32) << 16) + as * 256;
#pragma unroll 1
      for (int c = chalf * 4; c < chalf * 4 + 4; ++c) {
        uint32_t raw[32];
        tmem_ld32(t_addr + c * 32, raw);
        tmem_wait_ld();
        if (c == chalf * 4 + 3) {
          tc_fence_before();
          mbar_arrive_cluster(tempty_leader[as]);
        }
        epilogue_chunk(p, raw, row, row_ok, orow, rb_on, rope_row, n0 + c * 32);
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();            // no CTA leaves (or frees TMEM) while its peer may still address it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

static int launch_gemm_pair(const m3r_gemm_args* a, cudaStream_t stream) {
  CUtensorMap tmA, tmW;
  if (make_tmap_2d(&tmA, a->A, a->is_bf16, (uint64_t)a->K, (uint64_t)a->M, (uint64_t)a->lda, BK, 128)) return 1;
  if (make_tmap_2d(&tmW, a->W, a->is_bf16, (uint64_t)a->K, (uint64_t)a->N, (uint64_t)a->ldw, BK, 128)) return 1;
  GemmParams p;
  p.M = a->M; p.N = a->N; p.K = a->K; p.is_bf16 = a->is_bf16;
  p.bias = a->bias; p.act = a->act; p.residual = a->residual; p.ldr = a->ldr;
  p.rowbias = a->rowbias; p.rb_period = a->rb_period > 0 ? a->rb_period : 1; p.rb_first = a->rb_first;
  p.rope_tab = a->rope_tab; p.rope_cols = a->rope_cols; p.rope_period = a->rope_period > 0 ? a->rope_period : 1;
  p.out = a->out; p.ldc = a->ldc; p.out_dtype = a->out_dtype;
  p.rows_per_batch = a->rows_per_batch; p.batch_stride_rows = a->batch_stride_rows;
  p.n_peer_out = a->n_peer_out;
  p.trace = trace_buffer();
  p.w_static = a->w_static;
  for (int i = 0; i < M3R_MAX_PEERS; ++i) p.peer_out[i] = i < a->n_peer_out ? a->peer_out[i] : nullptr;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES);
    if (e != cudaSuccess) return set_error("gemm(pair): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int tiles = ((a->M + 255) / 256) * (a->N / 256);
  int pairs = num_sms() / 2;
  if (tiles < pairs) pairs = tiles;
  {
    ProfScope prof(PROF_GEMM256, 2.0 * a->M * (double)a->N * a->K, 2.0 * ((double)a->M * a->K + (double)a->N * a->K) + (double)a->M * a->N * (a->out_dtype ? 2 : 4), stream);
    cudaError_t le = launch_pdl(gemm_pair_kernel, dim3(2 * pairs), dim3(GEMM_THREADS), P_SMEM_BYTES, stream, tmA, tmW, p);
    if (le != cudaSuccess) return set_error("gemm(pair) launch: %s", cudaGetErrorString(le));
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("gemm(pair) launch: %s", cudaGetErrorString(e));
  count_launch();
  return 0;
}

}  // namespace m3r

extern "C" int m3r_gemm(const m3r_gemm_args* a, void* stream) {
  using namespace m3r;
  if (!a || !a->A || !a->W || !a->out) return set_error("gemm: null pointer");
  if (a->M <= 0) return 0;
  if (a->N % 64 || a->K % 64 || a->N <= 0 || a->K <= 0) return set_error("gemm: N (%d) and K (%d) must be multiples of 64", a->N, a->K);
  if (a->lda % 8 || a->ldw % 8) return set_error("gemm: lda/ldw must be multiples of 8 elements (16 B)");
  if ((a->out_dtype == M3R_OUT_F32 && a->ldc % 4) || (a->out_dtype == M3R_OUT_16 && a->ldc % 8)) return set_error("gemm: ldc alignment");
  if (a->residual && a->ldr % 4) return set_error("gemm: ldr alignment");
  if (a->rope_tab && (a->rope_cols % 64)) return set_error("gemm: rope_cols must be a multiple of 64");
  if (a->n_peer_out < 0 || a->n_peer_out > M3R_MAX_PEERS || (a->n_peer_out > 0 && a->out_dtype != M3R_OUT_16)) return set_error("gemm: peer outputs need 0..%d pointers and a 16-bit output", M3R_MAX_PEERS);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // Tile-width heuristic: widest BN that still yields about one wave of CTAs.
  const int tiles_m = (a->M + BM - 1) / BM;
  const int sms = num_sms();
  // (profiles/r01_small_gemm_tiles.txt: at M=768, BN=128 wins from ~100 tiles up, BN=64 below)
  int bn = 64;
  if (a->N % 256 == 0 && tiles_m * (a->N / 256) >= sms) bn = 256;
  else if (a->N % 128 == 0 && tiles_m * (a->N / 128) >= (sms * 2) / 3) bn = 128;
  {
    // CTA-pair kernel for problems with at least one 256x256 tile per SM pair.  Measured (profiles/r01_run13_gemm_pair.log):
    // +7..+20 % for K >= 1024, -1..-4 % for K = 768 (epilogue-bound tiles) -> used for K >= 1024; M3R_GEMM_PAIR=0/2
    // force it off / on for every eligible shape.
    static int pair_mode = -1;
    if (pair_mode < 0) { const char* e = getenv("M3R_GEMM_PAIR"); pair_mode = e ? atoi(e) : 1; }
    const bool eligible = a->N % 256 == 0 && ((a->M + 255) / 256) * (a->N / 256) >= sms / 2 && !getenv("M3R_GEMM_BN");
    if (eligible && (pair_mode == 2 || (pair_mode == 1 && a->K >= 1024))) return launch_gemm_pair(a, s);
  }
  const char* force = getenv("M3R_GEMM_BN");
  if (force) { int f = atoi(force); if ((f == 64 || f == 128 || f == 256) && a->N % f == 0) bn = f; }
  switch (bn) {
    case 256: return launch_gemm<256>(a, s);
    case 128: return launch_gemm<128>(a, s);
    default: return launch_gemm<64>(a, s);
  }
}
