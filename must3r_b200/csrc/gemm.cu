// Persistent warp-specialised tcgen05 GEMM for sm_100a:  out = epilogue(A[M,K] * W[N,K]^T).
//
//   warp 0      : TMA producer (one elected lane) -- A and W tiles, 128B-swizzled, mbarrier ring
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (128 x BN x 16 per instruction)
//   warps 2..9  : epilogue -- tcgen05.ld the fp32 accumulator (one row per thread, half the columns per warp), bias / RoPE / GELU /
//                 residual / image2-embed row bias, vectorised global stores
// The accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the MMAs of
// tile i+1.  Tiles are walked m-fastest so that co-resident CTAs share the weight tile in L2.
#include <map>
#include <mutex>
#include <utility>
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
  static constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES;   // 256:4  192:4  160:5  128:6  64:8
  static constexpr int TMEM_COLS = 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;   // power of two
  static constexpr int NCHUNK = BN / 32;                     // 32-column epilogue chunks
  static constexpr int CH0 = (NCHUNK + 1) / 2;               // chunks drained by the first warp of a lane quarter; the second takes the rest
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int PEER_STAGE_BYTES = 8 * 4096;          // per epilogue warp 32 rows x 128 B: transposes peer stores (below)
  static constexpr int SMEM_BYTES_PEER = SMEM_BYTES + PEER_STAGE_BYTES;
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
  // ---- grouped GEMM: `groups` independent problems of M rows each that share the shapes: A rows [g*M, (g+1)*M) of one
  // [groups*M, K] matrix, W rows [g*w_group_rows, +N), bias + g*bias_group, outputs from GemmGroupTab (groups == 1: `out`)
  int groups;
  long long w_group_rows, bias_group;
  // ---- LayerNorm emitted by the epilogue (EMIT kernels): besides `out` (fp32 x = acc + bias + residual) the CTAs of one
  // m-tile exchange per-row (mean, M2) partials through `stats`, meet at a device-scope counter and write the normalised
  // row (x - mean) * rstd as 16-bit into norm_out - the A operand of the next GEMM, whose weights carry the affine.
  void* norm_out;
  long long ldn;
  float norm_eps;
  float2* stats;               // [tiles_m][2 * tiles_n][128]
  unsigned int* sync;          // [tiles_m][2]: arrivals, departures (both zero between launches)
  // split-K of the LayerNorm-emitting kernel (K >= 2048, i.e. fc2): CTA pairs (tile, ks = 0 / 1) each stream half of K;
  // the ks = 1 CTA hands its fp32 accumulator tile to the ks = 0 CTA through `kpart` and a flag, the ks = 0 CTA adds it and
  // runs the epilogue.  Halves the bytes ONE SM pulls from L2, which is what bounds these one-wave GEMMs.
  int ksplit;                  // 1 or 2
  float4* kpart;               // [tiles][2 column halves][8][128] float4
  unsigned int* kflag;         // [tiles], zero between launches
};

// output pointers of a grouped GEMM (passed by value: kernel parameter space)
struct GemmGroupTab {
  void* out[M3R_MAX_GROUPS];
  void* peer[M3R_MAX_GROUPS][M3R_MAX_PEERS];
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// Epilogue arithmetic of one 32-column chunk of one accumulator row: bias, RoPE, row bias, GELU, residual -> v[32].
__device__ __forceinline__ void epilogue_math(const GemmParams& p, const float* bias, const uint32_t (&raw)[32], float (&v)[32], int row,
                                              bool row_ok, bool rb_on, const float* rope_row, int col0) {
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
  if (bias) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col0 + i));
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
  if (row_ok && p.residual) {
    const float4* r4 = reinterpret_cast<const float4*>(p.residual + (long long)row * p.ldr + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 r = r4[i];
      v[4 * i] += r.x; v[4 * i + 1] += r.y; v[4 * i + 2] += r.z; v[4 * i + 3] += r.w;
    }
  }
}

// Store of one chunk: fp32 or 16-bit row segment; the 16-bit form optionally also goes to the peers' buffers.
// A peer store instruction of a warp covers 32 rows x 16 B; the four uint4 of a thread complete a 64-byte segment per
// row back to back, so NVLink sees 64 B write bursts (the fused GEMM -> all-gather path).
__device__ __forceinline__ void epilogue_store(const GemmParams& p, void* out, void* const* peers, const float (&v)[32], long long orow, int col0) {
  if (p.out_dtype == 0) {
    float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + orow * p.ldc + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) o4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
    uint4 w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      w[i].x = pack16(v[8 * i], v[8 * i + 1], p.is_bf16);
      w[i].y = pack16(v[8 * i + 2], v[8 * i + 3], p.is_bf16);
      w[i].z = pack16(v[8 * i + 4], v[8 * i + 5], p.is_bf16);
      w[i].w = pack16(v[8 * i + 6], v[8 * i + 7], p.is_bf16);
    }
    uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(out) + orow * p.ldc + col0);
#pragma unroll
    for (int i = 0; i < 4; ++i) o4[i] = w[i];
    for (int pr = 0; pr < p.n_peer_out; ++pr) {
      uint4* q4 = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(peers[pr]) + orow * p.ldc + col0);
#pragma unroll
      for (int i = 0; i < 4; ++i) q4[i] = w[i];
    }
  }
}

__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, const uint32_t (&raw)[32], int row, bool row_ok, long long orow,
                                               bool rb_on, const float* rope_row, int col0) {
  float v[32];
  epilogue_math(p, p.bias, raw, v, row, row_ok, rb_on, rope_row, col0);
  if (row_ok) epilogue_store(p, p.out, p.peer_out, v, orow, col0);
}

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int atom_add_release_u32(unsigned int* p, unsigned int v) {
  unsigned int o;
  asm volatile("atom.add.release.gpu.global.u32 %0, [%1], %2;" : "=r"(o) : "l"(p), "r"(v) : "memory");
  return o;
}

constexpr int MODE_PLAIN = 0, MODE_EMIT = 1, MODE_GROUPED = 2;
struct NoTab {};
template <int MODE> struct TabOf { using type = NoTab; };
template <> struct TabOf<MODE_GROUPED> { using type = GemmGroupTab; };

template <int BN, int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmParams p,
            const __grid_constant__ typename TabOf<MODE>::type tab) {
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
  const int tiles_mg = (p.M + BM - 1) / BM;          // m-tiles per group
  const int tiles_m = tiles_mg * p.groups;
  const int tiles_n = p.N / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb_all = p.K / BK;
  // LayerNorm-emitting kernels own exactly one tile per CTA (or per CTA pair with split-K); the others walk tiles persistently
  const int ksplit = (MODE == MODE_EMIT) ? p.ksplit : 1;
  const int ks = (MODE == MODE_EMIT) ? (int)blockIdx.x / num_tiles : 0;
  const int tile0 = (MODE == MODE_EMIT) ? (int)blockIdx.x % num_tiles : (int)blockIdx.x;
  const int tile_step = (MODE == MODE_EMIT) ? num_tiles : (int)gridDim.x;
  const int num_kb = num_kb_all / ksplit;
  const int kb0 = ks * num_kb;
  // tile t -> (group g, m-tile inside the group, n-tile); m-fastest so that co-resident CTAs share the weight tile in L2
  auto a_row0 = [&](int tm) { return (tm / tiles_mg) * p.M + (tm % tiles_mg) * BM; };      // first A row of m-tile tm
  auto w_row0 = [&](int tm, int tn) { return (long long)(tm / tiles_mg) * p.w_group_rows + (long long)tn * BN; };

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
      if (p.w_static && tile0 < num_tiles) {
        const int n0 = (int)w_row0(tile0 % tiles_m, tile0 / tiles_m);
        pre = num_kb < STAGES ? num_kb : STAGES;
        for (int kb = 0; kb < pre; ++kb) {
          mbar_arrive_expect_tx(&full[kb], Cfg::STAGE_BYTES);
          tma_load_2d(sB + kb * BN * BK * 2, &tmW, &full[kb], (kb0 + kb) * BK, n0);
        }
      }
      griddep_wait();
      griddep_launch();
      M3R_TR(if (tr) tr[2] = gtime_ns();)
      int stage = 0; uint32_t phase = 0;
      for (int t = tile0; t < num_tiles; t += tile_step) {
        const int m0 = a_row0(t % tiles_m);
        const int n0 = (int)w_row0(t % tiles_m, t / tiles_m);
        for (int kb = 0; kb < num_kb; ++kb) {
          if (pre > 0) {                      // stage already armed, W half in flight
            --pre;
            tma_load_2d(sA + stage * BM * BK * 2, &tmA, &full[stage], (kb0 + kb) * BK, m0);
          } else {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES);
            tma_load_2d(sA + stage * BM * BK * 2, &tmA, &full[stage], (kb0 + kb) * BK, m0);
            tma_load_2d(sB + stage * BN * BK * 2, &tmW, &full[stage], (kb0 + kb) * BK, n0);
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
    for (int t = tile0; t < num_tiles; t += tile_step) {
      mbar_wait(&tempty[as], aphase ^ 1);      // epilogue has drained this accumulator buffer
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        M3R_TR(if (tr && kb == 0 && t == tile0 && (threadIdx.x & 31) == 0) tr[4] = gtime_ns();)
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
    const int chalf = (warp - 2) >> 2;            // which share of the tile's 32-column chunks this warp drains
    const int c_lo = chalf == 0 ? 0 : Cfg::CH0, c_hi = chalf == 0 ? Cfg::CH0 : Cfg::NCHUNK;
    const int lane = threadIdx.x & 31;
    int as = 0; uint32_t aphase = 0;
    for (int t = tile0; t < num_tiles; t += tile_step) {
      const int tm = t % tiles_m, tn = t / tiles_m;
      const int g = tm / tiles_mg;
      const int lrow = (tm % tiles_mg) * BM + quarter * 32 + lane;      // row inside the group
      const int n0 = tn * BN;
      const bool row_ok = lrow < p.M;
      long long orow = lrow;
      if (p.rows_per_batch > 0) orow = (long long)(lrow / p.rows_per_batch) * p.batch_stride_rows + lrow % p.rows_per_batch;
      const bool rb_on = p.rowbias != nullptr && (lrow % p.rb_period) >= p.rb_first;
      const float* rope_row = p.rope_tab ? p.rope_tab + (long long)(lrow % p.rope_period) * 64 : nullptr;
      const float* bias = p.bias ? p.bias + (long long)g * p.bias_group : nullptr;

      mbar_wait(&tfull[as], aphase);
      M3R_TR(if (tr && threadIdx.x == 64) tr[6] = gtime_ns();)
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t(quarter * 32) << 16) + as * BN;
      if (MODE == MODE_EMIT) {
        // ---- x = acc + bias + residual -> out (fp32); the thread keeps its BN/2 values for the normalisation
        static_assert(MODE != MODE_EMIT || BN == 64, "LayerNorm-emitting epilogue: one 32-column chunk per thread");
        uint32_t raw[32];
        tmem_ld32(t_addr + c_lo * 32, raw);
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(&tempty[as]);
        if (ksplit > 1) {
          // [tile][column half][8 float4][128 rows]: lanes are consecutive rows -> 512 B per warp request
          float4* kp = p.kpart + ((long long)(t * 2 + chalf) * 8) * BM + (quarter * 32 + lane);
          if (ks == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              kp[(long long)i * BM] = make_float4(__uint_as_float(raw[4 * i]), __uint_as_float(raw[4 * i + 1]), __uint_as_float(raw[4 * i + 2]), __uint_as_float(raw[4 * i + 3]));
            __threadfence();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (warp == 2 && lane == 0) atom_add_release_u32(p.kflag + t, 1u);
            continue;                                           // the partner CTA finishes the tile
          }
          if (warp == 2 && lane == 0) {
            unsigned int spins = 0;
            while (ld_acquire_u32(p.kflag + t) == 0u) { if (++spins > (1u << 26)) __trap(); }
            p.kflag[t] = 0u;                                    // re-armed for the next launch (nobody else touches it any more)
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 o = __ldcg(kp + (long long)i * BM);
            raw[4 * i] = __float_as_uint(__uint_as_float(raw[4 * i]) + o.x); raw[4 * i + 1] = __float_as_uint(__uint_as_float(raw[4 * i + 1]) + o.y);
            raw[4 * i + 2] = __float_as_uint(__uint_as_float(raw[4 * i + 2]) + o.z); raw[4 * i + 3] = __float_as_uint(__uint_as_float(raw[4 * i + 3]) + o.w);
          }
        }
        float v[32];
        epilogue_math(p, bias, raw, v, lrow, row_ok, rb_on, rope_row, n0 + c_lo * 32);
        if (!row_ok) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
        // two-pass statistics of the thread's 32 values, combined across the row's 2 * tiles_n partials with Chan's
        // formula (as accurate as a two-pass LayerNorm over the whole row)
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i += 2) { s0 += v[i]; s1 += v[i + 1]; }
        const float mu = (s0 + s1) * (1.0f / 32.0f);
        float q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i += 2) { const float a = v[i] - mu, b = v[i + 1] - mu; q0 = fmaf(a, a, q0); q1 = fmaf(b, b, q1); }
        const int parts = 2 * tiles_n;
        const int rit = quarter * 32 + lane;                       // row in tile
        float2* st = p.stats + ((long long)tm * parts) * BM;
        st[(long long)(tn * 2 + chalf) * BM + rit] = make_float2(mu, q0 + q1);
        __threadfence();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        unsigned int* cnt = p.sync + 2 * tm;
        if (warp == 2 && lane == 0) atom_add_release_u32(cnt, 1u);
        // the fp32 residual stream goes out while the other CTAs of the row block arrive (its 32 KB per CTA would otherwise
        // sit in front of the fence above)
        if (row_ok) epilogue_store(p, p.out, p.peer_out, v, orow, n0 + c_lo * 32);
        if (warp == 2 && lane == 0) {
          unsigned int spins = 0;
          while (ld_acquire_u32(cnt) < (unsigned int)tiles_n) { if (++spins > (1u << 26)) __trap(); }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        float msum = 0.f;
        float2 pr[24];
        const int np = parts <= 24 ? parts : 24;                   // host guarantees parts <= 24
#pragma unroll
        for (int i = 0; i < 24; ++i) if (i < np) { pr[i] = __ldcg(st + (long long)i * BM + rit); msum += pr[i].x; }
        const float mean = msum / (float)np;
        float m2 = 0.f;
#pragma unroll
        for (int i = 0; i < 24; ++i) if (i < np) { const float d = pr[i].x - mean; m2 += pr[i].y + 32.0f * d * d; }
        const float rstd = rsqrtf(m2 / (32.0f * (float)np) + p.norm_eps);
        if (row_ok) {
          uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.norm_out) + (long long)lrow * p.ldn + n0 + c_lo * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 w;
            w.x = pack16((v[8 * i] - mean) * rstd, (v[8 * i + 1] - mean) * rstd, p.is_bf16);
            w.y = pack16((v[8 * i + 2] - mean) * rstd, (v[8 * i + 3] - mean) * rstd, p.is_bf16);
            w.z = pack16((v[8 * i + 4] - mean) * rstd, (v[8 * i + 5] - mean) * rstd, p.is_bf16);
            w.w = pack16((v[8 * i + 6] - mean) * rstd, (v[8 * i + 7] - mean) * rstd, p.is_bf16);
            o4[i] = w;
          }
        }
        // departure: the last CTA of the m-tile to have read the partials re-arms both counters for the next launch
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (warp == 2 && lane == 0) {
          if (atomicAdd(cnt + 1, 1u) == (unsigned int)tiles_n - 1u) { cnt[1] = 0u; __threadfence(); cnt[0] = 0u; }
        }
      } else {
        void* out_g = p.out;
        void* const* peers_g = p.peer_out;
        if constexpr (MODE == MODE_GROUPED) { out_g = tab.out[g]; peers_g = tab.peer[g]; }
        if (c_lo == c_hi) {                       // narrow tiles: this warp has no chunk, it only returns the buffer
          tc_fence_before();
          mbar_arrive(&tempty[as]);
        }
        if (p.n_peer_out > 0 && p.out_dtype != 0) {
          // ---- fused GEMM -> all-gather epilogue.  A thread owns one ROW of the tile, so direct peer stores would be 32
          // scattered 16-byte pieces per instruction - partial sectors over NVLink (measured 180 GB/s at 8 GPUs, r02_run6).
          // The warp's rows x (up to) 64 columns go through a swizzled shared-memory tile and leave as 128-byte row segments:
          // one store instruction = 4 rows x 128 B (8 lanes per row), full sectors, to every peer's buffer.
          uint4* stg = reinterpret_cast<uint4*>(smem + STAGES * Cfg::STAGE_BYTES + 256) + (warp - 2) * 256;      // [32 rows][8 x 16 B]
          const int row_base = (tm % tiles_mg) * BM + quarter * 32;
#pragma unroll 1
          for (int c = c_lo; c < c_hi; c += 2) {
            const int nch = (c_hi - c) < 2 ? (c_hi - c) : 2;
            for (int u = 0; u < nch; ++u) {
              uint32_t raw[32];
              tmem_ld32(t_addr + (c + u) * 32, raw);
              tmem_wait_ld();
              if (c + u == c_hi - 1) { tc_fence_before(); mbar_arrive(&tempty[as]); }
              float v[32];
              epilogue_math(p, bias, raw, v, lrow, row_ok, rb_on, rope_row, n0 + (c + u) * 32);
              uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(out_g) + orow * p.ldc + n0 + (c + u) * 32);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint4 w;
                w.x = pack16(v[8 * i], v[8 * i + 1], p.is_bf16); w.y = pack16(v[8 * i + 2], v[8 * i + 3], p.is_bf16);
                w.z = pack16(v[8 * i + 4], v[8 * i + 5], p.is_bf16); w.w = pack16(v[8 * i + 6], v[8 * i + 7], p.is_bf16);
                if (row_ok) o4[i] = w;                                  // this GPU's copy: the thread's own row
                stg[lane * 8 + ((u * 4 + i) ^ (lane & 7))] = w;
              }
            }
            __syncwarp();
            const int lpr = 4 * nch, rpi = 32 / lpr;                    // lanes per row, rows per store instruction
            for (int k = 0; k < lpr; ++k) {
              const int r = k * rpi + lane / lpr, piece = lane % lpr;
              const int row_r = row_base + r;
              if (row_r < p.M) {
                long long orow_r = row_r;
                if (p.rows_per_batch > 0) orow_r = (long long)(row_r / p.rows_per_batch) * p.batch_stride_rows + row_r % p.rows_per_batch;
                const uint4 val = stg[r * 8 + (piece ^ (r & 7))];
                for (int pr = 0; pr < p.n_peer_out; ++pr)
                  reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(peers_g[pr]) + orow_r * p.ldc + n0 + c * 32)[piece] = val;
              }
            }
            __syncwarp();
          }
        } else
#pragma unroll 1
        for (int c = c_lo; c < c_hi; ++c) {
          uint32_t raw[32];
          tmem_ld32(t_addr + c * 32, raw);
          tmem_wait_ld();
          if (c == c_hi - 1) {  // this warp's share of the accumulator is read: hand it back early
            tc_fence_before();
            mbar_arrive(&tempty[as]);
          }
          float v[32];
          epilogue_math(p, bias, raw, v, lrow, row_ok, rb_on, rope_row, n0 + c * 32);
          if (row_ok) epilogue_store(p, out_g, peers_g, v, orow, n0 + c * 32);
        }
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

static void fill_params(GemmParams& p, const m3r_gemm_args* a) {
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
  p.groups = 1; p.w_group_rows = 0; p.bias_group = 0;
  p.norm_out = nullptr; p.ldn = 0; p.norm_eps = 0.f; p.stats = nullptr; p.sync = nullptr;
  p.ksplit = 1; p.kpart = nullptr; p.kflag = nullptr;
}

// Scratch of the LayerNorm-emitting epilogue: partial row statistics + arrival / departure counters (+ split-K hand-over).
// The CTAs of a launch spin on each other, so launches sharing a scratch must be ordered: the model code emits on the
// caller's stream only, never on its side streams.
struct EmitScratch { float2* stats = nullptr; unsigned int* sync = nullptr; float4* kpart = nullptr; unsigned int* kflag = nullptr; };
constexpr int EMIT_MAX_KSPLIT_TILES = 80;
constexpr int EMIT_MAX_TILES_M = 16, EMIT_MAX_PARTS = 24;
static EmitScratch* emit_scratch(cudaStream_t stream) {
  // one scratch per (device, stream): LayerNorm-emitting launches of different streams (two host threads driving the model,
  // as the reference's SLAM worker does, slam.py:533) must not share counters.  Entries live for the process lifetime.
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, EmitScratch> table;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  EmitScratch& e = table[std::make_pair(dev, stream)];
  if (!e.stats) {
    void* a = nullptr; void* b = nullptr;
    if (cudaMalloc(&a, sizeof(float2) * EMIT_MAX_TILES_M * EMIT_MAX_PARTS * BM) != cudaSuccess) return nullptr;
    if (cudaMalloc(&b, sizeof(unsigned int) * 2 * EMIT_MAX_TILES_M) != cudaSuccess) { cudaFree(a); return nullptr; }
    if (cudaMemset(b, 0, sizeof(unsigned int) * 2 * EMIT_MAX_TILES_M) != cudaSuccess) { cudaFree(a); cudaFree(b); return nullptr; }
    void* c = nullptr; void* d = nullptr;
    if (cudaMalloc(&c, sizeof(float) * EMIT_MAX_KSPLIT_TILES * BM * 64) != cudaSuccess) { cudaFree(a); cudaFree(b); return nullptr; }
    if (cudaMalloc(&d, sizeof(unsigned int) * EMIT_MAX_KSPLIT_TILES) != cudaSuccess || cudaMemset(d, 0, sizeof(unsigned int) * EMIT_MAX_KSPLIT_TILES) != cudaSuccess) {
      cudaFree(a); cudaFree(b); cudaFree(c); return nullptr;
    }
    e.stats = reinterpret_cast<float2*>(a); e.sync = reinterpret_cast<unsigned int*>(b);
    e.kpart = reinterpret_cast<float4*>(c); e.kflag = reinterpret_cast<unsigned int*>(d);
  }
  return &e;
}

template <int BN, int MODE>
static int launch_gemm(const m3r_gemm_args* a, cudaStream_t stream, const m3r_gemm_group* grp = nullptr) {
  using Cfg = GemmCfg<BN>;
  const int groups = (MODE == MODE_GROUPED) ? grp->groups : 1;
  CUtensorMap tmA, tmW;
  if (make_tmap_2d(&tmA, a->A, a->is_bf16, (uint64_t)a->K, (uint64_t)a->M * groups, (uint64_t)a->lda, BK, BM)) return 1;
  if (make_tmap_2d(&tmW, a->W, a->is_bf16, (uint64_t)a->K, MODE == MODE_GROUPED ? (uint64_t)((groups - 1) * grp->w_group_rows + a->N) : (uint64_t)a->N, (uint64_t)a->ldw, BK, BN)) return 1;
  GemmParams p;
  fill_params(p, a);
  typename TabOf<MODE>::type tab;
  const int tiles = ((a->M + BM - 1) / BM) * groups * (a->N / BN);
  if constexpr (MODE == MODE_GROUPED) {
    p.groups = groups; p.w_group_rows = grp->w_group_rows; p.bias_group = grp->bias_group;
    for (int g = 0; g < M3R_MAX_GROUPS; ++g) {
      tab.out[g] = g < groups ? grp->out[g] : nullptr;
      for (int r = 0; r < M3R_MAX_PEERS; ++r) tab.peer[g][r] = (g < groups && r < a->n_peer_out) ? grp->peer_out[g * M3R_MAX_PEERS + r] : nullptr;
    }
  }
  if constexpr (MODE == MODE_EMIT) {
    EmitScratch* e = emit_scratch(stream);
    if (!e) return set_error("gemm: LayerNorm-emit scratch allocation failed");
    p.norm_out = a->norm_out; p.ldn = a->ldn; p.norm_eps = a->norm_eps; p.stats = e->stats; p.sync = e->sync;
    static int ks_env = -1;
    if (ks_env < 0) { const char* v = getenv("M3R_KSPLIT"); ks_env = (v && v[0] == '0') ? 0 : 1; }
    if (ks_env && a->K >= 2048 && (a->K / BK) % 2 == 0 && 2 * tiles <= num_sms() && tiles <= EMIT_MAX_KSPLIT_TILES) {
      p.ksplit = 2; p.kpart = e->kpart; p.kflag = e->kflag;
    }
  }
  static bool attr_set[64] = {};
  int dev = 0; cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES_PEER);
    if (e != cudaSuccess) return set_error("gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set[dev] = true;
  }
  int grid = tiles < num_sms() ? tiles : num_sms();
  if (MODE == MODE_EMIT) grid = tiles * p.ksplit;
  if (MODE == MODE_GROUPED && grp->max_ctas > 0 && grid > grp->max_ctas) grid = grp->max_ctas;
  if (MODE == MODE_EMIT && grid > num_sms()) return set_error("gemm: LayerNorm-emitting epilogue needs one tile per CTA (%d tiles, %d SMs)", tiles, grid);
  {
    const int cat = BN >= 256 ? PROF_GEMM256 : (BN >= 128 ? PROF_GEMM128 : PROF_GEMM64);
    ProfScope prof(cat, 2.0 * a->M * groups * (double)a->N * a->K, 2.0 * ((double)a->M * groups * a->K + (double)a->N * groups * a->K) + (double)a->M * groups * a->N * (a->out_dtype ? 2 : 4), stream);
    const size_t smem_bytes = (a->n_peer_out > 0 && a->out_dtype != 0) ? Cfg::SMEM_BYTES_PEER : Cfg::SMEM_BYTES;
    cudaError_t le = launch_pdl(gemm_kernel<BN, MODE>, dim3(grid), dim3(GEMM_THREADS), smem_bytes, stream, tmA, tmW, p, tab);
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
  fill_params(p, a);
  static bool attr_set[64] = {};
  int dev = 0; cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES);
    if (e != cudaSuccess) return set_error("gemm(pair): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set[dev] = true;
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

static int check_args(const m3r_gemm_args* a) {
  if (!a || !a->A || !a->W) return set_error("gemm: null pointer");
  if (a->N % 32 || a->K % 64 || a->N <= 0 || a->K <= 0) return set_error("gemm: N (%d) must be a multiple of 32 and K (%d) of 64", a->N, a->K);
  if (a->lda % 8 || a->ldw % 8) return set_error("gemm: lda/ldw must be multiples of 8 elements (16 B)");
  if ((a->out_dtype == M3R_OUT_F32 && a->ldc % 4) || (a->out_dtype == M3R_OUT_16 && a->ldc % 8)) return set_error("gemm: ldc alignment");
  if (a->residual && a->ldr % 4) return set_error("gemm: ldr alignment");
  if (a->rope_tab && (a->rope_cols % 64)) return set_error("gemm: rope_cols must be a multiple of 64");
  if (a->n_peer_out < 0 || a->n_peer_out > M3R_MAX_PEERS || (a->n_peer_out > 0 && a->out_dtype != M3R_OUT_16)) return set_error("gemm: peer outputs need 0..%d pointers and a 16-bit output", M3R_MAX_PEERS);
  return 0;
}

// Tile width for the 1-CTA kernel.  One-wave problems (the one-view chain) are bound by the per-SM L2->smem ingest, i.e. by
// the bytes ONE CTA has to pull: K * (128 + BN) * 2 - so the narrowest tile that still fits in one wave wins (most SMs
// busy, least bytes per SM).  Multi-wave problems take the widest tile that keeps every SM busy (fewer re-reads of A).
static int pick_bn(int M, int N, int groups) {
  const int tiles_m = ((M + BM - 1) / BM) * groups;
  const int sms = num_sms();
  static const int cand[] = {64, 128, 160, 192, 256};
  for (int bn : cand)
    if (N % bn == 0 && tiles_m * (N / bn) <= sms) return bn;
  if (N % 256 == 0 && tiles_m * (N / 256) >= sms) return 256;
  if (N % 128 == 0 && tiles_m * (N / 128) >= (sms * 2) / 3) return 128;
  if (N % 64 == 0) return 64;
  if (N % 160 == 0) return 160;
  return 32;
}

}  // namespace m3r

extern "C" int m3r_gemm(const m3r_gemm_args* a, void* stream) {
  using namespace m3r;
  if (check_args(a)) return 1;
  if (!a->out) return set_error("gemm: null pointer");
  if (a->M <= 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int sms = num_sms();
  if (a->norm_out) {
    // LayerNorm-emitting epilogue: fp32 out = the new residual stream, norm_out = its normalised 16-bit rows
    if (a->N % 64 || a->N > 32 * EMIT_MAX_PARTS || a->out_dtype != M3R_OUT_F32 || a->ldn % 8)
      return set_error("gemm: norm_out needs N %% 64 == 0, N <= %d, an fp32 output and ldn %% 8 == 0", 32 * EMIT_MAX_PARTS);
    const int tiles_m = (a->M + BM - 1) / BM;
    if (tiles_m > EMIT_MAX_TILES_M || tiles_m * (a->N / 64) > sms)
      return set_error("gemm: norm_out needs one 128x64 tile per SM (M=%d N=%d); normalise with m3r_layernorm instead", a->M, a->N);
    return launch_gemm<64, MODE_EMIT>(a, s);
  }
  {
    // CTA-pair kernel for problems with at least one 256x256 tile per SM pair.  Measured (profiles/r01_run13_gemm_pair.log):
    // +7..+20 % for K >= 1024, -1..-4 % for K = 768 (epilogue-bound tiles) -> used for K >= 1024; M3R_GEMM_PAIR=0/2
    // force it off / on for every eligible shape.
    static int pair_mode = -1;
    if (pair_mode < 0) { const char* e = getenv("M3R_GEMM_PAIR"); pair_mode = e ? atoi(e) : 1; }
    const bool eligible = a->N % 256 == 0 && ((a->M + 255) / 256) * (a->N / 256) >= sms / 2 && !getenv("M3R_GEMM_BN");
    if (eligible && (pair_mode == 2 || (pair_mode == 1 && a->K >= 1024))) return launch_gemm_pair(a, s);
  }
  int bn = pick_bn(a->M, a->N, 1);
  const char* force = getenv("M3R_GEMM_BN");
  if (force) { int f = atoi(force); if ((f == 32 || f == 64 || f == 128 || f == 160 || f == 192 || f == 256) && a->N % f == 0) bn = f; }
  switch (bn) {
    case 256: return launch_gemm<256, MODE_PLAIN>(a, s);
    case 192: return launch_gemm<192, MODE_PLAIN>(a, s);
    case 160: return launch_gemm<160, MODE_PLAIN>(a, s);
    case 128: return launch_gemm<128, MODE_PLAIN>(a, s);
    case 64: return launch_gemm<64, MODE_PLAIN>(a, s);
    default: return launch_gemm<32, MODE_PLAIN>(a, s);
  }
}

// `groups` GEMMs of identical shape in one launch (the post-feedback K|V projections of all decoder levels,
// must3r/model/decoder.py:323-330): A = [groups * M, K] (group g = rows [g*M, (g+1)*M)), W / bias stacked with strides
// w_group_rows / bias_group, one output (and one set of peer outputs) per group.
extern "C" int m3r_gemm_grouped(const m3r_gemm_args* a, const m3r_gemm_group* grp, void* stream) {
  using namespace m3r;
  if (check_args(a)) return 1;
  if (!grp || grp->groups < 1 || grp->groups > M3R_MAX_GROUPS) return set_error("gemm_grouped: 1..%d groups", M3R_MAX_GROUPS);
  if (a->norm_out || a->residual) return set_error("gemm_grouped: no residual / norm_out");
  for (int g = 0; g < grp->groups; ++g) if (!grp->out[g]) return set_error("gemm_grouped: out[%d] is null", g);
  if (a->M <= 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int bn = pick_bn(a->M, a->N, grp->groups);
  if (bn != 256 && bn != 128) bn = a->N % 128 == 0 ? 128 : 64;
  if (a->N % bn) return set_error("gemm_grouped: N (%d) must be a multiple of 64", a->N);
  switch (bn) {
    case 256: return launch_gemm<256, MODE_GROUPED>(a, s, grp);
    case 128: return launch_gemm<128, MODE_GROUPED>(a, s, grp);
    default: return launch_gemm<64, MODE_GROUPED>(a, s, grp);
  }
}
