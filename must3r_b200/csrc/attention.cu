// FlashAttention-style softmax(Q K^T * scale) V for head_dim 64 on sm_100a tensor cores (tcgen05 + TMEM).
//
// One CTA = one (batch, head, 128-query tile); two CTAs are co-resident per SM (80 KB smem, 256 TMEM
// columns each) so that one CTA's softmax overlaps the other's MMAs.
//   warp 0     : TMA producer -- Q once, then K and V tiles (128 keys x 64) through 2-deep mbarrier rings
//   warp 1     : TMEM allocator + single-thread MMA issuer:  S = Q K^T (SS form), O_tile = P V (A = P from TMEM,
//                B = V as an MN-major smem operand, so V is consumed in its natural [keys, d] layout)
//   warps 2..5 : softmax -- one query row per thread (TMEM lane == row, no shuffles needed), two passes over
//                S in TMEM (row max, then exp2 / row sum / pack P), online-softmax state (m, l) and the output
//                accumulator O[64] in registers.
// TMEM buffer b (128 columns) holds S_b; after the softmax has consumed it, P_b overwrites columns [0,64)
// (packed 16-bit pairs) and the P V product lands in columns [64,128).
//
// Keys come from two segments (stored memory + this step's new tokens) so the reference's torch.cat of the
// memory (decoder.py:306) never happens; a per-batch skip range implements make_mem_mask (decoder.py:119-139).
#include <math.h>
#include "ptx.cuh"
#include "m3r_internal.h"

namespace m3r {

constexpr int AT_BM = 128;
constexpr int AT_BN = 128;
constexpr int HD = 64;
constexpr int KS = 2;                        // K / V ring depth
constexpr int AT_THREADS = 192;
constexpr int TILE_BYTES = 128 * HD * 2;     // 16 KB
constexpr int AT_SMEM = (1 + 2 * KS) * TILE_BYTES + 1024 + 256;
constexpr int AT_TMEM_COLS = 256;

struct AttnParams {
  int Nq, Nk0, Nk1;
  int kv_group;
  int skip_lo, skip_step, skip_len;
  int is_bf16;
  float sl2;                 // scale * log2(e)
  void* O;
  long long ldo;
};

struct TileIt {
  int seg, t;          // current tile
  int g0;              // global index of its first key
  int nvalid;          // valid local keys in the tile (<=128)
  bool mask;           // needs element masking
};

// Enumerate key tiles, skipping the ones entirely inside the skip range. Every warp role runs the same walk.
struct TileWalk {
  int nk[2], lo, hi;
  int seg, t;
  __device__ TileWalk(int nk0, int nk1, int lo_, int hi_) : lo(lo_), hi(hi_), seg(0), t(0) { nk[0] = nk0; nk[1] = nk1; }
  __device__ bool next(TileIt& it) {
    while (seg < 2) {
      const int n = nk[seg];
      if (t * AT_BN >= n) { ++seg; t = 0; continue; }
      const int l0 = t * AT_BN;
      const int l1 = min(l0 + AT_BN, n);
      const int base = seg == 0 ? 0 : nk[0];
      const int g0 = base + l0, g1 = base + l1;
      const int cur_t = t++;
      if (g0 >= lo && g1 <= hi) continue;     // fully masked: never loaded nor multiplied
      it.seg = seg; it.t = cur_t; it.g0 = g0; it.nvalid = l1 - l0;
      it.mask = (l1 - l0 < AT_BN) || (g0 < hi && g1 > lo);
      return true;
    }
    return false;
  }
};

__global__ void __launch_bounds__(AT_THREADS, 2)
attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK0,
            const __grid_constant__ CUtensorMap tmV0, const __grid_constant__ CUtensorMap tmK1,
            const __grid_constant__ CUtensorMap tmV1, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + TILE_BYTES;
  uint8_t* sV = smem + (1 + KS) * TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (1 + 2 * KS) * TILE_BYTES);
  uint64_t* q_full = bars;             // [1]
  uint64_t* k_full = bars + 1;         // [KS]
  uint64_t* k_empty = k_full + KS;     // [KS]
  uint64_t* v_full = k_empty + KS;     // [KS]
  uint64_t* v_empty = v_full + KS;     // [KS]
  uint64_t* s_full = v_empty + KS;     // [2]   MMA -> softmax : S_b ready
  uint64_t* p_full = s_full + 2;       // [2]   softmax -> MMA : P_b stored in TMEM
  uint64_t* o_full = p_full + 2;       // [2]   MMA -> softmax : (P V)_b ready
  uint64_t* s_empty = o_full + 2;      // [2]   softmax -> MMA : buffer b drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int kvb = b / p.kv_group;
  const int lo = p.skip_len > 0 ? p.skip_lo + (b % p.kv_group) * p.skip_step : 0;
  const int hi = p.skip_len > 0 ? lo + p.skip_len : 0;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK0); tma_prefetch_desc(&tmV0);
    if (p.Nk1 > 0) { tma_prefetch_desc(&tmK1); tma_prefetch_desc(&tmV1); }
    mbar_init(q_full, 1);
    for (int s = 0; s < KS; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 128); mbar_init(&o_full[s], 1); mbar_init(&s_empty[s], 128); }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, AT_TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, TILE_BYTES);
      tma_load_3d(sQ, &tmQ, q_full, h * HD, qt * AT_BM, b);
      TileWalk walk(p.Nk0, p.Nk1, lo, hi);
      TileIt it;
      int i = 0;
      while (walk.next(it)) {
        const int st = i % KS;
        const uint32_t ph = (i / KS) & 1;
        const CUtensorMap* mk = it.seg == 0 ? &tmK0 : &tmK1;
        const CUtensorMap* mv = it.seg == 0 ? &tmV0 : &tmV1;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
        tma_load_3d(sK + st * TILE_BYTES, mk, &k_full[st], h * HD, it.t * AT_BN, kvb);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], TILE_BYTES);
        tma_load_3d(sV + st * TILE_BYTES, mv, &v_full[st], h * HD, it.t * AT_BN, kvb);
        ++i;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t bf = p.is_bf16 ? 1u : 0u;
    const uint32_t idesc_qk = make_idesc(AT_BM, AT_BN, bf, 0, 0);   // S[128 x 128] = Q (K-major) * K^T (K-major)
    const uint32_t idesc_pv = make_idesc(AT_BM, HD, bf, 0, 1);      // O[128 x 64]  = P (TMEM)   * V (MN-major)
    int n_tiles = 0;
    { TileWalk w(p.Nk0, p.Nk1, lo, hi); TileIt it; while (w.next(it)) ++n_tiles; }
    const uint64_t qdesc = smem_desc_sw128(smem_u32(sQ));
    auto issue_qk = [&](int i) {
      const int st = i % KS;
      const int buf = i & 1;
      mbar_wait(&k_full[st], (i / KS) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t kdesc = smem_desc_sw128(smem_u32(sK + st * TILE_BYTES));
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_ss(tmem_base + buf * 128, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k ? 1u : 0u);
        umma_commit(&k_empty[st]);
        umma_commit(&s_full[buf]);
      }
      __syncwarp();
    };
    if (n_tiles > 0) {
      mbar_wait(q_full, 0);
      tc_fence_after();
      issue_qk(0);
      if (n_tiles > 1) issue_qk(1);
      for (int i = 0; i < n_tiles; ++i) {
        const int st = i % KS;
        const int buf = i & 1;
        mbar_wait(&p_full[buf], (i >> 1) & 1);
        mbar_wait(&v_full[st], (i / KS) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t vdesc = smem_desc_sw128(smem_u32(sV + st * TILE_BYTES));
#pragma unroll
          for (int k = 0; k < AT_BN / 16; ++k) {
            // 16 keys per MMA: P advances 8 TMEM columns (packed pairs), V advances 16 rows = 2048 B
            umma_ts(tmem_base + buf * 128 + 64, tmem_base + buf * 128 + k * 8, vdesc + (uint64_t)(k * 128), idesc_pv, k ? 1u : 0u);
          }
          umma_commit(&v_empty[st]);
          umma_commit(&o_full[buf]);
        }
        __syncwarp();
        if (i + 2 < n_tiles) {
          mbar_wait(&s_empty[buf], (i >> 1) & 1);   // softmax has read O_i: buffer may be overwritten
          tc_fence_after();
          issue_qk(i + 2);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps
    const int quarter = warp & 3;
    const int lane = threadIdx.x & 31;
    const int row = quarter * 32 + lane;
    const int q_idx = qt * AT_BM + row;
    const uint32_t lane_addr = tmem_base + (uint32_t(quarter * 32) << 16);
    float m_run = -INFINITY, l_run = 0.f, alpha_pending = 1.f;
    float acc[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] = 0.f;

    auto drain_o = [&](int i, float alpha) {
      const int buf = i & 1;
      mbar_wait(&o_full[buf], (i >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t raw[32];
        tmem_ld32(lane_addr + buf * 128 + 64 + c * 32, raw);
        tmem_wait_ld();
#pragma unroll
        for (int d = 0; d < 32; ++d) acc[c * 32 + d] = fmaf(acc[c * 32 + d], alpha, __uint_as_float(raw[d]));
      }
      tc_fence_before();
      mbar_arrive(&s_empty[buf]);
    };

    TileWalk walk(p.Nk0, p.Nk1, lo, hi);
    TileIt it;
    int i = 0;
    while (walk.next(it)) {
      const int buf = i & 1;
      mbar_wait(&s_full[buf], (i >> 1) & 1);
      tc_fence_after();
      const uint32_t s_addr = lane_addr + buf * 128;
      // ---- pass 1: row max
      float mx = m_run;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t raw[32];
        tmem_ld32(s_addr + c * 32, raw);
        tmem_wait_ld();
        if (it.mask) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = c * 32 + j, g = it.g0 + col;
            const bool ok = col < it.nvalid && !(g >= lo && g < hi);
            mx = fmaxf(mx, ok ? __uint_as_float(raw[j]) : -INFINITY);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(raw[j]));
        }
      }
      const float m_safe = (mx == -INFINITY) ? 0.f : mx;
      const float alpha = exp2f((m_run - m_safe) * p.sl2);    // m_run = -inf -> 0
      const float moff = m_safe * p.sl2;
      // ---- pass 2: P = exp2(S*sl2 - m*sl2), row sum, pack to 16-bit, store over S_b[0,64)
      float rsum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t raw[32];
        tmem_ld32(s_addr + c * 32, raw);
        tmem_wait_ld();
        float pv[32];
        if (it.mask) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = c * 32 + j, g = it.g0 + col;
            const bool ok = col < it.nvalid && !(g >= lo && g < hi);
            pv[j] = ok ? exp2f(fmaf(__uint_as_float(raw[j]), p.sl2, -moff)) : 0.f;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) pv[j] = exp2f(fmaf(__uint_as_float(raw[j]), p.sl2, -moff));
        }
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          rsum += pv[2 * j] + pv[2 * j + 1];
          pk[j] = pack16(pv[2 * j], pv[2 * j + 1], p.is_bf16);
        }
        tmem_st16(s_addr + c * 16, pk);
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&p_full[buf]);
      l_run = l_run * alpha + rsum;
      m_run = mx;
      // ---- fold the previous tile's P V product into the register accumulator
      if (i > 0) drain_o(i - 1, alpha_pending);
      alpha_pending = alpha;
      ++i;
    }
    if (i > 0) drain_o(i - 1, alpha_pending);

    if (q_idx < p.Nq) {
      const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
      uint16_t* orow = reinterpret_cast<uint16_t*>(p.O) + ((long long)b * p.Nq + q_idx) * p.ldo + h * HD;
      uint4* o4 = reinterpret_cast<uint4*>(orow);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint4 w;
        w.x = pack16(acc[8 * j] * inv, acc[8 * j + 1] * inv, p.is_bf16);
        w.y = pack16(acc[8 * j + 2] * inv, acc[8 * j + 3] * inv, p.is_bf16);
        w.z = pack16(acc[8 * j + 4] * inv, acc[8 * j + 5] * inv, p.is_bf16);
        w.w = pack16(acc[8 * j + 6] * inv, acc[8 * j + 7] * inv, p.is_bf16);
        o4[j] = w;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, AT_TMEM_COLS);
  }
}

}  // namespace m3r

extern "C" int m3r_attention(const m3r_attn_args* a, void* stream) {
  using namespace m3r;
  if (!a || !a->Q || !a->K0 || !a->V0 || !a->O) return set_error("attention: null pointer");
  if (a->B <= 0 || a->H <= 0 || a->Nq <= 0) return 0;
  if (a->kv_group < 1 || a->B % a->kv_group) return set_error("attention: B (%d) must be a multiple of kv_group (%d)", a->B, a->kv_group);
  if (a->Nk0 < 0 || a->Nk1 < 0 || a->Nk0 + a->Nk1 <= 0) return set_error("attention: no keys");
  if (a->Nk1 > 0 && (!a->K1 || !a->V1)) return set_error("attention: segment 1 pointers missing");
  if (a->ldq % 8 || a->ldk0 % 8 || (a->Nk1 > 0 && a->ldk1 % 8) || a->ldo % 8) return set_error("attention: leading dims must be multiples of 8");
  {
    // a row whose keys are all skipped would be 0/0; the reference never produces one (decoder.py:291-296)
    const int tot = a->Nk0 + a->Nk1;
    if (a->skip_len >= tot && a->skip_len > 0) return set_error("attention: skip range covers every key");
  }
  const int Bkv = a->B / a->kv_group;
  CUtensorMap tmQ, tmK0, tmV0, tmK1, tmV1;
  auto mk3 = [&](CUtensorMap* m, const void* base, int64_t ld, int64_t rows, int64_t bstride_rows, int64_t nb) {
    return make_tmap_3d(m, base, a->is_bf16, (uint64_t)a->H * HD, (uint64_t)rows, (uint64_t)nb, (uint64_t)ld,
                        (uint64_t)bstride_rows * (uint64_t)ld, HD, 128);
  };
  if (mk3(&tmQ, a->Q, a->ldq, a->Nq, a->Nq, a->B)) return 1;
  if (a->Nk0 > 0) {
    if (mk3(&tmK0, a->K0, a->ldk0, a->Nk0, a->kv_bstride0, Bkv)) return 1;
    if (mk3(&tmV0, a->V0, a->ldk0, a->Nk0, a->kv_bstride0, Bkv)) return 1;
  } else { tmK0 = tmQ; tmV0 = tmQ; }
  if (a->Nk1 > 0) {
    if (mk3(&tmK1, a->K1, a->ldk1, a->Nk1, a->kv_bstride1, Bkv)) return 1;
    if (mk3(&tmV1, a->V1, a->ldk1, a->Nk1, a->kv_bstride1, Bkv)) return 1;
  } else { tmK1 = tmK0; tmV1 = tmV0; }
  AttnParams p;
  p.Nq = a->Nq; p.Nk0 = a->Nk0; p.Nk1 = a->Nk1; p.kv_group = a->kv_group;
  p.skip_lo = a->skip_lo; p.skip_step = a->skip_step; p.skip_len = a->skip_len;
  p.is_bf16 = a->is_bf16; p.sl2 = a->scale * 1.4426950408889634f;
  p.O = a->O; p.ldo = a->ldo;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    if (e != cudaSuccess) return set_error("attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid((a->Nq + AT_BM - 1) / AT_BM, a->H, a->B);
  {
    const double nk_eff = (double)(a->Nk0 + a->Nk1 - a->skip_len);
    ProfScope prof(PROF_ATTN, 4.0 * a->B * (double)a->H * a->Nq * nk_eff * HD,
                   2.0 * ((double)a->B * a->Nq * a->H * HD * 2 + (double)(a->B / a->kv_group) * (a->Nk0 + a->Nk1) * a->H * HD * 2),
                   reinterpret_cast<cudaStream_t>(stream));
    attn_kernel<<<grid, AT_THREADS, AT_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(tmQ, tmK0, tmV0, tmK1, tmV1, p);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("attention launch: %s", cudaGetErrorString(e));
  count_launch();
  return 0;
}
