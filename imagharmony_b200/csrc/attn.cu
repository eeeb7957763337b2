// tcgen05 attention for sm_100a, head_dim 64, no mask, softmax scale 1/8 (folded into exp2).
//
// One CTA = 128 query rows of one (batch, head).  warp 0: TMA producer (Q once, K/V double-buffered per 128-key
// block); warp 1: TMEM allocator + single-thread tcgen05.mma issuer (S = Q K^T into TMEM, O_j = P_j V_j into TMEM);
// warps 2..5: one thread per query row -- reads S from TMEM (tcgen05.ld), row max / exp2 / row sum without any
// cross-thread shuffle, writes P as fp16 into 128B-swizzled shared memory (the A operand of the PV MMA), then folds
// the per-block O_j into fp32 register accumulators with the online-softmax rescale.  Two CTAs are co-resident per SM
// (112 KiB smem, 256 TMEM columns each) so one CTA's MMAs overlap the other's softmax.
//
// Decoupled IP cross-attention (n_ip > 0, Nk <= 128): keys/values are [text ; ip]; the row thread runs two separate
// softmaxes over the two column ranges of the same S tile, writes [P_t / l_t | ip_scale * P_ip / l_ip] and a single
// PV MMA yields  softmax(q k_t^T) v_t + ip_scale * softmax(q k_ip^T) v_ip  -- attention_processor.py:423-450.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdlib.h>
#include <math.h>

#include "../../include/ih_api.h"
#include "host_util.h"
#include "ptx.cuh"

namespace ih {

constexpr int ATT_THREADS = 192;
constexpr int ATT_TILE = 128 * 64 * 2;  // 16 KiB: 128 rows x 64 fp16
constexpr int ATT_SMEM_TILES = ATT_TILE * 5 + 2 * ATT_TILE;  // Q, K0, K1, V0, V1, P(2 halves)
constexpr int ATT_SMEM_BYTES = ATT_SMEM_TILES + 128;         // + barriers

struct AttnParams {
  int Nq, Nk, n_ip;
  int num_kv_blocks;
  float scale_log2;  // softmax scale * log2(e)
  float ip_scale;
  __half* out;
  long long ldo;
  // attn2 only: work decomposition.  CTAs [0, n_whole) process a whole (batch, head, 256-query pair); CTA
  // n_whole + i processes KV part (i % split) of pair n_whole + i / split and writes an un-normalised partial result
  // (fp32 O rows, running max, running sum) to the workspace, merged by attn_combine_kernel.
  int H, qpairs, n_whole, split;
  float* ws_o;   // [slots][256][64]
  float* ws_ml;  // [slots][256][2]
  int* ws_cnt;   // [split pairs] arrival counters (zero between launches) for the in-kernel merge; nullptr = separate
                 // attn_combine_kernel launch
  long long* trace;  // optional clock64 stamps of CTA 0, KV blocks 4..7 (ih_attention_set_trace); nullptr in production
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(ATT_THREADS, 2) attn_f16_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                   const __grid_constant__ CUtensorMap tmK,
                                                                   const __grid_constant__ CUtensorMap tmV,
                                                                   const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;
  uint8_t* sK = smem + ATT_TILE;      // 2 stages
  uint8_t* sV = smem + 3 * ATT_TILE;  // 2 stages
  uint8_t* sP = smem + 5 * ATT_TILE;  // 2 K-halves of 64 keys
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATT_SMEM_TILES);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* k_empty = bars + 3;  // [2]
  uint64_t* v_full = bars + 5;   // [2]
  uint64_t* v_empty = bars + 7;  // [2]
  uint64_t* s_full = bars + 9;
  uint64_t* p_full = bars + 10;
  uint64_t* o_full = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int nb = p.num_kv_blocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem_S = tmem_base;        // 128 columns
  const uint32_t tmem_O = tmem_base + 128;  // 64 columns

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, ATT_TILE);
      tma_load_3d(sQ, &tmQ, q_full, head * 64, q0, b);
      for (int j = 0; j < nb; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[s], ATT_TILE);
        tma_load_3d(sK + s * ATT_TILE, &tmK, &k_full[s], head * 64, j * 128, b);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[s], ATT_TILE);
        tma_load_3d(sV + s * ATT_TILE, &tmV, &v_full[s], head * 64, j * 128, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_f16(128, 128, false, false);  // S[128q x 128k] = Q(K-major) K(K-major)^T
      constexpr uint32_t idesc_o = umma_idesc_f16(128, 64, false, true);    // O[128q x 64d] = P(K-major) V(MN-major)
      const uint64_t q_desc = umma_desc_sw128(smem_u32(sQ));
      mbar_wait(q_full, 0);
      // S_0
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      {
        const uint64_t k_desc = umma_desc_sw128(smem_u32(sK));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_S, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0);
        umma_commit(s_full);
        umma_commit(&k_empty[0]);
      }
      for (int j = 0; j < nb; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(p_full, j & 1);
        mbar_wait(&v_full[s], ph);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(sV + s * ATT_TILE);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          // P: K-major, two 64-key halves 16 KiB apart, +32 B per 16 keys inside a half.
          const uint64_t p_desc = umma_desc_sw128(smem_u32(sP) + (kk >> 2) * ATT_TILE) + 2 * (kk & 3);
          // V: MN-major (d contiguous), 16 keys = 16 rows of 128 B = 2048 B per step.
          const uint64_t v_desc = umma_desc_sw128(v_addr + kk * 2048);
          umma_f16_ss(tmem_O, p_desc, v_desc, idesc_o, kk != 0);
        }
        umma_commit(o_full);
        umma_commit(&v_empty[s]);
        if (j + 1 < nb) {
          const int s1 = (j + 1) & 1;
          const uint32_t ph1 = ((j + 1) >> 1) & 1;
          mbar_wait(&k_full[s1], ph1);
          tc_fence_after();
          const uint64_t k_desc = umma_desc_sw128(smem_u32(sK + s1 * ATT_TILE));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_S, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0);
          umma_commit(s_full);
          umma_commit(&k_empty[s1]);
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------- softmax / epilogue: one thread per query row -------------------------------
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    uint8_t* p_row = sP + r * 128;
    const int rx = r & 7;
    const float sl2 = p.scale_log2;
    const bool single = (nb == 1);

    float m_run = -INFINITY, l_run = 0.f;
    float acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = 0.f;

    for (int j = 0; j < nb; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int valid = min(128, p.Nk - j * 128);
      const int n_text = single ? (valid - p.n_ip) : valid;  // columns [0,n_text) text, [n_text,valid) ip

      // pass 1: row maxima (text / ip segments)
      float mx_t = -INFINITY, mx_i = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_S + lane_base + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int col = c * 32 + e;
          const float s = __uint_as_float(v[e]);
          if (col < n_text) mx_t = fmaxf(mx_t, s);
          else if (col < valid) mx_i = fmaxf(mx_i, s);
        }
      }
      float m_new, alpha, inv_t = 1.f, inv_i = 0.f, m_ip = 0.f;
      if (!single) {
        m_new = fmaxf(m_run, mx_t * sl2);
        alpha = ex2f(m_run - m_new);
      } else {
        // single block: normalise before the PV MMA (two independent softmaxes)
        m_new = mx_t * sl2;
        m_ip = mx_i * sl2;
        alpha = 0.f;
        float l_t = 0.f, l_i = 0.f;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_S + lane_base + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int col = c * 32 + e;
            const float s = __uint_as_float(v[e]);
            if (col < n_text) l_t += ex2f(s * sl2 - m_new);
            else if (col < valid) l_i += ex2f(s * sl2 - m_ip);
          }
        }
        inv_t = 1.f / l_t;
        inv_i = (p.n_ip > 0) ? p.ip_scale / l_i : 0.f;
      }

      // pass 2: probabilities -> fp16 P tile in swizzled smem
      float rowsum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_S + lane_base + c * 32, v);
        tmem_ld_wait();
        float pr[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int col = c * 32 + e;
          const float s = __uint_as_float(v[e]);
          float pv = 0.f;
          if (col < n_text) {
            pv = ex2f(s * sl2 - m_new);
            rowsum += pv;
            pv *= inv_t;
          } else if (col < valid) {
            pv = ex2f(s * sl2 - m_ip) * inv_i;
          }
          pr[e] = pv;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int chunk = c * 4 + g;  // 16-byte chunk index along the 128 keys (8 keys each)
          uint4 o;
          o.x = pack_half2(pr[g * 8 + 0], pr[g * 8 + 1]);
          o.y = pack_half2(pr[g * 8 + 2], pr[g * 8 + 3]);
          o.z = pack_half2(pr[g * 8 + 4], pr[g * 8 + 5]);
          o.w = pack_half2(pr[g * 8 + 6], pr[g * 8 + 7]);
          uint8_t* dst = p_row + (chunk >> 3) * ATT_TILE + (((chunk & 7) ^ rx) << 4);
          *reinterpret_cast<uint4*>(dst) = o;
        }
      }
      l_run = l_run * alpha + rowsum;
      m_run = m_new;
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(p_full);

      // fold O_j into the register accumulators
      mbar_wait(o_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_O + lane_base + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) acc[c * 32 + e] = acc[c * 32 + e] * alpha + __uint_as_float(v[e]);
      }
      tc_fence_before();
    }

    const int qrow = q0 + r;
    if (qrow < p.Nq) {
      const float inv = single ? 1.f : 1.f / l_run;
      __half* dst = p.out + ((long long)b * p.Nq + qrow) * p.ldo + head * 64;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        uint4 o;
        o.x = pack_half2(acc[g * 8 + 0] * inv, acc[g * 8 + 1] * inv);
        o.y = pack_half2(acc[g * 8 + 2] * inv, acc[g * 8 + 3] * inv);
        o.z = pack_half2(acc[g * 8 + 4] * inv, acc[g * 8 + 5] * inv);
        o.w = pack_half2(acc[g * 8 + 6] * inv, acc[g * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(dst + g * 8) = o;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem_base);
}


// ================================================================================================================
// Cross-attention with a short key axis (Nk <= 96: 77 text tokens, optionally + the image-prompt tokens).
// One CTA = 128 query rows of one (batch, head); ~73 KiB of shared memory and 128 TMEM columns so that THREE CTAs
// share an SM and a whole SDXL layer (320 / 640 CTAs) runs in one / two waves -- the op is a latency chain
// (TMA -> S MMA -> softmax -> PV MMA -> store), not a throughput problem.  S = Q K^T is a single 128x96 MMA tile; the
// row thread runs the text softmax over columns [0, n_text) and, for the 10 IMAGHarmony layers, an independent softmax
// over the image columns [n_text, Nk), writes [P_t/l_t | scale*P_ip/l_ip] as the fp16 A operand and ONE PV MMA yields
// softmax(q k_t^T) v_t + scale * softmax(q k_ip^T) v_ip  (attention_processor.py:423-450).  O overlays S in TMEM.
// ================================================================================================================
constexpr int AX_THREADS = 192;
constexpr int AX_KV_TILE = 96 * 64 * 2;                                   // 12 KiB: 96 keys x 64 fp16
constexpr int AX_SMEM_TILES = ATT_TILE + 2 * AX_KV_TILE + 2 * ATT_TILE;   // Q | K | V | P (2 key halves)
constexpr int AX_SMEM_BYTES = AX_SMEM_TILES + 128;

__global__ void __launch_bounds__(AX_THREADS, 3) attnx_f16_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                   const __grid_constant__ CUtensorMap tmK,
                                                                   const __grid_constant__ CUtensorMap tmV,
                                                                   const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;
  uint8_t* sK = smem + ATT_TILE;
  uint8_t* sV = sK + AX_KV_TILE;
  uint8_t* sP = sV + AX_KV_TILE;   // 1024-aligned: 16384 + 2 * 12288 = 40960
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AX_SMEM_TILES);
  uint64_t* qk_full = bars + 0;
  uint64_t* v_full = bars + 1;
  uint64_t* s_full = bars + 2;
  uint64_t* p_full = bars + 3;
  uint64_t* o_full = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int head = blockIdx.y;
  const int b = blockIdx.z;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(qk_full, 1);
    mbar_init(v_full, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<128>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(qk_full, ATT_TILE + AX_KV_TILE);
      tma_load_3d(sQ, &tmQ, qk_full, head * 64, q0, b);
      tma_load_3d(sK, &tmK, qk_full, head * 64, 0, b);
      mbar_arrive_expect_tx(v_full, AX_KV_TILE);
      tma_load_3d(sV, &tmV, v_full, head * 64, 0, b);
    }
  } else if (warp == 1) {
    // whole warp, uniform control flow; one elected lane issues (see elect_one() in ptx.cuh)
    constexpr uint32_t idesc_s = umma_idesc_f16(128, 96, false, false);
    constexpr uint32_t idesc_o = umma_idesc_f16(128, 64, false, true);
    mbar_wait(qk_full, 0);
    tc_fence_after();
    const uint64_t q_desc = umma_desc_sw128(smem_u32(sQ));
    const uint64_t k_desc = umma_desc_sw128(smem_u32(sK));
    if (elect_one()) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0);
      umma_commit(s_full);
    }
    __syncwarp();
    mbar_wait(v_full, 0);
    mbar_wait(p_full, 0);
    tc_fence_after();
    const uint64_t p_desc = umma_desc_sw128(smem_u32(sP));
    const uint64_t v_desc = umma_desc_sw128(smem_u32(sV));
    // Two accumulating MMAs over UN-normalised probabilities (the row thread evaluates every exponential once):
    //   O_text (TMEM columns [0, 64))   = P_text V   -- P_text is zero at the image-prompt keys
    //   O_ip   (TMEM columns [64, 128)) = P_ip   V   -- P_ip covers the (<= 64) keys from ip_base on, in the Q buffer
    // both overlay the (already consumed) S columns; the epilogue forms O_text / l_text + scale * O_ip / l_ip.
    const int n_text = p.Nk - p.n_ip;
    const int ksteps_t = (n_text + 15) >> 4;
    const int ip_k0 = n_text >> 4;                       // first 16-key step that holds an image-prompt key
    const int ksteps_i = p.n_ip > 0 ? ((p.Nk + 15) >> 4) - ip_k0 : 0;
    if (elect_one()) {
      for (int kk = 0; kk < ksteps_t; ++kk)
        umma_f16_ss(tmem_base, p_desc + ((kk >> 2) * (ATT_TILE >> 4) + 2 * (kk & 3)), v_desc + kk * (2048 >> 4), idesc_o,
                    kk != 0);
      for (int kk = 0; kk < ksteps_i; ++kk)
        umma_f16_ss(tmem_base + 64, q_desc + 2 * kk, v_desc + (ip_k0 + kk) * (2048 >> 4), idesc_o, kk != 0);
      umma_commit(o_full);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t tS = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint8_t* p_row = sP + r * 128;
    uint8_t* pi_row = sQ + r * 128;      // P_ip reuses the Q tile (free once S has been computed)
    const int rx = r & 7;
    const float sl2 = p.scale_log2;
    const int valid = p.Nk;              // <= 96
    const int n_text = valid - p.n_ip;
    const int ip_chunk0 = (n_text >> 4) << 1;   // first 8-key chunk of the P_ip tile (16-key aligned)

    mbar_wait(s_full, 0);
    tc_fence_after();
    // pass 1: maxima of the two segments
    float mx_t = -INFINITY, mx_i = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tS + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int col = c * 32 + e;
        const float s = __uint_as_float(v[e]);
        if (col < n_text) mx_t = fmaxf(mx_t, s);
        else if (col < valid) mx_i = fmaxf(mx_i, s);
      }
    }
    const float m_t = mx_t * sl2, m_i = mx_i * sl2;
    // pass 2: every exponential ONCE: p = 2^(s * scale - m_segment), un-normalised, into the fp16 P tiles; the row sums
    // normalise the output rows in the epilogue (keys 96..127 of the second half are never read by the MMA)
    float l_t = 0.f, l_i = 0.f;
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tS + c * 32, v);
      tmem_ld_wait();
      float pt[32], pi[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int col = c * 32 + e;
        const float s = __uint_as_float(v[e]);
        const bool is_t = col < n_text, is_i = !is_t && col < valid;
        const float ex = ex2_approx(fmaf(s, sl2, is_t ? -m_t : -m_i));
        pt[e] = is_t ? ex : 0.f;
        pi[e] = is_i ? ex : 0.f;
        l_t += pt[e];
        l_i += pi[e];
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int chunk = c * 4 + g;
        uint4 o;
        o.x = pack_half2(pt[g * 8 + 0], pt[g * 8 + 1]);
        o.y = pack_half2(pt[g * 8 + 2], pt[g * 8 + 3]);
        o.z = pack_half2(pt[g * 8 + 4], pt[g * 8 + 5]);
        o.w = pack_half2(pt[g * 8 + 6], pt[g * 8 + 7]);
        *reinterpret_cast<uint4*>(p_row + (chunk >> 3) * ATT_TILE + (((chunk & 7) ^ rx) << 4)) = o;
        const int ic = chunk - ip_chunk0;          // chunk of the P_ip tile (keys ip_base + 8 ic ...)
        if (p.n_ip > 0 && ic >= 0 && ic < 8) {
          uint4 oi;
          oi.x = pack_half2(pi[g * 8 + 0], pi[g * 8 + 1]);
          oi.y = pack_half2(pi[g * 8 + 2], pi[g * 8 + 3]);
          oi.z = pack_half2(pi[g * 8 + 4], pi[g * 8 + 5]);
          oi.w = pack_half2(pi[g * 8 + 6], pi[g * 8 + 7]);
          *reinterpret_cast<uint4*>(pi_row + ((ic ^ rx) << 4)) = oi;
        }
      }
    }
    const float inv_t = 1.f / l_t;
    const float inv_i = (p.n_ip > 0) ? p.ip_scale / l_i : 0.f;
    tc_fence_before();
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(p_full);

    mbar_wait(o_full, 0);
    tc_fence_after();
    const int qrow = q0 + r;
    __half* dst = p.out + ((long long)b * p.Nq + qrow) * p.ldo + head * 64;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t ov[32], oi[32];
      tmem_ld_32x32b_x32(tS + c * 32, ov);
      if (p.n_ip > 0) tmem_ld_32x32b_x32(tS + 64 + c * 32, oi);
      tmem_ld_wait();
      if (qrow < p.Nq) {
        float y[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          y[e] = __uint_as_float(ov[e]) * inv_t;
          if (p.n_ip > 0) y[e] = fmaf(__uint_as_float(oi[e]), inv_i, y[e]);
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          o.x = pack_half2(y[g * 8 + 0], y[g * 8 + 1]);
          o.y = pack_half2(y[g * 8 + 2], y[g * 8 + 3]);
          o.z = pack_half2(y[g * 8 + 4], y[g * 8 + 5]);
          o.w = pack_half2(y[g * 8 + 6], y[g * 8 + 7]);
          *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<128>(tmem_base);
}

// ================================================================================================================
// v2: ping-pong flash attention.  One CTA = 256 query rows (two 128-row tiles A/B) of one (batch, head); 10 warps:
// warp 0 TMA, warp 1 MMA issuer, warps 2-5 softmax of tile A, warps 6-9 softmax of tile B.  While one tile's
// softmax runs on the CUDA cores / MUFU, the tensor core executes the other tile's S = Q K^T and O += P V.
// O stays in TMEM across KV blocks (accumulating MMA); the running max is only raised when it grew by more than 2^8
// ("lazy rescale"), in which case the row thread rescales its O row in TMEM (tcgen05.ld / tcgen05.st).
// ================================================================================================================
#ifndef IH_A2_POLY
#define IH_A2_POLY 1
#endif
// which of every 8 exponentials are evaluated on the FMA pipe instead of MUFU (IH_A2_POLY of 8)
#if IH_A2_POLY == 0
#define A2_POLY(e) (false)
#elif IH_A2_POLY == 1
#define A2_POLY(e) ((e) == 3)
#elif IH_A2_POLY == 2
#define A2_POLY(e) ((e) == 2 || (e) == 6)
#else
#define A2_POLY(e) ((e) == 2 || (e) == 5 || (e) == 7)
#endif
// Optional clock64 trace of CTA 0 (KV blocks 4..7), compiled in with -DIH_ATTN_TRACE=1 (tools/attn_trace.py).
#ifndef IH_ATTN_TRACE
#define IH_ATTN_TRACE 0
#endif
#if IH_ATTN_TRACE
#define A2_STAMP(cond, slot) do { if (cond) p.trace[slot] = clock64(); } while (0)
#else
#define A2_STAMP(cond, slot) do { } while (0)
#endif
constexpr int A2_THREADS = 576;  // warp 0 TMA, warp 1 MMA, 8 softmax warps per query tile (two threads per row)
constexpr int A2_KS = 3;  // K / V stages
constexpr int A2_SMEM_TILES = ATT_TILE * (2 + 2 * A2_KS) + 4 * ATT_TILE;  // Q_A Q_B | K[KS] | V[KS] | P_A P_B
constexpr int A2_XCH_BYTES = 2 * 2 * 2 * 128 * 4;   // [parity][tile][half][row] fp32: row max / row sum exchange
constexpr int A2_SMEM_BYTES = A2_SMEM_TILES + 256 + A2_XCH_BYTES;

__global__ void __launch_bounds__(A2_THREADS, 1) attn2_f16_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                   const __grid_constant__ CUtensorMap tmK,
                                                                   const __grid_constant__ CUtensorMap tmV,
                                                                   const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;                                  // 2 tiles
  uint8_t* sK = smem + 2 * ATT_TILE;                   // A2_KS stages
  uint8_t* sV = sK + A2_KS * ATT_TILE;                 // A2_KS stages
  uint8_t* sP = sV + A2_KS * ATT_TILE;                 // 2 tiles x 2 key halves
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A2_SMEM_TILES);
  uint64_t* q_full = bars;                  // [2]
  uint64_t* k_full = bars + 2;              // [KS]
  uint64_t* k_empty = k_full + A2_KS;       // [KS]
  uint64_t* v_full = k_empty + A2_KS;       // [KS]
  uint64_t* v_empty = v_full + A2_KS;       // [KS]
  uint64_t* s_full = v_empty + A2_KS;       // [2]
  uint64_t* p_full = s_full + 2;            // [2]
  uint64_t* o_full = p_full + 2;            // [2]
  uint64_t* s_free = o_full + 2;            // [2] all row threads hold their scores of the block in registers
  uint64_t* pv_done = s_free + 2;           // [2] PV MMA of the block has completed: P may be rewritten, O rescaled
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);
  float* xch = reinterpret_cast<float*>(smem + A2_SMEM_TILES + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  int pair = blockIdx.x, jb = 0, je = p.num_kv_blocks, slot = -1, sp = 0;
  if (pair >= p.n_whole) {
    slot = pair - p.n_whole;
    sp = slot / p.split;
    const int part = slot - sp * p.split;
    pair = p.n_whole + sp;
    jb = part * p.num_kv_blocks / p.split;          // balanced contiguous partition of the KV blocks
    je = (part + 1) * p.num_kv_blocks / p.split;
  }
  const int q0 = (pair % p.qpairs) * 256;
  const int head = (pair / p.qpairs) % p.H;
  const int b = pair / (p.qpairs * p.H);
  const int nb = je - jb;
  const bool tileB_active = (q0 + 128) < p.Nq;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int t = 0; t < 2; ++t) {
      mbar_init(&q_full[t], 1);
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 256);
      mbar_init(&s_free[t], 256);
      mbar_init(&pv_done[t], 1);
      mbar_init(&o_full[t], 1);
    }
    for (int s = 0; s < A2_KS; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&q_full[0], ATT_TILE);
      tma_load_3d(sQ, &tmQ, &q_full[0], head * 64, q0, b);
      if (tileB_active) {
        mbar_arrive_expect_tx(&q_full[1], ATT_TILE);
        tma_load_3d(sQ + ATT_TILE, &tmQ, &q_full[1], head * 64, q0 + 128, b);
      }
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nb; ++j) {
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[s], ATT_TILE);
        tma_load_3d(sK + s * ATT_TILE, &tmK, &k_full[s], head * 64, (jb + j) * 128, b);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[s], ATT_TILE);
        tma_load_3d(sV + s * ATT_TILE, &tmV, &v_full[s], head * 64, (jb + j) * 128, b);
        if (++s == A2_KS) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // The whole warp runs the issue loop (uniform control flow); one elected lane issues the MMAs and commits.
    {
      constexpr uint32_t idesc_s = umma_idesc_f16(128, 128, false, false);
      constexpr uint32_t idesc_o = umma_idesc_f16(128, 64, false, true);
      const int ntiles = tileB_active ? 2 : 1;
      for (int t = 0; t < ntiles; ++t) mbar_wait(&q_full[t], 0);
      const uint32_t q_lo = smem_u32(sQ), k_lo = smem_u32(sK), v_lo = smem_u32(sV);
      auto issue_s = [&](int t, int stage) {
        const uint64_t q_desc = umma_desc_sw128(q_lo + t * ATT_TILE);
        const uint64_t k_desc = umma_desc_sw128(k_lo + stage * ATT_TILE);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(tmem_base + t * 128, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0);
          umma_commit(&s_full[t]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int t, int stage, bool accumulate) {
        // P(t) lives in TENSOR memory: fp16 pairs in columns [384 + 64 t, +64) (8 columns per 16-key step), written
        // by the softmax threads with tcgen05.st -- the A operand costs no shared-memory bandwidth.
        // V: +128 descriptor units (2 KiB) per 16 keys of the MN-major tile.
        const uint64_t v_desc = umma_desc_sw128(v_lo + stage * ATT_TILE);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_f16_ts(tmem_base + 256 + t * 64, tmem_base + 384 + t * 64 + 8 * kk, v_desc + kk * (2048 >> 4), idesc_o,
                        (accumulate || kk != 0) ? 1u : 0u);
        }
        __syncwarp();
      };
      auto commit = [&](uint64_t* bar) {
        if (elect_one()) umma_commit(bar);
        __syncwarp();
      };
      // prologue: S(0) for both tiles
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      for (int t = 0; t < ntiles; ++t) issue_s(t, 0);
      commit(&k_empty[0]);
      // Event loop.  The row threads keep their scores in registers, so S(t, j+1) = Q K_{j+1}^T may overwrite the S
      // columns as soon as every thread has LOADED block j (s_free) -- the next scores are ready long before the
      // softmax of block j ends; PV(t, j) follows p_full(t, j) and reads P from its own TMEM columns.  Nothing here
      // blocks: an event is taken only when its operand tile has landed too, so a tile that runs ahead can never
      // starve the other tile's MMAs (and with them the producer's stage recycling).
      auto ready = [&](uint64_t* bar, uint32_t parity) { return __any_sync(0xffffffffu, mbar_test_wait(bar, parity)) != 0; };
      int js[2] = {1, 1};   // next S block to issue
      int jp[2] = {0, 0};   // next PV block to issue
      if (ntiles == 1) js[1] = jp[1] = nb;
      uint32_t idle = 0;
      while (jp[0] < nb || jp[1] < nb) {
        bool progressed = false;
        for (int t = 0; t < ntiles; ++t) {
          if (js[t] < nb && ready(&s_free[t], (js[t] - 1) & 1) && ready(&k_full[js[t] % A2_KS], (js[t] / A2_KS) & 1)) {
            const int j = js[t], s = j % A2_KS;
            tc_fence_after();
            issue_s(t, s);
            ++js[t];
            if (js[t ^ 1] > j) commit(&k_empty[s]);     // the other tile's S of block j was issued earlier
            progressed = true;
          }
          if (jp[t] < nb && ready(&p_full[t], jp[t] & 1) && ready(&v_full[jp[t] % A2_KS], (jp[t] / A2_KS) & 1)) {
            const int j = jp[t], s = j % A2_KS;
            tc_fence_after();
            issue_pv(t, s, j != 0);
            commit(&pv_done[t]);
            if (j == nb - 1) commit(&o_full[t]);
            ++jp[t];
            if (jp[t ^ 1] > j) commit(&v_empty[s]);
            progressed = true;
          }
        }
        if (progressed) idle = 0;
        else if (++idle > IH_SPIN_LIMIT) __trap();
      }
    }
  } else {
    // Two threads per query row: thread (r, hh) owns columns [64 hh, 64 hh + 64) of each 128-key S block, half of the O
    // row, and half of the exponentials.  Four softmax warps per SM sub-partition (instead of two) hide the dependent-
    // issue latency of the exp / max / sum chains, which is what bounded the one-thread-per-row version (ncu: 0.44
    // IPC, 5 cycles between a warp's issues).  The partners (warps w and w + 4: same TMEM lane quarter) meet once per
    // block on a 64-thread named barrier to exchange their half-row maxima.
    const int idx = warp - 2;
    const int t = idx >> 3;            // tile
    const int hh = (idx >> 2) & 1;     // column half
    if (t == 0 || tileB_active) {
      const int q = warp & 3;
      const int r = q * 32 + lane;
      const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
      const uint32_t tS = tmem_base + t * 128 + hh * 64 + lane_base;
      const uint32_t tO = tmem_base + 256 + t * 64 + hh * 32 + lane_base;
      const uint32_t tP = tmem_base + 384 + t * 64 + hh * 32 + lane_base;   // fp16 pairs of my 64 keys
      const float sl2 = p.scale_log2;
      const uint32_t pair_bar = 1 + t * 4 + q;          // named barrier of the two partner warps
      float* x_mine = xch + (t * 2 + hh) * 128 + r;     // + parity * 512
      float* x_peer = xch + (t * 2 + (hh ^ 1)) * 128 + r;
      float m_used = -INFINITY, l_run = 0.f;

      for (int j = 0; j < nb; ++j) {
#if IH_ATTN_TRACE
        const bool trj = p.trace && blockIdx.x == 0 && q == 0 && lane == 0 && hh == 0 && j >= 4 && j < 8;
        const int tb = t * 64 + (j - 4) * 8;
#endif
        A2_STAMP(trj, tb + 0);
        mbar_wait(&s_full[t], j & 1);
        tc_fence_after();
        A2_STAMP(trj, tb + 1);
        const int valid = p.Nk - (jb + j) * 128 - hh * 64;   // valid keys among my 64 columns
        // My 64 scores are read from TMEM exactly ONCE and stay in registers for both the max and the exp pass: TMEM
        // reads run at 64 B/clk/SM, so a second sweep over the 128 x 128 fp32 S tiles of both query tiles would cost
        // more (2 x 2048 cycles per KV block) than all the arithmetic.
        uint32_t va[32], vb[32];
        tmem_ld_32x32b_x32(tS, va);
        tmem_ld_32x32b_x32(tS + 32, vb);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&s_free[t]);   // the S columns may be overwritten with the next block's scores
        A2_STAMP(trj, tb + 2);
        if (valid < 64) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            if (e >= valid) va[e] = 0xff800000u;  // -inf
            if (32 + e >= valid) vb[e] = 0xff800000u;
          }
        }
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          mx0 = fmax3(mx0, __uint_as_float(va[e]), __uint_as_float(va[e + 1]));
          mx1 = fmax3(mx1, __uint_as_float(vb[e]), __uint_as_float(vb[e + 1]));
        }
        const float mx_mine = fmaxf(mx0, mx1);
        x_mine[(j & 1) * 512] = mx_mine;
        A2_STAMP(trj, tb + 3);
        named_bar_sync(pair_bar, 64);
        A2_STAMP(trj, tb + 4);
        const float m_blk = fmaxf(mx_mine, x_peer[(j & 1) * 512]) * sl2;
        if (j == 0) {
          m_used = m_blk;
        } else {
          const bool need = m_blk > m_used + 8.0f;      // same decision in both partners (same m_blk, same m_used)
          if (__any_sync(0xffffffffu, need)) {
            const float m_new = need ? m_blk : m_used;
            const float alpha = ex2_approx(m_used - m_new);
            m_used = m_new;
            l_run *= alpha;
            mbar_wait(&pv_done[t], (j - 1) & 1);        // O must hold every block up to j - 1 before it is rescaled
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {               // my half (32 columns) of the O row, 16 at a time
              uint32_t ov[16];
              tmem_ld_32x32b_x16(tO + c * 16, ov);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 16; ++e) ov[e] = __float_as_uint(__uint_as_float(ov[e]) * alpha);
              tmem_st_32x32b_x16(tO + c * 16, ov);
            }
            tmem_st_wait();
          }
        }
        // p = 2^(s*scale - m) -> fp16 pairs, stored to the tile's P columns in TMEM (the A operand of the PV MMA): my 64
        // keys = columns [32 hh, 32 hh + 32) of the 64.
        float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t pk[16];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t* src = (hf == 0 ? va : vb) + g * 8;
            float pr[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float xarg = fmaf(__uint_as_float(src[e]), sl2, -m_used);
              // a share of the exponentials runs on the FMA pipe instead of MUFU (IH_A2_POLY of every 8)
              pr[e] = A2_POLY(e) ? ex2_poly(xarg) : ex2_approx(xarg);
            }
            sum0 += (pr[0] + pr[1]) + (pr[2] + pr[3]);
            sum1 += (pr[4] + pr[5]) + (pr[6] + pr[7]);
            pk[g * 4 + 0] = pack_half2(pr[0], pr[1]);
            pk[g * 4 + 1] = pack_half2(pr[2], pr[3]);
            pk[g * 4 + 2] = pack_half2(pr[4], pr[5]);
            pk[g * 4 + 3] = pack_half2(pr[6], pr[7]);
          }
          // PV(t, j - 1) must have consumed the previous P before it is overwritten; half a block of exponentials after
          // it was issued it normally has.  (This also keeps p_full at most one phase ahead of the MMA warp.)
          if (hf == 0 && j > 0) mbar_wait(&pv_done[t], (j - 1) & 1);
          tmem_st_32x32b_x16(tP + hf * 16, pk);
        }
        tmem_st_wait();
        l_run += sum0 + sum1;
        A2_STAMP(trj, tb + 5);
        tc_fence_before();
        mbar_arrive(&p_full[t]);
        A2_STAMP(trj, tb + 6);
      }

      // epilogue: the partners add their half-row sums; each writes its 32 of the 64 output columns
      x_mine[(nb & 1) * 512] = l_run;
      named_bar_sync(pair_bar, 64);
      const float l_tot = l_run + x_peer[(nb & 1) * 512];
      mbar_wait(&o_full[t], 0);
      tc_fence_after();
      const int qrow = q0 + t * 128 + r;
      uint32_t ov[32];
      tmem_ld_32x32b_x32(tO, ov);
      tmem_ld_wait();
      if (slot >= 0) {
        // the fp32 O tile goes through this tile's (now idle) P buffer and leaves as ONE 32 KiB bulk copy; 16-byte chunk
        // c of row r sits at chunk position c ^ (r & 15) (bank-conflict-free staging; attn_combine_kernel undoes it)
        const long long wrow = (long long)slot * 256 + t * 128 + r;
        uint8_t* stage = sP + t * 2 * ATT_TILE;
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *reinterpret_cast<uint4*>(stage + r * 256 + (((hh * 8 + g) ^ (r & 15)) << 4)) =
              make_uint4(ov[g * 4 + 0], ov[g * 4 + 1], ov[g * 4 + 2], ov[g * 4 + 3]);
        if (hh == 0) reinterpret_cast<float2*>(p.ws_ml)[wrow] = make_float2(m_used, l_tot);
        fence_proxy_async_smem();
        named_bar_sync(9 + t, 256);
        if (hh == 0 && q == 0 && lane == 0) {
          bulk_store_linear(p.ws_o + ((long long)slot * 256 + t * 128) * 64, stage, 2 * ATT_TILE);
          tma_store_commit();
          tma_store_wait_all();
        }
        if (p.ws_cnt) {
          // In-kernel merge: the LAST of the `split` parts of this query pair to arrive (device-scope counter) combines
          // all partials -- in part order, so the result does not depend on who arrives last -- and writes the
          // output rows; no second launch.  M = max_i m_i, out = sum_i O_i 2^(m_i - M) / sum_i l_i 2^(m_i - M).
          volatile int* last_flag = reinterpret_cast<volatile int*>(tmem_slot + 2);
          const int nsoft = tileB_active ? 512 : 256;
          __threadfence();                       // my (m, l) row / the elected thread's completed bulk store
          named_bar_sync(11, nsoft);
          if (idx == 0 && lane == 0) {
            __threadfence();
            const int old = atomicAdd(&p.ws_cnt[sp], 1);
            const int last = (old == p.split - 1) ? 1 : 0;
            if (last) p.ws_cnt[sp] = 0;          // ready for the next launch
            *last_flag = last;
          }
          named_bar_sync(11, nsoft);
          if (*last_flag && qrow < p.Nq) {
            __threadfence();
            const int row = t * 128 + r;
            const float2* ml = reinterpret_cast<const float2*>(p.ws_ml);
            float mmax = -INFINITY;
            for (int i = 0; i < p.split; ++i) mmax = fmaxf(mmax, __ldcg(&ml[((long long)sp * p.split + i) * 256 + row]).x);
            float acc[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) acc[e] = 0.f;
            float l = 0.f;
            for (int i = 0; i < p.split; ++i) {
              const long long wrow = ((long long)sp * p.split + i) * 256 + row;
              const float2 v = __ldcg(&ml[wrow]);
              const float w = ex2f(v.x - mmax);
              l = fmaf(v.y, w, l);
              const float4* src = reinterpret_cast<const float4*>(p.ws_o + wrow * 64);
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 o = __ldcg(&src[(hh * 8 + g) ^ (r & 15)]);   // chunk permutation of the staging above
                acc[g * 4 + 0] = fmaf(o.x, w, acc[g * 4 + 0]);
                acc[g * 4 + 1] = fmaf(o.y, w, acc[g * 4 + 1]);
                acc[g * 4 + 2] = fmaf(o.z, w, acc[g * 4 + 2]);
                acc[g * 4 + 3] = fmaf(o.w, w, acc[g * 4 + 3]);
              }
            }
            const float inv = 1.f / l;
            __half* dst = p.out + ((long long)b * p.Nq + qrow) * p.ldo + head * 64 + hh * 32;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 o;
              o.x = pack_half2(acc[g * 8 + 0] * inv, acc[g * 8 + 1] * inv);
              o.y = pack_half2(acc[g * 8 + 2] * inv, acc[g * 8 + 3] * inv);
              o.z = pack_half2(acc[g * 8 + 4] * inv, acc[g * 8 + 5] * inv);
              o.w = pack_half2(acc[g * 8 + 6] * inv, acc[g * 8 + 7] * inv);
              *reinterpret_cast<uint4*>(dst + g * 8) = o;
            }
          }
        }
      } else if (qrow < p.Nq) {
        const float inv = 1.f / l_tot;
        __half* dst = p.out + ((long long)b * p.Nq + qrow) * p.ldo + head * 64 + hh * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          o.x = pack_half2(__uint_as_float(ov[g * 8 + 0]) * inv, __uint_as_float(ov[g * 8 + 1]) * inv);
          o.y = pack_half2(__uint_as_float(ov[g * 8 + 2]) * inv, __uint_as_float(ov[g * 8 + 3]) * inv);
          o.z = pack_half2(__uint_as_float(ov[g * 8 + 4]) * inv, __uint_as_float(ov[g * 8 + 5]) * inv);
          o.w = pack_half2(__uint_as_float(ov[g * 8 + 6]) * inv, __uint_as_float(ov[g * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + g * 8) = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// Merge the `split` KV parts of the query pairs that attn2_f16_kernel processed piecewise:
//   M = max_i m_i,  out = sum_i O_i 2^(m_i - M) / sum_i l_i 2^(m_i - M)   (fixed order: deterministic).
// grid = (#split pairs * 4), block = 256: 64 rows per block, 4 threads (16 columns each) per row.
__global__ void __launch_bounds__(256) attn_combine_kernel(const AttnParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int sp = blockIdx.x >> 2;
  const int row = (blockIdx.x & 3) * 64 + (threadIdx.x >> 2);
  const int c0 = (threadIdx.x & 3) * 16;
  const int pair = p.n_whole + sp;
  const int qrow = (pair % p.qpairs) * 256 + row;
  if (qrow >= p.Nq) return;
  const int head = (pair / p.qpairs) % p.H;
  const int b = pair / (p.qpairs * p.H);
  const float2* ml = reinterpret_cast<const float2*>(p.ws_ml);
  float mmax = -INFINITY;
  for (int i = 0; i < p.split; ++i) mmax = fmaxf(mmax, ml[((long long)sp * p.split + i) * 256 + row].x);
  float acc[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) acc[e] = 0.f;
  float l = 0.f;
  for (int i = 0; i < p.split; ++i) {
    const long long wrow = ((long long)sp * p.split + i) * 256 + row;
    const float2 v = ml[wrow];
    const float w = ex2f(v.x - mmax);
    l = fmaf(v.y, w, l);
    const float4* src = reinterpret_cast<const float4*>(p.ws_o + wrow * 64);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float4 o = src[((c0 >> 2) + g) ^ (row & 15)];   // chunk permutation of the producer's staging
      acc[g * 4 + 0] = fmaf(o.x, w, acc[g * 4 + 0]);
      acc[g * 4 + 1] = fmaf(o.y, w, acc[g * 4 + 1]);
      acc[g * 4 + 2] = fmaf(o.z, w, acc[g * 4 + 2]);
      acc[g * 4 + 3] = fmaf(o.w, w, acc[g * 4 + 3]);
    }
  }
  const float inv = 1.f / l;
  __half* dst = p.out + ((long long)b * p.Nq + qrow) * p.ldo + head * 64 + c0;
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    uint4 o;
    o.x = pack_half2(acc[g * 8 + 0] * inv, acc[g * 8 + 1] * inv);
    o.y = pack_half2(acc[g * 8 + 2] * inv, acc[g * 8 + 3] * inv);
    o.z = pack_half2(acc[g * 8 + 4] * inv, acc[g * 8 + 5] * inv);
    o.w = pack_half2(acc[g * 8 + 6] * inv, acc[g * 8 + 7] * inv);
    *reinterpret_cast<uint4*>(dst + g * 8) = o;
  }
}

// KV-split plan of the ping-pong kernel: with P query pairs on G SMs (one CTA per SM), the last P mod G pairs would
// run as a nearly empty extra wave; they are cut into `split` KV parts each so that (P mod G) * split <= G CTAs share
// that wave.  Returns split (1 = none) and the number of pairs processed whole.
constexpr long long A2_CNT_BYTES = 4096;   // 1024 arrival counters at the start of the attention workspace
static long long* g_attn_trace = nullptr;
static int g_split_policy = 0;   // 0: cost model, 1: split whenever a split plan exists (tests)

static void attn2_plan(int pairs, int nb, int* n_whole, int* split) {
  const int G = num_sms();
  const int rem = pairs % G;
  *n_whole = pairs;
  *split = 1;
  if (rem == 0 || nb < 2) return;
  int s = G / rem;
  if (s > nb) s = nb;
  if (s < 2) return;
  // measured (tools/attn_probe.py): a KV iteration takes ~2.7 us with every SM busy but ~2.0 us in a sparsely filled
  // last wave; a part pays ~11 us on top of its iterations (CTA set-up, partial epilogue, merge kernel)
  const float t_whole = 2.0f * nb, t_split = 2.7f * ((nb + s - 1) / s) + 11.0f;
  if (t_split >= t_whole && g_split_policy == 0) return;
  *n_whole = pairs - rem;
  *split = s;
}

}  // namespace ih

using namespace ih;

extern "C" void ih_attention_set_split_policy(int policy) { g_split_policy = policy; }
extern "C" void ih_attention_set_trace(void* device_buffer) { g_attn_trace = (long long*)device_buffer; }

extern "C" long long ih_attention_workspace_bytes(int B, int H, int Nq, int Nk, int n_ip) {
  if (B <= 0 || H <= 0 || Nq <= 0 || Nk <= 96 || n_ip != 0) return 0;
  int n_whole, split;
  const int pairs = B * H * ((Nq + 255) / 256);
  attn2_plan(pairs, (Nk + 127) / 128, &n_whole, &split);
  // arrival counters (one per split pair, FIXED place at the start of the buffer so that calls with different shapes
  // share them: the caller provides zeroed memory once, every launch leaves them at zero), then partial O rows + (m, l)
  if (pairs == n_whole) return 0;
  return A2_CNT_BYTES + (long long)(pairs - n_whole) * split * 256 * (64 + 2) * (long long)sizeof(float);
}

extern "C" int ih_attention_f16(const void* q, long long ldq, const void* k, long long ldk, const void* v,
                                long long ldv, void* out, long long ldo, int B, int H, int Nq, int Nk, int n_ip,
                                float ip_scale, void* stream) {
  return ih_attention_ws_f16(q, ldq, k, ldk, v, ldv, out, ldo, B, H, Nq, Nk, n_ip, ip_scale, nullptr, 0, stream);
}

extern "C" int ih_attention_ws_f16(const void* q, long long ldq, const void* k, long long ldk, const void* v,
                                   long long ldv, void* out, long long ldo, int B, int H, int Nq, int Nk, int n_ip,
                                   float ip_scale, void* workspace, long long workspace_bytes, void* stream) {
  IH_CHECK(q && k && v && out, IH_ERR_ARG, "ih_attention_f16: null pointer");
  IH_CHECK(B > 0 && H > 0 && Nq > 0 && Nk > 0, IH_ERR_SHAPE, "ih_attention_f16: bad shape");
  IH_CHECK(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, IH_ERR_ALIGN,
           "ih_attention_f16: row strides must be multiples of 8 elements");
  IH_CHECK(n_ip >= 0 && n_ip < Nk, IH_ERR_ARG, "ih_attention_f16: n_ip out of range");
  IH_CHECK(n_ip == 0 || Nk <= 128, IH_ERR_SHAPE, "ih_attention_f16: decoupled IP path needs Nk <= 128");

  CUtensorMap tq, tk, tv;
  const uint32_t box[3] = {64u, 128u, 1u};
  {
    const uint64_t dims[3] = {(uint64_t)H * 64, (uint64_t)Nq, (uint64_t)B};
    const uint64_t str[2] = {(uint64_t)ldq * 2, (uint64_t)Nq * ldq * 2};
    int rc = get_tmap_f16(&tq, q, 3, dims, str, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)H * 64, (uint64_t)Nk, (uint64_t)B};
    const uint64_t str[2] = {(uint64_t)ldk * 2, (uint64_t)Nk * ldk * 2};
    int rc = get_tmap_f16(&tk, k, 3, dims, str, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)H * 64, (uint64_t)Nk, (uint64_t)B};
    const uint64_t str[2] = {(uint64_t)ldv * 2, (uint64_t)Nk * ldv * 2};
    int rc = get_tmap_f16(&tv, v, 3, dims, str, box);
    if (rc) return rc;
  }
  AttnParams p{};
  p.Nq = Nq;
  p.Nk = Nk;
  p.n_ip = n_ip;
  p.num_kv_blocks = (Nk + 127) / 128;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  p.ip_scale = ip_scale;
  p.out = (__half*)out;
  p.ldo = ldo;
  p.trace = g_attn_trace;

  static bool configured = false;
  if (!configured) {
    IH_CUDA(cudaFuncSetAttribute(attn_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
    IH_CUDA(cudaFuncSetAttribute(attn_f16_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared));
    IH_CUDA(cudaFuncSetAttribute(attn2_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A2_SMEM_BYTES));
    configured = true;
  }
  if (Nk <= 96) {
    static bool cfg_x = false;
    if (!cfg_x) {
      IH_CUDA(cudaFuncSetAttribute(attnx_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AX_SMEM_BYTES));
      cfg_x = true;
    }
    CUtensorMap tkx, tvx;
    const uint32_t boxkv[3] = {64u, 96u, 1u};
    {
      const uint64_t dims[3] = {(uint64_t)H * 64, (uint64_t)Nk, (uint64_t)B};
      const uint64_t strk[2] = {(uint64_t)ldk * 2, (uint64_t)Nk * ldk * 2};
      const uint64_t strv[2] = {(uint64_t)ldv * 2, (uint64_t)Nk * ldv * 2};
      int rc = get_tmap_f16(&tkx, k, 3, dims, strk, boxkv);
      if (rc) return rc;
      rc = get_tmap_f16(&tvx, v, 3, dims, strv, boxkv);
      if (rc) return rc;
    }
    dim3 gridx((Nq + 127) / 128, H, B);
    IH_CUDA(launch_kernel(attnx_f16_kernel, gridx, dim3(AX_THREADS), (size_t)AX_SMEM_BYTES, (cudaStream_t)stream, tq, tkx,
                          tvx, p));
    return 0;
  }
  if (n_ip == 0) {
    p.H = H;
    p.qpairs = (Nq + 255) / 256;
    const int pairs = B * H * p.qpairs;
    attn2_plan(pairs, p.num_kv_blocks, &p.n_whole, &p.split);
    const long long slots = (long long)(pairs - p.n_whole) * p.split;
    const long long need = A2_CNT_BYTES + slots * 256 * 66 * (long long)sizeof(float);
    if (slots > 0 && (!workspace || workspace_bytes < need)) {
      p.n_whole = pairs;   // no (or too small a) workspace: every pair runs whole
      p.split = 1;
    }
    const int n_split_ctas = (pairs - p.n_whole) * p.split;
    p.ws_o = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + A2_CNT_BYTES);
    p.ws_ml = p.ws_o + (long long)n_split_ctas * 256 * 64;
    // IH_ATTN_FUSED_MERGE=1: the last part of a pair to finish merges all parts inside attn2_f16_kernel instead of the
    // attn_combine_kernel launch.  Measured on B200 (profiles/r2_notes.md): correct but SLOWER -- one SM pulls the 0.5 MB
    // of partials of a pair through its own L2 port (self-attention family 3.36 -> 5.28 ms per step) while the merge
    // kernel spreads a pair over four blocks -- so the separate launch stays the default.
    static const bool fused_merge = [] {
      const char* e = getenv("IH_ATTN_FUSED_MERGE");
      return e && e[0] == '1';
    }();
    p.ws_cnt = (fused_merge && n_split_ctas > 0 && pairs - p.n_whole <= A2_CNT_BYTES / (int)sizeof(int))
                   ? reinterpret_cast<int*>(workspace) : nullptr;
    IH_CUDA(launch_kernel(attn2_f16_kernel, dim3(p.n_whole + n_split_ctas), dim3(A2_THREADS), (size_t)(A2_SMEM_BYTES),
                          (cudaStream_t)stream, tq, tk, tv, p));
    if (n_split_ctas > 0 && !p.ws_cnt)
      IH_CUDA(launch_kernel(attn_combine_kernel, dim3((pairs - p.n_whole) * 4), dim3(256), (size_t)0,
                            (cudaStream_t)stream, p));
    return 0;
  }
  dim3 grid((Nq + 127) / 128, H, B);
  IH_CUDA(launch_kernel(attn_f16_kernel, dim3(grid), dim3(ATT_THREADS), (size_t)(ATT_SMEM_BYTES), (cudaStream_t)stream, tq, tk, tv, p));
  return 0;
}
