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

}  // namespace ih

using namespace ih;

extern "C" int ih_attention_f16(const void* q, long long ldq, const void* k, long long ldk, const void* v,
                                long long ldv, void* out, long long ldo, int B, int H, int Nq, int Nk, int n_ip,
                                float ip_scale, void* stream) {
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

  static bool configured = false;
  if (!configured) {
    IH_CUDA(cudaFuncSetAttribute(attn_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
    configured = true;
  }
  dim3 grid((Nq + 127) / 128, H, B);
  attn_f16_kernel<<<grid, ATT_THREADS, ATT_SMEM_BYTES, (cudaStream_t)stream>>>(tq, tk, tv, p);
  IH_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
