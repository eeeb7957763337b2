// Fused cross-attention front half for sm_100a:  O = CrossAttn( LayerNorm(h) W_q^T , K, V )  in ONE kernel.
//
// Replaces, for every attn2 of the SDXL UNet (attention_processor.py:396 `query = attn.to_q(hidden_states)`, :423-425
// text SDPA, :440-442 image SDPA, :450 `hidden + scale * ip_hidden`; AttnProcessor2_0 :292-314 for plain layers), the
// sequence  q-projection GEMM -> q round trip through HBM -> short-key attention kernel.
//
// One CTA = 128 query rows x one group of 4 heads (256 q columns):
//   phase 1  tcgen05 GEMM  acc[128 x 256] = h_tile[128 x C] . Wq_group[256 x C]^T  (TMA-fed 4-stage ring, fp32 in TMEM);
//            the block's LayerNorm (norm2) is folded in: Wq is gamma-scaled and row-centred, the epilogue applies
//            rstd[row] * acc + c[col] (see gemm.cu) -- so the kernel consumes the RAW residual stream.
//   phase 2  per head (two heads in flight, one per 4-warp slot): q_h -> fp16 swizzled smem tile -> S = q_h K_h^T
//            (128 x 96, Nk <= 96 keys) -> row softmax over the text columns and, independently, over the image-prompt
//            columns -> P = [P_t / l_t | ip_scale * P_ip / l_ip] (fp16 smem) -> O_h = P V_h (TMEM, overlays S) -> global.
//            K_h / V_h tiles (step-invariant, precomputed once per generate) are TMA-loaded into ring stages as soon as
//            the main loop has released them; the q / P tiles live in the two stages that are released last.
// TMEM: columns [0,256) accumulator, [256,384) slot 0 S/O, [384,512) slot 1 S/O.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/ih_api.h"
#include "host_util.h"
#include "ptx.cuh"

namespace ih {

constexpr int QX_THREADS = 320;  // warp 0 TMA, warp 1 MMA (main loop), warps 2-5 slot 0, warps 6-9 slot 1
constexpr int QX_STAGES = 4;
constexpr int QX_BM = 128, QX_BN = 256, QX_BK = 64;
constexpr int QX_A_BYTES = QX_BM * QX_BK * 2;              // 16 KiB
constexpr int QX_B_BYTES = QX_BN * QX_BK * 2;              // 32 KiB
constexpr int QX_STAGE_BYTES = QX_A_BYTES + QX_B_BYTES;    // 48 KiB
constexpr int QX_KV_TILE = 96 * 64 * 2;                    // 12 KiB: 96 keys x 64 fp16
constexpr int QX_Q_TILE = 128 * 64 * 2;                    // 16 KiB
constexpr int QX_SMEM_BYTES = QX_STAGES * QX_STAGE_BYTES + 256 + 1024;
static_assert(4 * QX_KV_TILE == QX_STAGE_BYTES, "K/V of two heads fill one ring stage");
static_assert(3 * QX_Q_TILE == QX_STAGE_BYTES, "q tile + two P halves fill one ring stage");

struct QxParams {
  int M, C, num_kb;         // rows of h, channels (= heads * 64), k-blocks of the projection
  int H, Nq, Nk, n_ip;
  float scale_log2, ip_scale;
  const __half* bias;       // [C] or nullptr (LN fold: W beta; plain: none)
  const float* ln_stats;    // [ln_slabs, M, 2] or nullptr
  int ln_slabs;
  float ln_inv_c, ln_eps;
  __half* out;
  long long ldo;
};

__global__ void __launch_bounds__(QX_THREADS, 1) qxattn_f16_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                    const __grid_constant__ CUtensorMap tmW,
                                                                    const __grid_constant__ CUtensorMap tmK,
                                                                    const __grid_constant__ CUtensorMap tmV,
                                                                    const QxParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + QX_STAGES * QX_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + QX_STAGES;
  uint64_t* acc_full = empty_bar + QX_STAGES;   // [1]
  uint64_t* kv_full = acc_full + 1;             // [2] K/V of heads {0,1} / {2,3} landed
  uint64_t* s_full = kv_full + 2;               // [2] per slot
  uint64_t* o_full = s_full + 2;                // [2] per slot
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int group = blockIdx.x;                 // 4-head group
  const int m0 = blockIdx.y * QX_BM;
  const int b = m0 / p.Nq;                      // Nq % 128 == 0: a tile never straddles two images
  const int head0 = group * 4;
  const int hg = min(4, p.H - head0);           // heads in this group (ragged last group: H = 10)
  const int nkb = p.num_kb;
  // ring stages reused by phase 2 (virtual k-blocks nkb .. nkb+3)
  const int st_kv0 = nkb % QX_STAGES, st_kv1 = (nkb + 1) % QX_STAGES;
  const int st_qp0 = (nkb + 2) % QX_STAGES, st_qp1 = (nkb + 3) % QX_STAGES;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int s = 0; s < QX_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&o_full[s], 1);
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
    // ------------------------------ TMA producer (whole warp; one elected lane issues) ------------------------------
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sA = smem + stage * QX_STAGE_BYTES;
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[stage], QX_STAGE_BYTES);
          tma_load_2d(sA, &tmA, &full_bar[stage], kb * QX_BK, m0);
          tma_load_2d(sA + QX_A_BYTES, &tmW, &full_bar[stage], kb * QX_BK, head0 * 64);
        }
        __syncwarp();
        if (++stage == QX_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      // K/V of this group's heads go into the next two ring slots as soon as the main loop has released them
      for (int pair = 0; pair < 2; ++pair) {
        const int nh = min(2, hg - 2 * pair);
        if (nh <= 0) break;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* base = smem + stage * QX_STAGE_BYTES;
        if (elect_one()) {
          mbar_arrive_expect_tx(&kv_full[pair], nh * 2 * QX_KV_TILE);
          for (int i = 0; i < nh; ++i) {
            const int head = head0 + 2 * pair + i;
            tma_load_3d(base + (2 * i) * QX_KV_TILE, &tmK, &kv_full[pair], head * 64, 0, b);
            tma_load_3d(base + (2 * i + 1) * QX_KV_TILE, &tmV, &kv_full[pair], head * 64, 0, b);
          }
        }
        __syncwarp();
        if (++stage == QX_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (projection main loop; whole warp, elected lane issues) ----------
    {
      constexpr uint32_t idesc = umma_idesc_f16(QX_BM, QX_BN, false, false);
      const uint32_t smem_lo = smem_u32(smem);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_lo + stage * QX_STAGE_BYTES;
        const uint64_t a_desc = umma_desc_sw128(a_addr);
        const uint64_t b_desc = umma_desc_sw128(a_addr + QX_A_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < QX_BK / 16; ++k)
            umma_f16_ss(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == QX_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) umma_commit(acc_full);
    }
    __syncwarp();
  } else {
    // ------------------------------ attention slots (one query row per thread) ------------------------------
    const int slot = (warp - 2) >> 2;
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;
    const bool issuer = (q == 0);                // this warp issues the slot's attention MMAs (elected lane)
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tAcc = tmem_base + lane_base;
    const uint32_t tS_mma = tmem_base + 256 + slot * 128;
    const uint32_t tS = tS_mma + lane_base;
    uint8_t* qp = smem + (slot == 0 ? st_qp0 : st_qp1) * QX_STAGE_BYTES;
    uint8_t* sQ = qp;                            // [128 x 64] fp16, SWIZZLE_128B K-major
    uint8_t* sP = qp + QX_Q_TILE;                // two key halves of [128 x 64]
    uint8_t* q_row = sQ + r * 128;
    uint8_t* p_row = sP + r * 128;
    const int rx = r & 7;
    const float sl2 = p.scale_log2;
    const int valid = p.Nk;                      // <= 96
    const int n_text = valid - p.n_ip;
    const long long orow = (long long)m0 + r;

    float ln_rstd = 1.f;
    if (p.ln_stats && orow < p.M) {              // hidden behind the projection main loop
      const float2* st = reinterpret_cast<const float2*>(p.ln_stats) + orow;
      float ssum = 0.f, ssq = 0.f;
      for (int i0 = 0; i0 < p.ln_slabs; i0 += 10) {
        float2 t[10];
#pragma unroll
        for (int i = 0; i < 10; ++i)
          t[i] = (i0 + i < p.ln_slabs) ? __ldg(st + (long long)(i0 + i) * p.M) : make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 10; ++i) {
          ssum += t[i].x;
          ssq += t[i].y;
        }
      }
      const float mean = ssum * p.ln_inv_c;
      ln_rstd = rsqrtf(fmaxf(ssq * p.ln_inv_c - mean * mean, 0.f) + p.ln_eps);
    }

    constexpr uint32_t idesc_s = umma_idesc_f16(128, 96, false, false);
    constexpr uint32_t idesc_o = umma_idesc_f16(128, 64, false, true);
    mbar_wait(acc_full, 0);
    tc_fence_after();

    for (int i = 0; i < 2; ++i) {
      const int hd = slot + 2 * i;               // head within the group (uniform over the slot)
      if (hd >= hg) break;
      const uint8_t* kv_base = smem + ((hd >> 1) == 0 ? st_kv0 : st_kv1) * QX_STAGE_BYTES + (hd & 1) * 2 * QX_KV_TILE;
      // ---- a. q_h = rstd * acc + c  -> fp16 A-operand tile
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tAcc + hd * 64 + c * 32, v);
        uint4 bv[4];
#pragma unroll
        for (int g = 0; g < 4; ++g)
          bv[g] = p.bias ? __ldg(reinterpret_cast<const uint4*>(p.bias + (head0 + hd) * 64 + c * 32 + g * 8))
                         : make_uint4(0u, 0u, 0u, 0u);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t bw[4] = {bv[g].x, bv[g].y, bv[g].z, bv[g].w};
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = unpack_half2(bw[e]);
            o[e] = pack_half2(fmaf(ln_rstd, __uint_as_float(v[g * 8 + 2 * e]), f.x),
                              fmaf(ln_rstd, __uint_as_float(v[g * 8 + 2 * e + 1]), f.y));
          }
          *reinterpret_cast<uint4*>(q_row + (((c * 4 + g) ^ rx) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      named_bar_sync(1 + slot, 128);
      if (issuer) {
        mbar_wait(&kv_full[hd >> 1], 0);
        tc_fence_after();
        const uint64_t q_desc = umma_desc_sw128(smem_u32(sQ));
        const uint64_t k_desc = umma_desc_sw128(smem_u32(kv_base));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss(tS_mma, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0);
          umma_commit(&s_full[slot]);
        }
        __syncwarp();
      }
      mbar_wait(&s_full[slot], i & 1);
      tc_fence_after();

      // ---- b. two independent softmaxes over [0, n_text) and [n_text, Nk)
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
      float l_t = 0.f, l_i = 0.f;
#pragma unroll 1
      for (int c = 0; c < 3; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tS + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int col = c * 32 + e;
          const float s = __uint_as_float(v[e]);
          if (col < n_text) l_t += ex2_approx(fmaf(s, sl2, -m_t));
          else if (col < valid) l_i += ex2_approx(fmaf(s, sl2, -m_i));
        }
      }
      const float inv_t = 1.f / l_t;
      const float inv_i = (p.n_ip > 0) ? p.ip_scale / l_i : 0.f;
#pragma unroll 1
      for (int c = 0; c < 3; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tS + c * 32, v);
        tmem_ld_wait();
        float pr[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int col = c * 32 + e;
          const float s = __uint_as_float(v[e]);
          float pv = 0.f;
          if (col < n_text) pv = ex2_approx(fmaf(s, sl2, -m_t)) * inv_t;
          else if (col < valid) pv = ex2_approx(fmaf(s, sl2, -m_i)) * inv_i;
          pr[e] = pv;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int chunk = c * 4 + g;
          uint4 o;
          o.x = pack_half2(pr[g * 8 + 0], pr[g * 8 + 1]);
          o.y = pack_half2(pr[g * 8 + 2], pr[g * 8 + 3]);
          o.z = pack_half2(pr[g * 8 + 4], pr[g * 8 + 5]);
          o.w = pack_half2(pr[g * 8 + 6], pr[g * 8 + 7]);
          *reinterpret_cast<uint4*>(p_row + (chunk >> 3) * QX_Q_TILE + (((chunk & 7) ^ rx) << 4)) = o;
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      named_bar_sync(1 + slot, 128);
      if (issuer) {
        tc_fence_after();
        const uint64_t p_desc = umma_desc_sw128(smem_u32(sP));
        const uint64_t v_desc = umma_desc_sw128(smem_u32(kv_base + QX_KV_TILE));
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 6; ++kk)   // O overlays the (already consumed) S columns
            umma_f16_ss(tS_mma, p_desc + ((kk >> 2) * (QX_Q_TILE >> 4) + 2 * (kk & 3)), v_desc + kk * (2048 >> 4),
                        idesc_o, kk != 0);
          umma_commit(&o_full[slot]);
        }
        __syncwarp();
      }
      mbar_wait(&o_full[slot], i & 1);
      tc_fence_after();

      // ---- c. O_h -> global (128 contiguous bytes per row)
      __half* dst = p.out + orow * p.ldo + (head0 + hd) * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t ov[32];
        tmem_ld_32x32b_x32(tS + c * 32, ov);
        tmem_ld_wait();
        if (orow < p.M) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 o;
            o.x = pack_half2(__uint_as_float(ov[g * 8 + 0]), __uint_as_float(ov[g * 8 + 1]));
            o.y = pack_half2(__uint_as_float(ov[g * 8 + 2]), __uint_as_float(ov[g * 8 + 3]));
            o.z = pack_half2(__uint_as_float(ov[g * 8 + 4]), __uint_as_float(ov[g * 8 + 5]));
            o.w = pack_half2(__uint_as_float(ov[g * 8 + 6]), __uint_as_float(ov[g * 8 + 7]));
            *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = o;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

}  // namespace ih

using namespace ih;

extern "C" int ih_xattn_q_fused_f16(const void* h, long long ldh, const void* wq, const void* bias,
                                    const void* ln_stats, int ln_slabs, float ln_eps, const void* k, long long ldk,
                                    const void* v, long long ldv, void* out, long long ldo, int B, int H, int Nq, int Nk,
                                    int n_ip, float ip_scale, int K, void* stream) {
  IH_CHECK(h && wq && k && v && out, IH_ERR_ARG, "ih_xattn_q_fused_f16: null pointer");
  IH_CHECK(B > 0 && H > 0 && Nq > 0 && Nk > 0 && K > 0, IH_ERR_SHAPE, "ih_xattn_q_fused_f16: bad shape");
  IH_CHECK(Nq % 128 == 0, IH_ERR_SHAPE, "ih_xattn_q_fused_f16: Nq must be a multiple of 128 (use ih_gemm_f16 + ih_attention_f16)");
  IH_CHECK(Nk <= 96, IH_ERR_SHAPE, "ih_xattn_q_fused_f16: Nk must be <= 96 (use ih_gemm_f16 + ih_attention_f16)");
  IH_CHECK(n_ip >= 0 && n_ip < Nk, IH_ERR_ARG, "ih_xattn_q_fused_f16: n_ip out of range");
  IH_CHECK(K % 8 == 0 && ldh % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, IH_ERR_ALIGN,
           "ih_xattn_q_fused_f16: K and row strides must be multiples of 8 elements");
  IH_CHECK(!ln_stats || ln_slabs > 0, IH_ERR_ARG, "ih_xattn_q_fused_f16: ln_stats needs ln_slabs > 0");
  const int C = H * 64;
  const int M = B * Nq;

  CUtensorMap ta, tw, tk, tv;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    const uint64_t str[1] = {(uint64_t)ldh * 2};
    const uint32_t box[2] = {(uint32_t)QX_BK, (uint32_t)QX_BM};
    int rc = get_tmap_f16(&ta, h, 2, dims, str, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)C};
    const uint64_t str[1] = {(uint64_t)K * 2};
    const uint32_t box[2] = {(uint32_t)QX_BK, (uint32_t)QX_BN};
    int rc = get_tmap_f16(&tw, wq, 2, dims, str, box);
    if (rc) return rc;
  }
  {
    const uint32_t box[3] = {64u, 96u, 1u};
    const uint64_t dims[3] = {(uint64_t)C, (uint64_t)Nk, (uint64_t)B};
    const uint64_t strk[2] = {(uint64_t)ldk * 2, (uint64_t)Nk * ldk * 2};
    const uint64_t strv[2] = {(uint64_t)ldv * 2, (uint64_t)Nk * ldv * 2};
    int rc = get_tmap_f16(&tk, k, 3, dims, strk, box);
    if (rc) return rc;
    rc = get_tmap_f16(&tv, v, 3, dims, strv, box);
    if (rc) return rc;
  }
  QxParams p{};
  p.M = M;
  p.C = C;
  p.num_kb = (K + QX_BK - 1) / QX_BK;
  p.H = H;
  p.Nq = Nq;
  p.Nk = Nk;
  p.n_ip = n_ip;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  p.ip_scale = ip_scale;
  p.bias = (const __half*)bias;
  p.ln_stats = (const float*)ln_stats;
  p.ln_slabs = ln_slabs;
  p.ln_inv_c = 1.0f / (float)K;
  p.ln_eps = ln_eps;
  p.out = (__half*)out;
  p.ldo = ldo;

  static bool configured = false;
  if (!configured) {
    IH_CUDA(cudaFuncSetAttribute(qxattn_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, QX_SMEM_BYTES));
    configured = true;
  }
  dim3 grid((H + 3) / 4, M / QX_BM);
  IH_CUDA(launch_kernel(qxattn_f16_kernel, grid, dim3(QX_THREADS), (size_t)QX_SMEM_BYTES, (cudaStream_t)stream, ta, tw,
                        tk, tv, p));
  return 0;
}
