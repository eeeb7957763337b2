// GroupNorm (NHWC, fp32/double statistics, optional SiLU, optional two-source channel concat) and LayerNorm.
// HBM-bound: every element is read twice (statistics + apply) and written once with 16-byte vector accesses.
// Replaces nn.GroupNorm / nn.SiLU / nn.LayerNorm of the diffusers UNet the reference drives
// (custom_pipelines.py:338-345); LayerNorm also serves ImageProjModel.norm (ip_adapter.py:47) and Resampler norms.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/ih_api.h"
#include "host_util.h"
#include "ptx.cuh"

namespace ih {

__device__ __forceinline__ void load8(const __half* p, float (&x)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = unpack_half2(w[e]);
    x[2 * e] = f.x;
    x[2 * e + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(__half* p, const float (&x)[8]) {
  uint4 o;
  o.x = pack_half2(x[0], x[1]);
  o.y = pack_half2(x[2], x[3]);
  o.z = pack_half2(x[4], x[5]);
  o.w = pack_half2(x[6], x[7]);
  *reinterpret_cast<uint4*>(p) = o;
}

// ---------------------------------------------------------------------------------------------------------------
// GroupNorm statistics: per (batch, group) sum and sum of squares, accumulated in double.
// block = CV * R threads (CV = C/8 channel vectors, R row lanes); grid = (blocks_per_image, B).
// ---------------------------------------------------------------------------------------------------------------
__global__ void gn_stats_kernel(const __half* __restrict__ x0, int C0, const __half* __restrict__ x1, int C1, int HW,
                                int groups, int R, double* __restrict__ stats) {
  extern __shared__ float s_acc[];  // [2][C]
  const int C = C0 + C1;
  const int CV = C >> 3;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int cv = threadIdx.x % CV;
  const int rl = threadIdx.x / CV;
  const int c = cv << 3;
  const __half* src;
  long long stride;
  if (c < C0) {
    src = x0 + (long long)b * HW * C0 + c;
    stride = C0;
  } else {
    src = x1 + (long long)b * HW * C1 + (c - C0);
    stride = C1;
  }
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.f;
  for (int row = blockIdx.x * R + rl; row < HW; row += gridDim.x * R) {
    float v[8];
    load8(src + (long long)row * stride, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      s[e] += v[e];
      q[e] += v[e] * v[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    atomicAdd(&s_acc[c + e], s[e]);
    atomicAdd(&s_acc[C + c + e], q[e]);
  }
  __syncthreads();
  const int cpg = C / groups;
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    double ds = 0.0, dq = 0.0;
    for (int i = 0; i < cpg; ++i) {
      ds += (double)s_acc[g * cpg + i];
      dq += (double)s_acc[C + g * cpg + i];
    }
    atomicAdd(&stats[((long long)b * groups + g) * 2 + 0], ds);
    atomicAdd(&stats[((long long)b * groups + g) * 2 + 1], dq);
  }
}

__global__ void gn_apply_kernel(const __half* __restrict__ x0, int C0, const __half* __restrict__ x1, int C1, int HW,
                                int groups, int R, const double* __restrict__ stats, const __half* __restrict__ gamma,
                                const __half* __restrict__ beta, float eps, int silu, __half* __restrict__ out) {
  extern __shared__ float s_ss[];  // scale[C], shift[C]
  const int C = C0 + C1;
  const int CV = C >> 3;
  const int b = blockIdx.y;
  const int cpg = C / groups;
  const double n = (double)HW * cpg;
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
    const int g = ch / cpg;
    const double mean = stats[((long long)b * groups + g) * 2 + 0] / n;
    double var = stats[((long long)b * groups + g) * 2 + 1] / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float ga = __half2float(gamma[ch]), be = __half2float(beta[ch]);
    s_ss[ch] = ga * rstd;
    s_ss[C + ch] = be - (float)mean * ga * rstd;
  }
  __syncthreads();
  const int cv = threadIdx.x % CV;
  const int rl = threadIdx.x / CV;
  const int c = cv << 3;
  const __half* src;
  long long stride;
  if (c < C0) {
    src = x0 + (long long)b * HW * C0 + c;
    stride = C0;
  } else {
    src = x1 + (long long)b * HW * C1 + (c - C0);
    stride = C1;
  }
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sc[e] = s_ss[c + e];
    sh[e] = s_ss[C + c + e];
  }
  __half* dst = out + (long long)b * HW * C + c;
  for (int row = blockIdx.x * R + rl; row < HW; row += gridDim.x * R) {
    float v[8];
    load8(src + (long long)row * stride, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float y = v[e] * sc[e] + sh[e];
      if (silu) y = silu_f(y);
      v[e] = y;
    }
    store8(dst + (long long)row * C, v);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, values held in registers (two-pass mean / variance), C <= 8 * 32 * LN_MAXV.
// ---------------------------------------------------------------------------------------------------------------
constexpr int LN_MAXV = 16;  // C up to 4096

__global__ void layernorm_kernel(const __half* __restrict__ x, const __half* __restrict__ gamma,
                                 const __half* __restrict__ beta, __half* __restrict__ out, int rows, int C,
                                 float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int CV = C >> 3;
  const __half* src = x + (long long)warp * C;
  float v[LN_MAXV][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int cv = lane + i * 32;
    if (cv < CV) {
      load8(src + cv * 8, v[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) sum += v[i][e];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int cv = lane + i * 32;
    if (cv < CV) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = v[i][e] - mean;
        sq += d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / (float)C + eps);
  __half* dst = out + (long long)warp * C;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int cv = lane + i * 32;
    if (cv < CV) {
      float g[8], bt[8], y[8];
      load8(gamma + cv * 8, g);
      load8(beta + cv * 8, bt);
#pragma unroll
      for (int e = 0; e < 8; ++e) y[e] = (v[i][e] - mean) * rstd * g[e] + bt[e];
      store8(dst + cv * 8, y);
    }
  }
}

}  // namespace ih

using namespace ih;

extern "C" int ih_groupnorm_f16(const void* x0, int C0, const void* x1, int C1, const void* gamma, const void* beta,
                                void* out, void* stats_ws, int B, int HW, int groups, float eps, int silu,
                                void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  IH_CHECK(x0 && gamma && beta && out && stats_ws, IH_ERR_ARG, "ih_groupnorm_f16: null pointer");
  if (!x1) C1 = 0;
  const int C = C0 + C1;
  IH_CHECK(C0 % 8 == 0 && C1 % 8 == 0 && C % groups == 0 && C > 0, IH_ERR_SHAPE,
           "ih_groupnorm_f16: C0=%d C1=%d groups=%d unsupported", C0, C1, groups);
  const int CV = C / 8;
  IH_CHECK(CV <= 1024, IH_ERR_SHAPE, "ih_groupnorm_f16: C too large");
  int R = 256 / CV;
  if (R < 1) R = 1;
  const int threads = CV * R;
  int blocks = (HW + R - 1) / R;
  const int cap = (num_sms() * 4 + B - 1) / B;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  IH_CUDA(cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * B * groups, stream));
  dim3 grid(blocks, B);
  gn_stats_kernel<<<grid, threads, 2 * C * sizeof(float), stream>>>((const __half*)x0, C0, (const __half*)x1, C1, HW,
                                                                     groups, R, (double*)stats_ws);
  IH_CUDA(cudaGetLastError());
  gn_apply_kernel<<<grid, threads, 2 * C * sizeof(float), stream>>>((const __half*)x0, C0, (const __half*)x1, C1, HW,
                                                                     groups, R, (const double*)stats_ws,
                                                                     (const __half*)gamma, (const __half*)beta, eps,
                                                                     silu, (__half*)out);
  IH_CUDA(cudaGetLastError());
  count_launch(2);
  return 0;
}

extern "C" int ih_layernorm_f16(const void* x, const void* gamma, const void* beta, void* out, int rows, int C,
                                float eps, void* stream) {
  IH_CHECK(x && gamma && beta && out, IH_ERR_ARG, "ih_layernorm_f16: null pointer");
  IH_CHECK(C % 8 == 0 && C <= 8 * 32 * LN_MAXV && rows > 0, IH_ERR_SHAPE, "ih_layernorm_f16: C=%d unsupported", C);
  const int warps_per_block = 8;
  const int blocks = (rows + warps_per_block - 1) / warps_per_block;
  layernorm_kernel<<<blocks, warps_per_block * 32, 0, (cudaStream_t)stream>>>(
      (const __half*)x, (const __half*)gamma, (const __half*)beta, (__half*)out, rows, C, eps);
  IH_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
