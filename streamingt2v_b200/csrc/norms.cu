// GroupNorm (32 groups, channel-last) and LayerNorm kernels — HBM-bound, 16-byte vectorised, fp32 statistics.
//
// Replaces: GroupNorm32 (reference code/models/svd/sgm/modules/diffusionmodules/util.py:274-276, eps 1e-5),
// Normalize (attention.py:132-135, eps 1e-6), the CAM joint (C/32,F,H,W) GroupNorm (code/models/cam/conditioning.py:57-59),
// nn.LayerNorm (attention.py:528-530, video_attention.py:59,87,101-102; controlnet.py:113-118).
#include <cuda_bf16.h>

#include "../../include/b200svd.h"
#include "common.h"
#include "ptx.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------------------------------
// GroupNorm statistics: x [N][P][C] bf16 (row stride ldx) -> sums [N][32][2] (double: sum, sum of squares).
// grid = (chunks, N); block = (C/8) * rows_per_iter threads; each thread owns one 8-channel vector column.
// ------------------------------------------------------------------------------------------------------------
__global__ void gn_stats_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, int P, int C, int rows_per_chunk,
                                double* __restrict__ sums, double* __restrict__ partial, int* __restrict__ counters) {
  // Deterministic (fixed summation order, no floating-point atomics):
  //   thread partials -> smem [rstep][2C] -> per-channel sums (fixed order over rstep) -> per-group chunk partial
  //   (double) -> global partial[n][chunk][64]; the LAST chunk block of sample n (atomic ticket) adds all chunk
  //   partials in chunk order and writes sums[n][32][2].
  extern __shared__ float sm[];  // [rstep][2][C]
  __shared__ int is_last;
  const int vecs = C >> 3;
  const int n = blockIdx.y;
  const int chunks = gridDim.x;
  const int p0 = blockIdx.x * rows_per_chunk;
  const int p1 = min(P, p0 + rows_per_chunk);
  const int v = threadIdx.x % vecs;
  const int r0 = threadIdx.x / vecs;
  const int rstep = blockDim.x / vecs;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  const __nv_bfloat16* base = x + ((int64_t)n * P) * ldx + v * 8;
  constexpr int U = 4;  // independent 16-byte loads in flight per thread
  int p = p0 + r0;
  for (; p + (U - 1) * rstep < p1; p += U * rstep) {
    uint4 u[U];
#pragma unroll
    for (int i = 0; i < U; ++i) u[i] = __ldg(reinterpret_cast<const uint4*>(base + (int64_t)(p + i * rstep) * ldx));
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const uint32_t w[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a = bf16_lo(w[j]), b = bf16_hi(w[j]);
        s[2 * j] += a;
        q[2 * j] += a * a;
        s[2 * j + 1] += b;
        q[2 * j + 1] += b * b;
      }
    }
  }
  for (; p < p1; p += rstep) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(base + (int64_t)p * ldx));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = bf16_lo(w[j]), b = bf16_hi(w[j]);
      s[2 * j] += a;
      q[2 * j] += a * a;
      s[2 * j + 1] += b;
      q[2 * j + 1] += b * b;
    }
  }
  float* mine = sm + (size_t)r0 * 2 * C;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    mine[v * 8 + j] = s[j];
    mine[C + v * 8 + j] = q[j];
  }
  __syncthreads();
  const int cpg = C >> 5;
  if (threadIdx.x < 64) {
    const int g = threadIdx.x & 31, which = threadIdx.x >> 5;  // which: 0 = sum, 1 = sum of squares
    double a = 0.0;
    for (int j = 0; j < cpg; ++j) {
      float c = 0.f;
      for (int r = 0; r < rstep; ++r) c += sm[(size_t)r * 2 * C + which * C + g * cpg + j];
      a += (double)c;
    }
    partial[(((int64_t)n * chunks + blockIdx.x) * 32 + g) * 2 + which] = a;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int ticket = atomicAdd(&counters[n], 1);
    is_last = (ticket == chunks - 1);
    if (is_last) counters[n] = 0;  // self-reset for the next launch
  }
  __syncthreads();
  if (is_last) {
    // fixed-order final reduction over the chunk partials, spread over the whole block: thread t owns pair t%64 and
    // the chunks congruent to t/64 modulo (blockDim/64); the (blockDim/64) strided sums are then added in order
    __threadfence();
    const int pairs = 64;
    const int parts = blockDim.x / pairs;  // >= 2 (blockDim >= 160)
    double* dsm = reinterpret_cast<double*>(sm);  // reuse the staging area: needs parts*64 doubles <= 2*C*rstep floats
    const int pr = threadIdx.x % pairs, part = threadIdx.x / pairs;
    if (part < parts) {
      // 16 independent partial sums keep 16 loads in flight per thread (the chain of dependent double adds over up
      // to ~1000 chunk partials, one L2 round trip each, used to dominate the launch for few, large samples); the
      // summation order stays fixed
      const double* src = partial + (int64_t)n * chunks * 64 + (pr & 31) * 2 + (pr >> 5);
      constexpr int W = 16;
      double a[W];
#pragma unroll
      for (int i = 0; i < W; ++i) a[i] = 0.0;
      int c = part;
      for (; c + (W - 1) * parts < chunks; c += W * parts) {
#pragma unroll
        for (int i = 0; i < W; ++i) a[i] += src[(int64_t)(c + i * parts) * 64];
      }
      for (; c < chunks; c += parts) a[0] += src[(int64_t)c * 64];
#pragma unroll
      for (int w = W / 2; w > 0; w >>= 1)
#pragma unroll
        for (int i = 0; i < w; ++i) a[i] += a[i + w];
      dsm[part * pairs + pr] = a[0];
    }
    __syncthreads();
    if (threadIdx.x < pairs) {
      double a = 0.0;
      for (int q = 0; q < parts; ++q) a += dsm[q * pairs + threadIdx.x];
      sums[((int64_t)n * 32 + (threadIdx.x & 31)) * 2 + (threadIdx.x >> 5)] = a;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// GroupNorm statistics from the per-quadrant partials written by the GEMM epilogue (mtgemm.cu, gn_part):
// part [n_slots][ld][2] fp32 (sum, sum of squares per channel over the <= 32 rows of a slot), slot_sample [n_slots].
// grid = (chunks of 64 slots, N); the block adds the slots of ITS sample in slot order (fixed), folds channels into
// groups in double, and the last block of the sample (ticket) adds the chunk partials in chunk order: deterministic.
// ------------------------------------------------------------------------------------------------------------
constexpr int GNP_SLOTS = 64;
constexpr int GNP_THREADS = 256;
__global__ void __launch_bounds__(GNP_THREADS)
gn_stats_partials_kernel(const float2* __restrict__ part, const int* __restrict__ slot_sample, int64_t n_slots,
                         int64_t ld, int C, double* __restrict__ sums, double* __restrict__ partial,
                         int* __restrict__ counters) {
  extern __shared__ float sm[];  // [2][C], reused as doubles by the last block
  __shared__ int is_last;
  __shared__ int hit[GNP_SLOTS];
  const int n = blockIdx.y;
  const int chunks = gridDim.x;
  const int64_t s0 = (int64_t)blockIdx.x * GNP_SLOTS;
  if (threadIdx.x < GNP_SLOTS) {
    const int64_t sl = s0 + threadIdx.x;
    hit[threadIdx.x] = (sl < n_slots && slot_sample[sl] == n) ? 1 : 0;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += GNP_THREADS) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < GNP_SLOTS; ++i) {
      if (hit[i]) {
        const float2 v = __ldg(part + (s0 + i) * ld + c);
        a += v.x;
        b += v.y;
      }
    }
    sm[c] = a;
    sm[C + c] = b;
  }
  __syncthreads();
  const int cpg = C >> 5;
  if (threadIdx.x < 64) {
    const int g = threadIdx.x & 31, which = threadIdx.x >> 5;
    double a = 0.0;
    for (int j = 0; j < cpg; ++j) a += (double)sm[which * C + g * cpg + j];
    partial[(((int64_t)n * chunks + blockIdx.x) * 32 + g) * 2 + which] = a;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int ticket = atomicAdd(&counters[n], 1);
    is_last = (ticket == chunks - 1);
    if (is_last) counters[n] = 0;
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    const int pairs = 64;
    const int parts = GNP_THREADS / pairs;  // 4
    double* dsm = reinterpret_cast<double*>(sm);  // parts * 64 doubles = 2 KB <= 2 * C floats (C >= 256) — checked on host
    const int pr = threadIdx.x % pairs, part_i = threadIdx.x / pairs;
    const double* src = partial + (int64_t)n * chunks * 64 + (pr & 31) * 2 + (pr >> 5);
    constexpr int W = 8;
    double a[W];
#pragma unroll
    for (int i = 0; i < W; ++i) a[i] = 0.0;
    int c = part_i;
    for (; c + (W - 1) * parts < chunks; c += W * parts) {
#pragma unroll
      for (int i = 0; i < W; ++i) a[i] += src[(int64_t)(c + i * parts) * 64];
    }
    for (; c < chunks; c += parts) a[0] += src[(int64_t)c * 64];
#pragma unroll
    for (int w = W / 2; w > 0; w >>= 1)
#pragma unroll
      for (int i = 0; i < w; ++i) a[i] += a[i + w];
    __syncthreads();
    dsm[part_i * pairs + pr] = a[0];
    __syncthreads();
    if (threadIdx.x < pairs) {
      double t = 0.0;
      for (int q = 0; q < parts; ++q) t += dsm[q * pairs + threadIdx.x];
      sums[((int64_t)n * 32 + (threadIdx.x & 31)) * 2 + (threadIdx.x >> 5)] = t;
    }
  }
}

// y = [silu]((x - mean) * rstd * gamma + beta), bf16 out (row stride ldy).
__global__ void gn_apply_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ y,
                                int64_t ldy, int P, int C, int rows_per_chunk, const double* __restrict__ sums,
                                const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                int apply_silu) {
  const int vecs = C >> 3;
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * rows_per_chunk;
  const int p1 = min(P, p0 + rows_per_chunk);
  const int v = threadIdx.x % vecs;
  const int r0 = threadIdx.x / vecs;
  const int rstep = blockDim.x / vecs;
  const int cpg = C >> 5;
  const double cnt = (double)P * (double)cpg;
  float sc[8], sh[8];
  int gprev = -1;
  float mean_f = 0.f, rstd = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = v * 8 + j;
    const int g = c / cpg;
    if (g != gprev) {  // the double-precision statistics are evaluated once per distinct group (<= 2 for C >= 256)
      const double su = sums[((int64_t)n * 32 + g) * 2], sq = sums[((int64_t)n * 32 + g) * 2 + 1];
      const double mean = su / cnt;
      double var = sq / cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      rstd = rsqrtf((float)var + eps);
      mean_f = (float)mean;
      gprev = g;
    }
    const float ga = __ldg(gamma + c), be = __ldg(beta + c);
    sc[j] = rstd * ga;
    sh[j] = be - mean_f * rstd * ga;
  }
  const __nv_bfloat16* xb = x + ((int64_t)n * P) * ldx + v * 8;
  __nv_bfloat16* yb = y + ((int64_t)n * P) * ldy + v * 8;
  auto apply8 = [&](const uint4& u) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = bf16_lo(w[j]) * sc[2 * j] + sh[2 * j];
      float b = bf16_hi(w[j]) * sc[2 * j + 1] + sh[2 * j + 1];
      if (apply_silu) {
        a = silu_fast(a);
        b = silu_fast(b);
      }
      o[j] = pack_bf16x2(a, b);
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
  };
  constexpr int U = 4;
  int p = p0 + r0;
  for (; p + (U - 1) * rstep < p1; p += U * rstep) {
    uint4 u[U];
#pragma unroll
    for (int i = 0; i < U; ++i) u[i] = __ldg(reinterpret_cast<const uint4*>(xb + (int64_t)(p + i * rstep) * ldx));
#pragma unroll
    for (int i = 0; i < U; ++i) *reinterpret_cast<uint4*>(yb + (int64_t)(p + i * rstep) * ldy) = apply8(u[i]);
  }
  for (; p < p1; p += rstep)
    *reinterpret_cast<uint4*>(yb + (int64_t)p * ldy) = apply8(__ldg(reinterpret_cast<const uint4*>(xb + (int64_t)p * ldx)));
}

// ------------------------------------------------------------------------------------------------------------
// LayerNorm over the channel dim: one warp per row.  y = ((x [+ fvec[row/rpf]]) - mean) * rstd * gamma + beta
// Optional: also write xsum = x + fvec (bf16) so the caller can use it as the residual stream, and fused SiLU.
// ------------------------------------------------------------------------------------------------------------
template <int MAXV>  // max 8-channel vectors per lane
__global__ void layernorm_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ y,
                                 int64_t ldy, int64_t rows, int C, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps, const float* __restrict__ fvec, int64_t ldf,
                                 int rows_per_frame, __nv_bfloat16* __restrict__ xsum, int64_t ldxs, int apply_silu) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int vecs = C >> 3;
  float val[MAXV][8];
  const __nv_bfloat16* xr = x + row * ldx;
  const float* fr = fvec ? fvec + (row / rows_per_frame) * ldf : nullptr;
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + i * 32;
    if (v < vecs) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(xr + v * 8));
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        val[i][2 * j] = bf16_lo(w[j]);
        val[i][2 * j + 1] = bf16_hi(w[j]);
      }
      if (fr) {
#pragma unroll
        for (int j = 0; j < 8; ++j) val[i][j] += __ldg(fr + v * 8 + j);
        if (xsum) {
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = pack_bf16x2(val[i][2 * j], val[i][2 * j + 1]);
          *reinterpret_cast<uint4*>(xsum + row * ldxs + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
          // the residual stream is the bf16-rounded sum; normalise exactly what the consumer will see
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            val[i][2 * j] = bf16_lo(o[j]);
            val[i][2 * j + 1] = bf16_hi(o[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += val[i][j];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + i * 32;
    if (v < vecs) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = val[i][j] - mean;
        sq += d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / (float)C + eps);
  __nv_bfloat16* yr = y + row * ldy;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + i * 32;
    if (v < vecs) {
      float o8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = v * 8 + j;
        float t = (val[i][j] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
        if (apply_silu) t = silu_fast(t);
        o8[j] = t;
      }
      *reinterpret_cast<uint4*>(yr + v * 8) = make_uint4(pack_bf16x2(o8[0], o8[1]), pack_bf16x2(o8[2], o8[3]),
                                                          pack_bf16x2(o8[4], o8[5]), pack_bf16x2(o8[6], o8[7]));
    }
  }
}

// Narrow rows (C <= 256: the ControlNet condition-embedding norms): LPR lanes own one row, one 16-byte vector each
// (lanes >= C/8 idle), 32/LPR rows per warp, so a warp issues full-width loads instead of one 64-byte row at a time.
template <int LPR>
__global__ void __launch_bounds__(256) layernorm_narrow_kernel(
    const __nv_bfloat16* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ y, int64_t ldy, int64_t rows, int C,
    const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int apply_silu) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, li = lane % LPR;
  const int64_t row = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + sub;
  const bool active = row < rows && li < (C >> 3);
  float val[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) val[j] = 0.f;
  if (active) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + row * ldx + li * 8));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      val[2 * j] = bf16_lo(w[j]);
      val[2 * j + 1] = bf16_hi(w[j]);
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) sum += val[j];
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)C;
  float sq = 0.f;
  if (active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = val[j] - mean;
      sq += d * d;
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / (float)C + eps);
  if (!active) return;
  const int c0 = li * 8;
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0) + 1);
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c0) + 1);
  const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  float o8[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float t = (val[j] - mean) * rstd * gg[j] + bb[j];
    if (apply_silu) t = silu_fast(t);
    o8[j] = t;
  }
  *reinterpret_cast<uint4*>(y + row * ldy + c0) = make_uint4(pack_bf16x2(o8[0], o8[1]), pack_bf16x2(o8[2], o8[3]),
                                                             pack_bf16x2(o8[4], o8[5]), pack_bf16x2(o8[6], o8[7]));
}

// Fast path for C = 40*LPR channels-vectors (C = 320, 640, 1280): LPR lanes own one row, 5 x 16-byte vectors each,
// all loads issued before first use; 32/LPR rows per warp.
template <int LPR>
__global__ void __launch_bounds__(256) layernorm5_kernel(
    const __nv_bfloat16* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ y, int64_t ldy, int64_t rows, int C,
    const float* __restrict__ gamma, const float* __restrict__ beta, float eps, const float* __restrict__ fvec,
    int64_t ldf, int rows_per_frame, __nv_bfloat16* __restrict__ xsum, int64_t ldxs, int apply_silu) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, li = lane % LPR;
  const int64_t row = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + sub;
  const bool active = row < rows;
  const int64_t rr = active ? row : rows - 1;
  const __nv_bfloat16* xr = x + rr * ldx;
  uint4 u[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) u[i] = __ldg(reinterpret_cast<const uint4*>(xr + (li + i * LPR) * 8));
  float val[5][8];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const uint32_t w[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      val[i][2 * j] = bf16_lo(w[j]);
      val[i][2 * j + 1] = bf16_hi(w[j]);
    }
  }
  if (fvec != nullptr) {
    const float* fr = fvec + (rr / rows_per_frame) * ldf;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const float4 f0 = __ldg(reinterpret_cast<const float4*>(fr + (li + i * LPR) * 8));
      const float4 f1 = __ldg(reinterpret_cast<const float4*>(fr + (li + i * LPR) * 8) + 1);
      val[i][0] += f0.x; val[i][1] += f0.y; val[i][2] += f0.z; val[i][3] += f0.w;
      val[i][4] += f1.x; val[i][5] += f1.y; val[i][6] += f1.z; val[i][7] += f1.w;
      if (xsum != nullptr) {
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = pack_bf16x2(val[i][2 * j], val[i][2 * j + 1]);
        if (active) *reinterpret_cast<uint4*>(xsum + row * ldxs + (li + i * LPR) * 8) = make_uint4(o[0], o[1], o[2], o[3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          val[i][2 * j] = bf16_lo(o[j]);
          val[i][2 * j + 1] = bf16_hi(o[j]);
        }
      }
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) sum += val[i][j];
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = val[i][j] - mean;
      sq += d * d;
    }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / (float)C + eps);
  if (!active) return;
  __nv_bfloat16* yr = y + row * ldy;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const int c0 = (li + i * LPR) * 8;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0) + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c0) + 1);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float o8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = (val[i][j] - mean) * rstd * gg[j] + bb[j];
      if (apply_silu) t = silu_fast(t);
      o8[j] = t;
    }
    *reinterpret_cast<uint4*>(yr + c0) = make_uint4(pack_bf16x2(o8[0], o8[1]), pack_bf16x2(o8[2], o8[3]),
                                                     pack_bf16x2(o8[4], o8[5]), pack_bf16x2(o8[6], o8[7]));
  }
}

static int gn_geometry(int C, int64_t n, int P, int* threads, int* rows_per_chunk, int* chunks) {
  if (C % 8 != 0 || C % 32 != 0) return 1;
  const int vecs = C / 8;
  if (vecs > 1024) return 1;
  int rpi = 256 / vecs;
  if (rpi < 1) rpi = 1;
  *threads = vecs * rpi;
  // Enough CTAs to cover the GPU several times over (these kernels are latency-bound otherwise): aim for
  // >= 8 CTAs per SM, but never less than one 4-deep unrolled batch of rows per thread, never more than 256 rows.
  const int min_rpc = rpi * 4;
  const int64_t target_ctas = (int64_t)sm_count() * 8;
  int64_t rpc = ((int64_t)P * n + target_ctas - 1) / target_ctas;
  rpc = ((rpc + min_rpc - 1) / min_rpc) * min_rpc;
  if (rpc < min_rpc) rpc = min_rpc;
  int64_t cap = 256;
  while ((int64_t)P / cap > 2048 && cap < 4096) cap *= 2;  // bound the number of chunk partials per sample
  if (rpc > cap) rpc = cap / min_rpc * min_rpc > 0 ? cap / min_rpc * min_rpc : min_rpc;
  if (rpc > P) rpc = P;
  *rows_per_chunk = (int)rpc;
  *chunks = (int)((P + rpc - 1) / rpc);
  return 0;
}

}  // namespace b200

extern "C" {

// sums: n*32*2 doubles (out).  scratch: at least b200svd_gn_scratch_doubles(n, p, c) doubles.  counters: n int32,
// zero-initialised once by the caller (the kernel leaves them zero).
int64_t b200svd_gn_scratch_doubles(int64_t n, int64_t p, int c) {
  using namespace b200;
  int threads, rpc, chunks;
  if (gn_geometry(c, n, (int)p, &threads, &rpc, &chunks)) return -1;
  return n * (int64_t)chunks * 64;
}

int b200svd_gn_stats_partials(const float* gn_part, const int32_t* gn_slot_sample, int64_t n_slots, int64_t gn_ld,
                              int c, int64_t n, void* sums, void* scratch, void* counters, void* stream) {
  using namespace b200;
  if (c % 32 != 0 || c < 256 || c > 8192 || gn_ld < c || n_slots < 1 || n < 1) {
    set_error("gn_stats_partials: need 256 <= c <= 8192, c %% 32 == 0, gn_ld >= c (c %d, ld %lld, slots %lld)", c,
              (long long)gn_ld, (long long)n_slots);
    return 1;
  }
  const int64_t chunks = (n_slots + GNP_SLOTS - 1) / GNP_SLOTS;
  dim3 grid((unsigned)chunks, (unsigned)n);
  const size_t smem = (size_t)2 * c * sizeof(float);
  static size_t smem_set_dev[B200_MAX_DEVICES] = {};
  size_t& smem_set = smem_set_dev[dev_slot()];
  if (smem > 48 * 1024 && smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(gn_stats_partials_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(gn_stats_partials)");
    smem_set = smem;
  }
  gn_stats_partials_kernel<<<grid, GNP_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2*>(gn_part), gn_slot_sample, n_slots, gn_ld, c, reinterpret_cast<double*>(sums),
      reinterpret_cast<double*>(scratch), reinterpret_cast<int*>(counters));
  B200_CHECK_LAUNCH("gn_stats_partials");
  return 0;
}

int b200svd_gn_stats(const void* x, int64_t ldx, int64_t n, int64_t p, int c, void* sums, void* scratch,
                     void* counters, void* stream) {
  using namespace b200;
  int threads, rpc, chunks;
  if (gn_geometry(c, n, (int)p, &threads, &rpc, &chunks)) {
    set_error("gn_stats: unsupported channel count %d (need multiple of 32, <= 8192)", c);
    return 1;
  }
  if (ldx % 8 != 0) {
    set_error("gn_stats: ldx must be a multiple of 8");
    return 1;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(chunks, (unsigned)n);
  const int rstep = threads / (c / 8);
  const size_t smem = (size_t)rstep * 2 * c * sizeof(float);
  static size_t smem_set_dev[B200_MAX_DEVICES] = {};
  size_t& smem_set = smem_set_dev[dev_slot()];
  if (smem > 48 * 1024 && smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(gn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "gn_stats smem attribute");
    smem_set = smem;
  }
  gn_stats_kernel<<<grid, threads, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx, (int)p, c, rpc,
                                             reinterpret_cast<double*>(sums), reinterpret_cast<double*>(scratch),
                                             reinterpret_cast<int*>(counters));
  B200_CHECK_LAUNCH("gn_stats");
  return 0;
}

int b200svd_gn_apply(const void* x, int64_t ldx, void* y, int64_t ldy, int64_t n, int64_t p, int c, const void* sums,
                     const float* gamma, const float* beta, float eps, int apply_silu, void* stream) {
  using namespace b200;
  int threads, rpc, chunks;
  if (gn_geometry(c, n, (int)p, &threads, &rpc, &chunks)) {
    set_error("gn_apply: unsupported channel count %d", c);
    return 1;
  }
  if (ldx % 8 != 0 || ldy % 8 != 0) {
    set_error("gn_apply: leading dims must be multiples of 8");
    return 1;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(chunks, (unsigned)n);
  gn_apply_kernel<<<grid, threads, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx,
                                            reinterpret_cast<__nv_bfloat16*>(y), ldy, (int)p, c, rpc,
                                            reinterpret_cast<const double*>(sums), gamma, beta, eps, apply_silu);
  B200_CHECK_LAUNCH("gn_apply");
  return 0;
}

int b200svd_layernorm(const void* x, int64_t ldx, void* y, int64_t ldy, int64_t rows, int c, const float* gamma,
                      const float* beta, float eps, const float* fvec, int64_t ldf, int rows_per_frame, void* xsum,
                      int64_t ldxs, int apply_silu, void* stream) {
  using namespace b200;
  if (c % 8 != 0 || c > 8 * 32 * 8) {
    set_error("layernorm: unsupported width %d (multiple of 8, <= 2048)", c);
    return 1;
  }
  if (ldx % 8 != 0 || ldy % 8 != 0 || (xsum && ldxs % 8 != 0)) {
    set_error("layernorm: leading dims must be multiples of 8");
    return 1;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int wpb = 8;
  const unsigned grid = (unsigned)((rows + wpb - 1) / wpb);
  const int vecs = c / 8;
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y);
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(xsum);
  if (rows_per_frame <= 0) rows_per_frame = 1;
  const bool f_ok = (fvec == nullptr) || (ldf % 4 == 0 && (reinterpret_cast<uintptr_t>(fvec) & 15) == 0);
  const bool gb_ok = ((reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0;
  if ((c == 320 || c == 640 || c == 1280) && f_ok && gb_ok) {
    const int lpr = c / 40;
    const int rpw = 32 / lpr;
    const unsigned g5 = (unsigned)((rows + (int64_t)wpb * rpw - 1) / ((int64_t)wpb * rpw));
    if (lpr == 8)
      layernorm5_kernel<8><<<g5, wpb * 32, 0, st>>>(xp, ldx, yp, ldy, rows, c, gamma, beta, eps, fvec, ldf, rows_per_frame, xs, ldxs, apply_silu);
    else if (lpr == 16)
      layernorm5_kernel<16><<<g5, wpb * 32, 0, st>>>(xp, ldx, yp, ldy, rows, c, gamma, beta, eps, fvec, ldf, rows_per_frame, xs, ldxs, apply_silu);
    else
      layernorm5_kernel<32><<<g5, wpb * 32, 0, st>>>(xp, ldx, yp, ldy, rows, c, gamma, beta, eps, fvec, ldf, rows_per_frame, xs, ldxs, apply_silu);
    B200_CHECK_LAUNCH("layernorm5");
    return 0;
  }
  if (vecs <= 32 && fvec == nullptr && xsum == nullptr && gb_ok) {
    const int lpr = vecs <= 4 ? 4 : vecs <= 8 ? 8 : vecs <= 16 ? 16 : 32;
    const int rpw = 32 / lpr;
    const unsigned gn = (unsigned)((rows + (int64_t)wpb * rpw - 1) / ((int64_t)wpb * rpw));
    if (lpr == 4)
      layernorm_narrow_kernel<4><<<gn, wpb * 32, 0, st>>>(xp, ldx, yp, ldy, rows, c, gamma, beta, eps, apply_silu);
    else if (lpr == 8)
      layernorm_narrow_kernel<8><<<gn, wpb * 32, 0, st>>>(xp, ldx, yp, ldy, rows, c, gamma, beta, eps, apply_silu);
    else if (lpr == 16)
      layernorm_narrow_kernel<16><<<gn, wpb * 32, 0, st>>>(xp, ldx, yp, ldy, rows, c, gamma, beta, eps, apply_silu);
    else
      layernorm_narrow_kernel<32><<<gn, wpb * 32, 0, st>>>(xp, ldx, yp, ldy, rows, c, gamma, beta, eps, apply_silu);
    B200_CHECK_LAUNCH("layernorm_narrow");
    return 0;
  }
#define LN_LAUNCH(MV)                                                                                              \
  layernorm_kernel<MV><<<grid, wpb * 32, 0, st>>>(xp, ldx, yp, ldy, rows, c, gamma, beta, eps, fvec, ldf,          \
                                                  rows_per_frame, xs, ldxs, apply_silu)
  if (vecs <= 32) LN_LAUNCH(1);
  else if (vecs <= 64) LN_LAUNCH(2);
  else if (vecs <= 128) LN_LAUNCH(4);
  else LN_LAUNCH(8);
#undef LN_LAUNCH
  B200_CHECK_LAUNCH("layernorm");
  return 0;
}

}  // extern "C"
