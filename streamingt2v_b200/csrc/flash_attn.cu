// FlashAttention forward for sm_100a, head dim 64: tcgen05.mma (S = Q K^T and O_blk = P V, accumulators in TMEM),
// TMA-fed Q/K/V tiles read straight from the fused QKV projection output [(n s), 3C] (no head split / transpose
// in HBM), online softmax in registers (one query row per thread).
//
// Replaces the spatial self-attention core of BasicTransformerBlock.attn1
// (reference code/models/svd/sgm/modules/attention.py:320-351 SDPA / :427-446 xformers), batch = frames,
// heads = C/64, sequence = H*W.
//
// CTA = 128 query rows of one (frame, head); 6 warps: warp0 TMA producer, warp1 MMA issuer + TMEM owner,
// warps 2..5 softmax/correction/epilogue.  Per 128-key block:
//   MMA : S[128x128] = Q K_j^T                      (4 x tcgen05.mma M128 N128 K16, K-major A and B)
//   SM  : m,l update; P = exp2(S*c - m*c) -> bf16 into smem (128B-swizzled K-major A tile)
//   MMA : O_blk[128x64] = P V_j                     (8 x tcgen05.mma M128 N64 K16, B = V is MN-major)
//   SM  : O = O*alpha + O_blk                        (registers)
// Two CTAs co-reside per SM so one CTA's softmax overlaps the other's MMAs.
#include <cuda.h>
#include <cuda_bf16.h>
#include <math.h>

#include "../../include/b200svd.h"
#include "common.h"
#include "ptx.cuh"

namespace b200 {

constexpr int FA_BQ = 128;
constexpr int FA_BK = 128;
constexpr int FA_D = 64;
constexpr int FA_KV_STAGES = 2;
constexpr int FA_Q_BYTES = FA_BQ * FA_D * 2;        // 16 KB
constexpr int FA_KV_TILE_BYTES = FA_BK * FA_D * 2;  // 16 KB each for K and V
constexpr int FA_P_BYTES = FA_BQ * FA_BK * 2;       // 32 KB (two 64-column swizzled sub-tiles)
constexpr int FA_SMEM_BYTES = FA_Q_BYTES + FA_KV_STAGES * 2 * FA_KV_TILE_BYTES + FA_P_BYTES + 128;  // 2 CTAs / SM
constexpr int FA_TMEM_COLS = 256;  // S: cols [0,128), O_blk: cols [128,192)

struct FaParams {
  __nv_bfloat16* out;
  int64_t ldo;
  int S, heads, C;
  float scale_log2;
};

__global__ void __launch_bounds__(192, 2)
flash_attn_kernel(const __grid_constant__ CUtensorMap tmQKV, const FaParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];  // SWIZZLE_128B tiles need 1024-byte alignment
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + FA_Q_BYTES;                          // stage s: K at sKV + s*32K, V at +16K
  uint8_t* sP = sKV + FA_KV_STAGES * 2 * FA_KV_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + FA_P_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                  // [2]
  uint64_t* kv_empty = bars + 3;                 // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x, head = blockIdx.y, n = blockIdx.z;
  const int q0 = q_tile * FA_BQ;
  const int nkb = (p.S + FA_BK - 1) / FA_BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQKV);
    mbar_init(q_full, 1);
    for (int s = 0; s < FA_KV_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, FA_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      mbar_expect_tx(q_full, FA_Q_BYTES);
      tma_load_3d(sQ, &tmQKV, q_full, head * FA_D, q0, n);
      for (int j = 0; j < nkb; ++j) {
        const int s = j % FA_KV_STAGES;
        const uint32_t ph = (j / FA_KV_STAGES) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1);
        uint8_t* sk = sKV + s * 2 * FA_KV_TILE_BYTES;
        mbar_expect_tx(&kv_full[s], 2 * FA_KV_TILE_BYTES);
        tma_load_3d(sk, &tmQKV, &kv_full[s], p.C + head * FA_D, j * FA_BK, n);
        tma_load_3d(sk + FA_KV_TILE_BYTES, &tmQKV, &kv_full[s], 2 * p.C + head * FA_D, j * FA_BK, n);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc_qk = make_idesc_f16(128, 128, 1, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_f16(128, 64, 1, 0, 1);  // B (=V) is MN-major
      const uint64_t qdesc = smem_desc_k_sw128(smem_u32(sQ));
      const uint64_t pdesc0 = smem_desc_k_sw128(smem_u32(sP));
      const uint64_t pdesc1 = smem_desc_k_sw128(smem_u32(sP + FA_BQ * 128));
      mbar_wait(q_full, 0);
      // S(0)
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      {
        const uint64_t kdesc = smem_desc_k_sw128(smem_u32(sKV));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_f16_ss(tmem_S, qdesc + kk * 2, kdesc + kk * 2, idesc_qk, kk > 0);
        umma_commit(s_full);
      }
      for (int j = 0; j < nkb; ++j) {
        const int s = j % FA_KV_STAGES;
        // P(j) ready (and S(j) fully read by the softmax warps)
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const uint64_t vdesc = smem_desc_mn_sw128(smem_u32(sKV + s * 2 * FA_KV_TILE_BYTES + FA_KV_TILE_BYTES));
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t pd = (kk < 4 ? pdesc0 : pdesc1) + (uint64_t)((kk & 3) * 2);
          // V: 16 keys per MMA = 2 swizzle row-groups of 1024 B -> +2048 B = +128 in the (addr>>4) field
          umma_f16_ss(tmem_O, pd, vdesc + (uint64_t)(kk * 128), idesc_pv, kk > 0);
        }
        umma_commit(o_full);
        umma_commit(&kv_empty[s]);
        if (j + 1 < nkb) {
          const int s1 = (j + 1) % FA_KV_STAGES;
          mbar_wait(&kv_full[s1], ((j + 1) / FA_KV_STAGES) & 1);
          tc_fence_after();
          const uint64_t kdesc = smem_desc_k_sw128(smem_u32(sKV + s1 * 2 * FA_KV_TILE_BYTES));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_f16_ss(tmem_S, qdesc + kk * 2, kdesc + kk * 2, idesc_qk, kk > 0);
          umma_commit(s_full);
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== softmax / correction / epilogue warps =====================
    const int qd = warp & 3;
    const int r = qd * 32 + lane;  // query row in tile == TMEM lane
    const uint32_t tl = ((uint32_t)(qd * 32)) << 16;
    float o_acc[FA_D];
#pragma unroll
    for (int d = 0; d < FA_D; ++d) o_acc[d] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    uint8_t* prow = sP + r * 128;
    const int sw = r & 7;

    for (int j = 0; j < nkb; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int kbase = j * FA_BK;
      const bool tail = (kbase + FA_BK > p.S);
      // pass 1: row max
      float mx = m_run;
#pragma unroll
      for (int c0 = 0; c0 < FA_BK; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_S + tl + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float sv = __uint_as_float(v[i]);
          if (tail && kbase + c0 + i >= p.S) sv = -INFINITY;
          mx = fmaxf(mx, sv);
        }
      }
      const float m_new = mx;  // finite: every block has >= 1 valid key
      const float alpha = exp2f((m_run - m_new) * p.scale_log2);
      const float mb = m_new * p.scale_log2;
      float lsum = 0.f;
      // pass 2: P = exp2(S*c - m*c) -> bf16 -> swizzled smem
#pragma unroll
      for (int c0 = 0; c0 < FA_BK; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_S + tl + c0, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float a = __uint_as_float(v[2 * i]), b = __uint_as_float(v[2 * i + 1]);
          a = exp2f(fmaf(a, p.scale_log2, -mb));
          b = exp2f(fmaf(b, p.scale_log2, -mb));
          if (tail) {
            if (kbase + c0 + 2 * i >= p.S) a = 0.f;
            if (kbase + c0 + 2 * i + 1 >= p.S) b = 0.f;
          }
          // accumulate the row sum from the bf16-rounded values that the PV MMA will actually use
          const uint32_t w = pack_bf16x2(a, b);
          pk[i] = w;
          lsum += bf16_lo(w) + bf16_hi(w);
        }
        // 32 columns = 4 x 16-byte chunks; sub-tile = c0/64, chunk index within the 128-byte row = (c0%64)/8 + q
        uint8_t* sub = prow + (c0 >> 6) * (FA_BQ * 128);
        const int cb = (c0 & 63) >> 3;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = (cb + q) ^ sw;
          *reinterpret_cast<uint4*>(sub + chunk * 16) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
      }
      l_run = l_run * alpha + lsum;
      m_run = m_new;
      // publish P (generic-proxy smem writes -> async proxy) and release S
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
      // rescale running output while the PV MMA runs
#pragma unroll
      for (int d = 0; d < FA_D; ++d) o_acc[d] *= alpha;
      mbar_wait(o_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < FA_D; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_O + tl + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[c0 + i] += __uint_as_float(v[i]);
      }
      tc_fence_before();  // order these TMEM reads before the next arrive (p_full) that lets PV(j+1) overwrite O
    }
    // epilogue: normalise and store 64 bf16 (128 B contiguous) per row
    if (q0 + r < p.S) {
      const float inv = 1.0f / l_run;
      __nv_bfloat16* dst = p.out + ((int64_t)n * p.S + q0 + r) * p.ldo + head * FA_D;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        reinterpret_cast<uint4*>(dst)[c] =
            make_uint4(pack_bf16x2(o_acc[8 * c] * inv, o_acc[8 * c + 1] * inv),
                       pack_bf16x2(o_acc[8 * c + 2] * inv, o_acc[8 * c + 3] * inv),
                       pack_bf16x2(o_acc[8 * c + 4] * inv, o_acc[8 * c + 5] * inv),
                       pack_bf16x2(o_acc[8 * c + 6] * inv, o_acc[8 * c + 7] * inv));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, FA_TMEM_COLS);
  }
}

}  // namespace b200

// qkv: [(n s), ldqkv] bf16 with columns [q | k | v], each C = heads*64 wide; out: [(n s), ldo] bf16 (C columns).
extern "C" int b200svd_flash_attn(const void* qkv, int64_t ldqkv, void* out, int64_t ldo, int n, int s, int heads,
                                  float scale, void* stream) {
  using namespace b200;
  if (ldqkv % 8 || ldo % 8) {
    set_error("flash_attn: leading dims must be multiples of 8");
    return 1;
  }
  const int C = heads * FA_D;
  CUtensorMap tm;
  uint64_t dims[3] = {(uint64_t)3 * C, (uint64_t)s, (uint64_t)n};
  uint64_t strides[2] = {(uint64_t)ldqkv * 2, (uint64_t)ldqkv * 2 * (uint64_t)s};
  uint32_t box[3] = {64, 128, 1};
  if (encode_tmap_bf16(&tm, qkv, 3, dims, strides, box)) return 1;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(flash_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_BYTES);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(flash_attn)");
    attr_set = true;
  }
  FaParams p;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.S = s;
  p.heads = heads;
  p.C = C;
  p.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((s + FA_BQ - 1) / FA_BQ, heads, n);
  flash_attn_kernel<<<grid, 192, FA_SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream)>>>(tm, p);
  B200_CHECK_LAUNCH("flash_attn");
  return 0;
}
