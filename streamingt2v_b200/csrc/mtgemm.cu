// Multi-tap tensor-core GEMM for sm_100a: TMA (rank-5 activation view) -> smem (128B swizzle) -> tcgen05.mma
// (fp32 accumulators in TMEM) -> fused epilogue.  See include/b200svd.h for the contract.
//
// Replaces, underneath StreamingWrapper.forward (reference code/models/diffusion/wrappers.py:23-78):
//   nn.Linear            code/models/svd/sgm/modules/attention.py:94-120,262-351, video_attention.py:23-168
//   Conv2d 3x3 (s1, s2)  code/models/svd/sgm/modules/diffusionmodules/openaimodel.py:107-207,257-305
//   Conv3d (3,1,1)       code/models/diffusion/video_model.py:46-59 (ResBlock dims=3)
//   + their elementwise neighbours (bias, emb add openaimodel.py:346-352, GEGLU attention.py:94-101,
//     residual / AlphaBlender diffusionmodules/util.py:358-370).
//
// CTA = 128 output rows x BN output columns; 6 warps: warp0 = TMA producer, warp1 = MMA issuer + TMEM owner,
// warps 2..5 = epilogue (one TMEM lane quadrant each).  K loop = taps x ceil(K/64) stages through a
// STAGES-deep mbarrier ring.  Two CTAs co-reside per SM (smem/TMEM sized for it) so one CTA's epilogue overlaps
// the other's main loop.
#include <cuda.h>
#include <cuda_bf16.h>
#include <string.h>

#include "../../include/b200svd.h"
#include "common.h"
#include "ptx.cuh"

namespace b200 {

struct GemmDev {
  int32_t tap_off[B200SVD_MAX_TAPS][5];
  uint32_t taps, kblocks;
  uint32_t n;            // GEMM N (weight rows)
  uint32_t n_tiles;
  uint32_t m_ext[3];
  uint32_t m_lb[3];      // log2 of box
  uint32_t m_tiles[3];
  uint32_t m_adim[3];
  int64_t out_rs[3];
  void* out;
  int64_t ldo;
  int32_t out_fp32;
  const float* bias;
  const float* fvec;
  int64_t ldf;
  uint32_t rows_per_frame;
  int32_t act;
  float s_acc;
  const __nv_bfloat16* res1;
  int64_t ld1;
  float s1;
  const __nv_bfloat16* res2;
  int64_t ld2;
  float s2;
};

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB

template <int BN>
struct TileCfg {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (BN >= 128) ? 3 : 4;
  static constexpr int TMEM_COLS = (BN <= 32) ? 32 : (BN <= 64) ? 64 : (BN <= 128) ? 128 : 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(192, 2)
mtgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmDev p) {
  using Cfg = TileCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment required by SWIZZLE_128B operands
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile coordinates: n tile fastest so that CTAs sharing an A tile run back to back (A is re-read from L2)
  const uint32_t n_tile = blockIdx.x % p.n_tiles;
  uint32_t mt = blockIdx.x / p.n_tiles;
  const uint32_t t1 = mt % p.m_tiles[0];
  mt /= p.m_tiles[0];
  const uint32_t t2 = mt % p.m_tiles[1];
  const uint32_t t3 = mt / p.m_tiles[1];
  const uint32_t mb1 = t1 << p.m_lb[0], mb2 = t2 << p.m_lb[1], mb3 = t3 << p.m_lb[2];
  const uint32_t n0 = n_tile * BN;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t total_iters = p.taps * p.kblocks;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int base[5] = {0, 0, 0, 0, 0};
      base[p.m_adim[0]] += (int)mb1;
      base[p.m_adim[1]] += (int)mb2;
      base[p.m_adim[2]] += (int)mb3;
      uint32_t it = 0;
      for (uint32_t tap = 0; tap < p.taps; ++tap) {
        const int c0 = p.tap_off[tap][0];
        const int c1 = base[1] + p.tap_off[tap][1];
        const int c2 = base[2] + p.tap_off[tap][2];
        const int c3 = base[3] + p.tap_off[tap][3];
        const int c4 = base[4] + p.tap_off[tap][4];
        for (uint32_t kb = 0; kb < p.kblocks; ++kb, ++it) {
          const uint32_t s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
          tma_load_5d(sa, &tmA, &full_bar[s], c0 + (int)(kb * BK), c1, c2, c3, c4);
          tma_load_3d(sb, &tmB, &full_bar[s], (int)(kb * BK), (int)n0, (int)tap);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = make_idesc_f16(BM, BN, /*bf16*/ 1, 0, 0);
      for (uint32_t it = 0; it < total_iters; ++it) {
        const uint32_t s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint64_t adesc = smem_desc_k_sw128(sa);
        const uint64_t bdesc = smem_desc_k_sw128(sa + A_STAGE_BYTES);
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
          // advance 16 elements (32 B) along K inside the 128B swizzle atom: +2 in the (addr>>4) field
          umma_f16_ss(tmem_base, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc,
                      (it > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs have read it
      }
      umma_commit(accum_bar);  // accumulator complete
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps (2..5) =====================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int r = q * 32 + lane;
    const uint32_t lb1 = p.m_lb[0], lb2 = p.m_lb[1];
    const uint32_t m1 = mb1 + (r & ((1u << lb1) - 1));
    const uint32_t m2 = mb2 + ((r >> lb1) & ((1u << lb2) - 1));
    const uint32_t m3 = mb3 + (r >> (lb1 + lb2));
    const bool valid = (m1 < p.m_ext[0]) && (m2 < p.m_ext[1]) && (m3 < p.m_ext[2]);
    const int64_t row = (int64_t)m1 * p.out_rs[0] + (int64_t)m2 * p.out_rs[1] + (int64_t)m3 * p.out_rs[2];
    const float* fv = nullptr;
    if (p.fvec != nullptr && valid) fv = p.fvec + (row / p.rows_per_frame) * p.ldf;

    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool geglu = (p.act == B200SVD_ACT_GEGLU);
    const int ncols = geglu ? BN / 2 : BN;
    const uint32_t n_out = geglu ? p.n / 2 : p.n;
    const uint32_t ocol0 = geglu ? n_tile * (BN / 2) : n0;

    for (int c0 = 0; c0 < ncols; c0 += 16) {
      if (ocol0 + c0 >= n_out) break;  // warp-uniform
      uint32_t va[16];
      uint32_t vg[16];
      tmem_ld16(tlane + (uint32_t)c0, va);
      if (geglu) tmem_ld16(tlane + (uint32_t)(BN / 2 + c0), vg);
      tmem_ld_wait();
      if (!valid) continue;
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(va[j]);
      const uint32_t wcol = n0 + c0;  // weight-row index of the (first) accumulator column
      const uint32_t ocol = ocol0 + c0;
      const bool full = (ocol + 16 <= n_out);
      if (p.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (full || ocol + j < n_out) v[j] += __ldg(p.bias + wcol + j);
      }
      if (fv != nullptr) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (full || ocol + j < n_out) v[j] += __ldg(fv + ocol + j);
      }
      if (p.act == B200SVD_ACT_SILU) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = silu_f(v[j]);
      } else if (p.act == B200SVD_ACT_GELU) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = gelu_f(v[j]);
      } else if (geglu) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float g = __uint_as_float(vg[j]);
          if (p.bias != nullptr) g += __ldg(p.bias + wcol + BN / 2 + j);
          v[j] = v[j] * gelu_f(g);
        }
      }
      if (p.s_acc != 1.0f) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] *= p.s_acc;
      }
      if (p.res1 != nullptr) {
        const __nv_bfloat16* rp = p.res1 + row * p.ld1 + ocol;
        if (full) {
          const uint4 a = __ldg(reinterpret_cast<const uint4*>(rp));
          const uint4 b = __ldg(reinterpret_cast<const uint4*>(rp) + 1);
          const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[2 * j] += p.s1 * bf16_lo(w[j]);
            v[2 * j + 1] += p.s1 * bf16_hi(w[j]);
          }
        } else {
          for (int j = 0; j < 16; ++j)
            if (ocol + j < n_out) v[j] += p.s1 * __bfloat162float(rp[j]);
        }
      }
      if (p.res2 != nullptr) {
        const __nv_bfloat16* rp = p.res2 + row * p.ld2 + ocol;
        if (full) {
          const uint4 a = __ldg(reinterpret_cast<const uint4*>(rp));
          const uint4 b = __ldg(reinterpret_cast<const uint4*>(rp) + 1);
          const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[2 * j] += p.s2 * bf16_lo(w[j]);
            v[2 * j + 1] += p.s2 * bf16_hi(w[j]);
          }
        } else {
          for (int j = 0; j < 16; ++j)
            if (ocol + j < n_out) v[j] += p.s2 * __bfloat162float(rp[j]);
        }
      }
      if (p.out_fp32) {
        float* op = reinterpret_cast<float*>(p.out) + row * p.ldo + ocol;
        if (full) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            reinterpret_cast<float4*>(op)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
          for (int j = 0; j < 16; ++j)
            if (ocol + j < n_out) op[j] = v[j];
        }
      } else {
        __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ldo + ocol;
        if (full) {
          uint4 o0, o1;
          o0.x = pack_bf16x2(v[0], v[1]);
          o0.y = pack_bf16x2(v[2], v[3]);
          o0.z = pack_bf16x2(v[4], v[5]);
          o0.w = pack_bf16x2(v[6], v[7]);
          o1.x = pack_bf16x2(v[8], v[9]);
          o1.y = pack_bf16x2(v[10], v[11]);
          o1.z = pack_bf16x2(v[12], v[13]);
          o1.w = pack_bf16x2(v[14], v[15]);
          reinterpret_cast<uint4*>(op)[0] = o0;
          reinterpret_cast<uint4*>(op)[1] = o1;
        } else {
          for (int j = 0; j < 16; ++j)
            if (ocol + j < n_out) op[j] = __float2bfloat16(v[j]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

static int ilog2_exact(uint32_t v) {
  if (v == 0 || (v & (v - 1)) != 0) return -1;
  int l = 0;
  while ((1u << l) < v) ++l;
  return l;
}

template <int BN>
static int launch(const b200svd_gemm_params* p, const CUtensorMap& tmA, const GemmDev& d, cudaStream_t st) {
  using Cfg = TileCfg<BN>;
  // weights [taps][n][k] -> TMA dims (k, n, taps)
  CUtensorMap tmB;
  uint64_t bd[3] = {p->k, p->n, p->taps};
  uint64_t bs[2] = {(uint64_t)p->k * 2, (uint64_t)p->k * 2 * p->n};
  uint32_t bb[3] = {64, (uint32_t)BN, 1};
  if (encode_tmap_bf16(&tmB, p->w_ptr, 3, bd, bs, bb)) return 1;
  GemmDev dd = d;
  dd.n_tiles = (p->n + BN - 1) / BN;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(mtgemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mtgemm)");
    attr_set = true;
  }
  const uint64_t grid = (uint64_t)dd.n_tiles * dd.m_tiles[0] * dd.m_tiles[1] * dd.m_tiles[2];
  if (grid == 0 || grid > 0x7FFFFFFFull) {
    set_error("mtgemm: bad grid size %llu", (unsigned long long)grid);
    return 1;
  }
  mtgemm_kernel<BN><<<(unsigned)grid, 192, Cfg::SMEM_BYTES, st>>>(tmA, tmB, dd);
  B200_CHECK_LAUNCH("mtgemm launch");
  return 0;
}

}  // namespace b200

extern "C" int b200svd_gemm(const b200svd_gemm_params* p, void* stream) {
  using namespace b200;
  if (p == nullptr) {
    set_error("b200svd_gemm: null params");
    return 1;
  }
  if (p->taps == 0 || p->taps > B200SVD_MAX_TAPS) {
    set_error("b200svd_gemm: taps=%u out of range", p->taps);
    return 1;
  }
  if (p->a_box[0] != 64 || (uint64_t)p->a_box[1] * p->a_box[2] * p->a_box[3] * p->a_box[4] != 128) {
    set_error("b200svd_gemm: A box must be 64 x (product 128), got [%u,%u,%u,%u,%u]", p->a_box[0], p->a_box[1],
              p->a_box[2], p->a_box[3], p->a_box[4]);
    return 1;
  }
  if (p->k % 8 != 0) {
    set_error("b200svd_gemm: K=%u must be a multiple of 8 (16-byte TMA rows)", p->k);
    return 1;
  }
  GemmDev d;
  memset(&d, 0, sizeof(d));
  for (uint32_t t = 0; t < p->taps; ++t)
    for (int i = 0; i < 5; ++i) d.tap_off[t][i] = p->tap_off[t][i];
  d.taps = p->taps;
  d.kblocks = (p->k + 63) / 64;
  d.n = p->n;
  uint32_t prod = 1;
  for (int i = 0; i < 3; ++i) {
    int lb = ilog2_exact(p->m_box[i]);
    if (lb < 0) {
      set_error("b200svd_gemm: m_box[%d]=%u is not a power of two", i, p->m_box[i]);
      return 1;
    }
    if (p->m_adim[i] < 1 || p->m_adim[i] > 4) {
      set_error("b200svd_gemm: m_adim[%d]=%u must be in 1..4", i, p->m_adim[i]);
      return 1;
    }
    if (p->a_box[p->m_adim[i]] != p->m_box[i]) {
      set_error("b200svd_gemm: a_box[%u]=%u != m_box[%d]=%u", p->m_adim[i], p->a_box[p->m_adim[i]], i, p->m_box[i]);
      return 1;
    }
    if (i > 0 && p->m_adim[i] <= p->m_adim[i - 1]) {
      set_error("b200svd_gemm: m_adim must be increasing");
      return 1;
    }
    prod *= p->m_box[i];
    d.m_ext[i] = p->m_ext[i];
    d.m_lb[i] = (uint32_t)lb;
    d.m_tiles[i] = (p->m_ext[i] + p->m_box[i] - 1) / p->m_box[i];
    d.m_adim[i] = p->m_adim[i];
    d.out_rs[i] = p->out_rs[i];
  }
  if (prod != 128) {
    set_error("b200svd_gemm: m_box product %u != 128", prod);
    return 1;
  }
  d.out = p->out;
  d.ldo = p->ldo;
  d.out_fp32 = p->out_fp32;
  d.bias = p->bias;
  d.fvec = p->fvec;
  d.ldf = p->ldf;
  d.rows_per_frame = p->rows_per_frame ? p->rows_per_frame : 1;
  d.act = p->act;
  d.s_acc = p->s_acc;
  d.res1 = reinterpret_cast<const __nv_bfloat16*>(p->res1);
  d.ld1 = p->ld1;
  d.s1 = p->s1;
  d.res2 = reinterpret_cast<const __nv_bfloat16*>(p->res2);
  d.ld2 = p->ld2;
  d.s2 = p->s2;

  int bn = p->bn;
  if (bn == 0) {
    if (p->n <= 32) bn = 32;
    else if (p->n <= 64) bn = 64;
    else if (p->n % 160 == 0) bn = 160;
    else bn = 128;
  }
  if (p->act == B200SVD_ACT_GEGLU && (p->n % bn) != 0) {
    set_error("b200svd_gemm: GEGLU needs n (%u) divisible by the N tile (%d)", p->n, bn);
    return 1;
  }
  const bool vec_ok = (p->ldo % 8 == 0) && (p->res1 == nullptr || p->ld1 % 8 == 0) &&
                      (p->res2 == nullptr || p->ld2 % 8 == 0);
  if (!vec_ok) {
    set_error("b200svd_gemm: leading dimensions must be multiples of 8 elements");
    return 1;
  }

  CUtensorMap tmA;
  if (encode_tmap_bf16(&tmA, p->a_ptr, 5, p->a_dims, p->a_strides, p->a_box)) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (bn) {
    case 32: return launch<32>(p, tmA, d, st);
    case 64: return launch<64>(p, tmA, d, st);
    case 128: return launch<128>(p, tmA, d, st);
    case 160: return launch<160>(p, tmA, d, st);
    default: set_error("b200svd_gemm: unsupported N tile %d", bn); return 1;
  }
}
