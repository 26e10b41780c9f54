// Multi-tap tensor-core GEMM for sm_100a (v4).
// TMA (rank-5 activation view) -> smem (128B swizzle) -> tcgen05.mma (fp32 accumulators in TMEM, double buffered)
// -> fused epilogue -> swizzled smem staging -> TMA store.  See include/b200svd.h for the contract.
//
// Replaces, underneath StreamingWrapper.forward (reference code/models/diffusion/wrappers.py:23-78):
//   nn.Linear            code/models/svd/sgm/modules/attention.py:94-120,262-351, video_attention.py:23-168
//   Conv2d 3x3 (s1, s2)  code/models/svd/sgm/modules/diffusionmodules/openaimodel.py:107-207,257-305
//   Conv3d (3,1,1)       code/models/diffusion/video_model.py:46-59 (ResBlock dims=3)
//   + their elementwise neighbours (bias, emb add openaimodel.py:346-352, GEGLU attention.py:94-101,
//     residual / AlphaBlender diffusionmodules/util.py:358-370).
//
// Persistent CTAs (one per SM; CTA pairs for the 2-SM tiles) walk a contiguous run of output tiles (N tile fastest).
// 18 warps:
//   warp 0       TMA producer for A/B (stage ring runs ahead across tiles)
//   warp 1       MMA issuer + TMEM owner (two accumulator buffers: tile i's epilogue overlaps tile i+1's main loop)
//   warps 2..17  epilogue: the CTA's output is a stream of (tile, 64-column sub-tile) blocks; each warp owns one TMEM
//                lane quadrant of every 4th block, start to finish, and never synchronises with another warp:
//                residual block -> (TMA) -> private 4 KB staging block; TMEM -> registers -> bias / activation /
//                residual (in place) -> 128B/64B-swizzled staging (bank-conflict free) -> one TMA store (hardware
//                clips ragged edges).
// 2-SM mode (template PAIR): a cluster of two CTAs computes a 256 x BN tile with tcgen05 cta_group::2 MMAs issued by
// the leader CTA; each CTA stages its own 128 A rows and HALF of the B tile, commits are multicast to both CTAs.
// Why (ncu, profiles/r01_ncu_gemm_*): v2 was bound by epilogue instruction issue and bank conflicts; v3's 16-column
// per-warp work items spent >90% of issued instructions on waits and index arithmetic and kept the four warps of a
// quadrant in lock-step through a named barrier; the large-K layers were bound by shared-memory bandwidth per FLOP,
// which the pair tiles halve for B.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/b200svd.h"
#include "common.h"
#include "ptx.cuh"

// Defaults of the round-2 epilogue specialisations (each also has an environment switch for A/B runs).  On since the GPU
// parity suite passed with them (tests/test_gemm_gpu.py: 216 cases; denoiser / chain / VAE goldens:
// profiles/r02_gputest_parity_newepi.log).
#define B200SVD_DEFAULT_LEAN_EPI 1
#define B200SVD_DEFAULT_GEGLU_EPI 1

namespace b200 {

struct GemmDev {
  int32_t tap_off[B200SVD_MAX_TAPS][5];
  uint32_t taps, kblocks;
  uint32_t n;  // GEMM N (weight rows)
  uint32_t n_tiles;
  uint32_t total_tiles, tiles_per_cta;
  uint32_t a_half_dim, a_half_size;  // QUAD: A-view dim whose box is halved, and the half extent
  uint32_t m_ext[3];
  uint32_t m_lb[3];  // log2 of box
  uint32_t m_tiles[3];
  uint32_t m_adim[3];
  int64_t out_rs[3];
  void* out;
  int64_t ldo;
  int32_t out_fp32;
  int32_t tma_epi;  // 1: staged TMA-store epilogue (bf16 out), 0: direct per-row stores (fp32 out)
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
  float* gn_part;           // GroupNorm partial sums of the output (EPI == 2 only), see b200svd.h
  int32_t* gn_slot_sample;
  int64_t gn_ld;
  uint32_t gn_rows;
};

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int EPI_WARPS = 16;
constexpr int FIRST_EPI_WARP = 2;
constexpr int NUM_THREADS = (FIRST_EPI_WARP + EPI_WARPS) * 32;  // 576 (register cap 112 per thread)
constexpr int OUT_BUF_BYTES = 32 * 64 * 2;   // one quadrant (32 rows) x 64 columns

template <int BN, int CL>
struct TileCfg {
  static constexpr bool PAIR = CL >= 2;
  // PAIR: two CTAs of a cluster compute a 256 x BN tile with one cta_group::2 MMA stream; each CTA stages its own
  // 128 A rows and HALF of the B tile (the tensor cores of the pair share B), which halves B traffic per FLOP.
  static constexpr int B_ROWS = PAIR ? BN / 2 : BN;
  static constexpr int B_STAGE_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  // BN = 320 (pair tiles only): every layer width of the network is a multiple of 320, so there is no ragged N tile,
  // and the tile has the fewest bytes into the SM per FLOP (36 KB per 128x320x64 MACs).  tcgen05.mma takes N <= 256:
  // two MMAs of N = 160 per k step, each CTA stages two 80-row pieces of B.  320 fp32 columns fit TMEM only once, so the
  // accumulator is single-buffered (the epilogue of tile i is exposed) — used where the main loop is long.
  static constexpr int NMMA = (BN > 256) ? 2 : 1;
  static constexpr int BN_MMA = BN / NMMA;
  static constexpr int NBUF = (BN > 256) ? 1 : 2;
  static constexpr int ACC_STRIDE = (BN <= 32) ? 32 : (BN <= 64) ? 64 : (BN <= 128) ? 128 : (BN <= 256) ? 256 : 512;
  static constexpr int TMEM_COLS = NBUF * ACC_STRIDE;
  static_assert(BN <= 256 || PAIR, "the 320-wide tile exists as a 2-SM tile only");
  static constexpr int NSUB = (BN + 63) / 64;
  static constexpr int OUT_BYTES = EPI_WARPS * OUT_BUF_BYTES;  // one private 32-row x 64-column staging block per warp
  static constexpr int FIXED_BYTES = OUT_BYTES + 512;
  static constexpr int STAGES_FIT = (232448 - FIXED_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 6 ? 6 : STAGES_FIT;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + FIXED_BYTES;
  static_assert(STAGES >= 2, "pipeline depth");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  static_assert(STAGE_BYTES % 1024 == 0, "stage alignment");
};

// `tile` enumerates (N tile fastest, then M tile) — in PAIR mode (M-tile PAIR); mrank selects this CTA's M tile of the
// pair.  An M tile index past the end decodes to coordinates beyond the extents (all rows out of bounds).
__device__ __forceinline__ uint32_t m_tile_index(const GemmDev& p, uint32_t tile, uint32_t mmul, uint32_t mrank) {
  return (tile / p.n_tiles) * mmul + mrank;
}
__device__ __forceinline__ void decode_tile(const GemmDev& p, uint32_t tile, uint32_t mmul, uint32_t mrank,
                                            uint32_t nmul, uint32_t nrank, uint32_t& n_tile, uint32_t& mb1,
                                            uint32_t& mb2, uint32_t& mb3) {
  n_tile = (tile % p.n_tiles) * nmul + nrank;  // p.n_tiles counts N work items (pairs of N tiles in QUAD mode)
  uint32_t mt = (tile / p.n_tiles) * mmul + mrank;
  const uint32_t t1 = mt % p.m_tiles[0];
  mt /= p.m_tiles[0];
  const uint32_t t2 = mt % p.m_tiles[1];
  const uint32_t t3 = mt / p.m_tiles[1];
  mb1 = t1 << p.m_lb[0];
  mb2 = t2 << p.m_lb[1];
  mb3 = t3 << p.m_lb[2];
}

// EPI: 0 = generic epilogue (bias / per-frame vector / activation / two residuals, all run-time switches);
//      1 = GEGLU with bias and nothing else (the FF1 projections, 17 % of the step): compile-time specialised, packed
//          FFMA2 math (ptx.cuh geglu2) — the generic loop spent ~30 issue slots per output there and bound the K = 320
//          layer at 0.45 of the tensor rate (profiles/r01_ncu_v4_geglu_details.txt).
template <int BN, int CL, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
mtgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAh,
              const __grid_constant__ CUtensorMap tmB,
              const __grid_constant__ CUtensorMap tmO64, const __grid_constant__ CUtensorMap tmO32,
              const __grid_constant__ CUtensorMap tmR64, const __grid_constant__ CUtensorMap tmR32, const GemmDev p) {
  using Cfg = TileCfg<BN, CL>;
  constexpr bool PAIR = CL >= 2;  // cta_group::2 tiles
  constexpr bool QUAD = CL == 4;  // two pairs on the same A rows (adjacent N tiles): A halves multicast between them
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ __align__(1024) uint8_t smem[];  // SWIZZLE_128B operands need 1024-byte alignment
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* out_base = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(out_base + Cfg::OUT_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;    // [2]
  uint64_t* acc_empty = acc_full + 2;         // [2]
  uint64_t* res_bar = acc_empty + 2;          // [EPI_WARPS] residual block landed in the warp's staging block
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + EPI_WARPS);
  static_assert((2 * STAGES + 4 + EPI_WARPS) * 8 + 4 <= 512, "barrier area");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const uint32_t rank4 = PAIR ? cluster_ctarank() : 0u;
  const uint32_t rank = rank4 & 1u;    // M tile of the pair; 0 = leader CTA (issues the MMAs)
  const uint32_t npair = rank4 >> 1;   // QUAD: which of the two N tiles this pair computes
  const uint32_t tile_begin = (blockIdx.x / CL) * p.tiles_per_cta;
  const uint32_t tile_end = min(tile_begin + p.tiles_per_cta, p.total_tiles);
  const bool geglu = (p.act == B200SVD_ACT_GEGLU);
  const uint32_t n_out = geglu ? p.n / 2 : p.n;
  const uint32_t tile_out_w = geglu ? (uint32_t)BN / 2 : (uint32_t)BN;  // output columns per tile
  const uint32_t nsub_out = (tile_out_w + 63) / 64;
  const bool use_res_ring = p.tma_epi && p.res1 != nullptr;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (p.tma_epi) {
      prefetch_tmap(&tmO64);
      prefetch_tmap(&tmO32);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], QUAD ? 2 : 1);  // QUAD: both pairs' MMAs must have read the stage (A is shared)
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], PAIR ? 2 * EPI_WARPS : EPI_WARPS);  // one arrive per epilogue warp (of both CTAs)
    }
    for (int b = 0; b < EPI_WARPS; ++b) mbar_init(&res_bar[b], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) {
      tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();  // peer barriers are initialised before any remote arrive / multicast commit
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t iters_per_tile = p.taps * p.kblocks;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer (A, B) =====================
      uint32_t it = 0;
      for (uint32_t tile = tile_begin; tile < tile_end; ++tile) {
        uint32_t n_tile, mb1, mb2, mb3;
        decode_tile(p, tile, PAIR ? 2u : 1u, rank, QUAD ? 2u : 1u, npair, n_tile, mb1, mb2, mb3);
        int base[5] = {0, 0, 0, 0, 0};
        base[p.m_adim[0]] += (int)mb1;
        base[p.m_adim[1]] += (int)mb2;
        base[p.m_adim[2]] += (int)mb3;
        const int n0 = (int)(n_tile * BN);
        for (uint32_t tap = 0; tap < p.taps; ++tap) {
          const int c0 = p.tap_off[tap][0];
          const int c1 = base[1] + p.tap_off[tap][1];
          const int c2 = base[2] + p.tap_off[tap][2];
          const int c3 = base[3] + p.tap_off[tap][3];
          const int c4 = base[4] + p.tap_off[tap][4];
          for (uint32_t kb = 0; kb < p.kblocks; ++kb, ++it) {
            const uint32_t s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait_parked(&empty_bar[s], ph ^ 1);
            uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
            if (PAIR) {
              // both CTAs' loads are credited to the leader's barrier; the leader expects the bytes of the pair
              if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * Cfg::STAGE_BYTES);
              if (QUAD) {
                // this CTA fetches one 64-row half of the A tile and multicasts it to itself and to its twin in the
                // other pair (same M tile, other N tile): half the L2->SM traffic for A
                int cc[5] = {c0 + (int)(kb * BK), c1, c2, c3, c4};
                cc[p.a_half_dim] += (int)(npair * p.a_half_size);
                tma_load_5d_2sm_mc(sa + npair * (A_STAGE_BYTES / 2), &tmAh, &full_bar[s], cc[0], cc[1], cc[2], cc[3],
                                   cc[4], (uint16_t)((1u << rank) | (1u << (rank + 2))));
              } else {
                tma_load_5d_2sm(sa, &tmA, &full_bar[s], c0 + (int)(kb * BK), c1, c2, c3, c4);
              }
#pragma unroll
              for (int hh = 0; hh < Cfg::NMMA; ++hh)
                tma_load_3d_2sm(sa + A_STAGE_BYTES + hh * (Cfg::BN_MMA / 2) * 128, &tmB, &full_bar[s], (int)(kb * BK),
                                n0 + hh * Cfg::BN_MMA + (int)(rank * (Cfg::BN_MMA / 2)), (int)tap);
            } else {
              mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
              tma_load_5d(sa, &tmA, &full_bar[s], c0 + (int)(kb * BK), c1, c2, c3, c4);
              tma_load_3d(sa + A_STAGE_BYTES, &tmB, &full_bar[s], (int)(kb * BK), n0, (int)tap);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ===================== MMA issuer (leader CTA only in PAIR mode) =====================
      constexpr uint32_t idesc = make_idesc_f16(PAIR ? 2 * BM : BM, Cfg::BN_MMA, /*bf16*/ 1, 0, 0);
      uint32_t it = 0, tcount = 0;
      for (uint32_t tile = tile_begin; tile < tile_end; ++tile, ++tcount) {
        const uint32_t b = Cfg::NBUF == 2 ? (tcount & 1) : 0u;
        const uint32_t bpar = Cfg::NBUF == 2 ? ((tcount >> 1) & 1) : (tcount & 1);
        mbar_wait_parked(&acc_empty[b], bpar ^ 1);  // epilogue has drained this accumulator buffer
        tc_fence_after();
        const uint32_t tacc = tmem_base + b * Cfg::ACC_STRIDE;
        for (uint32_t i = 0; i < iters_per_tile; ++i, ++it) {
          const uint32_t s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
          const uint64_t adesc = smem_desc_k_sw128(sa);
#pragma unroll
          for (int hh = 0; hh < Cfg::NMMA; ++hh) {
            const uint64_t bdesc = smem_desc_k_sw128(sa + A_STAGE_BYTES + hh * (PAIR ? Cfg::BN_MMA / 2 : Cfg::BN_MMA) * 128);
            const uint32_t tacc_h = tacc + (uint32_t)(hh * Cfg::BN_MMA);
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) {
              // advance 16 elements (32 B) along K inside the 128B swizzle atom: +2 in the (addr>>4) field
              if (PAIR)
                umma_f16_ss_2sm(tacc_h, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (i > 0 || kk > 0) ? 1u : 0u);
              else
                umma_f16_ss(tacc_h, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (i > 0 || kk > 0) ? 1u : 0u);
            }
          }
          // frees the smem stage (in both CTAs of a pair) once these MMAs have read it
          if (QUAD) umma_commit_2sm(&empty_bar[s], (uint16_t)0xF);
          else if (PAIR) umma_commit_2sm(&empty_bar[s]);
          else umma_commit(&empty_bar[s]);
        }
        // accumulator complete (signalled to the epilogue warps of both CTAs of a pair)
        if (PAIR) umma_commit_2sm(&acc_full[b], (uint16_t)(3u << (2 * npair)));
        else umma_commit(&acc_full[b]);
      }
    }
    __syncwarp();
  } else if (warp >= FIRST_EPI_WARP) {
    // ===================== epilogue warps (2..17) =====================
    // The output of the CTA is a stream of blocks (tile, 64-column sub-tile); warp (q, g) owns TMEM lane quadrant q of
    // every block n with n % 4 == g, start to finish: residual block -> (TMA) -> private staging block; TMEM ->
    // registers -> math (+ residual, in place) -> staging -> TMA store.  Warps never synchronise with each other, so
    // the four warps of a scheduler hide each other's latencies.  Every warp observes acc_full and arrives on
    // acc_empty for EVERY tile (also those it owns no block of): mbarrier parity waits are only unambiguous for a
    // waiter that is at most one phase behind.
    const int e = warp - FIRST_EPI_WARP;
    const int q = warp & 3;  // TMEM lane quadrant (hardware: warp % 4)
    const int g = e >> 2;    // block phase; warps 4g+2 .. 4g+5 cover the four quadrants once
    const int r = q * 32 + lane;
    const uint32_t lb1 = p.m_lb[0], lb2 = p.m_lb[1];
    const uint32_t qoff = (uint32_t)(q * 32);
    const uint32_t qo1 = qoff & ((1u << lb1) - 1);
    const uint32_t qo2 = (qoff >> lb1) & ((1u << lb2) - 1);
    const uint32_t qo3 = qoff >> (lb1 + lb2);
    uint8_t* ob = out_base + e * OUT_BUF_BYTES;
    uint64_t* my_res = &res_bar[e];
    const bool bias_vec = p.bias != nullptr && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0;
    const bool fvec_vec = p.fvec != nullptr && (reinterpret_cast<uintptr_t>(p.fvec) & 15) == 0 && (p.ldf & 3) == 0;
    const uint32_t num_tiles = tile_end > tile_begin ? tile_end - tile_begin : 0u;

    auto acc_buf = [](uint32_t t) -> uint32_t { return Cfg::NBUF == 2 ? (t & 1u) : 0u; };
    auto acc_par = [](uint32_t t) -> uint32_t { return Cfg::NBUF == 2 ? ((t >> 1) & 1u) : (t & 1u); };
    uint32_t tcount = 0, s = (uint32_t)g, tpass = 0, rcount = 0;
    while (s >= nsub_out) {
      s -= nsub_out;
      ++tcount;
    }
    uint32_t n_tile = 0, mb1 = 0, mb2 = 0, mb3 = 0;
    // residual block of (tile, s) -> staging block; the previous store must have finished reading it
    auto prefetch_residual = [&]() {
      const uint32_t ocol0 = n_tile * tile_out_w + 64 * s;
      if (use_res_ring && ocol0 < n_out && lane == 0) {
        const uint32_t rem = tile_out_w - 64 * s;
        const uint32_t subw = rem >= 64 ? 64u : rem;
        tma_store_wait_read<0>();
        mbar_expect_tx(my_res, subw * 2 * 32);
        tma_load_5d(ob, subw == 64 ? &tmR64 : &tmR32, my_res, (int)ocol0, (int)(mb1 + qo1), (int)(mb2 + qo2),
                    (int)(mb3 + qo3), 0);
      }
    };
    if (tcount < num_tiles) {
      decode_tile(p, tile_begin + tcount, PAIR ? 2u : 1u, rank, QUAD ? 2u : 1u, npair, n_tile, mb1, mb2, mb3);
      prefetch_residual();
    }
    while (tcount < num_tiles) {
      const uint32_t n0 = n_tile * BN;
      const uint32_t otile0 = n_tile * tile_out_w;
      const uint32_t m1 = mb1 + (r & ((1u << lb1) - 1));
      const uint32_t m2 = mb2 + ((r >> lb1) & ((1u << lb2) - 1));
      const uint32_t m3 = mb3 + (r >> (lb1 + lb2));
      const bool valid = (m1 < p.m_ext[0]) && (m2 < p.m_ext[1]) && (m3 < p.m_ext[2]);
      const int64_t row = (int64_t)m1 * p.out_rs[0] + (int64_t)m2 * p.out_rs[1] + (int64_t)m3 * p.out_rs[2];
      const float* fv = nullptr;
      if (p.fvec != nullptr && valid) fv = p.fvec + (int64_t)((uint32_t)row / p.rows_per_frame) * p.ldf;
      const uint32_t ocol0 = otile0 + 64 * s;
      const bool skip = ocol0 >= n_out;  // block beyond the ragged N edge (warp-uniform)
      const uint32_t rem = tile_out_w - 64 * s;
      const uint32_t subw = rem >= 64 ? 64u : rem;  // 64 or 32 (direct path: also narrower tiles)

      // release the accumulators of the tiles this warp owns no block of
      for (; tpass < tcount; ++tpass) {
        mbar_wait_parked(&acc_full[acc_buf(tpass)], acc_par(tpass));
        if (lane == 0) {
          if (PAIR) mbar_arrive_leader(&acc_empty[acc_buf(tpass)]);
          else mbar_arrive(&acc_empty[acc_buf(tpass)]);
        }
      }
      const uint32_t b = acc_buf(tcount);
      mbar_wait_parked(&acc_full[b], acc_par(tcount));
      tc_fence_after();
      const uint32_t tlane = tmem_base + b * Cfg::ACC_STRIDE + ((uint32_t)(q * 32) << 16);

      // per-16-column math shared by both epilogue flavours; v = accumulators on entry, finished values on exit
      auto finish16 = [&](float (&v)[16], const uint32_t (&vg)[16], uint32_t wcol, uint32_t ocol, bool full) {
        if (p.bias != nullptr) {
          if (full && bias_vec) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + wcol) + j4);
              v[4 * j4] += bb.x; v[4 * j4 + 1] += bb.y; v[4 * j4 + 2] += bb.z; v[4 * j4 + 3] += bb.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (full || ocol + j < n_out) v[j] += __ldg(p.bias + wcol + j);
          }
        }
        if (fv != nullptr) {
          if (full && fvec_vec) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(fv + ocol) + j4);
              v[4 * j4] += bb.x; v[4 * j4 + 1] += bb.y; v[4 * j4 + 2] += bb.z; v[4 * j4 + 3] += bb.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (full || ocol + j < n_out) v[j] += __ldg(fv + ocol + j);
          }
        }
        if (p.act == B200SVD_ACT_SILU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = silu_fast(v[j]);
        } else if (p.act == B200SVD_ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = gelu_fast(v[j]);
        } else if (geglu) {
          float gt[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) gt[j] = __uint_as_float(vg[j]);
          if (p.bias != nullptr) {
            if (full && bias_vec) {
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + wcol + BN / 2) + j4);
                gt[4 * j4] += bb.x; gt[4 * j4 + 1] += bb.y; gt[4 * j4 + 2] += bb.z; gt[4 * j4 + 3] += bb.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (full || ocol + j < n_out) gt[j] += __ldg(p.bias + wcol + BN / 2 + j);
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = v[j] * gelu_fast(gt[j]);
        }
        if (p.s_acc != 1.0f) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] *= p.s_acc;
        }
      };

      if (EPI == 1) {
        if (!skip) {
          // host guarantees: act == GEGLU, 16-byte aligned bias, no fvec / residuals, s_acc == 1, TMA-store epilogue,
          // n % BN == 0 (every 64-column block is full)
          if (lane == 0) tma_store_wait_read<0>();  // the previous TMA store has finished reading the staging block
          __syncwarp();
          uint8_t* orow = ob + lane * 128;
          const uint32_t x = (uint32_t)(lane & 7);
          const float4* bvp = reinterpret_cast<const float4*>(p.bias + n0 + 64 * s);
          const float4* bgp = reinterpret_cast<const float4*>(p.bias + n0 + (uint32_t)(BN / 2) + 64 * s);
#pragma unroll
          for (uint32_t c = 0; c < 4; ++c) {
            uint32_t va[16], vg[16];
            tmem_ld16(tlane + 64 * s + 16 * c, va);
            tmem_ld16(tlane + (uint32_t)(BN / 2) + 64 * s + 16 * c, vg);
            tmem_ld_wait();
            uint32_t o[8];
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 bv = __ldg(bvp + 4 * c + j4);
              const float4 bg = __ldg(bgp + 4 * c + j4);
              const f32x2 v01 = add2(pack2u(va[4 * j4], va[4 * j4 + 1]), pack2(bv.x, bv.y));
              const f32x2 v23 = add2(pack2u(va[4 * j4 + 2], va[4 * j4 + 3]), pack2(bv.z, bv.w));
              const f32x2 g01 = add2(pack2u(vg[4 * j4], vg[4 * j4 + 1]), pack2(bg.x, bg.y));
              const f32x2 g23 = add2(pack2u(vg[4 * j4 + 2], vg[4 * j4 + 3]), pack2(bg.z, bg.w));
              float r0, r1, r2, r3;
              unpack2(geglu2(v01, g01), r0, r1);
              unpack2(geglu2(v23, g23), r2, r3);
              o[2 * j4] = pack_bf16x2(r0, r1);
              o[2 * j4 + 1] = pack_bf16x2(r2, r3);
            }
            *reinterpret_cast<uint4*>(orow + (((2 * c) ^ x) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uint4*>(orow + (((2 * c + 1) ^ x) << 4)) = make_uint4(o[4], o[5], o[6], o[7]);
          }
          fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA store
        }
      } else if (EPI == 2) {
        if (!skip) {
          // linear family without activation: acc (+ bias) (+ per-frame vector) (* s_acc) (+ s1 res1) (+ s2 res2), all
          // optional parts behind warp-uniform branches, arithmetic on packed pairs.  Host guarantees: TMA-store
          // epilogue, n_out % 16 == 0 (a 16-column chunk is either complete or entirely past the edge), 16-byte
          // aligned bias / fvec (ldf % 4 == 0).  ~4 issue slots per output instead of ~20.
          if (use_res_ring) {
            mbar_wait(my_res, rcount & 1);  // residual block has landed in the staging block
            ++rcount;
          } else {
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
          }
          uint8_t* orow;
          uint32_t x;
          if (subw == 64) {
            orow = ob + lane * 128;
            x = (uint32_t)(lane & 7);
          } else {
            orow = ob + lane * 64;
            x = (uint32_t)((lane >> 1) & 3);
          }
          const uint32_t nchunks = subw >> 4;
          const f32x2 s1v = splat2(p.s1), s2v = splat2(p.s2), sav = splat2(p.s_acc);
          const bool has_res2 = p.res2 != nullptr && valid;
#pragma unroll
          for (uint32_t c = 0; c < 4; ++c) {
            if (c < nchunks) {
              const uint32_t acol = 64 * s + 16 * c;
              const uint32_t ocol = ocol0 + 16 * c;
              uint32_t va[16];
              tmem_ld16(tlane + acol, va);
              tmem_ld_wait();
              f32x2 v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = pack2u(va[2 * j], va[2 * j + 1]);
              const bool inb = (ocol + 16) <= n_out;  // else: the whole chunk is past the N edge (TMA clips it)
              if (p.bias != nullptr && inb) {
                const float4* bp = reinterpret_cast<const float4*>(p.bias + n0 + acol);
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                  const float4 bb = __ldg(bp + j4);
                  v[2 * j4] = add2(v[2 * j4], pack2(bb.x, bb.y));
                  v[2 * j4 + 1] = add2(v[2 * j4 + 1], pack2(bb.z, bb.w));
                }
              }
              if (fv != nullptr && inb) {
                const float4* fp = reinterpret_cast<const float4*>(fv + ocol);
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                  const float4 bb = __ldg(fp + j4);
                  v[2 * j4] = add2(v[2 * j4], pack2(bb.x, bb.y));
                  v[2 * j4 + 1] = add2(v[2 * j4 + 1], pack2(bb.z, bb.w));
                }
              }
              if (p.s_acc != 1.0f) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = mul2(v[j], sav);
              }
              const uint32_t c_lo = 2 * c, c_hi = 2 * c + 1;
              if (use_res_ring) {
                const uint4 a = *reinterpret_cast<const uint4*>(orow + ((c_lo ^ x) << 4));
                const uint4 bq = *reinterpret_cast<const uint4*>(orow + ((c_hi ^ x) << 4));
                const uint32_t w[8] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = fma2(pack2(bf16_lo(w[j]), bf16_hi(w[j])), s1v, v[j]);
              }
              if (has_res2 && inb) {
                const uint4* rp = reinterpret_cast<const uint4*>(p.res2 + row * p.ld2 + ocol);
                const uint4 a = __ldg(rp);
                const uint4 bq = __ldg(rp + 1);
                const uint32_t w[8] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = fma2(pack2(bf16_lo(w[j]), bf16_hi(w[j])), s2v, v[j]);
              }
              uint32_t o[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float lo, hi;
                unpack2(v[j], lo, hi);
                o[j] = pack_bf16x2(lo, hi);
              }
              *reinterpret_cast<uint4*>(orow + ((c_lo ^ x) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
              *reinterpret_cast<uint4*>(orow + ((c_hi ^ x) << 4)) = make_uint4(o[4], o[5], o[6], o[7]);
              if (p.gn_part != nullptr && inb) {
                // per-column sum / sum of squares over the 32 rows of this quadrant (of the bf16-ROUNDED values, the
                // ones the consumer reads): 32-lane transpose-reduce, 16 + 16 shuffles; lane L ends with column
                // (L >> 1) & 15, even lanes store 16 x (sum, sumsq) = 128 contiguous bytes
                float a[16], b[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float lo = valid ? bf16_lo(o[j]) : 0.f, hi = valid ? bf16_hi(o[j]) : 0.f;
                  a[2 * j] = lo;
                  a[2 * j + 1] = hi;
                  b[2 * j] = lo * lo;
                  b[2 * j + 1] = hi * hi;
                }
#pragma unroll
                for (int w = 8, bit = 16; w >= 1; w >>= 1, bit >>= 1) {
                  const bool up = (lane & bit) != 0;
#pragma unroll
                  for (int i = 0; i < w; ++i) {
                    const float keep_a = up ? a[w + i] : a[i], send_a = up ? a[i] : a[w + i];
                    const float keep_b = up ? b[w + i] : b[i], send_b = up ? b[i] : b[w + i];
                    a[i] = keep_a + __shfl_xor_sync(0xffffffffu, send_a, bit);
                    b[i] = keep_b + __shfl_xor_sync(0xffffffffu, send_b, bit);
                  }
                }
                a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
                b[0] += __shfl_xor_sync(0xffffffffu, b[0], 1);
                if ((lane & 1) == 0) {
                  const uint32_t slot = m_tile_index(p, tile_begin + tcount, PAIR ? 2u : 1u, rank) * 4u + (uint32_t)q;
                  float2* dst = reinterpret_cast<float2*>(p.gn_part) + (int64_t)slot * p.gn_ld + ocol + ((lane >> 1) & 15);
                  *dst = make_float2(a[0], b[0]);
                }
              }
            }
          }
          if (p.gn_part != nullptr && lane == 0 && n_tile == 0 && s == 0) {
            // lane 0 owns the first row of the quadrant: if it is out of range, so is every row of the quadrant
            const uint32_t slot = m_tile_index(p, tile_begin + tcount, PAIR ? 2u : 1u, rank) * 4u + (uint32_t)q;
            p.gn_slot_sample[slot] = valid ? (int32_t)((uint32_t)row / p.gn_rows) : -1;
          }
          fence_proxy_async_smem();
        }
      } else if (!skip && p.tma_epi) {
        if (use_res_ring) {
          mbar_wait(my_res, rcount & 1);  // residual block has landed in the staging block
          ++rcount;
        } else {
          if (lane == 0) tma_store_wait_read<0>();  // the previous TMA store has finished reading the staging block
          __syncwarp();
        }
        uint8_t* orow;
        uint32_t x;  // swizzle phase of this thread's staging row
        if (subw == 64) {
          orow = ob + lane * 128;
          x = (uint32_t)(lane & 7);
        } else {
          orow = ob + lane * 64;
          x = (uint32_t)((lane >> 1) & 3);
        }
        const uint32_t nchunks = subw >> 4;
#pragma unroll 1
        for (uint32_t c = 0; c < nchunks; ++c) {
          const uint32_t acol = 64 * s + 16 * c;  // accumulator column (value half for GEGLU)
          uint32_t va[16], vg[16];
          tmem_ld16(tlane + acol, va);
          if (geglu) tmem_ld16(tlane + (uint32_t)(BN / 2) + acol, vg);
          tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(va[j]);
          const uint32_t ocol = ocol0 + 16 * c;
          const bool full = (ocol + 16) <= n_out;
          finish16(v, vg, n0 + acol, ocol, full);
          const uint32_t c_lo = 2 * c, c_hi = 2 * c + 1;  // 16-byte chunks of the row
          if (use_res_ring) {
            // residual chunk sits where the result will go (same box, same swizzle): read, add, overwrite
            const uint4 a = *reinterpret_cast<const uint4*>(orow + ((c_lo ^ x) << 4));
            const uint4 bq = *reinterpret_cast<const uint4*>(orow + ((c_hi ^ x) << 4));
            const uint32_t w[8] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              v[2 * j] += p.s1 * bf16_lo(w[j]);
              v[2 * j + 1] += p.s1 * bf16_hi(w[j]);
            }
          }
          if (p.res2 != nullptr && valid) {
            // second residual (only the temporal AlphaBlender GEMMs): read straight from global, own row
            const __nv_bfloat16* rp = p.res2 + row * p.ld2 + ocol;
            if (full) {
              const uint4 a = __ldg(reinterpret_cast<const uint4*>(rp));
              const uint4 bq = __ldg(reinterpret_cast<const uint4*>(rp) + 1);
              const uint32_t w[8] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w};
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
          *reinterpret_cast<uint4*>(orow + ((c_lo ^ x) << 4)) =
              make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
          *reinterpret_cast<uint4*>(orow + ((c_hi ^ x) << 4)) =
              make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]),
                         pack_bf16x2(v[14], v[15]));
        }
        fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA store
      } else if (!skip) {
        // direct epilogue: fp32 outputs (tiny GEMMs: embeddings, conv_out, attention scores)
        for (uint32_t c0 = 0; c0 < subw; c0 += 16) {
          const uint32_t acol = 64 * s + c0;
          if (otile0 + acol >= n_out) break;
          uint32_t va[16], vg[16];
          tmem_ld16(tlane + acol, va);
          if (geglu) tmem_ld16(tlane + (uint32_t)(BN / 2) + acol, vg);
          tmem_ld_wait();
          if (!valid) continue;
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(va[j]);
          const uint32_t ocol = otile0 + acol;
          const bool full = (ocol + 16 <= n_out);
          finish16(v, vg, n0 + acol, ocol, full);
          if (p.res1 != nullptr) {
            const __nv_bfloat16* rp = p.res1 + row * p.ld1 + ocol;
            for (int j = 0; j < 16; ++j)
              if (full || ocol + j < n_out) v[j] += p.s1 * __bfloat162float(rp[j]);
          }
          if (p.res2 != nullptr) {
            const __nv_bfloat16* rp = p.res2 + row * p.ld2 + ocol;
            for (int j = 0; j < 16; ++j)
              if (full || ocol + j < n_out) v[j] += p.s2 * __bfloat162float(rp[j]);
          }
          if (p.out_fp32) {
            float* op = reinterpret_cast<float*>(p.out) + row * p.ldo + ocol;
            for (int j = 0; j < 16; ++j)
              if (full || ocol + j < n_out) op[j] = v[j];
          } else {
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ldo + ocol;
            for (int j = 0; j < 16; ++j)
              if (full || ocol + j < n_out) op[j] = __float2bfloat16(v[j]);
          }
        }
      }
      // this block's TMEM reads are done; the accumulator is handed back when the warp LEAVES the tile (with more than
      // four 64-column sub-blocks per tile — the 320-wide tile — a warp owns two blocks of the same tile, and a second
      // arrival for one tile would complete the barrier phase early)
      tc_fence_before();
      __syncwarp();
      const uint32_t s_cur = s;
      const uint32_t tprev = tcount;
      s += 4;
      while (s >= nsub_out) {
        s -= nsub_out;
        ++tcount;
      }
      if (lane == 0) {
        if (tcount != tprev) {
          if (PAIR) mbar_arrive_leader(&acc_empty[b]);
          else mbar_arrive(&acc_empty[b]);
        }
        if (!skip && p.tma_epi) {
          tma_store_5d(subw == 64 ? &tmO64 : &tmO32, ob, (int)ocol0, (int)(mb1 + qo1), (int)(mb2 + qo2),
                       (int)(mb3 + qo3), 0);
          tma_store_commit();
        }
      }
      (void)s_cur;
      tpass = (tcount != tprev) ? tprev + 1 : tprev;
      if (tcount < num_tiles) {
        if (tcount != tprev) decode_tile(p, tile_begin + tcount, PAIR ? 2u : 1u, rank, QUAD ? 2u : 1u, npair, n_tile, mb1, mb2, mb3);
        prefetch_residual();
      }
    }
    for (; tpass < num_tiles; ++tpass) {
      mbar_wait_parked(&acc_full[acc_buf(tpass)], acc_par(tpass));
      if (lane == 0) {
        if (PAIR) mbar_arrive_leader(&acc_empty[acc_buf(tpass)]);
        else mbar_arrive(&acc_empty[acc_buf(tpass)]);
      }
    }
    if (p.tma_epi && lane == 0) tma_store_wait_all();  // all bulk stores of this thread have landed
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all();  // the peer may still be reading its accumulators / smem stages
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

static int ilog2_exact(uint32_t v) {
  if (v == 0 || (v & (v - 1)) != 0) return -1;
  int l = 0;
  while ((1u << l) < v) ++l;
  return l;
}

// rank-5 view of a row-major [rows][ld] bf16 matrix addressed through the output-pixel space:
// dims (cols, M1, M2, M3, 1), element (col, m1, m2, m3) at base + (m1*rs0 + m2*rs1 + m3*rs2)*ld + col
static int encode_rows_view(CUtensorMap* tm, const void* base, int64_t ld, uint32_t cols, const uint32_t* m_ext,
                            const int64_t* rs, const uint32_t* box_rows, uint32_t box_cols) {
  uint64_t dims[5] = {cols, m_ext[0], m_ext[1], m_ext[2], 1};
  uint64_t str[4];
  for (int i = 0; i < 3; ++i) {
    // a size-1 dim may carry any stride; keep it a positive multiple of 16 bytes
    const int64_t s = (m_ext[i] > 1 ? rs[i] : 1) * ld * 2;
    if (s <= 0) {
      set_error("b200svd_gemm: output row strides must be positive");
      return 1;
    }
    str[i] = (uint64_t)s;
  }
  str[3] = str[2] * (m_ext[2] ? m_ext[2] : 1);
  uint32_t box[5] = {box_cols, box_rows[0], box_rows[1], box_rows[2], 1};
  return box_cols == 64 ? encode_tmap_bf16(tm, base, 5, dims, str, box) : encode_tmap_bf16_sw64(tm, base, 5, dims, str, box);
}

template <int BN, int CL, int EPI = 0>
static int launch(const b200svd_gemm_params* p, const CUtensorMap& tmA, const GemmDev& d, cudaStream_t st) {
  using Cfg = TileCfg<BN, CL>;
  constexpr bool PAIR = CL >= 2;
  constexpr bool QUAD = CL == 4;
  // weights [taps][n][k] -> TMA dims (k, n, taps); in PAIR mode each CTA loads half of the N tile
  CUtensorMap tmB;
  uint64_t bd[3] = {p->k, p->n, p->taps};
  uint64_t bs[2] = {(uint64_t)p->k * 2, (uint64_t)p->k * 2 * p->n};
  uint32_t bb[3] = {64, (uint32_t)(Cfg::B_ROWS / Cfg::NMMA), 1};
  if (encode_tmap_bf16(&tmB, p->w_ptr, 3, bd, bs, bb)) return 1;
  GemmDev dd = d;
  dd.n_tiles = (p->n + BN - 1) / BN;
  if (QUAD) dd.n_tiles = (dd.n_tiles + 1) / 2;  // N work items = pairs of N tiles
  // QUAD: second activation map whose box is one 64-row half of the A tile (outermost non-unit row dim halved)
  CUtensorMap tmAh = tmA;
  dd.a_half_dim = 0;
  dd.a_half_size = 0;
  if (QUAD) {
    int hd = 2;
    while (hd > 0 && p->m_box[hd] < 2) --hd;
    if (p->m_box[hd] < 2) {
      set_error("mtgemm: cannot halve the A tile");
      return 1;
    }
    uint32_t hb[5];
    for (int i = 0; i < 5; ++i) hb[i] = p->a_box[i];
    dd.a_half_dim = p->m_adim[hd];
    dd.a_half_size = p->m_box[hd] / 2;
    hb[dd.a_half_dim] = dd.a_half_size;
    if (encode_tmap_bf16(&tmAh, p->a_ptr, 5, p->a_dims, p->a_strides, hb)) return 1;
  }

  // epilogue tensor maps: output and residual-1, both in quadrant boxes of 32 rows
  CUtensorMap tmO64, tmO32, tmR64, tmR32;
  memset(&tmO64, 0, sizeof(tmO64));
  tmO32 = tmO64;
  tmR64 = tmO64;
  tmR32 = tmO64;
  if (dd.tma_epi) {
    const uint32_t n_out = p->act == B200SVD_ACT_GEGLU ? p->n / 2 : p->n;
    const uint32_t b1 = p->m_box[0], b2 = p->m_box[1];
    const uint32_t c1 = b1 < 32 ? b1 : 32;
    const uint32_t c2 = (b2 < 32 / c1) ? b2 : 32 / c1;
    const uint32_t c3 = 32 / (c1 * c2);
    const uint32_t qbox[3] = {c1, c2, c3};
    if (encode_rows_view(&tmO64, p->out, p->ldo, n_out, p->m_ext, p->out_rs, qbox, 64)) return 1;
    if (encode_rows_view(&tmO32, p->out, p->ldo, n_out, p->m_ext, p->out_rs, qbox, 32)) return 1;
    if (p->res1 != nullptr) {
      if (encode_rows_view(&tmR64, p->res1, p->ld1, n_out, p->m_ext, p->out_rs, qbox, 64)) return 1;
      if (encode_rows_view(&tmR32, p->res1, p->ld1, n_out, p->m_ext, p->out_rs, qbox, 32)) return 1;
    }
  }
  const int slot = dev_slot();
  static bool attr_set[B200_MAX_DEVICES] = {};
  if (!attr_set[slot]) {
    cudaError_t e = cudaFuncSetAttribute(mtgemm_kernel<BN, CL, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mtgemm)");
    attr_set[slot] = true;
  }
  const uint64_t m_tiles = (uint64_t)dd.m_tiles[0] * dd.m_tiles[1] * dd.m_tiles[2];
  const uint64_t total = (uint64_t)dd.n_tiles * (PAIR ? (m_tiles + 1) / 2 : m_tiles);  // (pair) tiles
  if (total == 0 || total > 0x7FFFFFFFull) {
    set_error("mtgemm: bad tile count %llu", (unsigned long long)total);
    return 1;
  }
  // CTAs, CTA pairs or CTA quads that can be co-resident (clusters must fit inside a GPC, so quads may not tile all SMs)
  static int max_clusters_dev[B200_MAX_DEVICES] = {};
  int& max_clusters = max_clusters_dev[slot];
  if (max_clusters == 0) {
    max_clusters = sm_count() / CL;
    if (CL > 1) {
      cudaLaunchConfig_t qc = {};
      qc.gridDim = dim3((unsigned)(sm_count() / CL * CL));
      qc.blockDim = dim3(NUM_THREADS);
      qc.dynamicSmemBytes = Cfg::SMEM_BYTES;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = CL;
      qa[0].val.clusterDim.y = 1;
      qa[0].val.clusterDim.z = 1;
      qc.attrs = qa;
      qc.numAttrs = 1;
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, mtgemm_kernel<BN, CL, EPI>, &qc) == cudaSuccess && nc > 0 && nc < max_clusters)
        max_clusters = nc;
      else
        (void)cudaGetLastError();
    }
  }
  const uint32_t workers = (uint32_t)max_clusters;
  const uint32_t nw = (uint32_t)(total < workers ? total : workers);
  dd.total_tiles = (uint32_t)total;
  dd.tiles_per_cta = (uint32_t)((total + nw - 1) / nw);
  const uint32_t used = (uint32_t)((total + dd.tiles_per_cta - 1) / dd.tiles_per_cta);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL * used);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, mtgemm_kernel<BN, CL, EPI>, tmA, tmAh, tmB, tmO64, tmO32, tmR64, tmR32, dd);
  if (e != cudaSuccess) return cuda_fail(e, "mtgemm launch");
  return 0;
}

// 0 = never, 1 = always (when the tile shape allows), 2 = auto (compute-heavy launches); env B200SVD_PAIR
static int g_pair_mode = -1;
static int pair_mode() {
  if (g_pair_mode < 0) {
    const char* e = getenv("B200SVD_PAIR");
    g_pair_mode = e ? atoi(e) : 2;
    if (g_pair_mode < 0 || g_pair_mode > 3) g_pair_mode = 2;
  }
  return g_pair_mode;
}

}  // namespace b200

extern "C" int b200svd_gemm_pair_mode(int mode) {
  const int prev = b200::pair_mode();
  if (mode >= 0 && mode <= 3) b200::g_pair_mode = mode;
  return prev;
}

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
  d.gn_part = p->gn_part;
  d.gn_slot_sample = p->gn_slot_sample;
  d.gn_ld = p->gn_ld;
  d.gn_rows = p->gn_rows ? p->gn_rows : 1;

  int bn = p->bn;
  if (p->act == B200SVD_ACT_GEGLU) {
    if (bn == 0) bn = 256;
    if (bn != 256 || (p->n % 256) != 0) {
      set_error("b200svd_gemm: GEGLU needs n (%u) divisible by 256 and the 256-wide tile (weights interleaved per tile)",
                p->n);
      return 1;
    }
  }
  if (bn == 0) {
    if (p->n <= 32) bn = 32;
    else if (p->n <= 64) bn = 64;
    else if (p->n <= 128) bn = 128;
    // the pair tiles are bound by L2->SM bytes per FLOP (tools/bench_bn.py): the 256-wide tile wins even with a
    // ragged last tile (its out-of-range weight rows are TMA zero fill, not traffic); N = 160 / 320 fit the 160 tile
    else if (p->n == 160 || p->n == 320) bn = 160;
    else bn = 256;
  }
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  d.tma_epi = (!p->out_fp32 && p->ldo % 8 == 0 && al16(p->out) &&
               (p->res1 == nullptr || (p->ld1 % 8 == 0 && al16(p->res1))) &&
               (p->res2 == nullptr || (p->ld2 % 8 == 0 && al16(p->res2))))
                  ? 1
                  : 0;
  if (!d.tma_epi && !p->out_fp32) {
    set_error("b200svd_gemm: bf16 outputs/residuals need 16-byte aligned bases and leading dims that are multiples of 8");
    return 1;
  }

  CUtensorMap tmA;
  if (encode_tmap_bf16(&tmA, p->a_ptr, 5, p->a_dims, p->a_strides, p->a_box)) return 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // 2-SM (cta_group::2) tiles whenever a wide tile has at least two M tiles: measured neutral-to-better on every
  // shape of the denoiser (profiles/r01_pair_shapes.txt); mode 1 additionally pairs the 128-wide tile
  const uint64_t m_tiles_all = (uint64_t)d.m_tiles[0] * d.m_tiles[1] * d.m_tiles[2];
  // a single M tile cannot pair: the 128-wide tile spreads such a (weight-streaming) GEMM over more SMs
  if (p->bn == 0 && bn == 256 && m_tiles_all < 2 && p->act != B200SVD_ACT_GEGLU) bn = 128;
  const int pm = pair_mode();
  const bool pair = m_tiles_all >= 2 && ((pm >= 1 && (bn == 256 || bn == 160)) || (pm == 1 && bn == 128));
  // mode 3: two pairs on adjacent N tiles share (multicast) their A tile when the N tile count is even
  const uint32_t n_tiles_all = (p->n + (uint32_t)bn - 1) / (uint32_t)bn;
  const bool quad = pair && pm == 3 && (n_tiles_all % 2) == 0;
  // 320-wide pair tile (B200SVD_BN320=1; single-buffered accumulator): long main loops only
  static int bn320 = -1;
  if (bn320 < 0) {
    const char* e = getenv("B200SVD_BN320");
    bn320 = e ? atoi(e) : 0;
  }
  // measured (profiles/r02_bench_vs_libs*.txt): wins for N = 320 (convs 0.69 -> 0.57 ms, 1.96 -> 1.58 ms; FF2 0.39 -> 0.355)
  // and for the N = 640 linears (0.353 -> 0.317); loses on the N = 640 / 1280 convolutions, whose 256-wide tiles are
  // double-buffered
  // mode 2 additionally takes the N = 640 convolutions (the 256-wide tile computes 768 columns for them: 17 % of the
  // tensor work is spent on zero-filled weight rows)
  if (bn320 && p->bn == 0 && (bn == 160 || bn == 256) &&
      (p->n == 320 || (p->n == 640 && (p->taps == 1 || bn320 >= 2))) && m_tiles_all >= 2 && pm >= 1 &&
      p->act == B200SVD_ACT_NONE && (uint64_t)p->taps * d.kblocks >= 15)
    bn = 320;
  // lean epilogue for everything without an activation (B200SVD_LEAN_EPI=0 keeps the generic loop: A/B knob)
  static int lean_epi = -1;
  if (lean_epi < 0) {
    const char* e = getenv("B200SVD_LEAN_EPI");
    lean_epi = e ? (atoi(e) != 0) : B200SVD_DEFAULT_LEAN_EPI;
  }
  const uint32_t n_out_h = p->act == B200SVD_ACT_GEGLU ? p->n / 2 : p->n;
  const bool lean = lean_epi && p->act == B200SVD_ACT_NONE && d.tma_epi && (n_out_h % 16) == 0 &&
                    (p->bias == nullptr || al16(p->bias)) &&
                    (p->fvec == nullptr || (al16(p->fvec) && (p->ldf % 4) == 0)) && !quad;
  if (p->gn_part != nullptr && (!lean || bn < 128 || p->gn_slot_sample == nullptr || p->gn_ld < (int64_t)p->n ||
                                 (p->gn_ld % 2) != 0 || (reinterpret_cast<uintptr_t>(p->gn_part) & 7) != 0)) {
    set_error("b200svd_gemm: gn_part needs the activation-free bf16 epilogue (n %% 16 == 0, n > 64), gn_slot_sample and "
              "an 8-byte aligned partial buffer with gn_ld >= n");
    return 1;
  }
  if (bn == 320) {
    if (m_tiles_all < 2 || p->n % 320 != 0 || p->act == B200SVD_ACT_GEGLU) {
      set_error("b200svd_gemm: the 320-wide tile is a 2-SM tile: needs >= 2 M tiles, n %% 320 == 0, no GEGLU");
      return 1;
    }
    return lean ? launch<320, 2, 2>(p, tmA, d, st) : launch<320, 2, 0>(p, tmA, d, st);
  }
  if (lean) {
    switch (bn) {
      case 128: return pair ? launch<128, 2, 2>(p, tmA, d, st) : launch<128, 1, 2>(p, tmA, d, st);
      case 160: return pair ? launch<160, 2, 2>(p, tmA, d, st) : launch<160, 1, 2>(p, tmA, d, st);
      case 256: return pair ? launch<256, 2, 2>(p, tmA, d, st) : launch<256, 1, 2>(p, tmA, d, st);
      default: break;
    }
  }
  switch (bn) {
    case 32: return launch<32, 1>(p, tmA, d, st);
    case 64: return launch<64, 1>(p, tmA, d, st);
    case 128: return pair ? launch<128, 2>(p, tmA, d, st) : launch<128, 1>(p, tmA, d, st);
    case 160:
      return quad ? launch<160, 4>(p, tmA, d, st) : pair ? launch<160, 2>(p, tmA, d, st) : launch<160, 1>(p, tmA, d, st);
    case 256: {
      // specialised GEGLU epilogue (bias only); B200SVD_GEGLU_EPI=0 keeps the generic loop (A/B knob)
      static int geglu_fast = -1;
      if (geglu_fast < 0) {
        const char* e = getenv("B200SVD_GEGLU_EPI");
        geglu_fast = e ? (atoi(e) != 0) : B200SVD_DEFAULT_GEGLU_EPI;
      }
      const bool g1 = geglu_fast && p->act == B200SVD_ACT_GEGLU && d.tma_epi && p->bias != nullptr && al16(p->bias) &&
                      p->fvec == nullptr && p->res1 == nullptr && p->res2 == nullptr && p->s_acc == 1.0f && !quad;
      if (g1) return pair ? launch<256, 2, 1>(p, tmA, d, st) : launch<256, 1, 1>(p, tmA, d, st);
      return quad ? launch<256, 4>(p, tmA, d, st) : pair ? launch<256, 2>(p, tmA, d, st) : launch<256, 1>(p, tmA, d, st);
    }
    default: set_error("b200svd_gemm: unsupported N tile %d", bn); return 1;
  }
}
