// K2 (large batches, unsharded database) - ONE retrieval kernel: the statistics pass folded into the apply pass.
//
// RANGE+ blends two softmaxes into one accumulator (reference: range/range.py:213-217,231-238), so both row normalisers
// must be known before the first P' tile is formed: until now a separate statistics kernel (retrieval_pc.cu:
// range_stats_pc_kernel) recomputed Q.K^T and every exponential, 18 % of the step.  In the producer/consumer apply kernel
// the PRODUCER pairs have what that pass needs already on chip - the K tile and the entry unit vectors in shared memory,
// idle tensor pipes (Q.K^T is a fifth of the flops, on a third of the SMs) and MUFU slack - while the consumers set the
// pace.  So a producer pair now works on two query-tile pairs at once:
//
//   pass p = 0 .. rounds     apply of the unit's work item p - 1   S  = Q(p-1) . K^T -> P' -> ring -> consumers   (p >= 1)
//                            statistics of its work item p         S' = Q(p)   . K^T -> row sums / maxima          (p < rounds)
//
// Both products read the same K stage.  Three softmax groups of four warps (as in retrieval_pc.cu: pass tile ix belongs to
// group ix mod 3) handle a tile's two products one after the other: S -> P' -> ring first (the consumers wait for it), then
// S' -> running row sum / maximum of the SEMANTIC softmax.  Two TMEM buffers per product (a dedicated statistics group of
// four warps - one warp per SM sub-partition working through every tile serially - could not keep the consumers' pace:
// 28.5 ms against 23 ms).  The GEOGRAPHIC normaliser needs no tensor core and no Q.K^T: a CUDA-core kernel over the
// unskipped tiles (range_geo_stats_kernel, ~1.3 ms) computes it beforehand, which also restores the tighter geo-skip mask
// of the apply side (known normalisers).  At the end of pass p the groups combine their partial sums and group 0 turns
// them into the per-row constants of work item p (the arithmetic of row_constants_kernel) for the apply side of pass
// p + 1 (shared memory) and the output scale for the consumers (global memory + a release flag).  Pass 0 is a
// statistics-only prologue (the consumers wait for the first P' tile), the last pass is apply-only.
//
// Consumers, ring, flags, accumulation windows and the cross-unit window barrier are those of retrieval_pc.cu (the
// consumer role is the same code; it only reads its output scale late, once the producer has published it).  An
// M-sharded database keeps the two-kernel path: its sums cross the ranks between the passes.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <cuda.h>
#include <cuda_runtime.h>

#include "ptx.cuh"
#include "range_kernels.h"

namespace {

constexpr int kBlockQ = 128, kKeys = 128, kXyzBytes = kKeys * 16;
constexpr int kGroups = 3;                                   // softmax groups of 4 warps; pass tile ix belongs to group ix % 3
constexpr int kSoftmaxWarps = 4 * kGroups;
constexpr int kWarpTma = kSoftmaxWarps, kWarpMma = kSoftmaxWarps + 1, kWarpPublish = kSoftmaxWarps + 2,
              kWarpXyz = kSoftmaxWarps + 3;
constexpr int kThreads = (kSoftmaxWarps + 4) * 32;
constexpr int kRing = 16;                                    // P' slots per producer CTA
constexpr int kPublishBatch = 4;                             // tiles per release of the `full` counter
constexpr int kWindow = 64;                                  // tiles per cross-unit synchronisation window
constexpr int kAccWindow = 128;                              // tiles per TMEM accumulation window (retrieval_pc.cu)
constexpr int kPiece = 16;                                   // S columns per TMEM load (both products share the registers)
#ifndef RANGE_FOLD_POLY
#define RANGE_FOLD_POLY 0
#endif
constexpr int kPolyEvery = RANGE_FOLD_POLY;                  // every kPolyEvery-th semantic exponential on the FMA pipe (0: never)
__device__ __forceinline__ float ex2_mixed(float x, int i) {
  return (kPolyEvery > 0 && (i % (kPolyEvery > 0 ? kPolyEvery : 1)) == kPolyEvery - 1) ? ptx::ex2_poly(x) : ptx::ex2(x);
}
template <int W>
__device__ __forceinline__ void tmem_ld_piece(uint32_t taddr, uint32_t (&v)[W]) {
  static_assert(W == 16 || W == 32, "piece width");
  if constexpr (W == 16) ptx::tmem_ld16(taddr, v);
  else ptx::tmem_ld32(taddr, v);
}
constexpr int kFlagStride = 32;                              // uint32 per flag line (128 B)
constexpr int kFlagsPerProducer = 4 * kFlagStride;           // full, done[0], done[1], scale_ready

struct ProdSmem {
  static constexpr int NS = 2, NX = 4;                       // K stages (32 KB: this CTA's 64 entries x 256 dims), xyz slots
  static constexpr int q = 0;                                // 2 x (4 x [128 rows x 64 dims] SW128): work items p - 1 / p
  static constexpr int stages = q + 2 * 65536;
  static constexpr int xyz = stages + NS * 32768;
  static constexpr int rowc = xyz + NX * kXyzBytes;          // [2][128 rows][2 float4] per-row constants of a work item
  static constexpr int red = rowc + 2 * kBlockQ * 32;        // [2][groups][128 rows] float2 partial (sum, max) of a pass
  static constexpr int bars = red + 2 * kGroups * kBlockQ * 8;
  static constexpr int b_q_full = 0, b_q_pair = 1, b_q_empty = 2;
  static constexpr int b_stage_full = 3;
  static constexpr int b_stage_empty = b_stage_full + NS;
  // "full" barriers are indexed by the tile number modulo lcm(buffers, groups): a barrier's consecutive phases then belong
  // to the SAME softmax group, which waits for them in order.  (Indexed by buffer, a group that runs two fills ahead of
  // the tensor core would take the completed phase before last for its own: mbarrier waits only know a parity.)
  static constexpr int NF = 6, NXF = 12;
  static constexpr int b_sa_full = b_stage_empty + NS;       // apply S: 2 buffers (tile ia & 1), NF full barriers (ia % 6)
  static constexpr int b_sa_empty = b_sa_full + NF;
  static constexpr int b_sb_full = b_sa_empty + 2;           // statistics S': 2 buffers, NF full barriers
  static constexpr int b_sb_empty = b_sb_full + NF;
  static constexpr int b_xyz_full = b_sb_empty + 2;          // xyz: NX slots (ia & 3), NXF full barriers (ia % 12)
  static constexpr int b_xyz_empty = b_xyz_full + NXF;
  static constexpr int b_rowc_ready = b_xyz_empty + NX;      // [2]
  static constexpr int b_slot_free = b_rowc_ready + 2;
  static constexpr int b_p_written = b_slot_free + kRing;
  static constexpr int n_bars = b_p_written + kRing;
  static constexpr int tmem_slot = bars + n_bars * 8;
  static constexpr int total = tmem_slot + 16;
};
struct ConsSmem {
  static constexpr int NV = 4, NP = 6, kStageV = 32768, kStageP = 16384;
  static constexpr int v = 0;
  static constexpr int p = v + NV * kStageV;
  static constexpr int bars = p + NP * kStageP;
  static constexpr int b_v_full = 0;
  static constexpr int b_v_empty = b_v_full + NV;
  static constexpr int b_p_full = b_v_empty + NV;
  static constexpr int b_p_empty = b_p_full + NP;
  static constexpr int b_o_full = b_p_empty + NP;
  static constexpr int b_o_empty = b_o_full + 1;
  static constexpr int n_bars = b_o_empty + 1;
  static constexpr int tmem_slot = bars + n_bars * 8;
  static constexpr int total = tmem_slot + 16;
};
constexpr int kDynamicSmem = (ProdSmem::total > ConsSmem::total ? ProdSmem::total : ConsSmem::total) + 1024;
static_assert(kDynamicSmem <= 232448, "shared memory per CTA");

struct PipeState {
  int idx = 0;
  uint32_t phase = 0;
  template <int N>
  __device__ __forceinline__ void advance() {
    if (++idx == N) {
      idx = 0;
      phase ^= 1;
    }
  }
};

// Work decomposition: identical to retrieval_pc.cu (unit u handles query-tile pair r * n_units + u in round r; leftover
// pairs of the last round are split over database ranges and their partial outputs summed).
struct FoldPlan {
  int n_units, full_rounds, tail_pairs, tail_split, tail_tiles;
  int tail_row0;
  size_t part_stride;
  int out_ld, out_f64;
  const int* perm;
  float a_geo, w_sem, w_geo, inv_vscale;      // blend weights: (beta, 1 - beta) for RANGE+, (1, 0) for RANGE
};
struct PcWork {
  int qp, t0, t1, split;
};
__device__ __forceinline__ int pc_rounds(const FoldPlan& p, int unit) {
  return p.full_rounds + (unit < p.tail_pairs * p.tail_split ? 1 : 0);
}
__device__ __forceinline__ PcWork pc_work(const FoldPlan& p, int unit, int r, int T) {
  if (r < p.full_rounds) return PcWork{r * p.n_units + unit, 0, T, -1};
  const int s = unit % p.tail_split;
  const int t0 = s * p.tail_tiles;
  return PcWork{p.full_rounds * p.n_units + unit / p.tail_split, t0, min(T, t0 + p.tail_tiles), p.tail_split > 1 ? s : -1};
}

template <bool kGeo>
__global__ void __launch_bounds__(kThreads, 1)
range_fold_pc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK64,
                     const __grid_constant__ CUtensorMap tmV128, const __grid_constant__ CUtensorMap tmP,
                     const float4* __restrict__ db_xyz, const float4* __restrict__ q_xyz, const float2* __restrict__ geo_sums,
                     const float2* __restrict__ geo_maxs, float4* rowc_out, int N, int M,
                     float a_sem, void* __restrict__ out, const uint32_t* __restrict__ geo_mask, int mask_words,
                     float* __restrict__ part, float4* __restrict__ acc_scratch, __half* __restrict__ ring,
                     uint32_t* __restrict__ flags, uint32_t* __restrict__ windows, const FoldPlan plan) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int cid = blockIdx.x >> 1;
  const int unit = cid / 3, role = cid % 3;                 // role 0: producer pair; 1, 2: consumer pairs
  const int n_units = plan.n_units;
  const bool active = unit < n_units;
  const bool producer = role == 0;
  const int T = (M + kKeys - 1) / kKeys;
  const int rounds = active ? pc_rounds(plan, unit) : 0;
  const int tail_items = plan.tail_pairs * plan.tail_split;
  const int tail_len = tail_items ? (plan.tail_split > 1 ? plan.tail_tiles : T) : 0;
  const int sync_units = plan.full_rounds > 0 ? n_units : tail_items;                      // units that have work
  // producer tiles of the longest unit: (its rounds) x T statistics / apply passes + the apply-only last pass
  const uint32_t longest = uint32_t(plan.full_rounds + (tail_items ? 1 : 0)) * uint32_t(T) + uint32_t(tail_items ? tail_len : T);
  const int n_windows = int((longest + kWindow - 1) / kWindow);
  // tiles the consumers of this unit receive over the whole launch (ring positions, `full` / `done` counters)
  const uint32_t my_tiles = uint32_t(plan.full_rounds) * uint32_t(T) +
                            (rounds > plan.full_rounds ? uint32_t(pc_work(plan, unit, plan.full_rounds, T).t1 -
                                                                  pc_work(plan, unit, plan.full_rounds, T).t0) : 0u);
  const int prod_id = unit * 2 + int(rank);
  uint32_t* full_flag = flags + size_t(prod_id) * kFlagsPerProducer;
  uint32_t* scale_flag = full_flag + 3 * kFlagStride;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (producer ? ProdSmem::bars : ConsSmem::bars));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (producer ? ProdSmem::tmem_slot : ConsSmem::tmem_slot));

  if (threadIdx.x == 0 && active) {
    if (producer) {
      using L = ProdSmem;
      ptx::mbar_init(&bars[L::b_q_full], 1);
      ptx::mbar_init(&bars[L::b_q_pair], 2);
      ptx::mbar_init(&bars[L::b_q_empty], 1);
      for (int i = 0; i < L::NS; ++i) {
        ptx::mbar_init(&bars[L::b_stage_full + i], 1);
        ptx::mbar_init(&bars[L::b_stage_empty + i], 1);
      }
      for (int i = 0; i < L::NF; ++i) {
        ptx::mbar_init(&bars[L::b_sa_full + i], 1);                      // MMA commit
        ptx::mbar_init(&bars[L::b_sb_full + i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        ptx::mbar_init(&bars[L::b_sa_empty + i], 2 * 4);                 // the 4 warps of the tile's group in both CTAs
        ptx::mbar_init(&bars[L::b_sb_empty + i], 2 * 4);
        ptx::mbar_init(&bars[L::b_rowc_ready + i], 4);                   // group 0's 4 warps, once the groups' partials are combined
      }
      for (int i = 0; i < L::NXF; ++i) ptx::mbar_init(&bars[L::b_xyz_full + i], 1);
      for (int i = 0; i < L::NX; ++i) ptx::mbar_init(&bars[L::b_xyz_empty + i], 4);     // the 4 warps of the tile's group
      for (int i = 0; i < kRing; ++i) {
        ptx::mbar_init(&bars[L::b_slot_free + i], 1);
        ptx::mbar_init(&bars[L::b_p_written + i], 4);
      }
    } else {
      using L = ConsSmem;
      for (int i = 0; i < L::NV; ++i) {
        ptx::mbar_init(&bars[L::b_v_full + i], 1);
        ptx::mbar_init(&bars[L::b_v_empty + i], 1);
      }
      for (int i = 0; i < L::NP; ++i) {
        ptx::mbar_init(&bars[L::b_p_full + i], 1);
        ptx::mbar_init(&bars[L::b_p_empty + i], 1);
      }
      ptx::mbar_init(&bars[L::b_o_full], 1);
      ptx::mbar_init(&bars[L::b_o_empty], 8);                            // 4 epilogue warps x 2 CTAs
    }
    ptx::fence_mbar_init();
  }
  if (active && warp == kWarpMma) ptx::tmem_alloc_2sm<512>(tmem_slot);
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = active ? __shfl_sync(0xffffffffu, *tmem_slot, 0) : 0u;
  const uint32_t smem_u = __shfl_sync(0xffffffffu, ptx::smem_u32(smem), 0);

  if (active && rounds > 0 && producer) {
    // =====================================================================================================
    // producer pair.  Pass p: apply of work item p - 1 (p >= 1), statistics of work item p (p < rounds).
    // Every pass but the last covers the whole database; the last one covers the last work item's range.
    // =====================================================================================================
    using L = ProdSmem;
    const uint32_t bars_u = smem_u + L::bars;
    const int passes = rounds + 1;
    const PcWork last_work = pc_work(plan, unit, rounds - 1, T);
    auto pass_lo = [&](int p) { return p == rounds ? last_work.t0 : 0; };
    auto pass_hi = [&](int p) { return p == rounds ? last_work.t1 : T; };
    auto mask_bit = [&](const uint32_t* row, int j) { return row != nullptr && ((__ldg(row + (j >> 5)) >> (j & 31)) & 1u); };
    auto mask_row_of = [&](int qp) {
      return (kGeo && geo_mask) ? geo_mask + size_t(2 * qp + int(rank)) * mask_words : static_cast<const uint32_t*>(nullptr);
    };

    if (warp == kWarpTma) {
      if (lane == 0) {
        ptx::prefetch_tmap(&tmQ);
        ptx::prefetch_tmap(&tmK64);
        PipeState st;
        uint32_t ix = 0;
        for (int p = 0; p < passes; ++p) {
          if (p < rounds) {                                               // Q of work item p -> q buffer p & 1
            const int qt = 2 * pc_work(plan, unit, p, T).qp + int(rank);
            if (p > 0) ptx::mbar_wait(&bars[L::b_q_empty], (p - 1) & 1);   // every product of pass p - 1 has read its Q tiles
            ptx::mbar_expect_tx(&bars[L::b_q_full], 65536);
            for (int c = 0; c < 4; ++c)
              ptx::tma_load_2d(smem + L::q + (p & 1) * 65536 + c * 16384, &tmQ, &bars[L::b_q_full], c * 64, qt * kBlockQ);
          }
          for (int j = pass_lo(p); j < pass_hi(p); ++j, ++ix) {
            if (leader && (ix % kWindow) == 0) {                          // keep the units on the same part of the database
              const uint32_t w = ix / kWindow;
              atomicAdd(&windows[w], 1u);
              if (w > 0) ptx::wait_flag_ge(&windows[w - 1], uint32_t(sync_units));
            }
            ptx::mbar_wait(&bars[L::b_stage_empty + st.idx], st.phase ^ 1);
            uint8_t* dst = smem + L::stages + st.idx * 32768;
            if (leader) ptx::mbar_expect_tx(&bars[L::b_stage_full + st.idx], 65536);
#pragma unroll
            for (int c = 0; c < 4; ++c)
              ptx::tma_load_2d_2sm(dst + c * 8192, &tmK64, &bars[L::b_stage_full + st.idx], c * 64, j * kKeys + int(rank) * 64);
            st.advance<L::NS>();
          }
        }
        if (leader)
          for (uint32_t w = (ix + kWindow - 1) / kWindow; w < uint32_t(n_windows); ++w) atomicAdd(&windows[w], 1u);
      }
    } else if (warp == kWarpXyz) {
      // ----- entry unit vectors of apply tile ia -> slot ia & 3 (only the apply side has geographic work) -----
      if (kGeo && lane == 0) {
        uint32_t ia = 0;
        for (int p = 1; p < passes; ++p) {
          const uint32_t* rowA = mask_row_of(pc_work(plan, unit, p - 1, T).qp);
          for (int j = pass_lo(p); j < pass_hi(p); ++j, ++ia) {
            const int x = ia & (L::NX - 1);
            ptx::mbar_wait(&bars[L::b_xyz_empty + x], ((ia / L::NX) & 1) ^ 1);
            uint64_t* full = &bars[L::b_xyz_full + ia % L::NXF];
            if (!mask_bit(rowA, j)) {
              ptx::mbar_expect_tx(full, kXyzBytes);
              ptx::bulk_load_1d(smem + L::xyz + x * kXyzBytes, db_xyz + j * kKeys, kXyzBytes, full);
            } else {
              ptx::mbar_arrive(full);
            }
          }
        }
      }
    } else if (warp == kWarpMma) {
      constexpr uint32_t idesc_qk = ptx::umma_idesc_f16(2 * kBlockQ, kKeys);
      PipeState st;
      uint32_t ia = 0, ib = 0;                                            // apply / statistics tiles issued so far
      for (int p = 0; p < passes; ++p) {
        const bool apply_on = p >= 1, stats_on = p < rounds;
        if (stats_on) {
          ptx::mbar_wait(&bars[L::b_q_full], p & 1);
          if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&bars[L::b_q_pair]), 0));
          __syncwarp();
        }
        if (!leader) continue;
        if (stats_on) ptx::mbar_wait_cluster(&bars[L::b_q_pair], p & 1);
        for (int j = pass_lo(p); j < pass_hi(p); ++j) {
          const int ba = ia & 1, bb = ib & 1;
          if (apply_on) ptx::mbar_wait_cluster(&bars[L::b_sa_empty + ba], ((ia >> 1) & 1) ^ 1);
          if (stats_on) ptx::mbar_wait_cluster(&bars[L::b_sb_empty + bb], ((ib >> 1) & 1) ^ 1);
          ptx::mbar_wait(&bars[L::b_stage_full + st.idx], st.phase);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t b_base = smem_u + L::stages + st.idx * 32768;
            if (apply_on) {
              const uint32_t a_base = smem_u + L::q + ((p - 1) & 1) * 65536;
#pragma unroll
              for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  ptx::umma_f16_ss_2sm(tmem_base + ba * kKeys, ptx::umma_desc_kmajor_sw128(a_base + c * 16384 + kk * 32),
                                       ptx::umma_desc_kmajor_sw128(b_base + c * 8192 + kk * 32), idesc_qk, (c | kk) != 0);
              ptx::umma_commit_2sm_u32(bars_u + 8 * (L::b_sa_full + ia % L::NF));
            }
            if (stats_on) {
              const uint32_t a_base = smem_u + L::q + (p & 1) * 65536;
#pragma unroll
              for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  ptx::umma_f16_ss_2sm(tmem_base + 256 + bb * kKeys, ptx::umma_desc_kmajor_sw128(a_base + c * 16384 + kk * 32),
                                       ptx::umma_desc_kmajor_sw128(b_base + c * 8192 + kk * 32), idesc_qk, (c | kk) != 0);
              ptx::umma_commit_2sm_u32(bars_u + 8 * (L::b_sb_full + ib % L::NF));
            }
            ptx::umma_commit_2sm_u32(bars_u + 8 * (L::b_stage_empty + st.idx));
          }
          __syncwarp();
          ia += apply_on ? 1u : 0u;
          ib += stats_on ? 1u : 0u;
          st.advance<L::NS>();
        }
        if (ptx::elect_one()) ptx::umma_commit_2sm_u32(bars_u + 8 * L::b_q_empty);
        __syncwarp();
      }
    } else if (warp == kWarpPublish) {
      // ----- ring bookkeeping: as retrieval_pc.cu (gate on the consumers' `done` counters, batched release of `full`) -----
      if (lane == 0) {
        const uint32_t total = my_tiles;
        uint32_t pub = 0, gate = 0, seen = 0, spins = 0;
        while (pub < total || gate < total) {
          bool progress = false;
          if (gate < total) {
            const uint32_t need = gate >= uint32_t(kRing) ? gate - kRing + 1 : 0u;
            if (seen < need) {
              const uint32_t d0 = ptx::ld_acquire_gpu(full_flag + kFlagStride), d1 = ptx::ld_acquire_gpu(full_flag + 2 * kFlagStride);
              seen = d0 < d1 ? d0 : d1;
            }
            if (seen >= need) {
              ptx::mbar_arrive(&bars[L::b_slot_free + (gate % kRing)]);
              ++gate;
              progress = true;
            }
          }
          if (pub < total && ptx::mbar_try_wait(&bars[L::b_p_written + (pub % kRing)], (pub / kRing) & 1)) {
            ++pub;
            if (pub % kPublishBatch == 0 || pub == total) ptx::st_release_gpu(full_flag, pub);
            progress = true;
          }
          if (progress) spins = 0;
          else if (++spins > (1u << 26)) {
            printf("range_b200: fold ring bookkeeping stalled (block %d, published %u gated %u of %u)\n", blockIdx.x, pub, gate, total);
            __trap();
          }
        }
      }
    } else {
      // ----- softmax groups (warps 0..11): pass tile ix belongs to group ix % 3; apply product first, then the statistics -----
      const int grp = warp >> 2, quarter = warp & 3;
      const int row = quarter * 32 + lane;
      const uint32_t sa_empty_leader0 = ptx::mapa(ptx::smem_u32(&bars[L::b_sa_empty]), 0);
      const uint32_t sb_empty_leader0 = ptx::mapa(ptx::smem_u32(&bars[L::b_sb_empty]), 0);
      __half* ring_row = ring + size_t(prod_id) * kRing * 128 * 128 + row * 8;
      const uint32_t lane_base = tmem_base + (uint32_t(quarter * 32) << 16);
      uint32_t bufA[kPiece], bufB[kPiece];
      constexpr int kPairs = 128 / (2 * kPiece);
      for (int p = 0; p < passes; ++p) {
        const bool apply_on = p >= 1, stats_on = p < rounds;
        const int lo = pass_lo(p), len = pass_hi(p) - lo;
        const uint32_t base_ix = uint32_t(p) * uint32_t(T);
        // apply side: work item p - 1, constants written by group 0 at the end of pass p - 1
        const uint32_t* mask_row = nullptr;
        float cs = -INFINITY, cg = -INFINITY, gx = 0.f, gy = 0.f, gz = 0.f;
        if (apply_on) {
          const int r = p - 1;
          mask_row = mask_row_of(pc_work(plan, unit, r, T).qp);
          ptx::mbar_wait(&bars[L::b_rowc_ready + (r & 1)], (r >> 1) & 1);
          const float4* rc = reinterpret_cast<const float4*>(smem + L::rowc + (r & 1) * kBlockQ * 32) + 2 * row;
          const float4 c0 = rc[0], c1 = rc[1];
          cs = c0.x; cg = c0.y; gx = c0.z; gy = c0.w; gz = c1.x;
        }
        float sum_s = 0.f, max_s = -2.f;                                  // statistics side: work item p
        for (int k = int((uint32_t(grp) + 3u - base_ix % 3u) % 3u); k < len; k += kGroups) {
          const int j = lo + k;
          const uint32_t ix = base_ix + uint32_t(k);
          if (apply_on) {
            // ======== S (TMEM fp32) -> P' (fp16) -> ring ========
            const uint32_t ia = ix - uint32_t(T);
            const int ba = ia & 1, x = ia & (L::NX - 1), slot = ia % kRing;
            const uint32_t taddr = lane_base + ba * kKeys;
            const bool with_geo = kGeo && !mask_bit(mask_row, j);
            __half* dst = ring_row + size_t(slot) * 128 * 128;
            ptx::mbar_wait(&bars[L::b_sa_full + ia % L::NF], (ia / L::NF) & 1);
            if (with_geo) ptx::mbar_wait(&bars[L::b_xyz_full + ia % L::NXF], (ia / L::NXF) & 1);
            ptx::tc_fence_after();
            tmem_ld_piece(taddr, bufA);
            auto piece = [&](const uint32_t (&cur)[kPiece], int h) {
              const uint32_t kxyz = smem_u + L::xyz + x * kXyzBytes + h * kPiece * 16;
              const int nvalid = M - (j * kKeys + h * kPiece);
              uint32_t packed[kPiece / 2];
              auto body = [&](auto masked, auto geo) {
                constexpr bool kM = decltype(masked)::value, kG = decltype(geo)::value;
#pragma unroll
                for (int w = 0; w < kPiece / 2; ++w) {
                  float pv[2];
#pragma unroll
                  for (int u = 0; u < 2; ++u) {
                    const int i = 2 * w + u;
                    float pr = ex2_mixed(fmaf(__uint_as_float(cur[i]), a_sem, cs), i);
                    if (kG) {
                      const float4 kk = ptx::lds_f4(kxyz + i * 16);
                      pr += ptx::ex2(fmaf(gx, kk.x, fmaf(gy, kk.y, fmaf(gz, kk.z, cg))));
                    }
                    if (kM && i >= nvalid) pr = 0.f;
                    pv[u] = pr;
                  }
                  packed[w] = ptx::pack_half2(pv[0], pv[1]);
                }
              };
              if (nvalid >= kPiece) {
                if (with_geo) body(std::false_type{}, std::true_type{}); else body(std::false_type{}, std::false_type{});
              } else {
                if (with_geo) body(std::true_type{}, std::true_type{}); else body(std::true_type{}, std::false_type{});
              }
              if (h == 0) ptx::mbar_wait(&bars[L::b_slot_free + slot], (ia / kRing) & 1);   // both consumers copied tile ia - kRing
#pragma unroll
              for (int e = 0; e < kPiece / 8; ++e)
                ptx::stg_u4(dst + ((h * (kPiece / 8) + e) * 128) * 8, packed[4 * e], packed[4 * e + 1], packed[4 * e + 2],
                            packed[4 * e + 3]);
            };
#pragma unroll 1
            for (int hp = 0; hp < kPairs; ++hp) {
              ptx::tmem_ld_wait();
              tmem_ld_piece(taddr + (2 * hp + 1) * kPiece, bufB);
              piece(bufA, 2 * hp);
              ptx::tmem_ld_wait();
              if (hp < kPairs - 1) {
                tmem_ld_piece(taddr + (2 * hp + 2) * kPiece, bufA);
              } else {                          // the whole S tile is in registers: the MMA warp may overwrite the buffer
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                  if (leader) ptx::mbar_arrive(&bars[L::b_sa_empty + ba]);
                  else ptx::mbar_arrive_cluster_relaxed(sa_empty_leader0 + 8 * ba);
                }
              }
              piece(bufB, 2 * hp + 1);
            }
            __syncwarp();
            if (lane == 0) {
              ptx::mbar_arrive(&bars[L::b_p_written + slot]);
              if (kGeo) ptx::mbar_arrive(&bars[L::b_xyz_empty + x]);
            }
          }
          if (stats_on) {
            // ======== S' (TMEM fp32) -> running sum 2^(a (s - 1)) and maximum of the semantic softmax ========
            const uint32_t ib = ix;
            const int bb = ib & 1;
            const uint32_t taddr = lane_base + 256 + bb * kKeys;
            ptx::mbar_wait(&bars[L::b_sb_full + ib % L::NF], (ib / L::NF) & 1);
            ptx::tc_fence_after();
            tmem_ld_piece(taddr, bufA);
            auto piece = [&](const uint32_t (&cur)[kPiece], int h) {
              const int nvalid = M - (j * kKeys + h * kPiece);
              auto body = [&](auto masked) {
                constexpr bool kM = decltype(masked)::value;
#pragma unroll
                for (int i = 0; i < kPiece; i += 2) {
                  float sv[2];
#pragma unroll
                  for (int u = 0; u < 2; ++u) {
                    const float sc = __uint_as_float(cur[i + u]);
                    const bool valid = !kM || (i + u < nvalid);
                    float es = ex2_mixed(fmaf(sc, a_sem, -a_sem), i + u);
                    if (!valid) es = 0.f;
                    sum_s += es;
                    sv[u] = valid ? sc : -2.f;
                  }
                  max_s = ptx::max3(max_s, sv[0], sv[1]);
                }
              };
              if (nvalid >= kPiece) body(std::false_type{}); else body(std::true_type{});
            };
#pragma unroll 1
            for (int hp = 0; hp < kPairs; ++hp) {
              ptx::tmem_ld_wait();
              tmem_ld_piece(taddr + (2 * hp + 1) * kPiece, bufB);
              piece(bufA, 2 * hp);
              ptx::tmem_ld_wait();
              if (hp < kPairs - 1) {
                tmem_ld_piece(taddr + (2 * hp + 2) * kPiece, bufA);
              } else {
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                  if (leader) ptx::mbar_arrive(&bars[L::b_sb_empty + bb]);
                  else ptx::mbar_arrive_cluster_relaxed(sb_empty_leader0 + 8 * bb);
                }
              }
              piece(bufB, 2 * hp + 1);
            }
          }
        }
        if (stats_on) {
          // ---- end of pass p: combine the groups' partials; group 0 forms the constants of work item p ----
          float2* red = reinterpret_cast<float2*>(smem + L::red) + (p & 1) * kGroups * kBlockQ;
          red[grp * kBlockQ + row] = make_float2(sum_s, max_s);
          asm volatile("bar.sync 2, %0;" ::"n"(kSoftmaxWarps * 32) : "memory");
          if (grp == 0) {
            const int n = (2 * pc_work(plan, unit, p, T).qp + int(rank)) * kBlockQ + row;
            float l_s = 0.f, m_s = -2.f;
#pragma unroll
            for (int g2 = 0; g2 < kGroups; ++g2) {                        // fixed order: repeatable
              const float2 o = red[g2 * kBlockQ + row];
              l_s += o.x;
              m_s = fmaxf(m_s, o.y);
            }
            float l_g = 1.f, top_g = 0.f, qgx = 0.f, qgy = 0.f, qgz = 0.f;
            const float a_geo = plan.a_geo;
            if (kGeo && n < N) {
              l_g = geo_sums[n].y;                                       // range_geo_stats_kernel
              top_g = plan.w_geo * ptx::ex2(a_geo * (geo_maxs[n].y - 1.f)) / l_g;
              const float4 qx = q_xyz[n];
              qgx = qx.x * a_geo; qgy = qx.y * a_geo; qgz = qx.z * a_geo;
            }
            // (retrieval.cu: row_constants_kernel)
            const float top_s = plan.w_sem * ptx::ex2(a_sem * (m_s - 1.f)) / l_s;
            const float B = top_s + top_g;
            const float C = 8192.f;
            float c_s = (plan.w_sem > 0.f) ? -a_sem + log2f(plan.w_sem * C / (B * l_s)) : -INFINITY;
            float c_g = (kGeo && plan.w_geo > 0.f) ? -a_geo + log2f(plan.w_geo * C / (B * l_g)) : -INFINITY;
            float scale = B / C * plan.inv_vscale;
            if (n >= N) { c_s = -INFINITY; c_g = -INFINITY; scale = 0.f; }
            float4* rc = reinterpret_cast<float4*>(smem + L::rowc + (p & 1) * kBlockQ * 32) + 2 * row;
            rc[0] = make_float4(c_s, c_g, qgx, qgy);
            rc[1] = make_float4(qgz, scale, 0.f, 0.f);
            if (n < N) rowc_out[2 * n + 1] = make_float4(qgz, scale, l_s, l_g);      // the consumers' output scale (+ the sums, for inspection)
            __threadfence();
            asm volatile("bar.sync 3, 128;" ::: "memory");                           // group 0's four warps
            if (warp == 0 && lane == 0) ptx::st_release_gpu(scale_flag, uint32_t(p + 1));
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&bars[L::b_rowc_ready + (p & 1)]);
          }
        }
      }
    }
  } else if (active && rounds > 0) {
    // =====================================================================================================
    // consumer pair: value dims [512 (role - 1), +512) of the unit's two query tiles (as retrieval_pc.cu)
    // =====================================================================================================
    using L = ConsSmem;
    const uint32_t bars_u = smem_u + L::bars;
    const int cp = role - 1;
    const int dimbase = cp * 512;
    if (warp == 0) {
      if (lane == 0) {                                                    // ----- Vt loader -----
        ptx::prefetch_tmap(&tmV128);
        PipeState st;
        for (int r = 0; r < rounds; ++r) {
          const PcWork wk = pc_work(plan, unit, r, T);
          for (int j = wk.t0; j < wk.t1; ++j) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const int key0 = j * kKeys + half * 64;
              ptx::mbar_wait(&bars[L::b_v_empty + st.idx], st.phase ^ 1);
              uint8_t* dst = smem + L::v + st.idx * L::kStageV;
              if (leader) ptx::mbar_expect_tx(&bars[L::b_v_full + st.idx], 2 * L::kStageV);
#pragma unroll
              for (int nb = 0; nb < 2; ++nb)
                ptx::tma_load_2d_2sm(dst + nb * 16384, &tmV128, &bars[L::b_v_full + st.idx], key0,
                                     dimbase + nb * 256 + int(rank) * 128);
              st.advance<L::NV>();
            }
          }
        }
      }
    } else if (warp == 6) {
      if (lane == 0) {                                                    // ----- P' loader: follows the producer's `full` counter -----
        ptx::prefetch_tmap(&tmP);
        PipeState st;
        uint32_t seen = 0;
        for (uint32_t it = 0; it < my_tiles; ++it) {
          const int slot = it % kRing;
          if (seen < it + 1) {
            uint32_t spins = 0;
            while ((seen = ptx::ld_acquire_gpu(full_flag)) < it + 1) {
              if (++spins > (1u << 26)) {                                 // (the statistics-only prologue precedes tile 0)
                printf("range_b200: fold consumer %d waits for tile %u, producer published %u\n", blockIdx.x, it, seen);
                __trap();
              }
            }
            ptx::fence_proxy_async_all();
          }
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            ptx::mbar_wait(&bars[L::b_p_empty + st.idx], st.phase ^ 1);
            if (leader) ptx::mbar_expect_tx(&bars[L::b_p_full + st.idx], 2 * L::kStageP);
            ptx::tma_load_2d_2sm(smem + L::p + st.idx * L::kStageP, &tmP, &bars[L::b_p_full + st.idx], 0,
                                 (prod_id * kRing + slot) * 16 + half * 8);
            st.advance<L::NP>();
          }
        }
      }
    } else if (warp == 1) {
      if (leader) {                                                       // ----- P' . Vt -----
        constexpr uint32_t idesc_pv = ptx::umma_idesc_f16(2 * kBlockQ, 256);
        uint32_t* done0 = flags + size_t(unit * 2) * kFlagsPerProducer + (1 + cp) * kFlagStride;
        uint32_t* done1 = flags + size_t(unit * 2 + 1) * kFlagsPerProducer + (1 + cp) * kFlagStride;
        PipeState sv, sp;
        uint32_t it = 0, ev = 0;
        for (int r = 0; r < rounds; ++r) {
          const PcWork wk = pc_work(plan, unit, r, T);
          const int nt = wk.t1 - wk.t0;
          for (int j = 0; j < nt; ++j, ++it) {
            const bool first = j % kAccWindow == 0;
            if (first && ev > 0) {
              ptx::mbar_wait_cluster(&bars[L::b_o_empty], (ev - 1) & 1);
              ptx::tc_fence_after();
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              ptx::mbar_wait(&bars[L::b_p_full + sp.idx], sp.phase);
              if (half == 1 && ptx::elect_one()) {
                ptx::st_relaxed_gpu(done0, it + 1);
                ptx::st_relaxed_gpu(done1, it + 1);
              }
              ptx::mbar_wait(&bars[L::b_v_full + sv.idx], sv.phase);
              ptx::tc_fence_after();
              if (ptx::elect_one()) {
                const uint32_t a_base = smem_u + L::p + sp.idx * L::kStageP;
                const uint32_t b_base = smem_u + L::v + sv.idx * L::kStageV;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                  for (int nb = 0; nb < 2; ++nb)
                    ptx::umma_f16_ss_2sm(tmem_base + nb * 256, ptx::umma_desc_kmajor_nosw(a_base + kk * 4096, 2048, 128),
                                         ptx::umma_desc_kmajor_sw128(b_base + nb * 16384 + kk * 32), idesc_pv,
                                         !(first && half == 0 && kk == 0));
                ptx::umma_commit_2sm_u32(bars_u + 8 * (L::b_v_empty + sv.idx));
                ptx::umma_commit_2sm_u32(bars_u + 8 * (L::b_p_empty + sp.idx));
              }
              __syncwarp();
              sv.advance<L::NV>();
              sp.advance<L::NP>();
            }
            if ((j + 1) % kAccWindow == 0 || j + 1 == nt) {
              if (ptx::elect_one()) ptx::umma_commit_2sm_u32(bars_u + 8 * L::b_o_full);
              __syncwarp();
              ++ev;
            }
          }
        }
      }
    } else if (warp >= 2 && warp < 6) {
      // ----- epilogue warps: O (TMEM, 128 lanes x 512 columns) -> global, scaled -----
      const int quarter = warp & 3;
      const int row = quarter * 32 + lane;
      const uint32_t o_empty_leader = ptx::mapa(ptx::smem_u32(&bars[L::b_o_empty]), 0);
      float4* scratch = acc_scratch + size_t((unit * 2 + cp) * 2 + int(rank)) * (128 * 128) + row;
      const uint64_t keep = ptx::l2_policy_evict_last();
      uint32_t ev = 0;
      for (int r = 0; r < rounds; ++r) {
        const PcWork wk = pc_work(plan, unit, r, T);
        const int qt = 2 * wk.qp + int(rank);
        const int n = qt * kBlockQ + row;
        float out_scale = 0.f;
        const bool direct = wk.split < 0;
        const size_t drow = direct ? size_t(plan.perm && n < N ? plan.perm[n] : n) * plan.out_ld : 0;
        float* orow32 = direct ? reinterpret_cast<float*>(out) + drow + dimbase
                               : part + size_t(wk.split) * plan.part_stride + size_t(n - plan.tail_row0) * 1024 + dimbase;
        double* orow64 = reinterpret_cast<double*>(out) + drow + dimbase;
        const bool f64 = direct && plan.out_f64;
        const int nseg = (wk.t1 - wk.t0 + kAccWindow - 1) / kAccWindow;
        for (int sg = 0; sg < nseg; ++sg, ++ev) {
          const bool last = sg == nseg - 1;
          const bool acc = sg > 0 && n < N;
          if (last) {
            // the output scale of work item r: published by the producer's statistics group at the end of pass r, long
            // before this round's products complete
            ptx::wait_flag_ge(scale_flag, uint32_t(r + 1));
            if (n < N) out_scale = __ldcg(reinterpret_cast<const float*>(rowc_out + 2 * n + 1) + 1);
          }
          float4 a0[8], a1[8];
          auto load_window = [&](int cc, float4 (&a)[8]) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = ptx::ldg_f4_hint(scratch + size_t(cc) * 8 * 128 + i * 128, keep);
          };
          auto flush_block = [&](int cc, uint32_t (&v)[32], const float4 (&a)[8]) {
            if (n >= N) return;
            if (acc) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                v[4 * i] = __float_as_uint(__uint_as_float(v[4 * i]) + a[i].x);
                v[4 * i + 1] = __float_as_uint(__uint_as_float(v[4 * i + 1]) + a[i].y);
                v[4 * i + 2] = __float_as_uint(__uint_as_float(v[4 * i + 2]) + a[i].z);
                v[4 * i + 3] = __float_as_uint(__uint_as_float(v[4 * i + 3]) + a[i].w);
              }
            }
            if (!last) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                ptx::stg_f4_hint(scratch + size_t(cc) * 8 * 128 + i * 128,
                                 make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                             __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])), keep);
            } else if (f64) {
#pragma unroll
              for (int i = 0; i < 32; i += 2)
                __stcs(reinterpret_cast<double2*>(orow64 + cc * 32 + i),
                       make_double2(double(__uint_as_float(v[i]) * out_scale), double(__uint_as_float(v[i + 1]) * out_scale)));
            } else {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                float4 o;
                o.x = __uint_as_float(v[i]) * out_scale;
                o.y = __uint_as_float(v[i + 1]) * out_scale;
                o.z = __uint_as_float(v[i + 2]) * out_scale;
                o.w = __uint_as_float(v[i + 3]) * out_scale;
                __stcs(reinterpret_cast<float4*>(orow32 + cc * 32 + i), o);
              }
            }
          };
          if (acc) load_window(0, a0);
          ptx::mbar_wait(&bars[L::b_o_full], ev & 1);
          ptx::tc_fence_after();
#pragma unroll 1
          for (int cc = 0; cc < 16; cc += 2) {
            uint32_t v[32];
            ptx::tmem_ld32(tmem_base + (uint32_t(quarter * 32) << 16) + cc * 32, v);
            if (acc) load_window(cc + 1, a1);
            ptx::tmem_ld_wait();
            flush_block(cc, v, a0);
            ptx::tmem_ld32(tmem_base + (uint32_t(quarter * 32) << 16) + (cc + 1) * 32, v);
            if (acc && cc + 2 < 16) load_window(cc + 2, a0);
            ptx::tmem_ld_wait();
            flush_block(cc + 1, v, a1);
          }
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(o_empty_leader);
        }
      }
    }
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (active && warp == kWarpMma) ptx::tmem_dealloc_2sm<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------
// Geographic row statistics on the CUDA cores (no Q.K^T, no tensor core): per row  sum_j 2^(a_g (g_j - 1))  and  max_j g_j
// over the tiles of one split that the bounding-cap mask does not skip (reference: the denominator of the geographic
// softmax, range/range.py:231-234).  Block = one 128-query tile, thread = one query; the entry unit vectors of a tile go
// through shared memory (broadcast LDS.128).  Partials [split][N] float2 = (0, sum) / (-2, max cos) in the layout of the
// statistics kernels, so reduce_stats_kernel merges them.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
range_geo_stats_kernel(const float4* __restrict__ db_xyz, const float4* __restrict__ q_xyz, int N, int M, int tiles_per_split,
                       float a_geo, float2* __restrict__ part_sum, float2* __restrict__ part_max,
                       const uint32_t* __restrict__ geo_mask, int mask_words) {
  __shared__ float4 xs[2][kKeys];
  const int qt = blockIdx.x, split = blockIdx.y;
  const int n = qt * kBlockQ + threadIdx.x;
  const int total_tiles = (M + kKeys - 1) / kKeys;
  const int t_begin = split * tiles_per_split, t_end = min(total_tiles, t_begin + tiles_per_split);
  const uint32_t* mask_row = geo_mask ? geo_mask + size_t(qt) * mask_words : nullptr;
  float4 qx = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n < N) qx = q_xyz[n];
  const float gx = qx.x * a_geo, gy = qx.y * a_geo, gz = qx.z * a_geo;
  float sum0 = 0.f, sum1 = 0.f, mx = -3.0e38f;
  int buf = 0;
  for (int t = t_begin; t < t_end; ++t) {
    if (mask_row != nullptr && ((__ldg(mask_row + (t >> 5)) >> (t & 31)) & 1u)) continue;     // block-uniform
    xs[buf][threadIdx.x] = db_xyz[size_t(t) * kKeys + threadIdx.x];
    __syncthreads();             // (two buffers: a thread can only overwrite buffer b after everyone passed the next barrier)
    const int nvalid = min(kKeys, M - t * kKeys);
    const uint32_t base = ptx::smem_u32(&xs[buf][0]);
    if (nvalid == kKeys) {
#pragma unroll 8
      for (int i = 0; i < kKeys; i += 2) {
        const float4 k0 = ptx::lds_f4(base + i * 16), k1 = ptx::lds_f4(base + i * 16 + 16);
        const float g0 = fmaf(gx, k0.x, fmaf(gy, k0.y, fmaf(gz, k0.z, -a_geo)));
        const float g1 = fmaf(gx, k1.x, fmaf(gy, k1.y, fmaf(gz, k1.z, -a_geo)));
        sum0 += ptx::ex2(g0);
        sum1 += ptx::ex2(g1);
        mx = ptx::max3(mx, g0, g1);
      }
    } else {
      for (int i = 0; i < nvalid; ++i) {
        const float4 k0 = ptx::lds_f4(base + i * 16);
        const float g0 = fmaf(gx, k0.x, fmaf(gy, k0.y, fmaf(gz, k0.z, -a_geo)));
        sum0 += ptx::ex2(g0);
        mx = fmaxf(mx, g0);
      }
    }
    buf ^= 1;
  }
  if (n < N) {
    part_sum[size_t(split) * N + n] = make_float2(0.f, sum0 + sum1);
    part_max[size_t(split) * N + n] = make_float2(-2.f, mx / a_geo + 1.f);      // mx holds a_geo (g - 1): store the raw cosine
  }
}

// partials [splits][rows][1024] fp32 -> rows row0.. of the caller's output
__global__ void fold_reduce_tail_kernel(const float4* __restrict__ part, size_t stride4, int splits, int rows, int row0,
                                        const int* __restrict__ perm, void* __restrict__ out, int out_ld, int out_f64) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= size_t(rows) * 256) return;
  float4 a = part[i];
  for (int k = 1; k < splits; ++k) {
    const float4 b = part[size_t(k) * stride4 + i];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  }
  const int n = row0 + int(i / 256), c = int(i % 256) * 4;
  const size_t o = size_t(perm ? perm[n] : n) * out_ld + c;
  if (out_f64) {
    double* d = reinterpret_cast<double*>(out) + o;
    d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w;
  } else {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + o) = a;
  }
}

}  // namespace

namespace rangeb200 {

static FoldPlan fold_plan(int sm_count, int64_t N, int64_t M) {
  FoldPlan p{};
  const int64_t qp = ((N + kBlockQ - 1) / kBlockQ + 1) / 2, T = (M + kKeys - 1) / kKeys;
  p.n_units = apply_pc_units(sm_count);
  p.full_rounds = int(qp / p.n_units);
  p.tail_pairs = int(qp % p.n_units);
  p.tail_split = 1;
  if (p.tail_pairs > 0) {          // spread the leftover pairs over the idle units, >= 32 tiles per range, <= 4 ranges
    int k = p.n_units / p.tail_pairs;
    if (k > 4) k = 4;
    while (k > 1 && T / k < 32) --k;
    p.tail_split = k < 1 ? 1 : k;
  }
  p.tail_tiles = int((T + p.tail_split - 1) / p.tail_split);
  p.tail_row0 = p.full_rounds * p.n_units * 2 * kBlockQ;
  const int64_t tail_rows = N - p.tail_row0 > 0 ? N - p.tail_row0 : 0;
  p.part_stride = size_t(tail_rows) * 1024;
  return p;
}
static size_t fold_window_count(const FoldPlan& p, int64_t M) {
  const int64_t T = (M + kKeys - 1) / kKeys;
  const int64_t tail_items = int64_t(p.tail_pairs) * p.tail_split;
  const int64_t tail_len = tail_items ? (p.tail_split > 1 ? p.tail_tiles : T) : 0;
  const int64_t longest = int64_t(p.full_rounds + (tail_items ? 1 : 0)) * T + (tail_items ? tail_len : T);
  return size_t((longest + kWindow - 1) / kWindow + 1);
}
static size_t fold_ring_flag_bytes(int sm_count) { return size_t(apply_pc_units(sm_count)) * 2 * kFlagsPerProducer * 4; }
size_t fold_pc_flag_bytes(int sm_count, int64_t N, int64_t M) {
  return fold_ring_flag_bytes(sm_count) + fold_window_count(fold_plan(sm_count, N, M), M) * 4;
}
size_t fold_pc_part_bytes(int sm_count, int64_t N, int64_t M) {
  const FoldPlan p = fold_plan(sm_count, N, M);
  return p.tail_split > 1 ? size_t(p.tail_split) * p.part_stride * 4 : 0;
}
size_t fold_pc_ring_bytes(int sm_count) { return size_t(apply_pc_units(sm_count)) * 2 * kRing * 128 * 128 * 2; }
int fold_pc_ring_rows(int sm_count) { return apply_pc_units(sm_count) * 2 * kRing * 16; }

// grid (query tiles, splits); partials in the statistics kernels' layout
cudaError_t launch_geo_stats(const RetrievalArgs& a, int splits, int tiles_per_split, float* part_sum, float* part_max,
                             cudaStream_t stream) {
  const int qtiles = (a.N + kBlockQ - 1) / kBlockQ;
  range_geo_stats_kernel<<<dim3(unsigned(qtiles), unsigned(splits)), 128, 0, stream>>>(
      a.db_xyz, a.q_xyz, a.N, a.M, tiles_per_split, a.a_geo, reinterpret_cast<float2*>(part_sum),
      reinterpret_cast<float2*>(part_max), a.geo_mask, a.mask_words);
  return cudaGetLastError();
}

cudaError_t launch_fold_pc(const RetrievalArgs& a, const CUtensorMap& tmP, float beta, float inv_vscale, const float* geo_sums,
                           const float* geo_maxs, float* rowc, void* out,
                           int out_ld, int out_f64, const int* perm, void* ring, void* flags, void* part, void* scratch,
                           int sm_count, cudaStream_t stream) {
  FoldPlan plan = fold_plan(sm_count, a.N, a.M);
  plan.out_ld = out_ld;
  plan.out_f64 = out_f64;
  plan.perm = perm;
  plan.a_geo = a.a_geo;
  plan.w_sem = a.geo ? beta : 1.f;
  plan.w_geo = a.geo ? 1.f - beta : 0.f;
  plan.inv_vscale = inv_vscale;
  cudaError_t e = cudaMemsetAsync(flags, 0, fold_pc_flag_bytes(sm_count, a.N, a.M), stream);
  if (e != cudaSuccess) return e;
  auto kern = a.geo ? range_fold_pc_kernel<true> : range_fold_pc_kernel<false>;
  if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynamicSmem)) != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(unsigned(sm_count / 2 * 2));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = size_t(kDynamicSmem);
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeCooperative;
  attr[1].val.cooperative = 1;
  cfg.attrs = attr;
  // cooperative by default (co-residency guaranteed or a clean launch error); RANGE_PC_COOP=0: plain cluster launch for profilers
  static const bool coop = !(getenv("RANGE_PC_COOP") && atoi(getenv("RANGE_PC_COOP")) == 0);
  cfg.numAttrs = 1;
  int resident_clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&resident_clusters, kern, &cfg) != cudaSuccess) resident_clusters = 0;
  (void)cudaGetLastError();
  if (resident_clusters < int(cfg.gridDim.x / 2)) {
    fprintf(stderr, "range_b200: only %d of %u CTA pairs can be resident; fused retrieval kernel not launched\n",
            resident_clusters, cfg.gridDim.x / 2);
    return cudaErrorCooperativeLaunchTooLarge;
  }
  cfg.numAttrs = coop ? 2 : 1;
  e = cudaLaunchKernelEx(&cfg, kern, a.tmQ, a.tmK64, a.tmV128, tmP, a.db_xyz, a.q_xyz, reinterpret_cast<const float2*>(geo_sums),
                         reinterpret_cast<const float2*>(geo_maxs), reinterpret_cast<float4*>(rowc), a.N, a.M,
                         a.a_sem, out, a.geo_mask, a.mask_words, reinterpret_cast<float*>(part),
                         reinterpret_cast<float4*>(scratch), reinterpret_cast<__half*>(ring), reinterpret_cast<uint32_t*>(flags),
                         reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(flags) + fold_ring_flag_bytes(sm_count)), plan);
  if (e != cudaSuccess) {
    fprintf(stderr, "range_b200: fused retrieval launch failed (%s); grid %u smem %d\n", cudaGetErrorString(e), cfg.gridDim.x,
            kDynamicSmem);
    return e;
  }
  if (plan.tail_split > 1) {
    const int rows = a.N - plan.tail_row0;
    fold_reduce_tail_kernel<<<unsigned((size_t(rows) * 256 + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const float4*>(part), plan.part_stride / 4, plan.tail_split, rows, plan.tail_row0, perm, out, out_ld,
        out_f64);
    return cudaGetLastError();
  }
  return cudaSuccess;
}

}  // namespace rangeb200
