// K2b (large batches) - apply as a role-specialised persistent kernel.
//
// The single-role apply kernel (retrieval.cu) is bound by TMEM capacity: a 128-query tile's 1024-wide fp32
// output needs 1024 of an SM's 512 TMEM columns, so four CTAs each own a 256-wide slice and each of them
// recomputes Q.K^T and the softmax (reference: range/range.py:213-215,231-234) - half of its tensor cycles and
// all of its MUFU cycles are redundant work.  Here the two halves of the computation live on different SMs:
//
//   producer pair  (2 CTAs, cta_group::2)  S = Q.K^T -> P' = 2^(a s + cs) + 2^(gam g + cg)  (fp16), ONCE per
//                                          (query tile, database tile); P' goes to a ring in global memory
//                                          (it stays in L2: 32 KB per tile, 16 slots per producer CTA, tile
//                                          layout [16 key chunks][128 rows][8 entries])
//   consumer pairs (2 x 2 CTAs)            O[128 queries x 512 dims] += P' . Vt   (all 512 TMEM columns are
//                                          accumulators); P' arrives by TMA as a no-swizzle K-major A operand,
//                                          Vt halves are shared by the pair as in retrieval.cu; the epilogue
//                                          writes the caller's (N,1280) rows directly (fused concat)
//
// One unit = 3 clusters of 2 CTAs = two query tiles x 1024 value dims; 148 SMs = 24 units (+ 2 idle clusters).
// Producers are MUFU-bound (1 ex2 per pair + the geo ex2 of the ~40 % unskipped tiles, 16/clk/SM), consumers
// tensor-bound (2048 clk per 128x128 tile); nothing is computed twice.  Hand-off through L2 is ordered with
// st.release.gpu / ld.acquire.gpu flags that count tiles (monotonic, zeroed by the host before the launch):
//   full[p]      tiles producer CTA p has published        (bookkeeping thread, once per 4 tiles, after the softmax
//                                                           warps' stores)
//   done[p][c]   tiles consumer pair c has copied to smem  (consumer leader, after the TMA load completed)
// A window barrier every 64 tiles keeps the 24 units on the same part of the database (L2 reuse).
// All CTAs must be co-resident: one CTA per SM, capacity checked by the launcher (launch_apply_pc).
//
// Measured on B200 at 100 000 x 100 000 (tools/time_apply.py; PROF=1 prints per-role wait cycles):
//   single-role CTA-pair kernel (retrieval.cu)                                  32.7 ms
//   first working version (row-major ring, release per tile, 16 lockstep warps) 35.3 ms   <- gpu-scope release: 3300 clk/tile
//   + batched release, separate Vt / P' loaders, relaxed `done`                 30.9 ms
//   + coalesced ring layout, no-swizzle A operand, pipelined TMEM loads,
//     6-stage P' ring in the consumer, four desynchronised softmax groups       24.5 ms
//   + cross-unit window barrier (DRAM reads 33 GB -> 5.5 GB), 4 Vt stages        22.9 ms
//   + Hilbert-ordered tiles (geo skip 44 % -> 52 %)                              22.6 ms
//   + three softmax groups over four S buffers                                  21.6 ms
//   + apply-pass geo mask from the known normalisers (skip 52 % -> 63 %)         20-21 ms (run-to-run +-1 ms: power cap)
// Dead ends: 64-entry S half tiles with double-buffered groups (UMMA N = 64 is operand-fetch bound: 24.7 ms);
// every 3rd-6th exponential on the FMA pipe (no gain: the SM is power-capped, not MUFU-issue-bound).
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
// Softmax warps: kGroups groups of 4 (one warp per TMEM lane quarter).  Tile it belongs to group it % kGroups and
// lives in S buffer it % 4.  With THREE groups over FOUR buffers there is always a spare buffer: when a group
// finishes tile it, Q.K^T of its next tile it + 3 went into a buffer that was released a whole tile earlier, so the
// group never waits for the tensor core (4 groups x 4 buffers: 395 clk per tile waiting for S).
#ifndef RANGE_PC_GROUPS
#define RANGE_PC_GROUPS 3
#endif
constexpr int kGroups = RANGE_PC_GROUPS;
constexpr int kSoftmaxWarps = 4 * kGroups;
constexpr int kWarpTma = kSoftmaxWarps, kWarpMma = kSoftmaxWarps + 1, kWarpPublish = kSoftmaxWarps + 2,
              kWarpXyz = kSoftmaxWarps + 3;
constexpr int kThreads = (kSoftmaxWarps + 4) * 32;
static_assert(kSoftmaxWarps + 4 >= 7, "the consumer roles use warps 0..6");
#ifndef RANGE_PC_RING
#define RANGE_PC_RING 16
#endif
constexpr int kRing = RANGE_PC_RING;             // P' slots per producer CTA
#ifndef RANGE_PC_BATCH
#define RANGE_PC_BATCH 4
#endif
constexpr int kPublishBatch = RANGE_PC_BATCH;                 // tiles per release of the `full` counter (kRing >= 3 batches)
#ifndef RANGE_PC_WINDOW
#define RANGE_PC_WINDOW 64
#endif
constexpr int kWindow = RANGE_PC_WINDOW;                      // tiles per cross-unit synchronisation window (8192 entries, 21 MB of database)
// Every kPolyEvery-th semantic exponential is evaluated on the FMA pipe (ptx::ex2_poly, rel. error 2.7e-6) instead of
// MUFU: the softmax warps are MUFU-bound with issue slots to spare.  0 = never.
#ifndef RANGE_PC_POLY
#define RANGE_PC_POLY 0
#endif
constexpr int kPolyEvery = RANGE_PC_POLY;
__device__ __forceinline__ float ex2_mixed(float x, int i) {
  return (kPolyEvery > 0 && (i % (kPolyEvery > 0 ? kPolyEvery : 1)) == kPolyEvery - 1) ? ptx::ex2_poly(x) : ptx::ex2(x);
}
// The tensor core adds every K = 16 product block into the fp32 accumulator with truncation, so a sum over the whole
// database loses ~M/16 * 2^-24 of its mass (measured with V = 1: -4e-4 at M = 100k, -6e-3 at M = 1M).  The consumers
// therefore accumulate in TMEM over kAccWindow tiles only, add that window into an fp32 scratch with rounded
// CUDA-core adds and restart from zero: bias <= 1024 * 2^-24 of a window's own contribution (-2e-5 measured).
#ifndef RANGE_PC_ACC_WINDOW
#define RANGE_PC_ACC_WINDOW 128
#endif
constexpr int kAccWindow = RANGE_PC_ACC_WINDOW;
// columns of S a softmax warp evaluates per step (one TMEM load in flight while the previous piece is evaluated)
// Measured: apply 16 (20.2 ms; 32 spills under the 128-register cap: 21.9 ms), statistics 32 (4.56 vs 4.65 ms).
constexpr int kPieceApply = 16, kPieceStats = 32;
template <int W>
__device__ __forceinline__ void tmem_ld_piece(uint32_t taddr, uint32_t (&v)[W]) {
  static_assert(W == 16 || W == 32, "piece width");
  if constexpr (W == 16) ptx::tmem_ld16(taddr, v);
  else ptx::tmem_ld32(taddr, v);
}
constexpr int kFlagStride = 32;                  // uint32 per flag line (128 B)
constexpr int kFlagsPerProducer = 3 * kFlagStride;   // full, done[0], done[1]

struct ProdSmem {
  static constexpr int NS = 4, NB = 4;           // K stages (32 KB: this CTA's 64 entries x 256 dims), S buffers
  static constexpr int q = 0;                    // 4 x [128 rows x 64 dims] SW128
  static constexpr int stages = q + 65536;
  static constexpr int xyz = stages + NS * 32768;
  static constexpr int bars = xyz + NB * kXyzBytes;
  static constexpr int b_q_full = 0, b_q_pair = 1, b_q_empty = 2;
  static constexpr int b_stage_full = 3;
  static constexpr int b_stage_empty = b_stage_full + NS;
  static constexpr int b_s_full = b_stage_empty + NS;
  static constexpr int b_s_empty = b_s_full + NB;
  static constexpr int b_xyz_empty = b_s_empty + NB;
  static constexpr int b_slot_free = b_xyz_empty + NB;
  static constexpr int b_p_written = b_slot_free + kRing;
  static constexpr int n_bars = b_p_written + kRing;
  static constexpr int tmem_slot = bars + n_bars * 8;
  static constexpr int total = tmem_slot + 16;
};
struct ConsSmem {
  // half tiles (64 entries).  Vt: 4 stages of [2 blocks][128 dims x 64 entries] SW128 (32 KB); P': 6 stages of
  // [8 key chunks][128 rows][16 B] (16 KB, no-swizzle K-major core matrices) - P' buffers three tiles ahead so the
  // L2 round trips of the hand-off stay off the tensor pipe's critical path.
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

// Work decomposition.  Unit u handles query-tile pair r * n_units + u in round r < full_rounds (whole database,
// output written directly).  The remaining tail_pairs pairs would keep only tail_pairs of the units busy for a
// whole round, so each of them is split over tail_split database ranges of tail_tiles tiles: tail_pairs *
// tail_split units work for 1 / tail_split of a round and write partial outputs that the host sums.
struct PcPlan {
  int n_units, full_rounds, tail_pairs, tail_split, tail_tiles;
  int tail_row0;              // first query row of the tail pairs
  size_t part_stride;         // floats between the partial outputs of two splits
  // where finished rows go: row n -> out + (perm ? perm[n] : n) * out_ld + column, as fp32 or fp64 (fused concat:
  // the caller's (N, 1280) result, range/range.py:222,240) - or the plain (N, 1024) fp32 O
  int out_ld, out_f64;
  const int* perm;
  // M-sharded database: finished (partial) rows go to the owner ranks' receive buffers over NVLink instead
  rangeb200::RowRoute route;
};
struct PcWork {
  int qp, t0, t1, split;      // query-tile pair, database tiles [t0, t1), split index or -1 (direct output)
};
__device__ __forceinline__ int pc_rounds(const PcPlan& p, int unit) {
  return p.full_rounds + (unit < p.tail_pairs * p.tail_split ? 1 : 0);
}
__device__ __forceinline__ PcWork pc_work(const PcPlan& p, int unit, int r, int T) {
  if (r < p.full_rounds) return PcWork{r * p.n_units + unit, 0, T, -1};
  const int s = unit % p.tail_split;
  const int t0 = s * p.tail_tiles;
  return PcWork{p.full_rounds * p.n_units + unit / p.tail_split, t0, min(T, t0 + p.tail_tiles), p.tail_split > 1 ? s : -1};
}

template <bool kGeo>
__global__ void __launch_bounds__(kThreads, 1)
range_apply_pc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK64,
                      const __grid_constant__ CUtensorMap tmV128, const __grid_constant__ CUtensorMap tmP,
                      const float4* __restrict__ db_xyz, const float4* __restrict__ rowc, int N, int M, float a_sem,
                      void* __restrict__ out, const uint32_t* __restrict__ geo_mask, int mask_words,
                      float* __restrict__ part, float4* __restrict__ acc_scratch, __half* __restrict__ ring,
                      uint32_t* __restrict__ flags,
                      uint32_t* __restrict__ windows, const PcPlan plan, int dbg, long long* __restrict__ prof) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  // developer instrumentation (tools/time_apply.py): wait-cycle accounting of unit 1's leader CTAs
  long long pacc[6] = {0, 0, 0, 0, 0, 0};
  const int cid_ = blockIdx.x >> 1;
  const int unit_ = cid_ / 3, role_ = cid_ % 3;        // a unit's producer pair and two consumer pairs are consecutive clusters
  // (measured: placing all producer pairs first and a unit's clusters n_units apart is 0.1-0.4 ms slower)
  const bool prof_on = prof != nullptr && unit_ == 1 && rank == 0;
#define PC_T0() long long _t0 = prof_on ? clock64() : 0
#define PC_ADD(k) do { if (prof_on) { long long _t1 = clock64(); pacc[k] += _t1 - _t0; _t0 = _t1; } } while (0)
#define PC_OUT(base) do { if (prof_on) { for (int _k = 0; _k < 6; ++_k) prof[(base) + _k] = pacc[_k]; } } while (0)
  const bool leader = rank == 0;
  const int unit = unit_, role = role_;                     // role 0: producer pair; 1, 2: consumer pairs
  const int n_units = plan.n_units;
  const bool active = unit < n_units;
  const bool producer = role == 0;
  const int T = (M + kKeys - 1) / kKeys;
  const int rounds = active ? pc_rounds(plan, unit) : 0;
  const int tail_items = plan.tail_pairs * plan.tail_split;
  const int tail_len = tail_items ? (plan.tail_split > 1 ? plan.tail_tiles : T) : 0;       // longest tail item, tiles
  const int sync_units = plan.full_rounds > 0 ? n_units : tail_items;                       // units that have work
  const int n_windows = int((uint32_t(plan.full_rounds) * uint32_t(T) + uint32_t(tail_len) + kWindow - 1) / kWindow);
  // tiles this unit processes over the whole launch (ring positions, `full` / `done` counters count these)
  const uint32_t my_tiles = uint32_t(plan.full_rounds) * uint32_t(T) +
                            (rounds > plan.full_rounds ? uint32_t(pc_work(plan, unit, plan.full_rounds, T).t1 -
                                                                  pc_work(plan, unit, plan.full_rounds, T).t0) : 0u);
  const int prod_id = unit * 2 + int(rank);                 // the producer CTA this CTA is / listens to
  uint32_t* full_flag = flags + size_t(prod_id) * kFlagsPerProducer;

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
      for (int i = 0; i < L::NB; ++i) {
        ptx::mbar_init(&bars[L::b_s_full + i], kGeo ? 2 : 1);            // MMA commit (+ this CTA's xyz bytes)
        ptx::mbar_init(&bars[L::b_s_empty + i], 2 * 4);                  // the owning group's 4 warps in both CTAs
        ptx::mbar_init(&bars[L::b_xyz_empty + i], 4);
      }
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
    // producer pair: query tiles 2 qp (leader) and 2 qp + 1 (peer)
    // =====================================================================================================
    using L = ProdSmem;
    const uint32_t bars_u = smem_u + L::bars;
    if (warp == kWarpTma) {
      if (lane == 0) {
        ptx::prefetch_tmap(&tmQ);
        ptx::prefetch_tmap(&tmK64);
        PipeState st;
        uint32_t it = 0;
        for (int r = 0; r < rounds; ++r) {
          const PcWork wk = pc_work(plan, unit, r, T);
          const int qt = 2 * wk.qp + int(rank);
          if (r > 0) ptx::mbar_wait(&bars[L::b_q_empty], (r - 1) & 1);        // every Q.K^T of the last round has read Q
          ptx::mbar_expect_tx(&bars[L::b_q_full], 65536);
          for (int c = 0; c < 4; ++c)
            ptx::tma_load_2d(smem + L::q + c * 16384, &tmQ, &bars[L::b_q_full], c * 64, qt * kBlockQ);
          for (int j = wk.t0; j < wk.t1; ++j, ++it) {
            const int key0 = j * kKeys;
            if (leader && (it % kWindow) == 0) {
              // Keep the units within two windows of each other: they all stream the same database, and only tiles
              // the other units touched recently are still in L2 (measured without this: 33 GB of DRAM reads per launch).
              const uint32_t w = it / kWindow;
              atomicAdd(&windows[w], 1u);
              if (w > 0) ptx::wait_flag_ge(&windows[w - 1], uint32_t(sync_units));
            }
            PC_T0();
            ptx::mbar_wait(&bars[L::b_stage_empty + st.idx], st.phase ^ 1);
            PC_ADD(0);
            uint8_t* dst = smem + L::stages + st.idx * 32768;
            if (leader) ptx::mbar_expect_tx(&bars[L::b_stage_full + st.idx], 65536);
#pragma unroll
            for (int c = 0; c < 4; ++c)       // this CTA's 64 entries x 256 dims, as 4 [64 x 64] boxes
              ptx::tma_load_2d_2sm(dst + c * 8192, &tmK64, &bars[L::b_stage_full + st.idx], c * 64, key0 + int(rank) * 64);
            st.advance<L::NS>();
          }
        }
        if (leader)       // units with fewer rounds: check in for the windows the others still have to pass
          for (uint32_t w = (it + kWindow - 1) / kWindow; w < uint32_t(n_windows); ++w) atomicAdd(&windows[w], 1u);
        PC_OUT(0);
      }
    } else if (warp == kWarpXyz) {
      // ----- xyz loader: entry unit vectors of tile it -> slot it & 3, bytes credited to the barrier S(it) arrives on.
      // Its own warp: the slot frees only when softmax(it - 4) is done, and no K load may queue behind that wait.
      if (kGeo && lane == 0) {
        uint32_t it = 0;
        for (int r = 0; r < rounds; ++r) {
          const PcWork wk = pc_work(plan, unit, r, T);
          const int qt = 2 * wk.qp + int(rank);
          const uint32_t* mask_row = geo_mask ? geo_mask + size_t(qt) * mask_words : nullptr;
          for (int j = wk.t0; j < wk.t1; ++j, ++it) {
            const int x = it & (L::NB - 1);
            PC_T0();
            ptx::mbar_wait(&bars[L::b_xyz_empty + x], ((it / L::NB) & 1) ^ 1);
            PC_ADD(2);
            const bool skip = mask_row != nullptr && ((__ldg(mask_row + (j >> 5)) >> (j & 31)) & 1u);
            if (skip) {
              ptx::mbar_arrive(&bars[L::b_s_full + x]);
            } else {
              ptx::mbar_expect_tx(&bars[L::b_s_full + x], kXyzBytes);
              ptx::bulk_load_1d(smem + L::xyz + x * kXyzBytes, db_xyz + j * kKeys, kXyzBytes, &bars[L::b_s_full + x]);
            }
          }
        }
      }
    } else if (warp == kWarpMma) {
      constexpr uint32_t idesc_qk = ptx::umma_idesc_f16(2 * kBlockQ, kKeys);
      PipeState st;
      uint32_t it = 0;
      for (int r = 0; r < rounds; ++r) {
        ptx::mbar_wait(&bars[L::b_q_full], r & 1);
        if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&bars[L::b_q_pair]), 0));
        __syncwarp();
        if (!leader) continue;
        ptx::mbar_wait_cluster(&bars[L::b_q_pair], r & 1);
        const PcWork wk = pc_work(plan, unit, r, T);
        for (int j = wk.t0; j < wk.t1; ++j, ++it) {
          const int b = it & (L::NB - 1);
          PC_T0();
          ptx::mbar_wait_cluster(&bars[L::b_s_empty + b], ((it / L::NB) & 1) ^ 1);   // S(it - 4) is in registers
          PC_ADD(0);
          ptx::mbar_wait(&bars[L::b_stage_full + st.idx], st.phase);
          PC_ADD(1);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t b_base = smem_u + L::stages + st.idx * 32768;
            const uint32_t a_base = smem_u + L::q;
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                ptx::umma_f16_ss_2sm(tmem_base + b * kKeys, ptx::umma_desc_kmajor_sw128(a_base + c * 16384 + kk * 32),
                                     ptx::umma_desc_kmajor_sw128(b_base + c * 8192 + kk * 32), idesc_qk, (c | kk) != 0);
            ptx::umma_commit_2sm_u32(bars_u + 8 * (L::b_stage_empty + st.idx));
            ptx::umma_commit_2sm_u32(bars_u + 8 * (L::b_s_full + b));
          }
          __syncwarp();
          PC_ADD(2);
          st.advance<L::NS>();
        }
        if (ptx::elect_one()) ptx::umma_commit_2sm_u32(bars_u + 8 * L::b_q_empty);
        __syncwarp();
      }
      if (lane == 0) PC_OUT(8);
    } else if (warp == kWarpPublish) {
      // ----- ring bookkeeping (one thread, two interleaved non-blocking streams) -----
      //   gate:    slot it % kRing is free once both consumer pairs have copied tile it - kRing (their `done` counters)
      //   publish: once the owning softmax group has stored tile it, advance the `full` counter.  A gpu-scope release
      //            costs ~1500 clk here (every store the SM has in flight must be acknowledged), so tiles are
      //            published in batches of kPublishBatch; the ring absorbs the added latency.
      if (lane == 0) {
        const uint32_t total = my_tiles;
        uint32_t pub = 0, gate = 0, seen = 0, spins = 0;
        while (pub < total || gate < total) {
          bool progress = false;
          if (gate < total) {
            const uint32_t need = gate >= uint32_t(kRing) && !(dbg & 2) ? gate - kRing + 1 : 0u;
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
          else if (++spins > (1u << 24)) {
            printf("range_b200: ring bookkeeping stalled (block %d, published %u gated %u of %u)\n", blockIdx.x, pub, gate, total);
            __trap();
          }
        }
      }
    } else {
      // ----- softmax warps: S (TMEM fp32) -> P' (fp16) -> ring in global memory -----
      // Four groups of four warps (one warp per TMEM lane quarter).  Group g owns S buffer g, i.e. the tiles
      // it = g (mod 4), and walks a tile in four 32-entry chunks.  The groups drift apart, so while one waits
      // for its next S tile, a TMEM load or the ring, the SM sub-partition's MUFU pipe is fed by the other three.
      const int grp = warp >> 2, quarter = warp & 3;
      const int row = quarter * 32 + lane;
      const uint32_t s_empty_leader0 = ptx::mapa(ptx::smem_u32(&bars[L::b_s_empty]), 0);
      // slot s of this producer: ring + (prod_id * kRing + s) * 128 * 128 halves, laid out [16 key chunks][128 rows][8]
      __half* ring_row = ring + size_t(prod_id) * kRing * 128 * 128 + row * 8;
      const uint32_t total = my_tiles, base_tail = uint32_t(plan.full_rounds) * uint32_t(T);
      int cur_r = -1, n = 0, tail_t0 = 0;
      const uint32_t* mask_row = nullptr;
      float cs = -INFINITY, cg = -INFINITY, gx = 0.f, gy = 0.f, gz = 0.f;
      uint32_t bufA[kPieceApply], bufB[kPieceApply];
      bool prefetched = false;        // the first piece of this tile's S was requested at the end of the last tile
      for (uint32_t it = uint32_t(grp); it < total; it += kGroups) {
        const int r = it < base_tail ? int(it / uint32_t(T)) : plan.full_rounds;
        const int b = it & (L::NB - 1);
        const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + b * kKeys;
        if (r != cur_r) {
          cur_r = r;
          const PcWork wk = pc_work(plan, unit, r, T);
          tail_t0 = wk.t0;
          const int qt = 2 * wk.qp + int(rank);
          n = qt * kBlockQ + row;
          mask_row = (kGeo && geo_mask) ? geo_mask + size_t(qt) * mask_words : nullptr;
          cs = -INFINITY; cg = -INFINITY; gx = gy = gz = 0.f;
          if (n < N) {
            const float4 c0 = rowc[2 * n], c1 = rowc[2 * n + 1];
            cs = c0.x; cg = c0.y; gx = c0.z; gy = c0.w; gz = c1.x;
          }
        }
        const int j = it < base_tail ? int(it - uint32_t(r) * uint32_t(T)) : tail_t0 + int(it - base_tail);   // database tile
        const int slot = it % kRing;
        const bool with_geo = kGeo && !(mask_row != nullptr && ((__ldg(mask_row + (j >> 5)) >> (j & 31)) & 1u));
        __half* dst = ring_row + size_t(slot) * 128 * 128;
        PC_T0();
        if (!prefetched) {
          ptx::mbar_wait(&bars[L::b_s_full + b], (it / L::NB) & 1);      // S(it) in TMEM and xyz(it) in smem
          ptx::tc_fence_after();
          tmem_ld_piece(taddr, bufA);
        }
        PC_ADD(0);
        // 128 / kPieceApply pieces; the TMEM load of piece h + 1 is in flight while piece h is evaluated
        auto piece = [&](const uint32_t (&cur)[kPieceApply], int h) {
          const uint32_t kxyz = smem_u + L::xyz + b * kXyzBytes + h * kPieceApply * 16;
          const int nvalid = M - (j * kKeys + h * kPieceApply);            // >= kPieceApply except in the last tile
          uint32_t packed[kPieceApply / 2];
          auto body = [&](auto masked, auto geo) {
            constexpr bool kM = decltype(masked)::value, kG = decltype(geo)::value;
#pragma unroll
            for (int w = 0; w < kPieceApply / 2; ++w) {
              float pv[2];
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const int i = 2 * w + u;
                float p = ex2_mixed(fmaf(__uint_as_float(cur[i]), a_sem, cs), i);
                if (kG) {
                  const float4 k = ptx::lds_f4(kxyz + i * 16);
                  p += ptx::ex2(fmaf(gx, k.x, fmaf(gy, k.y, fmaf(gz, k.z, cg))));
                }
                if (kM && i >= nvalid) p = 0.f;
                pv[u] = p;
              }
              packed[w] = ptx::pack_half2(pv[0], pv[1]);
            }
          };
          if (dbg & 8) {                                  // developer switch: no exponentials (consumer-bound run)
#pragma unroll
            for (int w = 0; w < kPieceApply / 2; ++w) packed[w] = cur[2 * w] & 0x3c003c00u;
          } else if (nvalid >= kPieceApply) {
            if (with_geo) body(std::false_type{}, std::true_type{}); else body(std::false_type{}, std::false_type{});
          } else {
            if (with_geo) body(std::true_type{}, std::true_type{}); else body(std::true_type{}, std::false_type{});
          }
          PC_ADD(2);
          if (h == 0) ptx::mbar_wait(&bars[L::b_slot_free + slot], (it / kRing) & 1);   // both consumers copied tile it - kRing
          PC_ADD(3);
          // ring tile layout [16 key chunks][128 rows][8 entries]: a warp's store covers 512 contiguous bytes
#pragma unroll
          for (int e = 0; e < kPieceApply / 8; ++e)
            ptx::stg_u4(dst + ((h * (kPieceApply / 8) + e) * 128) * 8, packed[4 * e], packed[4 * e + 1], packed[4 * e + 2],
                        packed[4 * e + 3]);
          PC_ADD(4);
        };
        constexpr int kPairs = 128 / (2 * kPieceApply);
#pragma unroll 1
        for (int hp = 0; hp < kPairs; ++hp) {
          ptx::tmem_ld_wait();
          PC_ADD(1);
          tmem_ld_piece(taddr + (2 * hp + 1) * kPieceApply, bufB);
          piece(bufA, 2 * hp);
          ptx::tmem_ld_wait();
          PC_ADD(1);
          if (hp < kPairs - 1) {
            tmem_ld_piece(taddr + (2 * hp + 2) * kPieceApply, bufA);
          } else {                              // the whole S tile is in registers: the MMA warp may overwrite the buffer
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (leader) ptx::mbar_arrive(&bars[L::b_s_empty + b]);
              else ptx::mbar_arrive_cluster_relaxed(s_empty_leader0 + 8 * b);
            }
            // the group's next tile normally sits in the spare buffer already: start its first TMEM load now, so the
            // barrier wait and the load latency hide behind the last piece of this tile
            const uint32_t nx = it + kGroups;
            prefetched = __all_sync(0xffffffffu, nx < total && ptx::mbar_try_wait(&bars[L::b_s_full + (nx & (L::NB - 1))],
                                                                                 (nx / L::NB) & 1));
            if (prefetched) {
              ptx::tc_fence_after();
              tmem_ld_piece(tmem_base + (uint32_t(quarter * 32) << 16) + (nx & (L::NB - 1)) * kKeys, bufA);
            }
          }
          piece(bufB, 2 * hp + 1);
        }
        __syncwarp();
        if (lane == 0) {
          ptx::mbar_arrive(&bars[L::b_p_written + slot]);
          if (kGeo) ptx::mbar_arrive(&bars[L::b_xyz_empty + b]);
        }
        PC_ADD(4);
      }
      if (threadIdx.x == 0) PC_OUT(32);
    }
  } else if (active && rounds > 0) {
    // =====================================================================================================
    // consumer pair: value dims [512 (role - 1), +512) of the unit's two query tiles
    // =====================================================================================================
    using L = ConsSmem;
    const uint32_t bars_u = smem_u + L::bars;
    const int cp = role - 1;
    const int dimbase = cp * 512;
    if (warp == 0) {
      // ----- Vt loader: never waits on the producer -----
      if (lane == 0) {
        ptx::prefetch_tmap(&tmV128);
        PipeState st;
        for (int r = 0; r < rounds; ++r) {
          const PcWork wk = pc_work(plan, unit, r, T);
          for (int j = wk.t0; j < wk.t1; ++j) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const int key0 = j * kKeys + half * 64;
              PC_T0();
              ptx::mbar_wait(&bars[L::b_v_empty + st.idx], st.phase ^ 1);
              PC_ADD(0);
              uint8_t* dst = smem + L::v + st.idx * L::kStageV;
              if (leader) ptx::mbar_expect_tx(&bars[L::b_v_full + st.idx], 2 * L::kStageV);
#pragma unroll
              for (int nb = 0; nb < 2; ++nb)     // this CTA's 128 dims of each 256-wide block x 64 entries
                ptx::tma_load_2d_2sm(dst + nb * 16384, &tmV128, &bars[L::b_v_full + st.idx], key0,
                                     dimbase + nb * 256 + int(rank) * 128);
              st.advance<L::NV>();
            }
          }
        }
        if (role == 1) PC_OUT(40);
      }
    } else if (warp == 6) {
      // ----- P' loader: follows the producer's `full` counter -----
      if (lane == 0) {
        ptx::prefetch_tmap(&tmP);
        PipeState st;
        uint32_t seen = 0;
        {
          for (uint32_t it = 0; it < my_tiles; ++it) {
            const int slot = it % kRing;
            PC_T0();
            if (seen < it + 1 && !(dbg & 1)) {               // the counter advances in batches: one poll + proxy fence per batch
              uint32_t spins = 0;
              while ((seen = ptx::ld_acquire_gpu(full_flag)) < it + 1) {
                if (++spins > (1u << 24)) {
                  printf("range_b200: consumer %d waits for tile %u, producer published %u\n", blockIdx.x, it, seen);
                  __trap();
                }
              }
              PC_ADD(0);
              ptx::fence_proxy_async_all();
              PC_ADD(1);
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              ptx::mbar_wait(&bars[L::b_p_empty + st.idx], st.phase ^ 1);
              PC_ADD(2);
              if (leader) ptx::mbar_expect_tx(&bars[L::b_p_full + st.idx], 2 * L::kStageP);
              // 8 key chunks x 2 KB: rows (prod * kRing + slot) * 16 + half * 8 .. + 8 of the ring viewed as [.][2 KB]
              ptx::tma_load_2d_2sm(smem + L::p + st.idx * L::kStageP, &tmP, &bars[L::b_p_full + st.idx], 0,
                                   (prod_id * kRing + slot) * 16 + half * 8);
              st.advance<L::NP>();
            }
          }
        }
        if (role == 1) PC_OUT(48);
      }
    } else if (warp == 1) {
      if (leader) {
        constexpr uint32_t idesc_pv = ptx::umma_idesc_f16(2 * kBlockQ, 256);
        uint32_t* done0 = flags + size_t(unit * 2) * kFlagsPerProducer + (1 + cp) * kFlagStride;       // producer rank 0
        uint32_t* done1 = flags + size_t(unit * 2 + 1) * kFlagsPerProducer + (1 + cp) * kFlagStride;   // producer rank 1
        PipeState sv, sp;
        uint32_t it = 0, ev = 0;          // ev: accumulation windows handed to the epilogue warps so far
        for (int r = 0; r < rounds; ++r) {
          const PcWork wk = pc_work(plan, unit, r, T);
          const int nt = wk.t1 - wk.t0;
          for (int j = 0; j < nt; ++j, ++it) {
            const bool first = j % kAccWindow == 0;                       // first tile of an accumulation window
            if (first && ev > 0) {
              ptx::mbar_wait_cluster(&bars[L::b_o_empty], (ev - 1) & 1);  // both epilogues have read the last window
              ptx::tc_fence_after();
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              PC_T0();
              ptx::mbar_wait(&bars[L::b_p_full + sp.idx], sp.phase);
              PC_ADD(0);
              if (half == 1 && ptx::elect_one()) {
                // both halves of both CTAs' P'(it) are in shared memory: the ring slots are free.  Relaxed is enough:
                // the TMA reads of the slot completed (mbarrier complete_tx) before these stores issue.
                ptx::st_relaxed_gpu(done0, it + 1);
                ptx::st_relaxed_gpu(done1, it + 1);
              }
              ptx::mbar_wait(&bars[L::b_v_full + sv.idx], sv.phase);
              PC_ADD(2);
              ptx::tc_fence_after();
              if (ptx::elect_one()) {
                const uint32_t a_base = smem_u + L::p + sp.idx * L::kStageP;
                const uint32_t b_base = smem_u + L::v + sv.idx * L::kStageV;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                  for (int nb = 0; nb < 2; ++nb)
                    if (!(dbg & 4))                          // developer switch: no P.V (producer-bound run)
                    ptx::umma_f16_ss_2sm(tmem_base + nb * 256, ptx::umma_desc_kmajor_nosw(a_base + kk * 4096, 2048, 128),
                                         ptx::umma_desc_kmajor_sw128(b_base + nb * 16384 + kk * 32), idesc_pv,
                                         !(first && half == 0 && kk == 0));
                ptx::umma_commit_2sm_u32(bars_u + 8 * (L::b_v_empty + sv.idx));
                ptx::umma_commit_2sm_u32(bars_u + 8 * (L::b_p_empty + sp.idx));
              }
              __syncwarp();
              PC_ADD(1);
              sv.advance<L::NV>();
              sp.advance<L::NP>();
            }
            if ((j + 1) % kAccWindow == 0 || j + 1 == nt) {               // window (or round) complete
              if (ptx::elect_one()) ptx::umma_commit_2sm_u32(bars_u + 8 * L::b_o_full);
              __syncwarp();
              ++ev;
            }
          }
        }
        if (lane == 0) PC_OUT(role == 1 ? 56 : 24);
      }
    } else if (warp >= 2 && warp < 6) {
      // ----- epilogue warps: O (TMEM, 128 lanes x 512 columns) -> global fp32, scaled -----
      const int quarter = warp & 3;
      const int row = quarter * 32 + lane;
      const uint32_t o_empty_leader = ptx::mapa(ptx::smem_u32(&bars[L::b_o_empty]), 0);
      // this CTA's window scratch: [16 column blocks][8 float4][128 rows] float4 (a warp's access is 512 contiguous bytes)
      float4* scratch = acc_scratch + size_t((unit * 2 + cp) * 2 + int(rank)) * (128 * 128) + row;
      const uint64_t keep = ptx::l2_policy_evict_last();
      uint32_t ev = 0;
      for (int r = 0; r < rounds; ++r) {
        const PcWork wk = pc_work(plan, unit, r, T);
        const int qt = 2 * wk.qp + int(rank);
        const int n = qt * kBlockQ + row;
        const float out_scale = n < N ? rowc[2 * n + 1].y : 0.f;
        // whole database: the final rows (caller's layout, dtype and row order); one range of a split tail pair: that
        // split's fp32 partial rows (summed and placed by the host's reduce kernel)
        const bool direct = wk.split < 0;
        const bool routed = plan.route.n_ranks > 0;
        const size_t drow = direct && !routed ? size_t(plan.perm && n < N ? plan.perm[n] : n) * plan.out_ld : 0;
        float* orow32 = !direct ? part + size_t(wk.split) * plan.part_stride + size_t(n - plan.tail_row0) * 1024 + dimbase
                        : routed ? rangeb200::route_row(plan.route, n) + dimbase       // peer memory: stores cross NVLink
                                 : reinterpret_cast<float*>(out) + drow + dimbase;
        double* orow64 = reinterpret_cast<double*>(out) + drow + dimbase;
        const bool f64 = direct && !routed && plan.out_f64;
        // routed: the tile's 128 rows belong to one owner (slab_rows is a multiple of 128): row 0 of the tile there
        float* peer_tile = routed ? rangeb200::route_row(plan.route, qt * kBlockQ) : nullptr;
        const int nseg = (wk.t1 - wk.t0 + kAccWindow - 1) / kAccWindow;
        for (int sg = 0; sg < nseg; ++sg, ++ev) {
          const bool last = sg == nseg - 1;
          // the windows accumulated so far come back from the scratch one column block ahead of the TMEM loads (the
          // first block before the window's last product has even completed): an L2 round trip per block would
          // otherwise be exposed sixteen times per flush, with the pair's tensor pipes idle meanwhile
          const bool acc = sg > 0 && n < N;
          float4 a0[8], a1[8];
          auto load_window = [&](int cc, float4 (&a)[8]) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = ptx::ldg_f4_hint(scratch + size_t(cc) * 8 * 128 + i * 128, keep);
          };
          auto flush_block = [&](int cc, uint32_t (&v)[32], const float4 (&a)[8]) {
            const bool to_peer = last && direct && routed;     // warp-collective below: no lane may leave early
            if (n >= N && !to_peer) return;
            if (acc) {                            // rounded fp32 adds
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
            } else if (to_peer) {
              // Partial rows of an M-sharded database leave for the owner rank's receive buffer over NVLink.  Peer stores
              // are not merged by the local L2, so a thread storing its own row 16 bytes at a time would send 16-byte
              // NVLink writes (measured: 60 GB/s at 8 GPUs).  The warp therefore transposes the 32 x 32 block with
              // shuffles and every store instruction writes four rows x 128 contiguous bytes (full lines).
              const int sub = lane >> 3, grp = lane & 7;
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * out_scale);
              float* tile = peer_tile + size_t(quarter * 32) * 1024 + dimbase + cc * 32 + grp * 4;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int src = 4 * i + sub;                  // row of the warp's 32 this lane stores for
                uint32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0;
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                  const uint32_t b0 = __shfl_sync(0xffffffffu, v[4 * g], src), b1 = __shfl_sync(0xffffffffu, v[4 * g + 1], src);
                  const uint32_t b2 = __shfl_sync(0xffffffffu, v[4 * g + 2], src), b3 = __shfl_sync(0xffffffffu, v[4 * g + 3], src);
                  if (grp == g) { o0 = b0; o1 = b1; o2 = b2; o3 = b3; }
                }
                if (qt * kBlockQ + quarter * 32 + src < N)
                  *reinterpret_cast<uint4*>(tile + size_t(src) * 1024) = make_uint4(o0, o1, o2, o3);
              }
            } else if (f64) {                     // result rows are written once and never read here: streaming stores
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
  ptx::cluster_sync();                           // neither CTA may free TMEM / exit while the pair is in flight
  if (active && warp == kWarpMma) ptx::tmem_dealloc_2sm<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------
// K2a (large batches): row statistics with the producer's structure - CTA pairs (two query tiles, half a K tile
// loaded per CTA), four S buffers, four desynchronised softmax groups, pipelined 16-column TMEM loads.  The
// exponentials are the whole cost of this pass (the MUFU pipe), so what matters is that some warp of every SM
// sub-partition always has exponentials to issue; the 16-warp lockstep kernel in retrieval.cu reaches 78 % of the
// MUFU peak, this one ~95 %.  Per row: sum_j 2^(a (s_j - 1)), max_j s_j (and the same for g) over the tiles
// [split * tiles_per_split, ...) - partials over splits / ranks merge with SUM and MAX (reference: the softmax
// denominators of range/range.py:215,234).
// ---------------------------------------------------------------------------------------------------
struct StatSmem {
  static constexpr int NS = 4, NB = 4;           // K stages (32 KB: this CTA's 64 entries x 256 dims), S buffers
  static constexpr int q = 0;
  static constexpr int stages = q + 65536;
  static constexpr int xyz = stages + NS * 32768;
  static constexpr int red = xyz + NB * kXyzBytes;             // cross-group reduction scratch [groups - 1][128] float4
  static constexpr int bars = red + 3 * kBlockQ * 16;
  static constexpr int b_q_full = 0, b_q_pair = 1;
  static constexpr int b_stage_full = 2;
  static constexpr int b_stage_empty = b_stage_full + NS;
  static constexpr int b_s_full = b_stage_empty + NS;
  static constexpr int b_s_empty = b_s_full + NB;
  static constexpr int b_xyz_empty = b_s_empty + NB;
  static constexpr int n_bars = b_xyz_empty + NB;
  static constexpr int tmem_slot = bars + n_bars * 8;
  static constexpr int total = tmem_slot + 16;
  static constexpr int dynamic_bytes = total + 1024;
};

template <bool kGeo>
__global__ void __launch_bounds__(kThreads, 1)
range_stats_pc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK64,
                      const float4* __restrict__ db_xyz, const float4* __restrict__ q_xyz, int N, int M,
                      int tiles_per_split, float a_sem, float a_geo, float* __restrict__ part_sum,
                      float* __restrict__ part_max, const uint32_t* __restrict__ geo_mask, int mask_words) {
  using L = StatSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::tmem_slot);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int qt = blockIdx.x;                       // cluster (2,1,1): blockIdx.x = 2 * pair + rank
  const int split = blockIdx.y;
  const int total_tiles = (M + kKeys - 1) / kKeys;
  const int t_begin = split * tiles_per_split;
  const int T = min(total_tiles, t_begin + tiles_per_split) - t_begin;
  const uint32_t* mask_row = (kGeo && geo_mask) ? geo_mask + size_t(qt) * mask_words : nullptr;
  auto skip_geo = [&](int t) { return mask_row != nullptr && ((__ldg(mask_row + (t >> 5)) >> (t & 31)) & 1u); };

  if (threadIdx.x == 0) {
    ptx::mbar_init(&bars[L::b_q_full], 1);
    ptx::mbar_init(&bars[L::b_q_pair], 2);
    for (int i = 0; i < L::NS; ++i) {
      ptx::mbar_init(&bars[L::b_stage_full + i], 1);
      ptx::mbar_init(&bars[L::b_stage_empty + i], 1);
    }
    for (int i = 0; i < L::NB; ++i) {
      ptx::mbar_init(&bars[L::b_s_full + i], kGeo ? 2 : 1);            // MMA commit (+ this CTA's xyz bytes)
      ptx::mbar_init(&bars[L::b_s_empty + i], 2 * 4);                  // the owning group's 4 warps in both CTAs
      ptx::mbar_init(&bars[L::b_xyz_empty + i], 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == kWarpMma) ptx::tmem_alloc_2sm<512>(tmem_slot);
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t smem_u = __shfl_sync(0xffffffffu, ptx::smem_u32(smem), 0);
  const uint32_t bars_u = smem_u + L::bars;

  if (warp == kWarpTma) {
    if (lane == 0 && T > 0) {
      ptx::prefetch_tmap(&tmQ);
      ptx::prefetch_tmap(&tmK64);
      ptx::mbar_expect_tx(&bars[L::b_q_full], 65536);
      for (int c = 0; c < 4; ++c) ptx::tma_load_2d(smem + L::q + c * 16384, &tmQ, &bars[L::b_q_full], c * 64, qt * kBlockQ);
      PipeState st;
      for (int j = 0; j < T; ++j) {
        ptx::mbar_wait(&bars[L::b_stage_empty + st.idx], st.phase ^ 1);
        uint8_t* dst = smem + L::stages + st.idx * 32768;
        if (leader) ptx::mbar_expect_tx(&bars[L::b_stage_full + st.idx], 65536);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          ptx::tma_load_2d_2sm(dst + c * 8192, &tmK64, &bars[L::b_stage_full + st.idx], c * 64,
                               (t_begin + j) * kKeys + int(rank) * 64);
        st.advance<L::NS>();
      }
    }
  } else if (warp == kWarpXyz) {
    if (kGeo && lane == 0) {
      for (int j = 0; j < T; ++j) {
        const int x = j & (L::NB - 1);
        ptx::mbar_wait(&bars[L::b_xyz_empty + x], ((j / L::NB) & 1) ^ 1);
        if (skip_geo(t_begin + j)) {
          ptx::mbar_arrive(&bars[L::b_s_full + x]);
        } else {
          ptx::mbar_expect_tx(&bars[L::b_s_full + x], kXyzBytes);
          ptx::bulk_load_1d(smem + L::xyz + x * kXyzBytes, db_xyz + (t_begin + j) * kKeys, kXyzBytes, &bars[L::b_s_full + x]);
        }
      }
    }
  } else if (warp == kWarpMma) {
    if (T > 0) {
      ptx::mbar_wait(&bars[L::b_q_full], 0);
      if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&bars[L::b_q_pair]), 0));
      __syncwarp();
    }
    if (leader && T > 0) {
      constexpr uint32_t idesc_qk = ptx::umma_idesc_f16(2 * kBlockQ, kKeys);
      ptx::mbar_wait_cluster(&bars[L::b_q_pair], 0);
      PipeState st;
      for (int j = 0; j < T; ++j) {
        const int b = j & (L::NB - 1);
        ptx::mbar_wait_cluster(&bars[L::b_s_empty + b], ((j / L::NB) & 1) ^ 1);
        ptx::mbar_wait(&bars[L::b_stage_full + st.idx], st.phase);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t b_base = smem_u + L::stages + st.idx * 32768;
          const uint32_t a_base = smem_u + L::q;
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              ptx::umma_f16_ss_2sm(tmem_base + b * kKeys, ptx::umma_desc_kmajor_sw128(a_base + c * 16384 + kk * 32),
                                   ptx::umma_desc_kmajor_sw128(b_base + c * 8192 + kk * 32), idesc_qk, (c | kk) != 0);
          ptx::umma_commit_2sm_u32(bars_u + 8 * (L::b_stage_empty + st.idx));
          ptx::umma_commit_2sm_u32(bars_u + 8 * (L::b_s_full + b));
        }
        __syncwarp();
        st.advance<L::NS>();
      }
    }
  } else if (warp < kSoftmaxWarps) {
    const int grp = warp >> 2, quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int n = qt * kBlockQ + row;
    const uint32_t s_empty_leader0 = ptx::mapa(ptx::smem_u32(&bars[L::b_s_empty]), 0);
    float4 qx = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kGeo && n < N) qx = q_xyz[n];
    const float gx = qx.x * a_geo, gy = qx.y * a_geo, gz = qx.z * a_geo;
    float sum_s = 0.f, sum_g = 0.f, max_s = -2.f, max_g = -3.0e38f;
    uint32_t bufA[kPieceStats], bufB[kPieceStats];
    bool prefetched = false;          // the first piece of this tile's S was requested at the end of the last tile
    for (int j = grp; j < T; j += kGroups) {
      const int t = t_begin + j;
      const int b = j & (L::NB - 1);
      const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + b * kKeys;
      const bool with_geo = kGeo && !skip_geo(t);
      if (!prefetched) {
        ptx::mbar_wait(&bars[L::b_s_full + b], (j / L::NB) & 1);      // S(j) in TMEM and xyz(j) in smem
        ptx::tc_fence_after();
        tmem_ld_piece(taddr, bufA);
      }
      auto piece = [&](const uint32_t (&cur)[kPieceStats], int h) {
        const uint32_t kxyz = smem_u + L::xyz + b * kXyzBytes + h * kPieceStats * 16;
        const int nvalid = M - (t * kKeys + h * kPieceStats);            // >= kPieceStats except in the last tile
        auto body = [&](auto masked, auto geo) {
          constexpr bool kM = decltype(masked)::value, kG = decltype(geo)::value;
#pragma unroll
          for (int i = 0; i < kPieceStats; i += 2) {
            float sv[2], gv[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const float s = __uint_as_float(cur[i + u]);
              const bool valid = !kM || (i + u < nvalid);
              float es = ex2_mixed(fmaf(s, a_sem, -a_sem), i + u);
              if (!valid) es = 0.f;
              sum_s += es;
              sv[u] = valid ? s : -2.f;
              if (kG) {
                const float4 k = ptx::lds_f4(kxyz + (i + u) * 16);
                const float g = fmaf(gx, k.x, fmaf(gy, k.y, fmaf(gz, k.z, -a_geo)));   // a_geo (g - 1)
                float eg = ptx::ex2(g);
                if (!valid) eg = 0.f;
                sum_g += eg;
                gv[u] = valid ? g : -3.0e38f;
              }
            }
            max_s = ptx::max3(max_s, sv[0], sv[1]);
            if (kG) max_g = ptx::max3(max_g, gv[0], gv[1]);
          }
        };
        if (nvalid >= kPieceStats) {
          if (with_geo) body(std::false_type{}, std::true_type{}); else body(std::false_type{}, std::false_type{});
        } else {
          if (with_geo) body(std::true_type{}, std::true_type{}); else body(std::true_type{}, std::false_type{});
        }
      };
      constexpr int kPairs = 128 / (2 * kPieceStats);
#pragma unroll 1
      for (int hp = 0; hp < kPairs; ++hp) {
        ptx::tmem_ld_wait();
        tmem_ld_piece(taddr + (2 * hp + 1) * kPieceStats, bufB);
        piece(bufA, 2 * hp);
        ptx::tmem_ld_wait();
        if (hp < kPairs - 1) {
          tmem_ld_piece(taddr + (2 * hp + 2) * kPieceStats, bufA);
        } else {                              // the whole S tile is in registers: the MMA warp may overwrite the buffer
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) ptx::mbar_arrive(&bars[L::b_s_empty + b]);
            else ptx::mbar_arrive_cluster_relaxed(s_empty_leader0 + 8 * b);
          }
          const int nx = j + kGroups;         // start the next tile's first TMEM load behind this tile's last piece
          prefetched = __all_sync(0xffffffffu, nx < T && ptx::mbar_try_wait(&bars[L::b_s_full + (nx & (L::NB - 1))],
                                                                             (nx / L::NB) & 1));
          if (prefetched) {
            ptx::tc_fence_after();
            tmem_ld_piece(tmem_base + (uint32_t(quarter * 32) << 16) + (nx & (L::NB - 1)) * kKeys, bufA);
          }
        }
        piece(bufB, 2 * hp + 1);
      }
      if (kGeo) {
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bars[L::b_xyz_empty + b]);
      }
    }
    // combine the four groups (each saw a quarter of the tiles), write this split's partials
    float4* red = reinterpret_cast<float4*>(smem + L::red);
    if (grp > 0) red[(grp - 1) * kBlockQ + row] = make_float4(sum_s, sum_g, max_s, max_g);
    asm volatile("bar.sync 1, %0;" ::"n"(kSoftmaxWarps * 32) : "memory");
    if (grp == 0 && n < N) {
#pragma unroll
      for (int g2 = 0; g2 < kGroups - 1; ++g2) {
        const float4 o = red[g2 * kBlockQ + row];
        sum_s += o.x;
        sum_g += o.y;
        max_s = fmaxf(max_s, o.z);
        max_g = fmaxf(max_g, o.w);
      }
      const float raw_g = kGeo ? (max_g / a_geo + 1.f) : 0.f;      // max_g holds a_geo (g - 1); store the raw cosine
      reinterpret_cast<float2*>(part_sum)[size_t(split) * N + n] = make_float2(sum_s, sum_g);
      reinterpret_cast<float2*>(part_max)[size_t(split) * N + n] = make_float2(max_s, raw_g);
    }
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (warp == kWarpMma) ptx::tmem_dealloc_2sm<512>(tmem_base);
}

}  // namespace

namespace rangeb200 {

extern long long* g_prof_buffer;      // retrieval.cu (developer instrumentation)

int apply_pc_units(int sm_count) { return (sm_count / 2) / 3; }
size_t apply_pc_ring_bytes(int sm_count) { return size_t(apply_pc_units(sm_count)) * 2 * kRing * 128 * 128 * 2; }
int apply_pc_ring_rows(int sm_count) { return apply_pc_units(sm_count) * 2 * kRing * 16; }   // rows of 2 KB

static PcPlan pc_plan(int sm_count, int64_t N, int64_t M) {
  PcPlan p{};
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
static size_t pc_window_count(const PcPlan& p, int64_t M) {
  const int64_t T = (M + kKeys - 1) / kKeys;
  const int64_t tail_len = p.tail_pairs ? (p.tail_split > 1 ? p.tail_tiles : T) : 0;
  return size_t((int64_t(p.full_rounds) * T + tail_len + kWindow - 1) / kWindow + 1);
}
static size_t pc_ring_flag_bytes(int sm_count) { return size_t(apply_pc_units(sm_count)) * 2 * kFlagsPerProducer * 4; }
size_t apply_pc_flag_bytes(int sm_count, int64_t N, int64_t M) {
  return pc_ring_flag_bytes(sm_count) + pc_window_count(pc_plan(sm_count, N, M), M) * 4;
}
// developer / test hook: {n_units, full_rounds, tail_pairs, tail_split, tail_tiles, tail_row0, windows}
void apply_pc_describe_plan(int sm_count, int64_t N, int64_t M, int32_t out[7]) {
  const PcPlan p = pc_plan(sm_count, N, M);
  out[0] = p.n_units; out[1] = p.full_rounds; out[2] = p.tail_pairs; out[3] = p.tail_split; out[4] = p.tail_tiles;
  out[5] = p.tail_row0; out[6] = int32_t(pc_window_count(p, M));
}
// window scratch of the consumers: [units * 4 CTAs][128 rows x 512 columns] fp32 (L2-resident, 25 MB on 148 SMs)
size_t apply_pc_scratch_bytes(int sm_count) { return size_t(apply_pc_units(sm_count)) * 4 * 128 * 512 * 4; }
// partial outputs of the split tail pairs: [tail_split][tail rows][1024] fp32 (0 when nothing is split)
size_t apply_pc_part_bytes(int sm_count, int64_t N, int64_t M) {
  const PcPlan p = pc_plan(sm_count, N, M);
  return p.tail_split > 1 ? size_t(p.tail_split) * p.part_stride * 4 : 0;
}

// partials [splits][rows][1024] fp32 -> rows row0.. of the caller's output (layout / dtype / row order of PcPlan)
__global__ void reduce_tail_kernel(const float4* __restrict__ part, size_t stride4, int splits, int rows, int row0,
                                   const int* __restrict__ perm, void* __restrict__ out, int out_ld, int out_f64,
                                   const RowRoute route) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= size_t(rows) * 256) return;
  float4 a = part[i];
  for (int k = 1; k < splits; ++k) {
    const float4 b = part[size_t(k) * stride4 + i];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  }
  const int n = row0 + int(i / 256), c = int(i % 256) * 4;
  if (route.n_ranks > 0) {
    *reinterpret_cast<float4*>(route_row(route, n) + c) = a;
    return;
  }
  const size_t o = size_t(perm ? perm[n] : n) * out_ld + c;
  if (out_f64) {
    double* d = reinterpret_cast<double*>(out) + o;
    d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w;
  } else {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + o) = a;
  }
}

cudaError_t launch_apply_pc(const RetrievalArgs& a, const CUtensorMap& tmP, const float* rowc, void* out, int out_ld,
                            int out_f64, const int* perm, const RowRoute* route, void* ring, void* flags, void* part,
                            void* scratch, int sm_count, cudaStream_t stream) {
  PcPlan plan = pc_plan(sm_count, a.N, a.M);
  plan.out_ld = out_ld;
  plan.out_f64 = out_f64;
  plan.perm = perm;
  if (route) plan.route = *route;
#ifdef RANGE_DEVELOPER_SWITCHES       // NVCC_EXTRA=-DRANGE_DEVELOPER_SWITCHES: RANGE_PC_DBG decouples the roles (results are then wrong)
  static const int dbg = getenv("RANGE_PC_DBG") ? atoi(getenv("RANGE_PC_DBG")) : 0;
#else
  constexpr int dbg = 0;
#endif
  cudaError_t e = cudaMemsetAsync(flags, 0, apply_pc_flag_bytes(sm_count, a.N, a.M), stream);
  if (e != cudaSuccess) return e;
  auto kern = a.geo ? range_apply_pc_kernel<true> : range_apply_pc_kernel<false>;
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
  // Every CTA must be resident (the roles wait on each other): the kernel is launched COOPERATIVELY, which either
  // guarantees co-residency or fails with a clean launch error the caller sees (no spinning CTAs, no poisoned context).
  // Round 1 measured the cooperative path 6 % slower; re-measured on the r2a build it is not (21.5 vs 22.1 ms, inside the
  // +-1 ms run-to-run spread of this power-capped kernel).  RANGE_PC_COOP=0 selects the plain cluster launch (capacity
  // checked below) for profilers that cannot replay cooperative launches; the kernel's bounded waits still trap rather
  // than hang if another kernel then holds SMs for long.
  static const bool coop = !(getenv("RANGE_PC_COOP") && atoi(getenv("RANGE_PC_COOP")) == 0);
  cfg.numAttrs = 1;
  // capacity per (device, kernel instantiation): the answer depends on both
  static int resident_cache[64][2];
  static bool resident_known[64][2] = {};
  int dev = 0;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  int resident_clusters = 0;
  if (dev >= 0 && dev < 64 && resident_known[dev][a.geo ? 1 : 0]) {
    resident_clusters = resident_cache[dev][a.geo ? 1 : 0];
  } else {
    if (cudaOccupancyMaxActiveClusters(&resident_clusters, kern, &cfg) != cudaSuccess) resident_clusters = 0;
    if (dev >= 0 && dev < 64) { resident_cache[dev][a.geo ? 1 : 0] = resident_clusters; resident_known[dev][a.geo ? 1 : 0] = true; }
  }
  if (resident_clusters < int(cfg.gridDim.x / 2)) {
    fprintf(stderr, "range_b200: only %d of %u CTA pairs can be resident; producer/consumer kernel not launched\n",
            resident_clusters, cfg.gridDim.x / 2);
    return cudaErrorCooperativeLaunchTooLarge;
  }
  cfg.numAttrs = coop ? 2 : 1;
  // Optional (RANGE_PC_PERSIST=1): a persisting-L2 access window over the consumers' window scratch (25 MB, read and
  // rewritten every 128 tiles; with evict_last hints alone part of it still goes to DRAM between flushes).  Measured on
  // the r2f build: no gain (apply 22.2 ms with the window, 21.8 ms without - inside the run-to-run spread), so it is
  // off by default; the device limit is raised once per device and the stream attribute restored after the launch.
  static const bool persist = getenv("RANGE_PC_PERSIST") && atoi(getenv("RANGE_PC_PERSIST")) == 1;
  bool window_set = false;
  if (persist && scratch) {
    static bool limit_set[64] = {};
    if (dev >= 0 && dev < 64 && !limit_set[dev]) {
      size_t cur = 0;
      const size_t want = apply_pc_scratch_bytes(sm_count) + (size_t(2) << 20);
      if (cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize) == cudaSuccess && cur < want)
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
      limit_set[dev] = true;
      (void)cudaGetLastError();
    }
    cudaStreamAttrValue av{};
    av.accessPolicyWindow.base_ptr = scratch;
    av.accessPolicyWindow.num_bytes = apply_pc_scratch_bytes(sm_count);
    av.accessPolicyWindow.hitRatio = 1.0f;
    av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    window_set = cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess;
    (void)cudaGetLastError();
  }
  auto clear_window = [&]() {
    if (!window_set) return;
    cudaStreamAttrValue av{};
    av.accessPolicyWindow.num_bytes = 0;
    cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &av);
    (void)cudaGetLastError();
  };
  e = cudaLaunchKernelEx(&cfg, kern, a.tmQ, a.tmK64, a.tmV128, tmP, a.db_xyz, reinterpret_cast<const float4*>(rowc),
                         a.N, a.M, a.a_sem, out, a.geo_mask, a.mask_words, reinterpret_cast<float*>(part),
                         reinterpret_cast<float4*>(scratch), reinterpret_cast<__half*>(ring),
                         reinterpret_cast<uint32_t*>(flags),
                         reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(flags) + pc_ring_flag_bytes(sm_count)), plan, dbg,
                         g_prof_buffer);
  clear_window();
  if (e != cudaSuccess) {
    fprintf(stderr, "range_b200: producer/consumer apply launch failed (%s); grid %u smem %d\n", cudaGetErrorString(e),
            cfg.gridDim.x, kDynamicSmem);
    return e;
  }
  if (plan.tail_split > 1) {    // sum the partial outputs of the split tail pairs into the last rows of out
    const int rows = a.N - plan.tail_row0;
    reduce_tail_kernel<<<unsigned((size_t(rows) * 256 + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const float4*>(part), plan.part_stride / 4, plan.tail_split, rows, plan.tail_row0, perm, out, out_ld,
        out_f64, plan.route);
    return cudaGetLastError();
  }
  return cudaSuccess;
}

// stats over pairs of query tiles; grid (2 * ceil(qtiles / 2), splits), cluster (2,1,1)
cudaError_t launch_stats_pc(const RetrievalArgs& a, float* part_sum, float* part_max, cudaStream_t stream) {
  auto kern = a.geo ? range_stats_pc_kernel<true> : range_stats_pc_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, StatSmem::dynamic_bytes);
  if (e != cudaSuccess) return e;
  const int qtiles = (a.N + kBlockQ - 1) / kBlockQ;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(unsigned((qtiles + 1) / 2 * 2), unsigned(a.stats_splits), 1);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = size_t(StatSmem::dynamic_bytes);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, a.tmQ, a.tmK64, a.db_xyz, a.q_xyz, a.N, a.M, a.stats_tiles_per_split, a.a_sem,
                            a.a_geo, part_sum, part_max, a.geo_mask, a.mask_words);
}

}  // namespace rangeb200
