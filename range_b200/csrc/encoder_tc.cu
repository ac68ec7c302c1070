// K1/K1b on the tensor cores: the SIREN layers as split-precision (3 x fp16) tcgen05 GEMMs (fp32-class accuracy).
//
// The reference runs the SIREN in fp64 (model_old.py:326-330), but its own spherical-harmonic input carries
// 1e-3 .. 5e-2 of fp64 rounding noise; an fp32-accurate SIREN moves the normalised
// embedding by ~1e-6 (measured, tests/test_gpu_parity.py::test_encoder_split_vs_fp64), far below that.
// Every operand x is split x = hi + lo into two fp16 values (11 + 11 significant bits, what a 3xTF32 split keeps too) and
//     A.B  ~=  A_hi.B_hi + A_hi.B_lo + A_lo.B_hi          (dropped term: 2^-22 relative)
// accumulated in fp32 in TMEM with kind::f16 MMAs - twice the TF32 rate and half the operand bytes (the TF32 version of
// this kernel was shared-memory-bandwidth bound: 156 B/clk of operand traffic).  Weights are pre-scaled by 2^10 so
// their low parts stay in fp16's normal range (|W| <= 1/1600 in the first layer); the epilogue scales back.  Bias add
// and the sine's argument reduction are done in fp64 in the epilogue.
//
//   sh_rounds_kernel     the analytic harmonics, lanes over Horner chains: the 820 (l,|m|) chains are sorted by length and
//                        cut into rounds of 32; a warp evaluates one round for four queries at a time (one coefficient
//                        load feeds four DFMAs), multiplies by the queries' cos / sin(|m| phi) and writes each round's 64
//                        columns (cos, sin per chain) as one coalesced 128-byte store per query: K0 = 64 * rounds columns,
//                        the first layer's columns are permuted (and zero-padded) to match.
//   sh_rowmajor_kernel   thread per query (closed-form harmonics, or tables too large for shared memory): features in
//                        PRODUCTION order (|m|-major) as hi/lo fp16, row-major [N][L*L], through a per-warp 32x32 smem
//                        transpose so global writes are coalesced although a thread owns a whole query.
//   split_weights_kernel W fp64 [H][K] -> hi, lo fp16 of 2^10 W, [H][K] (optionally with the column permutation)
//   siren_tc_kernel      CTA = 128 queries x 256 outputs, K-blocks of 64: TMA (SWIZZLE_128B) -> smem ring ->
//                        12 x tcgen05.mma kind::f16 (128x256x16) per block -> TMEM -> epilogue warps.
#include <cstdint>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "ptx.cuh"
#include "range_kernels.h"
#include "sh_closed_form.cuh"
#include "split_f16.cuh"

namespace {

constexpr double kDeg2Rad = 0.017453292519943295769236907684886;

constexpr float kWScale = 1024.f;        // weights are stored as 2^10 W (hi + lo fp16)

using rangeb200::split_f16;
using rangeb200::split_f16x2;

// ---------------------------------------------------------------------------------------------------
// SH features, row-major hi/lo fp32, production order:  for am: for l >= am: [cos] then [sin] (am > 0)
// ---------------------------------------------------------------------------------------------------
constexpr int kShWarps = 4;
__global__ void __launch_bounds__(kShWarps * 32)
sh_rowmajor_kernel(const double* __restrict__ lonlat, int N, int L, const double* __restrict__ pref,
                   const int* __restrict__ off, const double* __restrict__ coef, const int* __restrict__ par,
                   int closed_form, const double* __restrict__ norm, __half* __restrict__ Yh, __half* __restrict__ Yl) {
  __shared__ __half tile[kShWarps][2][32][34];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = (blockIdx.x * kShWarps + warp) * 32;
  if (n0 >= N) return;
  const int n = n0 + lane;
  const int F = L * L;
  double lon = 0.0, lat = 0.0;
  if (n < N) { lon = lonlat[2 * n]; lat = lonlat[2 * n + 1]; }
  const double phi = (lon + 180.0) * kDeg2Rad;
  const double theta = (lat + 90.0) * kDeg2Rad;
  const double c = cos(theta);
  const double c2 = c * c;
  const double s = sqrt(1.0 - c2);
  double spow = 1.0;
  int e = 0, cnt = 0, f0 = 0;
  auto emit = [&](double v) {
    __half hi, lo;
    split_f16(v, hi, lo);
    tile[warp][0][lane][cnt] = hi;
    tile[warp][1][lane][cnt] = lo;
    if (++cnt == 32) {
      __syncwarp();
      const int rows = min(32, N - n0);
      for (int r = 0; r < rows; ++r) {                       // row r of the tile = query n0 + r; lanes = features
        Yh[size_t(n0 + r) * F + f0 + lane] = tile[warp][0][r][lane];
        Yl[size_t(n0 + r) * F + f0 + lane] = tile[warp][1][r][lane];
      }
      __syncwarp();
      cnt = 0;
      f0 += 32;
    }
  };
  rangeb200::ClosedFormLegendre cf;
  cf.init(c);
  for (int am = 0; am < L; ++am) {
    double cm = 1.0, sm = 0.0;
    if (am > 0) {
      spow *= s;
      sincos(double(am) * phi, &sm, &cm);
    }
    if (closed_form) {                              // spherical_harmonics_closed_form.py:32-40, same production order
      cf.start_order(am);
      for (int l = am; l < L; ++l, ++e) {
        const double p = cf.next(l, am), nf = __ldg(norm + e);
        if (am == 0) {
          emit(nf * p);
        } else {
          emit((nf * cm) * p);
          emit((nf * sm) * p);
        }
      }
      continue;
    }
    // Horner chains of consecutive degrees l = am + 2k, am + 2k + 1 have the same length (k + 1 terms): several
    // independent chains per iteration hide the fp64 FMA and table-load latency (a thread owns a whole query)
    auto finish = [&](double acc, int ee) {
      if (__ldg(par + ee)) acc *= c;
      if (am == 0) {
        emit(acc);
      } else {
        const double leg = (__ldg(pref + ee) * spow) * acc;
        emit(leg * cm);
        emit(leg * sm);
      }
    };
    int l = am;
    // four chains per iteration (degrees l .. l + 3: lengths n, n, n + 1, n + 1), then the two-chain step for a remainder
    for (; l + 3 < L; l += 4, e += 4) {
      const int k0 = __ldg(off + e), k1 = __ldg(off + e + 1), k2 = __ldg(off + e + 2), k3 = __ldg(off + e + 3);
      const int n0 = k1 - k0, n1 = k2 - k1, n2 = k3 - k2, n3 = __ldg(off + e + 4) - k3;
      double a0 = __ldg(coef + k0), a1 = __ldg(coef + k1), a2 = __ldg(coef + k2), a3 = __ldg(coef + k3);
      const int nmin = min(min(n0, n1), min(n2, n3));
      for (int t = 1; t < nmin; ++t) {
        a0 = fma(a0, c2, __ldg(coef + k0 + t));
        a1 = fma(a1, c2, __ldg(coef + k1 + t));
        a2 = fma(a2, c2, __ldg(coef + k2 + t));
        a3 = fma(a3, c2, __ldg(coef + k3 + t));
      }
      for (int t = nmin; t < n0; ++t) a0 = fma(a0, c2, __ldg(coef + k0 + t));
      for (int t = nmin; t < n1; ++t) a1 = fma(a1, c2, __ldg(coef + k1 + t));
      for (int t = nmin; t < n2; ++t) a2 = fma(a2, c2, __ldg(coef + k2 + t));
      for (int t = nmin; t < n3; ++t) a3 = fma(a3, c2, __ldg(coef + k3 + t));
      finish(a0, e);
      finish(a1, e + 1);
      finish(a2, e + 2);
      finish(a3, e + 3);
    }
    for (; l + 1 < L; l += 2, e += 2) {
      int k0 = __ldg(off + e);
      int k1 = __ldg(off + e + 1);
      const int n0 = k1 - k0, n1 = __ldg(off + e + 2) - k1;
      double a0 = __ldg(coef + k0), a1 = __ldg(coef + k1);
      const int nmin = min(n0, n1);
      for (int t = 1; t < nmin; ++t) {
        a0 = fma(a0, c2, __ldg(coef + k0 + t));
        a1 = fma(a1, c2, __ldg(coef + k1 + t));
      }
      for (int t = nmin; t < n0; ++t) a0 = fma(a0, c2, __ldg(coef + k0 + t));
      for (int t = nmin; t < n1; ++t) a1 = fma(a1, c2, __ldg(coef + k1 + t));
      finish(a0, e);
      finish(a1, e + 1);
    }
    if (l < L) {
      int k = __ldg(off + e);
      const int kend = __ldg(off + e + 1);
      double acc = __ldg(coef + k);
      for (++k; k < kend; ++k) acc = fma(acc, c2, __ldg(coef + k));
      finish(acc, e);
      ++e;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// SH features by rounds (ShTable::rounds): lane = Horner chain, kRoundQ queries per warp pass
// ---------------------------------------------------------------------------------------------------
// Same operations per feature as sh_rowmajor_kernel / the raster kernels, in the same association:
//   acc = Horner in c^2 (a chain shorter than its round starts with fma(0, c2, 0) = 0 steps: exact), acc *= c for odd
//   parity, leg = (pref * s^|m|) * acc, feature = leg * cos(|m| phi) | leg * sin(|m| phi);  |m| = 0 has pref = 1 in the
//   round table (1 * 1 * acc * 1 = acc exactly; its sin column is 0 and the first layer's weight column there is zero).
// A thread-per-query kernel executes ~3 400 warp instructions per query (coefficient loads, loop overhead and the
// transposing stores around 5 950 DFMAs) and waits on the coefficient loads (L1 hit rate 57 %); here the round table sits
// in shared memory, a coefficient load feeds kRoundQ DFMAs and nothing is transposed: 1 670 warp instructions per query
// (676 -> 244 us per 100 000 queries; ncu: issue slots 68 % busy, conversion pipe 53 %, fp64 pipe 35 %).
#ifndef RANGE_SH_WARPS
#define RANGE_SH_WARPS 24
#endif
#ifndef RANGE_SH_Q
#define RANGE_SH_Q 4
#endif
// Measured, whole encoder per 100 000 queries (warps, queries per pass): (24, 4) 0.986 ms, (16, 8) 0.977, (32, 4) 0.979,
// (16, 4) 0.982, (20, 6) 0.968, (24, 2) 1.026 - flat once a coefficient load feeds four DFMAs.
constexpr int kRoundWarps = RANGE_SH_WARPS;      // warps per CTA (one CTA per SM: the round table is staged once per CTA)
constexpr int kRoundQ = RANGE_SH_Q;              // queries per warp pass

__host__ __device__ constexpr size_t round_scratch_doubles(int L) { return size_t(kRoundQ) * 4 + size_t(3) * kRoundQ * L; }

__global__ void __launch_bounds__(kRoundWarps * 32, 1)
sh_rounds_kernel(const double* __restrict__ lonlat, int N, int L, int R, const double* __restrict__ g_tab, int tab_doubles,
                 const int* __restrict__ g_meta, const int* __restrict__ g_roff, int K0, uint32_t* __restrict__ Yh,
                 uint32_t* __restrict__ Yl) {
  extern __shared__ __align__(16) uint8_t sh_smem[];
  double* tab = reinterpret_cast<double*>(sh_smem);
  int* meta = reinterpret_cast<int*>(tab + tab_doubles);
  int* roff = meta + R * 32;
  const size_t scratch0 = (size_t(tab_doubles) * 8 + size_t(R) * 128 + size_t(R + 1) * 4 + 15) & ~size_t(15);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* qv = reinterpret_cast<double*>(sh_smem + scratch0) + size_t(warp) * round_scratch_doubles(L);   // [Q][4]: c, c2, s, phi
  double* cmw = qv + kRoundQ * 4;          // [Q][L] cos(am phi)
  double* smw = cmw + kRoundQ * L;         // [Q][L] sin(am phi)
  double* spw = smw + kRoundQ * L;         // [Q][L] s^am (sequential products, like the per-query kernels)
  {
    const double2* src = reinterpret_cast<const double2*>(g_tab);
    double2* dst = reinterpret_cast<double2*>(tab);
    for (int i = threadIdx.x; i < tab_doubles / 2; i += blockDim.x) dst[i] = __ldg(src + i);
    for (int i = threadIdx.x; i < R * 32; i += blockDim.x) meta[i] = __ldg(g_meta + i);
    for (int i = threadIdx.x; i <= R; i += blockDim.x) roff[i] = __ldg(g_roff + i);
  }
  __syncthreads();
  const int nb = (N + kRoundQ - 1) / kRoundQ;
  const uint32_t row_words = uint32_t(K0) / 2;
  for (int b = blockIdx.x * kRoundWarps + warp; b < nb; b += gridDim.x * kRoundWarps) {
    __syncwarp();
    if (lane < kRoundQ) {
      const int n = b * kRoundQ + lane;
      double lon = 0.0, lat = 0.0;
      if (n < N) { lon = lonlat[2 * n]; lat = lonlat[2 * n + 1]; }
      const double phi = __dmul_rn(__dadd_rn(lon, 180.0), kDeg2Rad);
      const double theta = __dmul_rn(__dadd_rn(lat, 90.0), kDeg2Rad);
      const double c = cos(theta);
      const double c2 = __dmul_rn(c, c);
      const double s = sqrt(__dsub_rn(1.0, c2));
      qv[lane * 4 + 0] = c; qv[lane * 4 + 1] = c2; qv[lane * 4 + 2] = s; qv[lane * 4 + 3] = phi;
      double sp = 1.0;
      spw[lane * L] = 1.0;
      for (int am = 1; am < L; ++am) {
        sp = __dmul_rn(sp, s);
        spw[lane * L + am] = sp;
      }
    }
    __syncwarp();
    for (int it = lane; it < kRoundQ * L; it += 32) {
      const int q = it / L, am = it - q * L;
      double sv, cv;
      sincos(__dmul_rn(double(am), qv[q * 4 + 3]), &sv, &cv);      // am = 0: (0, 1) exactly
      cmw[it] = cv;
      smw[it] = sv;
    }
    __syncwarp();
    double c[kRoundQ], c2[kRoundQ];
    uint32_t row[kRoundQ];
#pragma unroll
    for (int q = 0; q < kRoundQ; ++q) {
      c[q] = qv[q * 4];
      c2[q] = qv[q * 4 + 1];
      row[q] = uint32_t(b * kRoundQ + q) * row_words + lane;
    }
    const int nvalid = min(kRoundQ, N - b * kRoundQ);
#pragma unroll 1
    for (int r = 0; r < R; ++r) {
      const int o = roff[r];
      const int nst = ((roff[r + 1] - o) >> 5) - 1;
      const double* t = tab + o + lane;
      const double pref = t[0];
      const int m = meta[r * 32 + lane];
      const int am = m & 0xff;
      double a[kRoundQ];
#pragma unroll
      for (int q = 0; q < kRoundQ; ++q) a[q] = 0.0;
#pragma unroll 2
      for (int st = 1; st <= nst; ++st) {
        const double cf = t[st * 32];
#pragma unroll
        for (int q = 0; q < kRoundQ; ++q) a[q] = fma(a[q], c2[q], cf);
      }
#pragma unroll
      for (int q = 0; q < kRoundQ; ++q) {
        double x = a[q];
        if (m & 0x100) x = __dmul_rn(x, c[q]);
        const double leg = __dmul_rn(__dmul_rn(pref, spw[q * L + am]), x);
        const double vc = __dmul_rn(leg, cmw[q * L + am]), vs = __dmul_rn(leg, smw[q * L + am]);
        uint32_t h, l;
        split_f16x2(vc, vs, h, l);
        if (q < nvalid) {
          Yh[row[q] + r * 32] = h;
          Yl[row[q] + r * 32] = l;
        }
      }
    }
  }
}

// W fp64 [H][K_in] -> hi/lo fp16 of 2^10 W, [H][K]; column f of the output reads column perm[f] of the input (zero when
// perm[f] < 0; perm may be null: identity)
__global__ void split_weights_kernel(const double* __restrict__ W, int H, int K_in, int K, const int* __restrict__ perm,
                                     __half* __restrict__ Wh, __half* __restrict__ Wl) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= size_t(H) * K) return;
  const int h = int(i / K), f = int(i % K);
  const int src = perm ? perm[f] : f;
  __half hi, lo;
  split_f16(src < 0 ? 0.0 : double(kWScale) * W[size_t(h) * K_in + src], hi, lo);
  Wh[i] = hi;
  Wl[i] = lo;
}

// ---------------------------------------------------------------------------------------------------
// out(128 x 256) = act(A(128 x K) . B(256 x K)^T + bias)
// ---------------------------------------------------------------------------------------------------
constexpr int kTcStages = 2;
constexpr int kTcStageBytes = 2 * 16384 + 2 * 32768;     // A_hi | A_lo | B_hi | B_lo
constexpr int kTcEpiWarps = 8;
constexpr int kTcThreads = (kTcEpiWarps + 2) * 32;
constexpr int kTcSmem = kTcStages * kTcStageBytes + 256 + 1024;

__global__ void __launch_bounds__(kTcThreads, 1)
siren_tc_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                const double* __restrict__ bias, int N, int K, int H, double act_w0, __half* __restrict__ out_hi,
                __half* __restrict__ out_lo, double* __restrict__ out_f64) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTcStages * kTcStageBytes);   // full[S] | empty[S] | done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kTcStages + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // blockIdx.x = column block: the CTAs that share a row tile's A operand are adjacent in launch order (second read from L2)
  const int row0 = blockIdx.y * 128, col0 = blockIdx.x * 256;
  const int KB = K / 64;                  // K blocks of 64 fp16 = one 128-byte swizzle row

  if (threadIdx.x == 0) {
    for (int i = 0; i < kTcStages; ++i) {
      ptx::mbar_init(&bars[i], 1);
      ptx::mbar_init(&bars[kTcStages + i], 1);
    }
    ptx::mbar_init(&bars[2 * kTcStages], 1);
    ptx::fence_mbar_init();
  }
  if (warp == kTcEpiWarps + 1) ptx::tmem_alloc<256>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t smem_u = __shfl_sync(0xffffffffu, ptx::smem_u32(smem), 0);
  const uint32_t bars_u = smem_u + kTcStages * kTcStageBytes;

  if (warp == kTcEpiWarps) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tmAh); ptx::prefetch_tmap(&tmAl); ptx::prefetch_tmap(&tmBh); ptx::prefetch_tmap(&tmBl);
      int idx = 0; uint32_t phase = 0;
      for (int kb = 0; kb < KB; ++kb) {
        ptx::mbar_wait(&bars[kTcStages + idx], phase ^ 1);
        uint8_t* st = smem + idx * kTcStageBytes;
        ptx::mbar_expect_tx(&bars[idx], kTcStageBytes);
        ptx::tma_load_2d(st, &tmAh, &bars[idx], kb * 64, row0);
        ptx::tma_load_2d(st + 16384, &tmAl, &bars[idx], kb * 64, row0);
        ptx::tma_load_2d(st + 32768, &tmBh, &bars[idx], kb * 64, col0);
        ptx::tma_load_2d(st + 65536, &tmBl, &bars[idx], kb * 64, col0);
        if (++idx == kTcStages) { idx = 0; phase ^= 1; }
      }
    }
  } else if (warp == kTcEpiWarps + 1) {
    constexpr uint32_t idesc = ptx::umma_idesc_f16(128, 256);
    int idx = 0; uint32_t phase = 0;
    for (int kb = 0; kb < KB; ++kb) {
      ptx::mbar_wait(&bars[idx], phase);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t st = smem_u + idx * kTcStageBytes;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t ah = ptx::umma_desc_kmajor_sw128(st + kk * 32);
          const uint64_t al = ptx::umma_desc_kmajor_sw128(st + 16384 + kk * 32);
          const uint64_t bh = ptx::umma_desc_kmajor_sw128(st + 32768 + kk * 32);
          const uint64_t bl = ptx::umma_desc_kmajor_sw128(st + 65536 + kk * 32);
          ptx::umma_f16_ss(tmem_base, al, bh, idesc, (kb | kk) != 0);   // small terms first
          ptx::umma_f16_ss(tmem_base, ah, bl, idesc, 1);
          ptx::umma_f16_ss(tmem_base, ah, bh, idesc, 1);
        }
        ptx::umma_commit_u32(bars_u + 8 * (kTcStages + idx));
        if (kb == KB - 1) ptx::umma_commit_u32(bars_u + 8 * (2 * kTcStages));
      }
      __syncwarp();
      if (++idx == kTcStages) { idx = 0; phase ^= 1; }
    }
  } else {
    // ===== epilogue: warp w -> TMEM lanes 32 (w%4).., columns 128 (w/4) .. +127 =====
    const int quarter = warp & 3, half = warp >> 2;
    const int n = row0 + quarter * 32 + lane;
    ptx::mbar_wait(&bars[2 * kTcStages], 0);
    ptx::tc_fence_after();
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      uint32_t v[32];
      const int c0 = half * 128 + cc * 32;
      ptx::tmem_ld32(tmem_base + (uint32_t(quarter * 32) << 16) + c0, v);
      ptx::tmem_ld_wait();
      if (n < N) {
        if (out_f64) {
          double* o = out_f64 + size_t(n) * H + col0 + c0;
#pragma unroll
          for (int i = 0; i < 32; i += 2)
            *reinterpret_cast<double2*>(o + i) =
                make_double2(double(__uint_as_float(v[i]) * (1.f / kWScale)) + __ldg(bias + col0 + c0 + i),
                             double(__uint_as_float(v[i + 1]) * (1.f / kWScale)) + __ldg(bias + col0 + c0 + i + 1));
        } else {
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float sv[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              // sin(w0 (acc + b)): argument reduced in fp64 (phases reach tens of radians), sine in fp32
              const double t = act_w0 * (double(__uint_as_float(v[i + u]) * (1.f / kWScale)) + __ldg(bias + col0 + c0 + i + u));
              const double r = fma(-rint(t * 0.15915494309189535), 6.283185307179586, t);
              sv[u] = sinf(float(r));
            }
            const __half2 h = __floats2half2_rn(sv[0], sv[1]);
            const __half2 l = __floats2half2_rn(sv[0] - __low2float(h), sv[1] - __high2float(h));
            hi[i / 2] = *reinterpret_cast<const uint32_t*>(&h);
            lo[i / 2] = *reinterpret_cast<const uint32_t*>(&l);
          }
          uint4* oh = reinterpret_cast<uint4*>(out_hi + size_t(n) * H + col0 + c0);
          uint4* ol = reinterpret_cast<uint4*>(out_lo + size_t(n) * H + col0 + c0);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            oh[i] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
            ol[i] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kTcEpiWarps + 1) ptx::tmem_dealloc<256>(tmem_base);
}


// ---------------------------------------------------------------------------------------------------
// The same GEMM + epilogue on CTA pairs, persistent: out(256 x 256 per pair) = act(A . B^T + bias)
// ---------------------------------------------------------------------------------------------------
// siren_tc_kernel's CTAs each re-read their 256 weight rows for every 128 queries: 98 KB of operands per 12 MMAs, 4 GB
// of L2 -> shared-memory traffic for the first layer (8.5 TB/s: the L2 is the bound, not the tensor core), and prologue,
// pipeline fill and epilogue are exposed once per tile (layers 1-2 have 8 k-blocks per tile).  Here a CTA pair
// (cta_group::2, M = 256) shares the weight tile - each CTA loads its own 128 query rows and HALF of the 256 weight rows,
// 64 KB per 12 MMAs - and a cluster walks over its tiles with the accumulator double-buffered in TMEM (2 x 256 columns):
// the epilogue warps of both CTAs drain tile i while the pair's tensor cores run tile i + 1, the TMA warp never stops.
constexpr int kSpStages = 3;
constexpr int kSpStageBytes = 4 * 16384;                 // A_hi | A_lo | B_hi | B_lo, 128 rows x 128 B each
constexpr int kSpEpiWarps = 8;
constexpr int kSpWarpTma = kSpEpiWarps, kSpWarpMma = kSpEpiWarps + 1;
constexpr int kSpThreads = (kSpEpiWarps + 2) * 32;
constexpr int kSpBiasBytes = kSpEpiWarps * 1024;         // per epilogue warp: its 128 columns' bias terms (fp32 w0 b, or fp64 b)
constexpr int kSpSmem = kSpStages * kSpStageBytes + 256 + kSpBiasBytes + 1024;

__global__ void __launch_bounds__(kSpThreads, 1)
siren_pair_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                  const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                  const double* __restrict__ bias, int N, int K, int H, double act_w0, __half* __restrict__ out_hi,
                  __half* __restrict__ out_lo, double* __restrict__ out_f64) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // full[S] | empty[S] | acc_full[2] | acc_empty[2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSpStages * kSpStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kSpStages;
  uint64_t* acc_full = bars + 2 * kSpStages;
  uint64_t* acc_empty = bars + 2 * kSpStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kSpStages + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int KB = K / 64, CB = H / 256;
  const int tiles = ((N + 255) / 256) * CB;              // tile t: row pair t / CB, column block t % CB (neighbours share A)

  if (threadIdx.x == 0) {
    for (int i = 0; i < kSpStages; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&acc_full[i], 1);
      ptx::mbar_init(&acc_empty[i], 2 * kSpEpiWarps);    // the epilogue warps of both CTAs
    }
    ptx::fence_mbar_init();
  }
  if (warp == kSpWarpMma) ptx::tmem_alloc_2sm<512>(tmem_slot);
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t smem_u = __shfl_sync(0xffffffffu, ptx::smem_u32(smem), 0);
  const uint32_t bars_u = smem_u + kSpStages * kSpStageBytes;

  if (warp == kSpWarpTma) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tmAh); ptx::prefetch_tmap(&tmAl); ptx::prefetch_tmap(&tmBh); ptx::prefetch_tmap(&tmBl);
      int idx = 0; uint32_t phase = 0;
      for (int t = cluster; t < tiles; t += n_clusters) {
        const int row0 = (t / CB) * 256 + int(rank) * 128, col0 = (t % CB) * 256 + int(rank) * 128;
        for (int kb = 0; kb < KB; ++kb) {
          ptx::mbar_wait(&empty[idx], phase ^ 1);
          uint8_t* st = smem + idx * kSpStageBytes;
          if (leader) ptx::mbar_expect_tx(&full[idx], 2 * kSpStageBytes);     // both CTAs' bytes are credited to the leader
          ptx::tma_load_2d_2sm(st, &tmAh, &full[idx], kb * 64, row0);
          ptx::tma_load_2d_2sm(st + 16384, &tmAl, &full[idx], kb * 64, row0);
          ptx::tma_load_2d_2sm(st + 32768, &tmBh, &full[idx], kb * 64, col0);
          ptx::tma_load_2d_2sm(st + 49152, &tmBl, &full[idx], kb * 64, col0);
          if (++idx == kSpStages) { idx = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == kSpWarpMma) {
    if (leader) {
      constexpr uint32_t idesc = ptx::umma_idesc_f16(256, 256);
      int idx = 0; uint32_t phase = 0;
      int i = 0;
      for (int t = cluster; t < tiles; t += n_clusters, ++i) {
        const int buf = i & 1;
        ptx::mbar_wait_cluster(&acc_empty[buf], ((i >> 1) & 1) ^ 1);          // both CTAs drained this accumulator
        ptx::tc_fence_after();
        const uint32_t d = tmem_base + buf * 256;
        for (int kb = 0; kb < KB; ++kb) {
          ptx::mbar_wait(&full[idx], phase);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t st = smem_u + idx * kSpStageBytes;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t ah = ptx::umma_desc_kmajor_sw128(st + kk * 32);
              const uint64_t al = ptx::umma_desc_kmajor_sw128(st + 16384 + kk * 32);
              const uint64_t bh = ptx::umma_desc_kmajor_sw128(st + 32768 + kk * 32);
              const uint64_t bl = ptx::umma_desc_kmajor_sw128(st + 49152 + kk * 32);
              ptx::umma_f16_ss_2sm(d, al, bh, idesc, (kb | kk) != 0);          // small terms first
              ptx::umma_f16_ss_2sm(d, ah, bl, idesc, 1);
              ptx::umma_f16_ss_2sm(d, ah, bh, idesc, 1);
            }
            ptx::umma_commit_2sm_u32(bars_u + 8 * (kSpStages + idx));
            if (kb == KB - 1) ptx::umma_commit_2sm_u32(bars_u + 8 * (2 * kSpStages + buf));
          }
          __syncwarp();
          if (++idx == kSpStages) { idx = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===== epilogue: warp w -> TMEM lanes 32 (w%4).., columns 128 (w/4) .. +127 of the current accumulator =====
    // Hidden layers: sin(w0 (acc + b)) entirely in fp32 - x = fma(acc, w0 / 2^10, w0 b) (half an ulp of |x| <= 64: 2e-6,
    // the size of the accumulator's own rounding times w0), two-constant Cody-Waite reduction to [-pi, pi], MUFU.SIN.
    // (siren_tc_kernel reduces in fp64: 5 fp64 + 3 conversion instructions per output, which left the 8 epilogue warps
    // waiting on the fp64 / conversion pipes for longer than a K = 512 tile's MMAs take.)
    const int quarter = warp & 3, half = warp >> 2;
    const uint32_t acc_empty_leader = ptx::mapa(ptx::smem_u32(acc_empty), 0);
    uint8_t* bias_w = smem + kSpStages * kSpStageBytes + 256 + warp * 1024;
    float* bias_f = reinterpret_cast<float*>(bias_w);
    double* bias_d = reinterpret_cast<double*>(bias_w);
    const float xscale = float(act_w0) * (1.f / kWScale);
    const uint32_t bias_u = smem_u + kSpStages * kSpStageBytes + 256 + warp * 1024;     // explicit LDS (generic loads -> LD.E)
    int i = 0;
    for (int t = cluster; t < tiles; t += n_clusters, ++i) {
      const int buf = i & 1;
      const int n = (t / CB) * 256 + int(rank) * 128 + quarter * 32 + lane;
      const int col0 = (t % CB) * 256;
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double b = __ldg(bias + col0 + half * 128 + j * 32 + lane);
        if (out_f64) bias_d[j * 32 + lane] = b;
        else bias_f[j * 32 + lane] = float(act_w0 * b);
      }
      __syncwarp();
      ptx::mbar_wait(&acc_full[buf], (i >> 1) & 1);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t v[32];
        const int c0 = half * 128 + cc * 32;
        ptx::tmem_ld32(tmem_base + (uint32_t(quarter * 32) << 16) + buf * 256 + c0, v);
        ptx::tmem_ld_wait();
        if (cc == 3) {                   // this warp's part of the accumulator is in registers: hand the buffer back
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) ptx::mbar_arrive(&acc_empty[buf]);
            else ptx::mbar_arrive_cluster_relaxed(acc_empty_leader + 8 * buf);
          }
        }
        if (n < N) {
          if (out_f64) {
            double* o = out_f64 + size_t(n) * H + col0 + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float4 braw = ptx::lds_f4(bias_u + (cc * 32 + j) * 8);
              const double2 b = make_double2(__hiloint2double(__float_as_int(braw.y), __float_as_int(braw.x)),
                                             __hiloint2double(__float_as_int(braw.w), __float_as_int(braw.z)));
              *reinterpret_cast<double2*>(o + j) = make_double2(double(__uint_as_float(v[j]) * (1.f / kWScale)) + b.x,
                                                                double(__uint_as_float(v[j + 1]) * (1.f / kWScale)) + b.y);
            }
          } else {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = ptx::lds_f4(bias_u + (cc * 32 + j) * 4);
              const float bw[4] = {b4.x, b4.y, b4.z, b4.w};
              float sv[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float x = fmaf(__uint_as_float(v[j + u]), xscale, bw[u]);
                // k = rint(x / 2 pi) by the 1.5 * 2^23 trick (|x / 2 pi| < 2^22), r = x - k 2 pi with 2 pi = hi + lo
                const float k = __fadd_rn(__fadd_rn(__fmul_rn(x, 0.15915494309189535f), 12582912.f), -12582912.f);
                const float r = fmaf(k, 1.7484555e-7f, fmaf(k, -6.2831855f, x));
                sv[u] = __sinf(r);
              }
              const __half2 h = __floats2half2_rn(sv[0], sv[1]);
              const __half2 l = __floats2half2_rn(sv[0] - __low2float(h), sv[1] - __high2float(h));
              hi[j / 2] = *reinterpret_cast<const uint32_t*>(&h);
              lo[j / 2] = *reinterpret_cast<const uint32_t*>(&l);
              const __half2 h2 = __floats2half2_rn(sv[2], sv[3]);
              const __half2 l2 = __floats2half2_rn(sv[2] - __low2float(h2), sv[3] - __high2float(h2));
              hi[j / 2 + 1] = *reinterpret_cast<const uint32_t*>(&h2);
              lo[j / 2 + 1] = *reinterpret_cast<const uint32_t*>(&l2);
            }
            uint4* oh = reinterpret_cast<uint4*>(out_hi + size_t(n) * H + col0 + c0);
            uint4* ol = reinterpret_cast<uint4*>(out_lo + size_t(n) * H + col0 + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              oh[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
              ol[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
            }
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (warp == kSpWarpMma) ptx::tmem_dealloc_2sm<512>(tmem_base);
}

}  // namespace

namespace rangeb200 {

size_t sh_rounds_smem_bytes(int L, int rounds, int rtab_doubles) {
  if (rounds <= 0) return 0;
  const size_t table = (size_t(rtab_doubles) * 8 + size_t(rounds) * 128 + size_t(rounds + 1) * 4 + 15) & ~size_t(15);
  return table + size_t(kRoundWarps) * round_scratch_doubles(L) * 8;
}

cudaError_t launch_sh_rowmajor(const ShTable& t, const double* lonlat, int N, void* Yh, void* Yl, int sm_count,
                               cudaStream_t s) {
  if (N <= 0) return cudaSuccess;
  if (t.rounds > 0) {
    const size_t smem = sh_rounds_smem_bytes(t.L, t.rounds, t.rtab_doubles);
    cudaError_t e = cudaFuncSetAttribute(sh_rounds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    const int nb = (N + kRoundQ - 1) / kRoundQ;
    const int grid = min(sm_count, (nb + kRoundWarps - 1) / kRoundWarps);
    sh_rounds_kernel<<<grid, kRoundWarps * 32, smem, s>>>(lonlat, N, t.L, t.rounds, t.rtab, t.rtab_doubles, t.rmeta, t.rroff,
                                                         t.K0, reinterpret_cast<uint32_t*>(Yh), reinterpret_cast<uint32_t*>(Yl));
    return cudaGetLastError();
  }
  const int per_block = kShWarps * 32;
  sh_rowmajor_kernel<<<(N + per_block - 1) / per_block, per_block, 0, s>>>(lonlat, N, t.L, t.pref, t.off, t.coef,
                                                                           t.par, t.closed_form, t.norm, reinterpret_cast<__half*>(Yh),
                                                                           reinterpret_cast<__half*>(Yl));
  return cudaGetLastError();
}

cudaError_t launch_split_weights(const double* W, int H, int K_in, int K, const int* perm, void* Wh, void* Wl,
                                 cudaStream_t s) {
  const size_t total = size_t(H) * K;
  split_weights_kernel<<<unsigned((total + 255) / 256), 256, 0, s>>>(W, H, K_in, K, perm, reinterpret_cast<__half*>(Wh),
                                                                     reinterpret_cast<__half*>(Wl));
  return cudaGetLastError();
}

cudaError_t launch_siren_pair(const CUtensorMap& tmAh, const CUtensorMap& tmAl, const CUtensorMap& tmBh,
                              const CUtensorMap& tmBl, const double* bias, int N, int K, int H, double act_w0,
                              void* out_hi, void* out_lo, double* out_f64, int sm_count, cudaStream_t s) {
  if (N <= 0) return cudaSuccess;
  if (K % 64 || H % 256) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(siren_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSpSmem);
  if (e != cudaSuccess) return e;
  const int tiles = ((N + 255) / 256) * (H / 256);
  const int clusters = tiles < sm_count / 2 ? tiles : sm_count / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kSpThreads);
  cfg.dynamicSmemBytes = kSpSmem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, siren_pair_kernel, tmAh, tmAl, tmBh, tmBl, bias, N, K, H, act_w0,
                            reinterpret_cast<__half*>(out_hi), reinterpret_cast<__half*>(out_lo), out_f64);
}

cudaError_t launch_siren_tc(const CUtensorMap& tmAh, const CUtensorMap& tmAl, const CUtensorMap& tmBh,
                            const CUtensorMap& tmBl, const double* bias, int N, int K, int H, double act_w0,
                            void* out_hi, void* out_lo, double* out_f64, cudaStream_t s) {
  if (N <= 0) return cudaSuccess;
  if (K % 64 || H % 256) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(siren_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem);
  if (e != cudaSuccess) return e;
  dim3 grid(H / 256, (N + 127) / 128);
  siren_tc_kernel<<<grid, kTcThreads, kTcSmem, s>>>(tmAh, tmAl, tmBh, tmBl, bias, N, K, H, act_w0,
                                                    reinterpret_cast<__half*>(out_hi), reinterpret_cast<__half*>(out_lo), out_f64);
  return cudaGetLastError();
}

}  // namespace rangeb200
