// K1  spherical harmonics, K1b SIREN layers (fp64 tensor-core GEMM + fused bias/sin), K3 normalise/concat.
//
// Reference: range/location_models/satclip/positional_encoding/spherical_harmonics.py:27-42 (+ the
// generated closed forms), location_encoder.py:98-151, range/range.py:212,222-229,240.
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "range_kernels.h"
#include "sh_closed_form.cuh"

namespace {

constexpr double kDeg2Rad = 0.017453292519943295769236907684886;   // torch.deg2rad's constant

// ---------------------------------------------------------------------------------------------------
// K1: one thread per query.  Feature l*l+l+-m = pref * (1-c^2)^(|m|/2) * Q_l|m|(c) * {cos,sin}(|m| phi) with
// Q evaluated by Horner in c^2 on the reference generator's 15-digit coefficients (range_b200/sh_table.py).
// Table reads are warp-uniform (broadcast); writes are coalesced across queries (feature-major output).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
sh_kernel(const double* __restrict__ lonlat, int N, int L, const double* __restrict__ pref,
          const int* __restrict__ off, const double* __restrict__ coef, const int* __restrict__ par,
          int closed_form, const double* __restrict__ norm, double* __restrict__ Yt, size_t ld) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const double lon = lonlat[2 * n], lat = lonlat[2 * n + 1];
  const double phi = (lon + 180.0) * kDeg2Rad;      // spherical_harmonics.py:31
  const double theta = (lat + 90.0) * kDeg2Rad;     // :32
  const double c = cos(theta);
  const double c2 = c * c;
  const double s2 = 1.0 - c2;
  const double s = sqrt(s2);
  double spow = 1.0;                                // (1 - c^2)^(am/2)
  int e = 0;
  rangeb200::ClosedFormLegendre cf;
  cf.init(c);
  for (int am = 0; am < L; ++am) {
    double cm = 1.0, sm = 0.0;
    if (am > 0) {
      spow *= s;
      sincos(double(am) * phi, &sm, &cm);
    }
    if (closed_form) {                              // spherical_harmonics_closed_form.py:32-40
      cf.start_order(am);
      for (int l = am; l < L; ++l, ++e) {
        const double p = cf.next(l, am), nf = __ldg(norm + e);
        const size_t f0 = size_t(l) * l + l;
        if (am == 0) {
          Yt[f0 * ld + n] = nf * p;
        } else {
          Yt[(f0 + am) * ld + n] = (nf * cm) * p;
          Yt[(f0 - am) * ld + n] = (nf * sm) * p;
        }
      }
      continue;
    }
    for (int l = am; l < L; ++l, ++e) {
      int k = __ldg(off + e);
      const int kend = __ldg(off + e + 1);
      double acc = __ldg(coef + k);
      for (++k; k < kend; ++k) acc = fma(acc, c2, __ldg(coef + k));
      if (__ldg(par + e)) acc *= c;
      const size_t f0 = size_t(l) * l + l;
      if (am == 0) {
        Yt[f0 * ld + n] = acc;
      } else {
        const double leg = (__ldg(pref + e) * spow) * acc;
        Yt[(f0 + am) * ld + n] = leg * cm;
        Yt[(f0 - am) * ld + n] = leg * sm;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// K1b: OUT (H x N) = act(W (H x K) . Xt (K x N) + b)   in fp64 on the DMMA pipe (mma.sync m8n8k4.f64).
// CTA tile 64 (H) x 128 (queries) x 16, 8 warps of 32 x 32, 3-stage cp.async pipeline.
// ---------------------------------------------------------------------------------------------------
constexpr int BM = 64, BN = 128, BK = 16, LDA = BK + 4, LDB = BN + 4, GEMM_STAGES = 3;
constexpr int GEMM_SMEM = GEMM_STAGES * (BM * LDA + BK * LDB) * 8;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(uint32_t(__cvta_generic_to_shared(smem))), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ void dmma(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256)
siren_layer_kernel(const double* __restrict__ W, const double* __restrict__ bias, const double* __restrict__ Xt,
                   size_t ldx, int H, int K, int N, double act_w0, double* __restrict__ out, size_t ldo,
                   int out_rowmajor) {
  extern __shared__ __align__(16) uint8_t gemm_smem[];
  double* As = reinterpret_cast<double*>(gemm_smem);
  double* Bs = As + GEMM_STAGES * BM * LDA;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int wm = (warp & 1) * 32, wn = (warp >> 1) * 32;
  const int g = lane >> 2, t = lane & 3;

  auto load_stage = [&](int stage, int k0) {
    // A: 64 rows x 16 doubles = 128 x 16-B chunks ... 2 doubles per chunk -> 64*8 = 512 chunks
    double* a = As + stage * BM * LDA;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int ch = tid + i * 256;          // 0..511
      const int r = ch >> 3, cc = (ch & 7) * 2;
      cp_async16(a + r * LDA + cc, W + size_t(h0 + r) * K + k0 + cc);
    }
    // B: 16 rows x 128 doubles = 16 * 64 = 1024 chunks
    double* b = Bs + stage * BK * LDB;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ch = tid + i * 256;
      const int r = ch >> 6, cc = (ch & 63) * 2;
      cp_async16(b + r * LDB + cc, Xt + size_t(k0 + r) * ldx + n0 + cc);
    }
  };

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int KT = K / BK;
  for (int s = 0; s < GEMM_STAGES - 1; ++s) {
    if (s < KT) load_stage(s, s * BK);
    cp_async_commit();
  }
  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<GEMM_STAGES - 2>();
    __syncthreads();
    const int nxt = kt + GEMM_STAGES - 1;
    if (nxt < KT) load_stage(nxt % GEMM_STAGES, nxt * BK);
    cp_async_commit();
    const double* a = As + (kt % GEMM_STAGES) * BM * LDA;
    const double* b = Bs + (kt % GEMM_STAGES) * BK * LDB;
#pragma unroll
    for (int k4 = 0; k4 < BK; k4 += 4) {
      double af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = a[(wm + i * 8 + g) * LDA + k4 + t];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = b[(k4 + t) * LDB + wn + j * 8 + g];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[i][j], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();

  // epilogue: bias + sin, C fragment: row = g (h), cols = 2 t + {0,1} (queries)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int h = h0 + wm + i * 8 + g;
    const double bv = bias[h];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + wn + j * 8 + 2 * t;
      double v0 = acc[i][j][0] + bv, v1 = acc[i][j][1] + bv;
      if (act_w0 != 0.0) {
        v0 = sin(act_w0 * v0);
        v1 = sin(act_w0 * v1);
      }
      if (out_rowmajor) {
        if (n < N) out[size_t(n) * ldo + h] = v0;
        if (n + 1 < N) out[size_t(n + 1) * ldo + h] = v1;
      } else {
        // feature-major buffers are padded to a multiple of BN columns
        *reinterpret_cast<double2*>(out + size_t(h) * ldo + n) = make_double2(v0, v1);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// K3: warp per query: L2-normalise (fp64), emit q64 / q16 and the unit vector of the query location
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
normalize_kernel(const double* __restrict__ e, const double* __restrict__ lonlat, int N, int D,
                 double* __restrict__ q64, size_t ldq, __half* __restrict__ q16, float4* __restrict__ qxyz) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const double* row = e + size_t(n) * D;
  double ss = 0.0;
  for (int i = lane; i < D; i += 32) {
    const double v = row[i];
    ss = fma(v, v, ss);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const double nrm = sqrt(ss);                                        // range.py:212
  for (int i = lane; i < D; i += 32) {
    const double v = row[i] / nrm;
    q64[size_t(n) * ldq + i] = v;
    q16[size_t(n) * D + i] = __double2half(v);
  }
  if (lane == 0) {
    // range.py:225-229 + utils/utils.py:11-16: (deg * pi) / 180 in fp64, cast to fp32 at :231
    const double lon = lonlat[2 * n] * CUDART_PI / 180.0, lat = lonlat[2 * n + 1] * CUDART_PI / 180.0;
    double sl, cl, sp, cp;
    sincos(lon, &sl, &cl);
    sincos(lat, &sp, &cp);
    qxyz[n] = make_float4(float(cp * cl), float(cp * sl), float(sp), 0.f);
  }
}

__global__ void __launch_bounds__(256)
concat_q_kernel(const double* __restrict__ q64, int N, int DQ, const int* __restrict__ perm, void* out, int ld, int col0,
                int dtype) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= size_t(N) * DQ) return;
  const size_t n = i / DQ;
  const int c = int(i - n * DQ);
  const size_t o = size_t(perm ? perm[n] : n) * ld + col0 + c;
  if (dtype == 0) reinterpret_cast<double*>(out)[o] = q64[i];
  else reinterpret_cast<float*>(out)[o] = float(q64[i]);
}

}  // namespace

namespace rangeb200 {

cudaError_t launch_sh(const ShTable& t, const double* lonlat, int N, double* Yt, size_t ld, cudaStream_t s) {
  if (N <= 0) return cudaSuccess;
  sh_kernel<<<(N + 127) / 128, 128, 0, s>>>(lonlat, N, t.L, t.pref, t.off, t.coef, t.par, t.closed_form, t.norm, Yt, ld);
  return cudaGetLastError();
}

cudaError_t launch_siren_layer(const double* W, const double* b, const double* Xt, size_t ldx, int H, int K, int N,
                               double act_w0, double* out, size_t ldo, int out_rowmajor, cudaStream_t s) {
  if (N <= 0) return cudaSuccess;
  if (H % BM || K % BK) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(siren_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
  if (e != cudaSuccess) return e;
  dim3 grid((N + BN - 1) / BN, H / BM);
  siren_layer_kernel<<<grid, 256, GEMM_SMEM, s>>>(W, b, Xt, ldx, H, K, N, act_w0, out, ldo, out_rowmajor);
  return cudaGetLastError();
}

cudaError_t launch_normalize(const double* e, const double* lonlat, int N, int D, double* q64, size_t ldq,
                             void* q16, float* qxyz, cudaStream_t s) {
  if (N <= 0) return cudaSuccess;
  normalize_kernel<<<(N + 7) / 8, 256, 0, s>>>(e, lonlat, N, D, q64, ldq, reinterpret_cast<__half*>(q16),
                                               reinterpret_cast<float4*>(qxyz));
  return cudaGetLastError();
}

cudaError_t launch_concat_q(const double* q64, int N, int DQ, const int* perm, void* out, int ld, int col0, int dtype,
                            cudaStream_t s) {
  if (N <= 0) return cudaSuccess;
  const size_t total = size_t(N) * DQ;
  concat_q_kernel<<<unsigned((total + 255) / 256), 256, 0, s>>>(q64, N, DQ, perm, out, ld, col0, dtype);
  return cudaGetLastError();
}

}  // namespace rangeb200
