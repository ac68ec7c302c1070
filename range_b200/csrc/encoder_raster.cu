// Spherical-harmonic features of a lat/lon raster (dense-grid embedding, BASELINE config 5; the reference builds such
// grids in range/evaluation/visualize_embeddings.py:29-45 and feeds them to SphericalHarmonics.forward point by point,
// positional_encoding/spherical_harmonics.py:27-42).
//
// Every feature of the analytic harmonics is a product of a latitude factor and a longitude factor,
//     Y(l, +-am) = [ (pref(l,am) * sin(theta)^am) * Q(l,am)(cos theta) ] * { cos(am phi) | sin(am phi) },
// and sh_rowmajor_kernel (encoder_tc.cu) evaluates it in exactly that association.  On an H x W raster the bracket
// takes only H distinct values per (l,am) and the trigonometric factor only W per am, so
//   raster_lat_kernel      one thread per DISTINCT latitude: the 820 Horner chains (5950 fp64 FMAs) -> leg[H][E]
//   raster_lon_kernel      one thread per (distinct longitude, am): sincos -> trig[W][L]
//   raster_combine_kernel  one warp per query: 1600 fp64 multiplies, hi/lo fp16 split, coalesced row-major stores in the
//                          encoder's feature layout (ShTable::fmap: column -> entry, |m|, cos / sin, or always zero)
// replace 5950 FMAs + 40 sincos per query.  Each operation is the one the per-point kernel performs, in the same order,
// so the features are bit-identical to it (tests/test_gpu_parity.py::test_raster_encoder_is_bit_identical).
// The combine kernel also materialises the queries' (lon, lat) rows for the normalise kernel's unit vectors.
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "range_kernels.h"
#include "split_f16.cuh"

namespace {

constexpr double kDeg2Rad = 0.017453292519943295769236907684886;

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

using rangeb200::split_f16;

// leg[i * E + e], entries e in the table's |m|-major order (for am: for l >= am)
__global__ void __launch_bounds__(64)
raster_lat_kernel(const double* __restrict__ lat, int H, int L, const double* __restrict__ pref,
                  const int* __restrict__ off, const double* __restrict__ coef, const int* __restrict__ par,
                  double* __restrict__ leg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H) return;
  const int E = L * (L + 1) / 2;
  const double theta = (lat[i] + 90.0) * kDeg2Rad;
  const double c = cos(theta);
  const double c2 = __dmul_rn(c, c);           // explicit roundings: what the per-point kernel's SASS does (DMUL, DADD)
  const double s = sqrt(__dsub_rn(1.0, c2));
  double spow = 1.0;
  double* out = leg + size_t(i) * E;
  int e = 0;
  for (int am = 0; am < L; ++am) {
    if (am > 0) spow *= s;
    for (int l = am; l < L; ++l, ++e) {
      int k = __ldg(off + e);
      const int kend = __ldg(off + e + 1);
      double acc = __ldg(coef + k);
      for (++k; k < kend; ++k) acc = fma(acc, c2, __ldg(coef + k));
      if (__ldg(par + e)) acc = __dmul_rn(acc, c);
      out[e] = am == 0 ? acc : __dmul_rn(__dmul_rn(__ldg(pref + e), spow), acc);
    }
  }
}

// trig[j * L + am] = (cos(am phi_j), sin(am phi_j)); am = 0 -> (1, 0)
__global__ void raster_lon_kernel(const double* __restrict__ lon, int W, int L, double2* __restrict__ trig) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= W * L) return;
  const int j = t / L, am = t % L;
  const double phi = (lon[j] + 180.0) * kDeg2Rad;
  double cm = 1.0, sm = 0.0;
  if (am > 0) sincos(double(am) * phi, &sm, &cm);
  trig[t] = make_double2(cm, sm);
}

// raster point p = i * W + j (lat-major, like coord_grid): its (latitude, longitude) indices and its coordinates, for the
// points p0 + (perm ? perm[n] : n) - the host never builds index or coordinate lists (range_raster_points)
__global__ void __launch_bounds__(256)
raster_points_kernel(long long p0, int N, const int* __restrict__ perm, int H, int W, const double* __restrict__ lat,
                     const double* __restrict__ lon, int2* __restrict__ ij, double2* __restrict__ lonlat) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const long long p = p0 + (perm ? perm[n] : n);
  const int i = int(p / W), j = int(p - (long long)i * W);
  if (ij) ij[n] = make_int2(i, j);
  if (lonlat) {
    const bool inside = p >= 0 && i < H;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    lonlat[n] = inside ? make_double2(lon[j], lat[i]) : make_double2(nan, nan);
  }
}

constexpr int kCombineWarps = 8;
__global__ void __launch_bounds__(kCombineWarps * 32)
raster_combine_kernel(const int2* __restrict__ ij, int N, int L, int F, int H, int W, const double* __restrict__ lat,
                      const double* __restrict__ lon, const double* __restrict__ leg, const double2* __restrict__ trig,
                      const int* __restrict__ fmap, __half* __restrict__ Yh, __half* __restrict__ Yl,
                      double* __restrict__ lonlat) {
  const int n = blockIdx.x * kCombineWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= N) return;
  const int2 idx = ij[n];                      // (latitude index, longitude index)
  const int E = L * (L + 1) / 2;               // F = columns per row in the encoder's feature layout (ShTable::K0, fmap)
  if (idx.x < 0 || idx.x >= H || idx.y < 0 || idx.y >= W) {     // not a raster point: a NaN row, visible in the result
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (lane == 0) reinterpret_cast<double2*>(lonlat)[n] = make_double2(nan, nan);
    const __half2 hn = __float2half2_rn(__int_as_float(0x7fc00000));
    for (int p = lane; 2 * p < F; p += 32) {
      reinterpret_cast<__half2*>(Yh + size_t(n) * F)[p] = hn;
      reinterpret_cast<__half2*>(Yl + size_t(n) * F)[p] = hn;
    }
    return;
  }
  const double* lrow = leg + size_t(idx.x) * E;
  const double2* trow = trig + size_t(idx.y) * L;
  if (lane == 0) reinterpret_cast<double2*>(lonlat)[n] = make_double2(lon[idx.y], lat[idx.x]);
  __half2* yh = reinterpret_cast<__half2*>(Yh + size_t(n) * F);
  __half2* yl = reinterpret_cast<__half2*>(Yl + size_t(n) * F);
  for (int p = lane; 2 * p < F; p += 32) {     // columns 2p, 2p + 1 (F % 64 == 0)
    const int2 m = reinterpret_cast<const int2*>(fmap)[p];
    double v[2];
    const int mm[2] = {m.x, m.y};
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int am = (mm[u] >> 16) & 0xff;
      if (mm[u] & rangeb200::kShZeroSlot) { v[u] = 0.0; continue; }
      const double a = __ldg(lrow + (mm[u] & 0xffff));
      if (am == 0) {
        v[u] = a;
      } else {
        const double2 t = __ldg(trow + am);
        v[u] = __dmul_rn(a, (mm[u] >> 24) ? t.y : t.x);
      }
    }
    __half h0, l0, h1, l1;
    split_f16(v[0], h0, l0);
    split_f16(v[1], h1, l1);
    yh[p] = __halves2half2(h0, h1);
    yl[p] = __halves2half2(l0, l1);
  }
}

}  // namespace

namespace rangeb200 {

size_t raster_tables_bytes(int L, int H, int W) {
  const size_t E = size_t(L) * (L + 1) / 2;
  return align_up(size_t(H) * E * 8, 256) + align_up(size_t(W) * L * 16, 256) + align_up(size_t(H) * 8, 256) +
         align_up(size_t(W) * 8, 256);
}

RasterTables raster_tables_layout(int L, int H, int W, void* buf) {
  const size_t E = size_t(L) * (L + 1) / 2;
  char* p = reinterpret_cast<char*>(buf);
  RasterTables t;
  t.H = H; t.W = W;
  t.leg = reinterpret_cast<double*>(p); p += align_up(size_t(H) * E * 8, 256);
  t.trig = p; p += align_up(size_t(W) * L * 16, 256);
  t.lat = reinterpret_cast<double*>(p); p += align_up(size_t(H) * 8, 256);
  t.lon = reinterpret_cast<double*>(p);
  return t;
}

cudaError_t launch_raster_tables(const ShTable& sh, const double* lat, const double* lon, const RasterTables& t,
                                 cudaStream_t s) {
  cudaError_t e = cudaMemcpyAsync(t.lat, lat, size_t(t.H) * 8, cudaMemcpyDeviceToDevice, s);
  if (e != cudaSuccess) return e;
  e = cudaMemcpyAsync(t.lon, lon, size_t(t.W) * 8, cudaMemcpyDeviceToDevice, s);
  if (e != cudaSuccess) return e;
  raster_lat_kernel<<<(t.H + 63) / 64, 64, 0, s>>>(t.lat, t.H, sh.L, sh.pref, sh.off, sh.coef, sh.par, t.leg);
  raster_lon_kernel<<<(t.W * sh.L + 255) / 256, 256, 0, s>>>(t.lon, t.W, sh.L, reinterpret_cast<double2*>(t.trig));
  return cudaGetLastError();
}

cudaError_t launch_raster_points(const RasterTables& t, long long p0, int N, const int32_t* perm, int32_t* ij,
                                 double* lonlat, cudaStream_t s) {
  if (N <= 0) return cudaSuccess;
  raster_points_kernel<<<(N + 255) / 256, 256, 0, s>>>(p0, N, perm, t.H, t.W, t.lat, t.lon, reinterpret_cast<int2*>(ij),
                                                      reinterpret_cast<double2*>(lonlat));
  return cudaGetLastError();
}

cudaError_t launch_raster_combine(const ShTable& sh, const RasterTables& t, const int32_t* ij, int N, void* Yh, void* Yl,
                                  double* lonlat, cudaStream_t s) {
  if (N <= 0) return cudaSuccess;
  raster_combine_kernel<<<(N + kCombineWarps - 1) / kCombineWarps, kCombineWarps * 32, 0, s>>>(
      reinterpret_cast<const int2*>(ij), N, sh.L, sh.K0, t.H, t.W, t.lat, t.lon, t.leg,
      reinterpret_cast<const double2*>(t.trig), sh.fmap, reinterpret_cast<__half*>(Yh), reinterpret_cast<__half*>(Yl), lonlat);
  return cudaGetLastError();
}

}  // namespace rangeb200
