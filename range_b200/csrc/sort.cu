// Spatial batching of the queries (no counterpart in the reference, which treats rows independently:
// range/range.py:213-240 are all row-wise).  The geographic softmax of RANGE+ is local - exp(40 (cos d - 1)) -
// so when the 128 queries of a tile are close to each other whole database tiles can skip the geo term
// (retrieval.cu: geo_mask_kernel).  Queries arrive in arbitrary order; this file computes a permutation that
// groups them by a cube-map Hilbert cell (deterministic counting sort), the encoder and retrieval run on the
// permuted rows and range_concat scatters the results back to the caller's order.
//
//   cell_hist_kernel    key[i] = face * 4^k + hilbert(u, v);  hist[key]++
//   cell_scan_kernel    start[c] = exclusive prefix sum (single CTA; <= 6 * 4^7 + 1 cells)
//   cell_scatter_kernel tmp[cursor[key[i]]++] = i                      (order inside a cell: arbitrary)
//   cell_rank_kernel    orders each cell's members by original index  (-> deterministic permutation)
#include <cstdint>
#include <cuda_runtime.h>

#include "range_kernels.h"

namespace {

constexpr int kMaxBits = 7;            // up to 6 * 4^7 = 98 304 cells
constexpr int kRankLimit = 64;         // cells larger than this keep the scatter order (still a valid permutation)

// position of grid cell (x, y) of a 2^bits x 2^bits grid along the Hilbert curve (the classic xy2d walk)
__device__ __forceinline__ uint32_t hilbert_index(uint32_t x, uint32_t y, int bits) {
  const uint32_t n = 1u << bits;
  uint32_t d = 0;
  for (uint32_t s = n >> 1; s > 0; s >>= 1) {
    const uint32_t rx = (x & s) ? 1u : 0u, ry = (y & s) ? 1u : 0u;
    d += s * s * ((3u * rx) ^ ry);
    if (ry == 0) {
      if (rx == 1) { x = n - 1 - x; y = n - 1 - y; }
      const uint32_t t = x; x = y; y = t;
    }
  }
  return d;
}

__device__ __forceinline__ uint32_t cell_key(double lon_deg, double lat_deg, int bits) {
  const float kRad = 0.017453292519943295f;
  float slon, clon, slat, clat;
  sincosf(float(lon_deg) * kRad, &slon, &clon);
  sincosf(float(lat_deg) * kRad, &slat, &clat);
  const float x = clat * clon, y = clat * slon, z = slat;
  const float ax = fabsf(x), ay = fabsf(y), az = fabsf(z);
  int face;
  float u, v, m;
  if (ax >= ay && ax >= az) { face = x >= 0.f ? 0 : 1; m = ax; u = y; v = z; }
  else if (ay >= az)        { face = y >= 0.f ? 2 : 3; m = ay; u = x; v = z; }
  else                      { face = z >= 0.f ? 4 : 5; m = az; u = x; v = y; }
  m = fmaxf(m, 1e-20f);
  // equal-angle cube map: cells of roughly equal area
  const float fu = atanf(u / m) * 1.2732395447351628f, fv = atanf(v / m) * 1.2732395447351628f;   // [-1, 1]
  const int g = 1 << bits;
  int iu = int((fu + 1.f) * 0.5f * float(g)), iv = int((fv + 1.f) * 0.5f * float(g));
  iu = min(max(iu, 0), g - 1);
  iv = min(max(iv, 0), g - 1);
  return (uint32_t(face) << (2 * bits)) | hilbert_index(uint32_t(iu), uint32_t(iv), bits);
}

__global__ void __launch_bounds__(256)
cell_hist_kernel(const double2* __restrict__ lonlat, int N, int bits, uint32_t* __restrict__ key,
                 uint32_t* __restrict__ hist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double2 p = lonlat[i];
  const uint32_t k = cell_key(p.x, p.y, bits);
  key[i] = k;
  atomicAdd(&hist[k], 1u);
}

// exclusive scan of hist[0..n) -> start[0..n], start[n] = total; cursor = copy of start.  One CTA of 1024.
__global__ void __launch_bounds__(1024)
cell_scan_kernel(const uint32_t* __restrict__ hist, int n, uint32_t* __restrict__ start, uint32_t* __restrict__ cursor) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const uint32_t v = i < n ? hist[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      warp_sums[lane] = w;
    }
    __syncthreads();
    const uint32_t excl = carry + (warp ? warp_sums[warp - 1] : 0u) + x - v;
    if (i < n) { start[i] = excl; cursor[i] = excl; }
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) start[n] = carry;
}

__global__ void __launch_bounds__(256)
cell_scatter_kernel(const uint32_t* __restrict__ key, int N, uint32_t* __restrict__ cursor, uint32_t* __restrict__ tmp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  tmp[atomicAdd(&cursor[key[i]], 1u)] = uint32_t(i);
}

__global__ void __launch_bounds__(256)
cell_rank_kernel(const uint32_t* __restrict__ key, const uint32_t* __restrict__ tmp, const uint32_t* __restrict__ start,
                 const double2* __restrict__ lonlat, int N, int32_t* __restrict__ perm, double2* __restrict__ sorted) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  const uint32_t i = tmp[p];
  const uint32_t c = key[i];
  const uint32_t lo = start[c], hi = start[c + 1];
  uint32_t dst = uint32_t(p);
  if (hi - lo <= uint32_t(kRankLimit)) {
    uint32_t rank = 0;
    for (uint32_t r = lo; r < hi; ++r) rank += tmp[r] < i;
    dst = lo + rank;
  }
  perm[dst] = int32_t(i);
  sorted[dst] = lonlat[i];
}

}  // namespace

namespace rangeb200 {

int sort_cell_bits(int N) {
  // about 16 queries per cell: 6 * 4^bits ~ N / 16
  int bits = 0;
  while (bits < kMaxBits && (6LL << (2 * (bits + 1))) * 16 <= N) ++bits;
  return bits;
}

size_t sort_workspace_bytes(int N) {
  const size_t cells = (size_t(6) << (2 * sort_cell_bits(N))) + 1;
  return (2 * size_t(N) + 3 * cells) * 4 + 1024;
}

cudaError_t launch_sort_queries(const double* lonlat, int N, double* lonlat_sorted, int32_t* perm, void* workspace,
                                cudaStream_t s) {
  if (N <= 0) return cudaSuccess;
  const int bits = sort_cell_bits(N);
  const int cells = 6 << (2 * bits);
  uint32_t* key = reinterpret_cast<uint32_t*>((reinterpret_cast<size_t>(workspace) + 255) / 256 * 256);
  uint32_t* tmp = key + N;
  uint32_t* hist = tmp + N;
  uint32_t* start = hist + (cells + 1);
  uint32_t* cursor = start + (cells + 1);
  cudaError_t e = cudaMemsetAsync(hist, 0, size_t(cells + 1) * 4, s);
  if (e != cudaSuccess) return e;
  const int blocks = (N + 255) / 256;
  const double2* ll = reinterpret_cast<const double2*>(lonlat);
  cell_hist_kernel<<<blocks, 256, 0, s>>>(ll, N, bits, key, hist);
  cell_scan_kernel<<<1, 1024, 0, s>>>(hist, cells, start, cursor);
  cell_scatter_kernel<<<blocks, 256, 0, s>>>(key, N, cursor, tmp);
  cell_rank_kernel<<<blocks, 256, 0, s>>>(key, tmp, start, ll, N, perm, reinterpret_cast<double2*>(lonlat_sorted));
  return cudaGetLastError();
}

}  // namespace rangeb200
