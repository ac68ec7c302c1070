// Spatial batching of the queries (no counterpart in the reference, which treats rows independently:
// range/range.py:213-240 are all row-wise).  The geographic softmax of RANGE+ is local - exp(40 (cos d - 1)) -
// so when the 128 queries of a tile are close to each other whole database tiles can skip the geo term
// (retrieval.cu: geo_mask_kernel).  Queries arrive in arbitrary order; this file computes a permutation that
// groups them by a cube-map Hilbert cell, the encoder and retrieval run on the permuted rows and range_concat
// scatters the results back to the caller's order.
//
// The permutation is a STABLE least-significant-digit radix sort of (cell key, original index): it is a pure function
// of the coordinates, whatever the cell occupancy (clustered query sets, raster chunks), so every rank of an
// M-sharded run derives the same row order and repeated calls are bit-identical.
//
//   cell_key_kernel       key[i] = face * 4^k + hilbert(u, v)
//   per 8-bit digit:      radix_hist_kernel (per-block digit counts, digit-major) -> cell_scan_kernel (single CTA)
//                         -> radix_scatter_kernel (stable: block order, then warp order, then lane order)
//   gather_sorted_kernel  perm / sorted coordinates
#include <cstdint>
#include <cuda_runtime.h>

#include "range_kernels.h"

namespace {

constexpr int kMaxBits = 7;            // up to 6 * 4^7 = 98 304 cells
constexpr int kDigitBits = 8, kDigits = 1 << kDigitBits;
constexpr int kSortThreads = 256, kSortItems = 4, kSortTile = kSortThreads * kSortItems;   // elements per block

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
cell_key_kernel(const double2* __restrict__ lonlat, int N, int bits, uint32_t* __restrict__ key, uint32_t* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double2 p = lonlat[i];
  key[i] = cell_key(p.x, p.y, bits);
  idx[i] = uint32_t(i);
}

// hist[d * nblocks + block] = elements of this block's tile whose digit is d
__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const uint32_t* __restrict__ key, int N, int shift, uint32_t* __restrict__ hist, int nblocks) {
  __shared__ uint32_t cnt[kDigits];
  cnt[threadIdx.x] = 0;
  __syncthreads();
  const int base = blockIdx.x * kSortTile;
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const int i = base + r * kSortThreads + threadIdx.x;
    if (i < N) atomicAdd(&cnt[(key[i] >> shift) & (kDigits - 1)], 1u);
  }
  __syncthreads();
  hist[size_t(threadIdx.x) * nblocks + blockIdx.x] = cnt[threadIdx.x];
}

// Stable scatter: element i of this block goes to start[digit][block] + (number of earlier elements of the block with
// the same digit).  "Earlier" = smaller i: rounds in order, warps in order, lanes in order.
__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const uint32_t* __restrict__ key, const uint32_t* __restrict__ idx, int N, int shift,
                     const uint32_t* __restrict__ start, int nblocks, uint32_t* __restrict__ key_out,
                     uint32_t* __restrict__ idx_out) {
  constexpr int kWarps = kSortThreads / 32;
  __shared__ uint32_t running[kDigits];
  __shared__ uint32_t wcnt[kWarps][kDigits];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  running[threadIdx.x] = start[size_t(threadIdx.x) * nblocks + blockIdx.x];
  const int base = blockIdx.x * kSortTile;
  for (int r = 0; r < kSortItems; ++r) {
#pragma unroll
    for (int w = 0; w < kWarps; ++w) wcnt[w][threadIdx.x] = 0;
    __syncthreads();
    const int i = base + r * kSortThreads + threadIdx.x;
    const bool valid = i < N;
    const uint32_t k = valid ? key[i] : 0u;
    const uint32_t d = valid ? ((k >> shift) & (kDigits - 1)) : uint32_t(kDigits);      // kDigits: matches only other tails
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const uint32_t before = __popc(peers & ((1u << lane) - 1u));
    if (valid && before == 0) wcnt[warp][d] = __popc(peers);
    __syncthreads();
    {   // thread t owns digit t: exclusive prefix over the warps, then advance the block's running offset
      uint32_t acc = running[threadIdx.x];
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        const uint32_t c = wcnt[w][threadIdx.x];
        wcnt[w][threadIdx.x] = acc;
        acc += c;
      }
      running[threadIdx.x] = acc;
    }
    __syncthreads();
    if (valid) {
      const uint32_t dst = wcnt[warp][d] + before;
      key_out[dst] = k;
      idx_out[dst] = idx[i];
    }
    __syncthreads();
  }
}

// exclusive scan of hist[0..n) -> start[0..n], start[n] = total.  One CTA of 1024.
__global__ void __launch_bounds__(1024)
cell_scan_kernel(const uint32_t* __restrict__ hist, int n, uint32_t* __restrict__ start) {
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
    if (i < n) start[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) start[n] = carry;
}

__global__ void __launch_bounds__(256)
gather_sorted_kernel(const uint32_t* __restrict__ idx, const double2* __restrict__ lonlat, int N, int32_t* __restrict__ perm,
                     double2* __restrict__ sorted) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  const uint32_t i = idx[p];
  perm[p] = int32_t(i);
  sorted[p] = lonlat[i];
}

}  // namespace

namespace rangeb200 {

int sort_cell_bits(int N) {
  // about 16 queries per cell: 6 * 4^bits ~ N / 16
  int bits = 0;
  while (bits < kMaxBits && (6LL << (2 * (bits + 1))) * 16 <= N) ++bits;
  return bits;
}

static int sort_passes(int bits) { return (3 + 2 * bits + kDigitBits - 1) / kDigitBits; }
static int sort_blocks(int N) { return (N + kSortTile - 1) / kSortTile; }

int sort_launches(int N) { return N > 0 ? 2 + 3 * sort_passes(sort_cell_bits(N)) : 0; }

size_t sort_workspace_bytes(int N) {
  // key / index ping-pong buffers, digit-major block histogram and its scan
  return (4 * size_t(N) + 2 * (size_t(kDigits) * sort_blocks(N) + 1)) * 4 + 1024;
}

cudaError_t launch_sort_queries(const double* lonlat, int N, double* lonlat_sorted, int32_t* perm, void* workspace,
                                cudaStream_t s) {
  if (N <= 0) return cudaSuccess;
  const int bits = sort_cell_bits(N);
  const int nblocks = sort_blocks(N), nh = kDigits * nblocks;
  uint32_t* key[2];
  uint32_t* idx[2];
  key[0] = reinterpret_cast<uint32_t*>((reinterpret_cast<size_t>(workspace) + 255) / 256 * 256);
  idx[0] = key[0] + N;
  key[1] = idx[0] + N;
  idx[1] = key[1] + N;
  uint32_t* hist = idx[1] + N;
  uint32_t* start = hist + (nh + 1);
  const int blocks = (N + 255) / 256;
  const double2* ll = reinterpret_cast<const double2*>(lonlat);
  cell_key_kernel<<<blocks, 256, 0, s>>>(ll, N, bits, key[0], idx[0]);
  int cur = 0;
  for (int pass = 0; pass < sort_passes(bits); ++pass, cur ^= 1) {
    const int shift = pass * kDigitBits;
    radix_hist_kernel<<<nblocks, kSortThreads, 0, s>>>(key[cur], N, shift, hist, nblocks);
    cell_scan_kernel<<<1, 1024, 0, s>>>(hist, nh, start);
    radix_scatter_kernel<<<nblocks, kSortThreads, 0, s>>>(key[cur], idx[cur], N, shift, start, nblocks, key[cur ^ 1],
                                                          idx[cur ^ 1]);
  }
  gather_sorted_kernel<<<blocks, 256, 0, s>>>(idx[cur], ll, N, perm, reinterpret_cast<double2*>(lonlat_sorted));
  return cudaGetLastError();
}

}  // namespace rangeb200
