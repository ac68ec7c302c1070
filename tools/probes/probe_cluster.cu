// Micro-probes for the v2 retrieval design (run on a B200):
//   1. DSMEM bulk copy (cp.async.bulk.shared::cluster.shared::cta) all-to-all in a 4-CTA cluster:
//      correctness + bytes/cycle per SM.
//   2. tcgen05.mma with A from TMEM (TS mode), A written with tcgen05.st 32x32b, + tcgen05.commit multicast to
//      a barrier in ANOTHER CTA of the cluster.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../range_b200/csrc probe_cluster.cu -o probe_cluster
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "ptx.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void dsmem_bulk_copy(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t mbar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(mbar_cluster) : "memory");
}

// ---------------------------------------------------------------- probe 1
constexpr int CH = 16384;
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(128, 1)
dsmem_kernel(int rounds, int chunks_per_round, long long* cycles, int* errors) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* src = smem;                       // 16 KB
  uint8_t* dst = smem + CH;                  // [4][16 KB]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 5 * CH);   // [4]
  const uint32_t me = cluster_ctarank();
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) ptx::mbar_init(&full[i], 1); ptx::fence_mbar_init(); }
  uint32_t* s32 = reinterpret_cast<uint32_t*>(src);
  for (int i = threadIdx.x; i < CH / 4; i += blockDim.x) s32[i] = me * 100000u + i;
  ptx::fence_proxy_async_smem();
  __syncthreads();
  cluster_sync();
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    uint32_t phase = 0;
    for (int r = 0; r < rounds; ++r) {
      for (int p = 0; p < 4; ++p) if (p != (int)me) ptx::mbar_expect_tx(&full[p], CH * chunks_per_round);
      for (int c = 0; c < chunks_per_round; ++c)
        for (int p = 1; p < 4; ++p) {
          const uint32_t peer = (me + p) & 3;
          dsmem_bulk_copy(mapa(ptx::smem_u32(dst + me * CH), peer), ptx::smem_u32(src), CH, mapa(ptx::smem_u32(&full[me]), peer));
        }
      for (int p = 0; p < 4; ++p) if (p != (int)me) ptx::mbar_wait(&full[p], phase);
      phase ^= 1;
    }
  }
  __syncthreads();
  long long t1 = clock64();
  cluster_sync();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  int err = 0;
  for (int p = 0; p < 4; ++p) if (p != (int)me) {
    const uint32_t* d = reinterpret_cast<const uint32_t*>(dst + p * CH);
    for (int i = threadIdx.x; i < CH / 4; i += blockDim.x) if (d[i] != p * 100000u + i) ++err;
  }
  if (err) atomicAdd(errors, err);
}

// ---------------------------------------------------------------- probe 2: TS-mode MMA + remote commit
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
         "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
         "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
         "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               :: "r"(ptx::smem_u32(bar)), "h"(mask) : "memory");
}

// A [128][64] fp16 row-major (global), B [128][64] fp16 row-major (N x K) -> D [128][128] fp32 = A B^T
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
ts_kernel(const __half* A, const __half* B, float* D, int* remote_flag) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sB = smem;                                   // 128 rows x 128 B, SW128
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 16384);   // [0] mma done (local), [1] signalled by the OTHER cta
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 16384 + 64);
  const uint32_t me = cluster_ctarank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { ptx::mbar_init(&bars[0], 1); ptx::mbar_init(&bars[1], 1); ptx::fence_mbar_init(); }
  if (warp == 0) ptx::tmem_alloc<256>(slot);
  // B -> swizzled smem (manual): row r, 16B chunk c at r*128 + ((c ^ (r&7))<<4)
  for (int i = threadIdx.x; i < 128 * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sB + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + r * 64 + c * 8);
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tm = *slot;
  const uint32_t tm_a = tm + 128;            // A: 128 lanes x 32 columns (64 fp16)
  // A row -> TMEM: thread t holds row (warp*32 + lane): 64 halves = 32 packed words
  {
    const int row = warp * 32 + lane;
    uint32_t v[32];
    const uint32_t* src = reinterpret_cast<const uint32_t*>(A + row * 64);
    for (int i = 0; i < 32; ++i) v[i] = src[i];
    tmem_st32(tm_a + (uint32_t(warp * 32) << 16), v);
    tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_f16(128, 128);
      for (int kk = 0; kk < 4; ++kk)
        umma_f16_ts(tm, tm_a + kk * 8, ptx::umma_desc_kmajor_sw128(ptx::smem_u32(sB) + kk * 32), idesc, kk != 0);
      ptx::umma_commit(&bars[0]);
      umma_commit_mc(&bars[1], uint16_t(1u << (me ^ 1)));     // arrive on bars[1] of the OTHER CTA
    }
    __syncwarp();
  }
  ptx::mbar_wait(&bars[0], 0);
  ptx::mbar_wait(&bars[1], 0);          // completes only if the peer's multicast commit reached us
  ptx::tc_fence_after();
  if (threadIdx.x == 0) atomicAdd(remote_flag, 1);
  if (me == 0) {
    const int row = warp * 32 + lane;
    for (int cc = 0; cc < 4; ++cc) {
      uint32_t v[32];
      ptx::tmem_ld32(tm + (uint32_t(warp * 32) << 16) + cc * 32, v);
      ptx::tmem_ld_wait();
      for (int i = 0; i < 32; ++i) D[row * 128 + cc * 32 + i] = __uint_as_float(v[i]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 0) ptx::tmem_dealloc<256>(tm);
}

int main() {
  // ---- probe 1
  {
    long long* cyc; int* err;
    CK(cudaMallocManaged(&cyc, 64 * sizeof(long long))); CK(cudaMallocManaged(&err, sizeof(int))); *err = 0;
    const int smem = 5 * CH + 64;
    CK(cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int grid : {4, 148 / 4 * 4}) {
      for (int cpr : {1, 6}) {
        const int rounds = 200;
        dsmem_kernel<<<grid, 128, smem>>>(rounds, cpr, cyc, err);
        CK(cudaDeviceSynchronize());
        double mx = 0; for (int i = 0; i < grid; ++i) mx = mx > cyc[i] ? mx : cyc[i];
        const double bytes_out = double(rounds) * cpr * 3 * CH;
        printf("DSMEM all-to-all: grid %3d, %d x 16KB chunks/round/peer: %.0f cycles/round, out %.1f B/cycle/SM (in = same), errors %d\n",
               grid, cpr, mx / rounds, bytes_out / mx, *err);
      }
    }
  }
  // ---- probe 2
  {
    std::vector<__half> hA(128 * 64), hB(128 * 64);
    std::vector<float> fA(128 * 64), fB(128 * 64);
    srand(1);
    for (int i = 0; i < 128 * 64; ++i) { fA[i] = (rand() % 2001 - 1000) / 1000.f; fB[i] = (rand() % 2001 - 1000) / 1000.f;
      hA[i] = __float2half(fA[i]); hB[i] = __float2half(fB[i]); fA[i] = __half2float(hA[i]); fB[i] = __half2float(hB[i]); }
    __half *dA, *dB; float* dD; int* flag;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, 128 * 128 * 4));
    CK(cudaMallocManaged(&flag, 4)); *flag = 0;
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 20000));
    ts_kernel<<<2, 128, 20000>>>(dA, dB, dD, flag);
    CK(cudaDeviceSynchronize());
    std::vector<float> hD(128 * 128);
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 128; ++n) {
      double ref = 0; for (int k = 0; k < 64; ++k) ref += double(fA[m * 64 + k]) * fB[n * 64 + k];
      maxerr = fmax(maxerr, fabs(ref - hD[m * 128 + n]));
    }
    printf("TS-mode MMA (A in TMEM via tcgen05.st): max abs err %.3e ; remote multicast commit reached both CTAs: %s\n",
           maxerr, *flag == 2 ? "yes" : "NO");
  }
  return 0;
}
