// Accuracy of packed half-precision exponentials for the softmax SUMS (statistics pass): bias and spread of
//   2^x ~ 2^floor(x) * ex2.approx.f16x2(frac(x))      against exp2 in double,  x in [-35, 0]
// next to the fp32 MUFU path (ex2.approx.ftz.f32).  A systematic bias of the f16x2 unit would shift every row
// normaliser by the same factor - that is what would decide whether the sums may use it.
// FINDING (cuobjdump -sass, sm_100a, CUDA 12.9): ex2.approx.f16x2 is NOT a packed MUFU operation here - ptxas emits two
// MUFU.EX2.F16 instructions (one per half: `MUFU.EX2.F16 R34, R7` / `MUFU.EX2.F16 R36, R7.H1`), so it issues as many
// MUFU operations as two fp32 exponentials.  Half-precision exponentials cannot halve the statistics pass's MUFU work
// on B200; the idea is closed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/probe_ex2 tools/probes/probe_ex2.cu && build/probe_ex2
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2_f32(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_h2(uint32_t x) { uint32_t y; asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }

__global__ void probe(int n, double* acc) {
  // acc: [0..2] f32 path: sum rel err, sum rel err^2, max |rel err|; [3..5] f16x2 path; [6..7] sums of values (f32, f16x2); [8] exact
  double s1 = 0, s2 = 0, m1 = 0, t1 = 0, t2 = 0, m2 = 0, v1 = 0, v2 = 0, v0 = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float xa = -35.0f * (float(i) + 0.5f) / float(n), xb = -35.0f * (float(n - 1 - i) + 0.25f) / float(n);
    const double ea = exp2(double(xa)), eb = exp2(double(xb));
    const float fa = ex2_f32(xa), fb = ex2_f32(xb);
    // f16x2 path: fractional parts in (-1, 0], packed; integer parts applied as exponent offsets in fp32
    const float ia = floorf(xa) + 1.f, ib = floorf(xb) + 1.f;          // xf = x - i in (-1, 0]
    const __half2 xf = __floats2half2_rn(xa - ia, xb - ib);
    const uint32_t r = ex2_h2(*reinterpret_cast<const uint32_t*>(&xf));
    const float2 rf = __half22float2(*reinterpret_cast<const __half2*>(&r));
    const float ha = rf.x * exp2f(ia), hb = rf.y * exp2f(ib);
    const double ra = (fa - ea) / ea, rb = (fb - eb) / eb, qa = (ha - ea) / ea, qb = (hb - eb) / eb;
    s1 += ra + rb; s2 += ra * ra + rb * rb; m1 = fmax(m1, fmax(fabs(ra), fabs(rb)));
    t1 += qa + qb; t2 += qa * qa + qb * qb; m2 = fmax(m2, fmax(fabs(qa), fabs(qb)));
    v1 += fa + fb; v2 += ha + hb; v0 += ea + eb;
  }
  atomicAdd(&acc[0], s1); atomicAdd(&acc[1], s2); atomicAdd(&acc[3], t1); atomicAdd(&acc[4], t2);
  atomicAdd(&acc[6], v1); atomicAdd(&acc[7], v2); atomicAdd(&acc[8], v0);
  // max via atomic on the bit pattern (non-negative doubles order like integers)
  atomicMax(reinterpret_cast<unsigned long long*>(&acc[2]), (unsigned long long)__double_as_longlong(m1));
  atomicMax(reinterpret_cast<unsigned long long*>(&acc[5]), (unsigned long long)__double_as_longlong(m2));
}

int main() {
  const int n = 1 << 24;
  double* d; cudaMalloc(&d, 9 * sizeof(double)); cudaMemset(d, 0, 9 * sizeof(double));
  probe<<<296, 256>>>(n, d);
  double h[9]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  if (cudaGetLastError() != cudaSuccess) { printf("probe failed\n"); return 1; }
  const double cnt = 2.0 * n;
  printf("ex2.approx.ftz.f32  : mean rel err %+.3e  rms %.3e  max %.3e  | sum ratio %.9f\n", h[0] / cnt, sqrt(h[1] / cnt), h[2], h[6] / h[8]);
  printf("ex2.approx.f16x2: mean rel err %+.3e  rms %.3e  max %.3e  | sum ratio %.9f\n", h[3] / cnt, sqrt(h[4] / cnt), h[5], h[7] / h[8]);
  return 0;
}
