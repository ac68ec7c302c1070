// x = hi + lo with hi, lo fp16 (11 + 11 significant bits): the operand split of the tensor-core encoder's first layer.
// One definition for every producer of harmonics features (per-query kernels, raster combine): they must agree bit for bit.
#pragma once
#include <cuda_fp16.h>

namespace rangeb200 {

__device__ __forceinline__ double f16_bits_to_double(unsigned short h) {
  double d;
  asm("cvt.f64.f16 %0, %1;" : "=d"(d) : "h"(h));
  return d;
}

// hi = rn_f16(x), lo = rn_f16(x - hi): two F2F.F16.F64, one F2F.F64.F16 and one DADD (going through fp32 costs three
// more instructions per conversion on sm_100: F2F.F32.F64 needs a subnormal fix-up)
__device__ __forceinline__ void split_f16(double x, __half& hi, __half& lo) {
  hi = __double2half(x);
  lo = __double2half(__dsub_rn(x, f16_bits_to_double(__half_as_ushort(hi))));
}

// (x, y) -> hi / lo as packed half2 bit patterns (x in the low half)
__device__ __forceinline__ void split_f16x2(double x, double y, uint32_t& hi, uint32_t& lo) {
  __half hx, lx, hy, ly;
  split_f16(x, hx, lx);
  split_f16(y, hy, ly);
  hi = uint32_t(__half_as_ushort(hx)) | uint32_t(__half_as_ushort(hy)) << 16;
  lo = uint32_t(__half_as_ushort(lx)) | uint32_t(__half_as_ushort(ly)) << 16;
}

}  // namespace rangeb200
