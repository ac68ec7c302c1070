// harmonics_calculation == 'closed-form': the unnormalised associated-Legendre recurrence of
// range/location_models/satclip/positional_encoding/spherical_harmonics_closed_form.py:8-26 (with Condon-Shortley
// phase), evaluated in the reference's operation order with explicitly rounded fp64 operations (no FMA contraction), so
// the features agree with the reference to the last bits (cos / sin of the library aside).
#pragma once

namespace rangeb200 {

struct ClosedFormLegendre {
  double x, somx2, pmm, fact, pm1, pm2;
  __device__ __forceinline__ void init(double cos_theta) {
    x = cos_theta;
    somx2 = sqrt(__dmul_rn(1.0 - x, 1.0 + x));       // closed_form.py:11
    pmm = 1.0;
    fact = 1.0;
  }
  // call once per |m| = am (ascending from 0): P_am^am                                           (:10-15)
  __device__ __forceinline__ void start_order(int am) {
    if (am > 0) {
      pmm = __dmul_rn(__dmul_rn(pmm, -fact), somx2);
      fact += 2.0;
    }
    pm1 = pm2 = 0.0;
  }
  // call for l = am, am + 1, ... in order: P_l^am                                                (:16-26)
  __device__ __forceinline__ double next(int l, int am) {
    double p;
    if (l == am) p = pmm;
    else if (l == am + 1) p = __dmul_rn(__dmul_rn(x, 2.0 * am + 1.0), pmm);
    else p = __ddiv_rn(__dsub_rn(__dmul_rn(__dmul_rn(2.0 * l - 1.0, x), pm1), __dmul_rn(double(l + am) - 1.0, pm2)),
                       double(l - am));
    pm2 = pm1;
    pm1 = p;
    return p;
  }
};

}  // namespace rangeb200
