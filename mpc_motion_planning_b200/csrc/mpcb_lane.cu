// Lane-per-scenario engine (mpcb_lane_kernel.cuh): one instantiation per compilation,
// nvcc -DMPCB_LANE_FAMILY=k -c mpcb_lane.cu (see build.py).
#include "mpcb_variants.h"
#include "mpcb_lane_kernel.cuh"

#ifndef MPCB_LANE_FAMILY
#error "compile with -DMPCB_LANE_FAMILY=0..9"
#endif

namespace mpcb {

template <int NR, int MO, bool RK4 = false>
static LaneVariant make_lane_variant() {
  LaneVariant v;
  v.launch = [](const KParams &p, double *ws, size_t nslot, int grid, cudaStream_t st) {
    kin_lane_kernel<NR, MO, RK4><<<grid, MPCB_LANE_BLOCK, 0, st>>>(p, ws, nslot);
    return cudaGetLastError();
  };
  v.kernel = (const void *)&kin_lane_kernel<NR, MO, RK4>;
  v.slot_doubles = [](int N) { return LaneLayout<NR, MO, RK4>::slot_doubles(N); };
  v.block = MPCB_LANE_BLOCK;
  return v;
}

#if MPCB_LANE_FAMILY == 0
LaneVariant lane_variant_kin_0_0() { return make_lane_variant<0, 0>(); }
#elif MPCB_LANE_FAMILY == 1
LaneVariant lane_variant_kin_1_0() { return make_lane_variant<1, 0>(); }
#elif MPCB_LANE_FAMILY == 2
LaneVariant lane_variant_kin_1_1() { return make_lane_variant<1, 1>(); }
#elif MPCB_LANE_FAMILY == 3
LaneVariant lane_variant_kin_1_2() { return make_lane_variant<1, 2>(); }
#elif MPCB_LANE_FAMILY == 4
LaneVariant lane_variant_kin_1_3() { return make_lane_variant<1, 3>(); }
#elif MPCB_LANE_FAMILY == 5
LaneVariant lane_variant_kin_1_4() { return make_lane_variant<1, 4>(); }
// Runge-Kutta shooting defects (cfg.integrator = MPCB_INTEGRATOR_RK4)
#elif MPCB_LANE_FAMILY == 6
LaneVariant lane_variant_rk4_0_0() { return make_lane_variant<0, 0, true>(); }
#elif MPCB_LANE_FAMILY == 7
LaneVariant lane_variant_rk4_1_0() { return make_lane_variant<1, 0, true>(); }
#elif MPCB_LANE_FAMILY == 8
LaneVariant lane_variant_rk4_1_1() { return make_lane_variant<1, 1, true>(); }
#elif MPCB_LANE_FAMILY == 9
LaneVariant lane_variant_rk4_1_2() { return make_lane_variant<1, 2, true>(); }
#endif

}  // namespace mpcb
