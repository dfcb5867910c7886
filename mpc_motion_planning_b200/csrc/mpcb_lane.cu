// Lane-per-scenario engine (mpcb_lane_kernel.cuh): one instantiation per compilation,
// nvcc -DMPCB_LANE_FAMILY=k -c mpcb_lane.cu (see build.py).
#include "mpcb_variants.h"
#include "mpcb_lane_kernel.cuh"

#ifndef MPCB_LANE_FAMILY
#error "compile with -DMPCB_LANE_FAMILY=0..5"
#endif

namespace mpcb {

template <int NR, int MO>
static LaneVariant make_lane_variant() {
  LaneVariant v;
  v.launch = [](const KParams &p, double *ws, size_t nslot, int grid, cudaStream_t st) {
    kin_lane_kernel<NR, MO><<<grid, MPCB_LANE_BLOCK, 0, st>>>(p, ws, nslot);
    return cudaGetLastError();
  };
  v.kernel = (const void *)&kin_lane_kernel<NR, MO>;
  v.slot_doubles = [](int N) { return LaneLayout<NR, MO>::slot_doubles(N); };
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
#endif

}  // namespace mpcb
