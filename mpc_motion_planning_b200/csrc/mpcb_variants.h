// Kernel families and the host-side description of one launchable variant.  Every family is its own
// translation unit (mpcb_variants.cu compiled with -DMPCB_FAMILY=k) so that the build runs in parallel.
#pragma once
#include "mpcb_kernel.cuh"

namespace mpcb {

typedef cudaError_t (*launch_fn)(const KParams &, int grid, size_t smem, cudaStream_t);
typedef const void *kernel_ptr;

struct Variant {
  launch_fn launch;
  kernel_ptr kernel;
  size_t (*smem_bytes)(int N);
  int nx, nbx;
  size_t slab_doubles;  // per resident warp; 0 = everything in shared memory, no slab
  int warps;            // warps (= scenarios in flight) per block
  // optional small-batch sibling (one warp per block, the whole iterate in shared memory): used when the
  // batch fits the SMs in one wave, where the L2 round trips of the slab are pure latency
  launch_fn lat_launch;
  kernel_ptr lat_kernel;
  size_t (*lat_smem_bytes)(int N);  // per block
  int lat_warps;
  // optional restoration-capable sibling (second pass over the scenarios whose line search failed, see KParams)
  bool rs_inline;  // the main kernels carry the restoration phase themselves (KParams.restoration switches it on)
  launch_fn resto_launch;
  kernel_ptr resto_kernel;
  size_t (*resto_smem_bytes)(int N);  // per warp
  size_t resto_slab_doubles;
  int resto_warps;
};

// second engine: one scenario per lane (mpcb_lane_kernel.cuh), kinematic families with plain rows
typedef cudaError_t (*lane_launch_fn)(const KParams &, double *ws, size_t nslot, int grid, cudaStream_t);
struct LaneVariant {
  lane_launch_fn launch;
  kernel_ptr kernel;
  size_t (*slot_doubles)(int N);  // workspace doubles per lane
  int block;
};
LaneVariant lane_variant_kin_0_0();
LaneVariant lane_variant_kin_1_0();
LaneVariant lane_variant_kin_1_1();
LaneVariant lane_variant_kin_1_2();
LaneVariant lane_variant_kin_1_3();
LaneVariant lane_variant_kin_1_4();
LaneVariant lane_variant_rk4_0_0();  // Runge-Kutta shooting defects
LaneVariant lane_variant_rk4_1_0();
LaneVariant lane_variant_rk4_1_1();
LaneVariant lane_variant_rk4_1_2();

// the candidate that keeps the most warps resident for horizon N (first one wins ties); mpcb_api.cu
Variant pick_by_occupancy(Variant *cand, int n, int N);

// variant_kin_<NR>_<MO>_<OBS>: rate rows, obstacles, obstacle-row mode (0 none, 1 h >= 0, 3 discrete-time CBF)
Variant variant_kin_0_0_0(int N);
Variant variant_kin_1_0_0(int N);
Variant variant_kin_1_1_1(int N);
Variant variant_kin_1_2_1(int N);
Variant variant_kin_1_3_1(int N);
Variant variant_kin_1_1_3(int N);
Variant variant_kin_1_2_3(int N);
Variant variant_kin_1_3_3(int N);
Variant variant_kin_1_4_1(int N);
Variant variant_kin_1_4_3(int N);
Variant variant_dyn(int N);
Variant variant_dyn_shipped(int N);  // dyn rows paired with the bound lists as shipped (cfg.dyn_rows)

}  // namespace mpcb
