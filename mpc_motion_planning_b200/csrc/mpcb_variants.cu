// One kernel family per compilation: nvcc -DMPCB_FAMILY=k -c mpcb_variants.cu (see build.py).
#include "mpcb_variants.h"

#ifndef MPCB_FAMILY
#error "compile with -DMPCB_FAMILY=0..11"
#endif

#if MPCB_FAMILY == 8 || MPCB_FAMILY == 9
#include "mpcb_dyn_kernel.cuh"
#endif

namespace mpcb {

#if MPCB_FAMILY < 8 || MPCB_FAMILY >= 10

// MPCB_RS_INLINE = 1: the main kernels of the families with inequality rows carry the restoration phase themselves (a
// failed line search continues in place); 0 (shipped): the failed scenarios are solved again by a sibling kernel.
// Measured at B = 10,000 kin-CBF (tests/tools/resto_check.py, profiles/r02_restoration.txt): inline costs the regular
// path 15 % (168-register cap: the restoration branches spill) - 24.3 ms with restoration off against 21.2 ms - and
// 30.9 ms with it on; the sibling leaves the regular path untouched and takes 33.5 ms with restoration on.
#ifndef MPCB_RS_INLINE
#define MPCB_RS_INLINE 0
#endif
template <int NR, int MO>
constexpr bool kRsMain = MPCB_RS_INLINE && (NR + MO > 0);

template <int NR, int MO, int OBS, int W, bool GS>
static cudaError_t launch_kin(const KParams &p, int grid, size_t smem, cudaStream_t st) {
  kin_solve_kernel<NR, MO, OBS, W, GS, false, kRsMain<NR, MO>><<<grid, 32 * W, smem, st>>>(p);
  return cudaGetLastError();
}

template <int NR, int MO, int OBS, int W, bool GS>
static Variant make_kin_variant_w() {
  Variant v;
  v.launch = &launch_kin<NR, MO, OBS, W, GS>;
  v.kernel = (const void *)&kin_solve_kernel<NR, MO, OBS, W, GS, false, kRsMain<NR, MO>>;
  v.smem_bytes = [](int N) { return KinLayout<NR, MO, OBS == 3, GS, false, kRsMain<NR, MO>>::bytes(N); };
  v.nx = 4;
  v.nbx = 2;
  v.slab_doubles = KinLayout<NR, MO, OBS == 3, GS, false, kRsMain<NR, MO>>::slab_doubles();
  v.rs_inline = kRsMain<NR, MO>;
  v.warps = W;
#ifndef MPCB_AS_W
#define MPCB_AS_W 1  // warps per block of the all-shared kernel (compile-time tuning knob)
#endif
  v.lat_launch = [](const KParams &p, int grid, size_t smem, cudaStream_t st) {
    kin_solve_kernel<NR, MO, OBS, MPCB_AS_W, false, true, kRsMain<NR, MO>><<<grid, 32 * MPCB_AS_W, smem, st>>>(p);
    return cudaGetLastError();
  };
  v.lat_kernel = (const void *)&kin_solve_kernel<NR, MO, OBS, MPCB_AS_W, false, true, kRsMain<NR, MO>>;
  v.lat_smem_bytes = [](int N) { return KinLayout<NR, MO, OBS == 3, false, true, kRsMain<NR, MO>>::bytes(N) * MPCB_AS_W; };
  v.lat_warps = MPCB_AS_W;
  // restoration-capable sibling: the step lives in the slab (smallest shared-memory record), 2 warps per block,
  // registers uncapped.  Only the families with inequality rows have one.
  v.resto_launch = nullptr;
  v.resto_kernel = nullptr;
  v.resto_smem_bytes = nullptr;
  v.resto_slab_doubles = 0;
  v.resto_warps = 0;
  if constexpr (NR + MO > 0 && !kRsMain<NR, MO>) {
    constexpr int RW = 2;
    v.resto_launch = [](const KParams &p, int grid, size_t smem, cudaStream_t st) {
      // programmatic dependent launch behind the main kernel (same stream): these blocks may become resident once every
      // block of the main grid has started, i.e. as its last wave drains; they never call griddepcontrol.wait - the
      // hand-over of scenarios goes through resto_list / resto_sync
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid);
      cfg.blockDim = dim3(32 * RW);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      return cudaLaunchKernelEx(&cfg, kin_solve_kernel<NR, MO, OBS, RW, true, false, true>, p);
    };
    v.resto_kernel = (const void *)&kin_solve_kernel<NR, MO, OBS, RW, true, false, true>;
    v.resto_smem_bytes = [](int N) { return KinLayout<NR, MO, OBS == 3, true, false, true>::bytes(N); };
    v.resto_slab_doubles = KinLayout<NR, MO, OBS == 3, true, false, true>::slab_doubles();
    v.resto_warps = RW;
  }
  return v;
}

// Layout and warps per block: the candidate that keeps the most warps resident for this horizon.
// Order = preference on ties: step in shared memory before step in the slab, larger W first.
template <int NR, int MO, int OBS>
static Variant make_kin_variant(int N) {
#ifndef MPCB_W0  // candidate warps-per-block values (compile-time tuning knob)
#define MPCB_W0 4
#define MPCB_W1 2
#define MPCB_W2 1
#endif
  Variant cand[6] = {make_kin_variant_w<NR, MO, OBS, MPCB_W0, false>(), make_kin_variant_w<NR, MO, OBS, MPCB_W1, false>(),
                     make_kin_variant_w<NR, MO, OBS, MPCB_W2, false>(), make_kin_variant_w<NR, MO, OBS, MPCB_W0, true>(),
                     make_kin_variant_w<NR, MO, OBS, MPCB_W1, true>(), make_kin_variant_w<NR, MO, OBS, MPCB_W2, true>()};
  return pick_by_occupancy(cand, 6, N);
}

#if MPCB_FAMILY == 0
Variant variant_kin_0_0_0(int N) { return make_kin_variant<0, 0, 0>(N); }
#elif MPCB_FAMILY == 1
Variant variant_kin_1_0_0(int N) { return make_kin_variant<1, 0, 0>(N); }
#elif MPCB_FAMILY == 2
Variant variant_kin_1_1_1(int N) { return make_kin_variant<1, 1, 1>(N); }
#elif MPCB_FAMILY == 3
Variant variant_kin_1_2_1(int N) { return make_kin_variant<1, 2, 1>(N); }
#elif MPCB_FAMILY == 4
Variant variant_kin_1_3_1(int N) { return make_kin_variant<1, 3, 1>(N); }
#elif MPCB_FAMILY == 5
Variant variant_kin_1_1_3(int N) { return make_kin_variant<1, 1, 3>(N); }
#elif MPCB_FAMILY == 6
Variant variant_kin_1_2_3(int N) { return make_kin_variant<1, 2, 3>(N); }
#elif MPCB_FAMILY == 7
Variant variant_kin_1_3_3(int N) { return make_kin_variant<1, 3, 3>(N); }
#elif MPCB_FAMILY == 10
Variant variant_kin_1_4_1(int N) { return make_kin_variant<1, 4, 1>(N); }
#elif MPCB_FAMILY == 11
Variant variant_kin_1_4_3(int N) { return make_kin_variant<1, 4, 3>(N); }
#endif

#else  // MPCB_FAMILY 8, 9: dynamic bicycle, rows aligned / rows as shipped

constexpr bool kShipped = MPCB_FAMILY == 9;

template <int W>
static cudaError_t launch_dyn(const KParams &p, int grid, size_t smem, cudaStream_t st) {
  dyn_solve_kernel<W, kShipped><<<grid, 32 * W, smem, st>>>(p);
  return cudaGetLastError();
}

template <int W>
static Variant make_dyn_variant_w() {
  Variant v;
  v.launch = &launch_dyn<W>;
  v.kernel = (const void *)&dyn_solve_kernel<W, kShipped>;
  v.smem_bytes = [](int N) { return DynLayout::bytes(N); };
  v.nx = 6;
  v.nbx = 3;
  v.slab_doubles = DynLayout::slab_doubles();
  v.warps = W;
  v.lat_launch = nullptr;
  v.lat_kernel = nullptr;
  v.lat_smem_bytes = nullptr;
  v.lat_warps = 0;
  v.rs_inline = false;
  v.resto_launch = nullptr;
  v.resto_kernel = nullptr;
  v.resto_smem_bytes = nullptr;
  v.resto_slab_doubles = 0;
  v.resto_warps = 0;
  return v;
}

#if MPCB_FAMILY == 8
Variant variant_dyn(int N) {
#else
Variant variant_dyn_shipped(int N) {
#endif
  Variant cand[3] = {make_dyn_variant_w<4>(), make_dyn_variant_w<2>(), make_dyn_variant_w<1>()};
  return pick_by_occupancy(cand, 3, N);
}

#endif

}  // namespace mpcb
