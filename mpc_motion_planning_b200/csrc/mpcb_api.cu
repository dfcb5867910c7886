// mpcb200 host side of the C ABI declared in include/mpcb200.h.  No torch types, no
// exceptions across the boundary.  There is no CPU fallback: without a CUDA device
// mpcb_create fails with MPCB_E_NODEVICE.
#include "../../include/mpcb200.h"
#include "mpcb_variants.h"
#include "dyn_model.cuh"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <new>

using namespace mpcb;

#ifndef MPCB_LANE_MIN_B_DEFAULT
#define MPCB_LANE_MIN_B_DEFAULT 20480  // MPCB_ENGINE_AUTO: batches of the row-free families at least this large go to the lane engine
#endif

namespace {

thread_local char g_cuda_err[256] = "";

bool cuda_ok(cudaError_t e, const char *what) {
  if (e == cudaSuccess) return true;
  snprintf(g_cuda_err, sizeof g_cuda_err, "%s: %s", what, cudaGetErrorString(e));
  return false;
}

}  // namespace

namespace mpcb {
// shared by the kernel-family translation units (mpcb_variants.cu)
Variant pick_by_occupancy(Variant *cand, int n, int N) {
  if (const char *w = getenv("MPCB_FORCE_W")) {  // tuning knob: warps per block
    for (int i = 0; i < n; i++)
      if (cand[i].warps == atoi(w)) return cand[i];
  }
  int best = 0, best_warps = -1;
  for (int i = 0; i < n; i++) {
    size_t smem = cand[i].smem_bytes(N) * cand[i].warps;
    int bps = 0;
    if (cudaFuncSetAttribute(cand[i].kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); continue; }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, cand[i].kernel, 32 * cand[i].warps, smem) != cudaSuccess) { cudaGetLastError(); continue; }
    if (bps * cand[i].warps > best_warps) { best_warps = bps * cand[i].warps; best = i; }
  }
  return cand[best];
}
}  // namespace mpcb

namespace {

bool select_variant(const mpcb_cfg &c, Variant &v) {
  const int M = c.obs_mode == MPCB_OBS_NONE ? 0 : c.M;
  if (c.model == MPCB_MODEL_DYN) {
    // the reference's dyn NLP: both rate rows (df, ax in that order), one obstacle, sqrt rows
    if (c.obs_mode == MPCB_OBS_SQRT && c.n_rate == 2 && M == 1 && c.rate_ctrl[0] == 0 && c.rate_ctrl[1] == 1) {
      v = c.dyn_rows == MPCB_DYN_ROWS_AS_SHIPPED ? variant_dyn_shipped(c.N) : variant_dyn(c.N);
      return true;
    }
    return false;
  }
  if (c.model == MPCB_MODEL_KIN) {
    if (c.obs_mode == MPCB_OBS_NONE && c.n_rate == 0) { v = variant_kin_0_0_0(c.N); return true; }
    if (c.obs_mode == MPCB_OBS_NONE && c.n_rate == 1) { v = variant_kin_1_0_0(c.N); return true; }
    if (c.obs_mode == MPCB_OBS_ELLIPSE && c.n_rate == 1 && M == 1) { v = variant_kin_1_1_1(c.N); return true; }
    if (c.obs_mode == MPCB_OBS_ELLIPSE && c.n_rate == 1 && M == 2) { v = variant_kin_1_2_1(c.N); return true; }
    if (c.obs_mode == MPCB_OBS_ELLIPSE && c.n_rate == 1 && M == 3) { v = variant_kin_1_3_1(c.N); return true; }
    if (c.obs_mode == MPCB_OBS_ELLIPSE && c.n_rate == 1 && M == 4) { v = variant_kin_1_4_1(c.N); return true; }
    if (c.obs_mode == MPCB_OBS_DCBF && c.n_rate == 1 && M == 1) { v = variant_kin_1_1_3(c.N); return true; }
    if (c.obs_mode == MPCB_OBS_DCBF && c.n_rate == 1 && M == 2) { v = variant_kin_1_2_3(c.N); return true; }
    if (c.obs_mode == MPCB_OBS_DCBF && c.n_rate == 1 && M == 3) { v = variant_kin_1_3_3(c.N); return true; }
    if (c.obs_mode == MPCB_OBS_DCBF && c.n_rate == 1 && M == 4) { v = variant_kin_1_4_3(c.N); return true; }
  }
  return false;
}

// lane-per-scenario engine: kinematic model, plain rows (none / h >= 0), one target per scenario, at most one rate row
// and that on the steering angle
bool select_lane_variant(const mpcb_cfg &c, LaneVariant &v) {
  if (c.model != MPCB_MODEL_KIN || c.ref_mode != MPCB_REF_TERMINAL) return false;
  const int M = c.obs_mode == MPCB_OBS_NONE ? 0 : c.M;
  if (c.obs_mode != MPCB_OBS_NONE && c.obs_mode != MPCB_OBS_ELLIPSE) return false;
  if (c.integrator == MPCB_INTEGRATOR_RK4) {
    if (c.n_rate == 0 && M == 0) { v = lane_variant_rk4_0_0(); return true; }
    if (c.n_rate != 1 || c.rate_ctrl[0] != 0) return false;
    switch (M) {
      case 0: v = lane_variant_rk4_1_0(); return true;
      case 1: v = lane_variant_rk4_1_1(); return true;
      case 2: v = lane_variant_rk4_1_2(); return true;
    }
    return false;
  }
  if (c.n_rate == 0 && M == 0) { v = lane_variant_kin_0_0(); return true; }
  if (c.n_rate != 1 || c.rate_ctrl[0] != 0) return false;
  switch (M) {
    case 0: v = lane_variant_kin_1_0(); return true;
    case 1: v = lane_variant_kin_1_1(); return true;
    case 2: v = lane_variant_kin_1_2(); return true;
    case 3: v = lane_variant_kin_1_3(); return true;
    case 4: v = lane_variant_kin_1_4(); return true;
  }
  return false;
}

double relax_lo(double b, double f) { return std::isfinite(b) ? b - f * std::fmax(1.0, std::fabs(b)) : b; }
double relax_hi(double b, double f) { return std::isfinite(b) ? b + f * std::fmax(1.0, std::fabs(b)) : b; }

}  // namespace

struct mpcb_handle {
  mpcb_cfg cfg;
  Variant var;
  KParams kp;  // everything except pointers and B
  size_t smem;
  int device;
  mpcb_launch_info info;
  int persistent_grid;  // resident blocks of the persistent kernel (0: grid = B)
  int lat_grid;         // resident blocks of the small-batch kernel (0: not available for this configuration)
  size_t lat_smem;
  mpcb_launch_info main_info, lat_info;  // static launch facts of the two kernels (info = the one last launched)
  double *d_slab;
  int *d_counter;
  // device buffers for the host-pointer entry point
  cudaStream_t stream;
  int cap_B;
  double *d_x0, *d_xs, *d_obs, *d_zin, *d_u0, *d_cost, *d_z, *d_lam, *d_lamg, *d_lamx;
  int32_t *d_status, *d_iters;
  // one solve at a time per handle: the work queue head and the slab belong to the launch in flight.  A solve
  // issued on another stream first waits (on the device) for the previous one.
  cudaEvent_t last_done;
  cudaStream_t last_stream;
  bool has_last;
  int order_n;                    // length of the installed processing order
  double *lam_g_out, *lam_x_out;  // optional dual outputs (mpcb_set_dual_outputs)
  int duals_on_host;
  // restoration pass: scenarios whose line search failed in the main kernel (list + count on the device), the
  // second work-queue head, and the launch geometry of the restoration-capable sibling kernel
  int32_t *d_resto_list;
  int *d_resto_count;  // resto_sync of KParams: [0] entries reserved, [1] consumer tickets, [2] main warps finished
  int resto_cap, resto_grid;
  size_t resto_smem;
  double *d_resto_slab;
  // lane-per-scenario engine: workspace of one slot per resident lane, used for batches of at least lane_min_B
  LaneVariant lane;
  double *d_lane_ws;
  size_t lane_nslot;
  int lane_grid, lane_min_B;
  bool lane_forced;  // cfg.engine = LANE or cfg.integrator = RK4: never fall back to the warp kernels
  mpcb_launch_info lane_info;
  int *d_debug_errors;  // -DMPCB_DEBUG_SLOTS builds: slot-ownership violations seen by the kernels
};

extern "C" {

int mpcb_version(void) { return MPCB_VERSION; }

const char *mpcb_strerror(int code) {
  switch (code) {
    case MPCB_OK: return "ok";
    case MPCB_E_ARG: return "bad argument or unsupported configuration";
    case MPCB_E_CUDA: return "CUDA runtime error";
    case MPCB_E_NOMEM: return "out of memory";
    case MPCB_E_NODEVICE: return "no CUDA device (this library has no CPU fallback)";
    default: return "unknown error";
  }
}

const char *mpcb_last_cuda_error(void) { return g_cuda_err; }

int mpcb_nx(const mpcb_cfg *cfg) { return cfg && cfg->model == MPCB_MODEL_DYN ? 6 : 4; }
int mpcb_nv(const mpcb_cfg *cfg) { return cfg ? 2 * cfg->N + mpcb_nx(cfg) * (cfg->N + 1) : 0; }

int mpcb_workspace_bytes(const mpcb_cfg *cfg, int B, size_t *bytes) {
  if (!cfg || !bytes || B < 0) return MPCB_E_ARG;
  // Independent of B: the kernels are persistent and keep the primal-dual iterate of each RESIDENT
  // warp in a global-memory slab (allocated by mpcb_create, sized by the occupancy of the device:
  // at most 32 warps x #SMs).
  Variant v;
  if (!select_variant(*cfg, v)) return MPCB_E_ARG;
  int ndev = 0, dev = 0, sms = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); *bytes = 0; return MPCB_OK; }
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  *bytes = v.slab_doubles * sizeof(double) * (size_t)sms * 32;
  // plus the lane-per-scenario engine's workspace: one slot per resident lane (at most 2048 threads per SM)
  LaneVariant lv;
  const char *eng = getenv("MPCB_ENGINE");
  if (!(eng && !strcmp(eng, "warp")) && cfg->engine != MPCB_ENGINE_WARP && select_lane_variant(*cfg, lv) &&
      (cfg->engine == MPCB_ENGINE_LANE || cfg->obs_mode == MPCB_OBS_NONE || (eng && !strcmp(eng, "lane"))))
    *bytes += lv.slot_doubles(cfg->N) * sizeof(double) * (size_t)sms * 2048;
  return MPCB_OK;
}

int mpcb_create(const mpcb_cfg *cfg, mpcb_handle **out) {
  if (!cfg || !out) return MPCB_E_ARG;
  *out = nullptr;
  const mpcb_cfg &c = *cfg;
  if (c.N < 2 || c.N > MPCB_NMAX || c.M < 0 || c.M > MPCB_MMAX || c.max_iter < 0) return MPCB_E_ARG;
  if (c.n_rate < 0 || c.n_rate > 2) return MPCB_E_ARG;
  for (int r = 0; r < c.n_rate; r++)
    if (c.rate_ctrl[r] < 0 || c.rate_ctrl[r] > 1 || !(c.rate_lo[r] < c.rate_hi[r])) return MPCB_E_ARG;
  if (!(c.T > 0) || !(c.tol > 0) || !(c.mu_init > 0) || !(c.bound_relax >= 0)) return MPCB_E_ARG;
  if (c.obs_mode != MPCB_OBS_NONE && c.M < 1) return MPCB_E_ARG;
  if (c.obs_input != MPCB_OBS_TRAJECTORY && c.obs_input != MPCB_OBS_INITIAL && c.obs_input != MPCB_OBS_STATIC) return MPCB_E_ARG;
  if (c.ref_mode != MPCB_REF_TERMINAL && c.ref_mode != MPCB_REF_TRAJECTORY) return MPCB_E_ARG;
  if (c.ref_mode == MPCB_REF_TRAJECTORY && c.model != MPCB_MODEL_KIN) return MPCB_E_ARG;
  if (c.obs_mode == MPCB_OBS_DCBF && !(c.cbf_gamma > 0.0 && c.cbf_gamma <= 1.0)) return MPCB_E_ARG;
  if (c.dyn_rows != MPCB_DYN_ROWS_ALIGNED && c.dyn_rows != MPCB_DYN_ROWS_AS_SHIPPED) return MPCB_E_ARG;
  if (c.dyn_rows == MPCB_DYN_ROWS_AS_SHIPPED && c.model != MPCB_MODEL_DYN) return MPCB_E_ARG;
  if (c.engine != MPCB_ENGINE_AUTO && c.engine != MPCB_ENGINE_WARP && c.engine != MPCB_ENGINE_LANE) return MPCB_E_ARG;
  if (c.integrator != MPCB_INTEGRATOR_EULER && c.integrator != MPCB_INTEGRATOR_RK4) return MPCB_E_ARG;
  Variant var;
  if (!select_variant(c, var)) return MPCB_E_ARG;
  // the kernels assume: both controls two-sided; the model's bounded states (kin: y, vx; dyn: + vy)
  // two-sided; every other state free
  for (int i = 0; i < 2; i++)
    if (!std::isfinite(c.u_lo[i]) || !std::isfinite(c.u_hi[i]) || !(c.u_lo[i] < c.u_hi[i])) return MPCB_E_ARG;
  for (int i = 0; i < var.nx; i++) {
    bool bounded = (i == 1 || i == 3 || (var.nx == 6 && i == 4));
    bool fin = std::isfinite(c.x_lo[i]) && std::isfinite(c.x_hi[i]);
    bool free_ = std::isinf(c.x_lo[i]) && c.x_lo[i] < 0 && std::isinf(c.x_hi[i]) && c.x_hi[i] > 0;
    if (bounded ? !(fin && c.x_lo[i] < c.x_hi[i]) : !free_) return MPCB_E_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
    cudaGetLastError();
    return MPCB_E_NODEVICE;
  }
  mpcb_handle *h = new (std::nothrow) mpcb_handle();
  if (!h) return MPCB_E_NOMEM;
  memset(h, 0, sizeof *h);
  h->cfg = c;
  h->var = var;
  if (!cuda_ok(cudaGetDevice(&h->device), "cudaGetDevice")) { mpcb_destroy(h); return MPCB_E_CUDA; }

  KParams &k = h->kp;
  memset(&k, 0, sizeof k);
  const int M = c.obs_mode == MPCB_OBS_NONE ? 0 : c.M;
  k.obs_input = c.obs_input;
  k.ref_mode = c.ref_mode;
  k.cbf_g1 = 1.0 - c.cbf_gamma;
  k.N = c.N; k.obs_mode = c.obs_mode; k.du0_cost = c.du0_cost; k.init_mode = c.init_mode; k.max_iter = c.max_iter;
  k.rate_ctrl[0] = c.rate_ctrl[0]; k.rate_ctrl[1] = c.rate_ctrl[1];
  const int n_obs_st = (c.obs_mode == MPCB_OBS_ELLIPSE || c.obs_mode == MPCB_OBS_DCBF) ? c.N : (c.obs_mode == MPCB_OBS_SQRT ? c.N + 1 : 0);
  k.n_eq = var.nx * (c.N + 1) + c.n_rate * (c.N - 1) + M * n_obs_st;
  k.n_bm = 4 * c.N + 2 * var.nbx * (c.N + 1) + 2 * c.n_rate * (c.N - 1) + M * n_obs_st;
  k.T = c.T;
  for (int i = 0; i < 6; i++) { k.Q[i] = c.Q[i]; k.x_lo[i] = relax_lo(c.x_lo[i], c.bound_relax); k.x_hi[i] = relax_hi(c.x_hi[i], c.bound_relax); }
  for (int i = 0; i < 2; i++) {
    k.R[i] = c.R[i]; k.DR[i] = c.DR[i];
    k.rate_lo[i] = relax_lo(c.rate_lo[i], c.bound_relax); k.rate_hi[i] = relax_hi(c.rate_hi[i], c.bound_relax);
    k.u_lo[i] = relax_lo(c.u_lo[i], c.bound_relax); k.u_hi[i] = relax_hi(c.u_hi[i], c.bound_relax);
  }
  k.obs_lo = relax_lo(c.obs_lo, c.bound_relax);
  k.ego_hl = c.ego_hl; k.ego_hw = c.ego_hw; k.safe_l = c.safe_l; k.safe_w = c.safe_w; k.dyn_sx = c.dyn_sx; k.dyn_sy = c.dyn_sy;
  k.Veh_l = c.Veh_l; k.lf = c.Veh_lf; k.lr = c.Veh_lr; k.m = c.Veh_m; k.Iz = c.Veh_Iz;
  k.aopt_f = c.aopt_f; k.aopt_r = c.aopt_r; k.Fymax_f = c.Fymax_f; k.Fymax_r = c.Fymax_r;
  k.tol = c.tol; k.mu_init = c.mu_init;

  h->smem = var.smem_bytes(c.N) * var.warps;
  if (const char *pad = getenv("MPCB_SMEM_PAD")) h->smem += (size_t)atoi(pad);  // tuning knob: lowers occupancy
  cudaDeviceProp prop;
  if (!cuda_ok(cudaGetDeviceProperties(&prop, h->device), "cudaGetDeviceProperties")) { mpcb_destroy(h); return MPCB_E_CUDA; }
  if (h->smem > (size_t)prop.sharedMemPerBlockOptin) { mpcb_destroy(h); return MPCB_E_ARG; }
  if (!cuda_ok(cudaFuncSetAttribute(var.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem), "cudaFuncSetAttribute")) {
    mpcb_destroy(h);
    return MPCB_E_CUDA;
  }
  cudaFuncAttributes fa;
  if (!cuda_ok(cudaFuncGetAttributes(&fa, var.kernel), "cudaFuncGetAttributes")) { mpcb_destroy(h); return MPCB_E_CUDA; }
  int bps = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, var.kernel, 32 * var.warps, h->smem);
  h->info.block = 32 * var.warps;
  h->info.smem_bytes = (int32_t)h->smem;
  h->info.regs_per_thread = fa.numRegs;
  h->info.blocks_per_sm = bps;
  h->info.num_sms = prop.multiProcessorCount;
  if (var.slab_doubles) {
    if (bps < 1) { mpcb_destroy(h); return MPCB_E_ARG; }
    h->persistent_grid = bps * prop.multiProcessorCount;
    size_t bytes = (size_t)h->persistent_grid * var.warps * var.slab_doubles * sizeof(double);
    if (!cuda_ok(cudaMalloc(&h->d_slab, bytes), "cudaMalloc slab") || !cuda_ok(cudaMalloc(&h->d_counter, sizeof(int)), "cudaMalloc counter")) {
      mpcb_destroy(h);
      return MPCB_E_NOMEM;
    }
    cudaMemset(h->d_slab, 0, bytes);
  }
  h->main_info = h->info;
  if (var.lat_kernel && !getenv("MPCB_NO_LATENCY_VARIANT")) {
    // small-batch sibling: one warp per block, everything in shared memory
    h->lat_smem = var.lat_smem_bytes(c.N);
    if (const char *pad = getenv("MPCB_LAT_SMEM_PAD")) h->lat_smem += (size_t)atoi(pad);  // tuning knob: lowers occupancy
    int lb = 0;
    if (h->lat_smem <= (size_t)prop.sharedMemPerBlockOptin &&
        cudaFuncSetAttribute(var.lat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->lat_smem) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&lb, var.lat_kernel, 32 * var.lat_warps, h->lat_smem) == cudaSuccess && lb >= 1) {
      h->lat_grid = lb * prop.multiProcessorCount;
      cudaFuncAttributes la;
      h->lat_info = h->info;
      if (cudaFuncGetAttributes(&la, var.lat_kernel) == cudaSuccess) h->lat_info.regs_per_thread = la.numRegs;
      h->lat_info.block = 32 * var.lat_warps;
      h->lat_info.smem_bytes = (int32_t)h->lat_smem;
      h->lat_info.blocks_per_sm = lb;
      if (!h->d_counter && !cuda_ok(cudaMalloc(&h->d_counter, sizeof(int)), "cudaMalloc counter")) { mpcb_destroy(h); return MPCB_E_NOMEM; }
    } else {
      cudaGetLastError();
    }
  }
  // ---- second engine (one scenario per lane).  cfg.engine, overridden by the environment for A/B runs: MPCB_ENGINE=warp
  // disables it, =lane forces it for every batch size, MPCB_LANE_MIN_B moves AUTO's switch-over.  AUTO uses it only for
  // the families without obstacle rows (it measures 19 % faster there at 100k scenarios and 2x slower on the CBF families,
  // whose iteration counts and line searches diverge between the lanes of a warp)
  {
    const char *eng = getenv("MPCB_ENGINE");
    int engine = c.engine;
    if (eng && !strcmp(eng, "warp")) engine = MPCB_ENGINE_WARP;
    if (eng && !strcmp(eng, "lane")) engine = MPCB_ENGINE_LANE;
    if (c.integrator == MPCB_INTEGRATOR_RK4) {  // only the lane engine integrates with Runge-Kutta
      if (engine == MPCB_ENGINE_WARP || c.restoration) { mpcb_destroy(h); return MPCB_E_ARG; }
      engine = MPCB_ENGINE_LANE;
    }
    LaneVariant lv;
    const bool has_lane = select_lane_variant(c, lv);
    if (engine == MPCB_ENGINE_LANE && !has_lane) { mpcb_destroy(h); return MPCB_E_ARG; }
    const bool rows = c.obs_mode != MPCB_OBS_NONE;
    if (has_lane && (engine == MPCB_ENGINE_LANE || (engine == MPCB_ENGINE_AUTO && !rows))) {
      int lb = 0;
      cudaFuncAttributes la;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&lb, lv.kernel, lv.block, 0) == cudaSuccess && lb >= 1 &&
          cudaFuncGetAttributes(&la, lv.kernel) == cudaSuccess) {
        if (const char *cap = getenv("MPCB_LANE_BLOCKS_PER_SM")) { int v_ = atoi(cap); if (v_ >= 1 && v_ < lb) lb = v_; }
        h->lane = lv;
        h->lane_grid = lb * prop.multiProcessorCount;
        h->lane_nslot = (size_t)h->lane_grid * lv.block;
        h->lane_forced = engine == MPCB_ENGINE_LANE;
        h->lane_min_B = engine == MPCB_ENGINE_LANE ? 1 : (getenv("MPCB_LANE_MIN_B") ? atoi(getenv("MPCB_LANE_MIN_B")) : MPCB_LANE_MIN_B_DEFAULT);
        size_t bytes = h->lane_nslot * lv.slot_doubles(c.N) * sizeof(double);
        if (!cuda_ok(cudaMalloc(&h->d_lane_ws, bytes), "cudaMalloc lane workspace")) { mpcb_destroy(h); return MPCB_E_NOMEM; }
        if (!h->d_counter && !cuda_ok(cudaMalloc(&h->d_counter, sizeof(int)), "cudaMalloc counter")) { mpcb_destroy(h); return MPCB_E_NOMEM; }
        h->lane_info = h->info;
        h->lane_info.block = lv.block;
        h->lane_info.smem_bytes = 0;
        h->lane_info.regs_per_thread = la.numRegs;
        h->lane_info.blocks_per_sm = lb;
      } else {
        cudaGetLastError();
      }
    }
  }
#ifdef MPCB_DEBUG_SLOTS
  if (!cuda_ok(cudaMalloc(&h->d_debug_errors, sizeof(int)), "cudaMalloc debug counter")) { mpcb_destroy(h); return MPCB_E_NOMEM; }
  cudaMemset(h->d_debug_errors, 0, sizeof(int));
  k.debug_errors = h->d_debug_errors;
  k.debug_selftest = getenv("MPCB_DEBUG_SLOTS_SELFTEST") ? 1 : 0;
#endif
  k.resto_max_calls = c.resto_max_calls;
  k.restoration = (c.restoration && var.rs_inline) ? 1 : 0;
  if (c.restoration && !var.rs_inline && var.resto_kernel) {
    h->resto_smem = var.resto_smem_bytes(c.N) * var.resto_warps;
    int rb = 0;
    if (h->resto_smem <= (size_t)prop.sharedMemPerBlockOptin &&
        cudaFuncSetAttribute(var.resto_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->resto_smem) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&rb, var.resto_kernel, 32 * var.resto_warps, h->resto_smem) == cudaSuccess && rb >= 1) {
      h->resto_grid = rb * prop.multiProcessorCount;
      // the restoration pass overlaps the main kernel's last wave: its own slab
      size_t need = (size_t)h->resto_grid * var.resto_warps * var.resto_slab_doubles * sizeof(double);
      if (!cuda_ok(cudaMalloc(&h->d_resto_slab, need), "cudaMalloc restoration slab")) { mpcb_destroy(h); return MPCB_E_NOMEM; }
      cudaMemset(h->d_resto_slab, 0, need);
      if (!cuda_ok(cudaMalloc(&h->d_resto_count, 4 * sizeof(int)), "cudaMalloc resto sync")) { mpcb_destroy(h); return MPCB_E_NOMEM; }
      if (!h->d_counter && !cuda_ok(cudaMalloc(&h->d_counter, sizeof(int)), "cudaMalloc counter")) { mpcb_destroy(h); return MPCB_E_NOMEM; }
    } else {
      cudaGetLastError();
      mpcb_destroy(h);
      return MPCB_E_ARG;  // restoration requested but its kernel does not fit this device / horizon
    }
  }
  if (!cuda_ok(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking), "cudaStreamCreate")) { mpcb_destroy(h); return MPCB_E_CUDA; }
  if (!cuda_ok(cudaEventCreateWithFlags(&h->last_done, cudaEventDisableTiming), "cudaEventCreate")) { mpcb_destroy(h); return MPCB_E_CUDA; }
  *out = h;
  return MPCB_OK;
}

static void free_bufs(mpcb_handle *h) {
  cudaFree(h->d_x0); cudaFree(h->d_xs); cudaFree(h->d_obs); cudaFree(h->d_zin); cudaFree(h->d_u0);
  cudaFree(h->d_cost); cudaFree(h->d_z); cudaFree(h->d_lam); cudaFree(h->d_status); cudaFree(h->d_iters);
  cudaFree(h->d_lamg); cudaFree(h->d_lamx);
  h->d_x0 = h->d_xs = h->d_obs = h->d_zin = h->d_u0 = h->d_cost = h->d_z = h->d_lam = h->d_lamg = h->d_lamx = nullptr;
  h->d_status = h->d_iters = nullptr;
  h->cap_B = 0;
}

void mpcb_destroy(mpcb_handle *h) {
  if (!h) return;
  if (h->stream) cudaStreamSynchronize(h->stream);  // a batch submitted with mpcb_submit_batch_host may still be in flight
  if (h->has_last) cudaEventSynchronize(h->last_done);  // ... or a solve on a caller's stream
  if (h->last_done) cudaEventDestroy(h->last_done);
  free_bufs(h);
  cudaFree(h->d_slab);
  cudaFree(h->d_counter);
  cudaFree(h->d_resto_list);
  cudaFree(h->d_resto_count);
  cudaFree(h->d_resto_slab);
  cudaFree(h->d_lane_ws);
  cudaFree(h->d_debug_errors);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int mpcb_solve_batch(mpcb_handle *h, int B, const double *x0, const double *xs, const double *obs,
                     const double *z_init, double *u0, double *cost, int32_t *status, int32_t *iters,
                     double *z_out, double *lam_out, void *stream) {
  if (!h || B < 0) return MPCB_E_ARG;
  if (B == 0) return MPCB_OK;  // empty batch: nothing to read or write
  if (!x0 || !xs || !u0 || !cost || !status || !iters) return MPCB_E_ARG;
  if (h->cfg.obs_mode != MPCB_OBS_NONE && !obs) return MPCB_E_ARG;
  if (h->kp.order && h->order_n != B) return MPCB_E_ARG;  // an installed order is a permutation of exactly this batch
  KParams k = h->kp;
  k.B = B;
  k.x0 = x0; k.xs = xs; k.obs = obs; k.z_init = z_init;
  k.u0 = u0; k.cost = cost; k.status = status; k.iters = iters; k.z_out = z_out; k.lam_out = lam_out;
  if (!h->duals_on_host) { k.lam_g_out = h->lam_g_out; k.lam_x_out = h->lam_x_out; }
  else if (stream == (void *)h->stream && h->d_lamg) { k.lam_g_out = h->lam_g_out ? h->d_lamg : nullptr; k.lam_x_out = h->lam_x_out ? h->d_lamx : nullptr; }
  // The queue head and the slab are per handle: a solve issued on a different stream than the previous one
  // queues behind it on the device (same stream: already ordered).
  if (h->has_last && h->last_stream != (cudaStream_t)stream &&
      !cuda_ok(cudaStreamWaitEvent((cudaStream_t)stream, h->last_done, 0), "cudaStreamWaitEvent")) return MPCB_E_CUDA;
  const bool with_resto = h->resto_grid > 0;
  if (with_resto) {
    if (B > h->resto_cap) {  // grows the list of scenarios for the second pass (allocation; mpcb_reserve sizes it up front)
      if (h->has_last) cudaEventSynchronize(h->last_done);
      cudaFree(h->d_resto_list);
      h->d_resto_list = nullptr;
      h->resto_cap = 0;
      if (!cuda_ok(cudaMalloc(&h->d_resto_list, (size_t)B * sizeof(int32_t)), "cudaMalloc resto list")) return MPCB_E_NOMEM;
      h->resto_cap = B;
    }
    if (!cuda_ok(cudaMemsetAsync(h->d_resto_count, 0, 4 * sizeof(int), (cudaStream_t)stream), "resto queue reset") ||
        !cuda_ok(cudaMemsetAsync(h->d_resto_list, 0xFF, (size_t)B * sizeof(int32_t), (cudaStream_t)stream), "resto list reset")) return MPCB_E_CUDA;
    k.resto_list = h->d_resto_list;
    k.resto_sync = h->d_resto_count;
  }
  int grid = B;
  cudaError_t e;
  static const int lat_force = getenv("MPCB_LAT_MAX_B") ? atoi(getenv("MPCB_LAT_MAX_B")) : 0;  // tuning knob
  const bool small_batch = h->lat_grid && (B <= h->lat_grid * h->var.lat_warps || B <= lat_force);
  const bool lane_engine = h->d_lane_ws && B >= h->lane_min_B && (!k.trace || h->lane_forced);  // (the lane kernel writes no trace)
  int main_warps = 0;
  if (lane_engine) {
    // one scenario per lane: the lanes of the resident grid pull scenarios from the queue
    // the lanes of the resident grid pull scenarios from the queue; without rows every scenario takes the same 17-18
    // iterations, the batch passes in ceil(need / lane_grid) waves whose duration depends only weakly on how full
    // they are (20.9 ms at 56,832 lanes, 16.8 ms at 32,768), and equally full waves beat a full one plus a remainder
    const int need = (B + h->lane.block - 1) / h->lane.block;
    grid = need < h->lane_grid ? need : h->lane_grid;
    static const bool balance = !getenv("MPCB_LANE_NO_BALANCE");
    if (balance && need > h->lane_grid && (h->cfg.obs_mode == MPCB_OBS_NONE || h->cfg.M == 0)) {
      const int waves = (need + h->lane_grid - 1) / h->lane_grid;
      grid = (need + waves - 1) / waves;
    }
    k.slab = nullptr;
    k.counter = h->d_counter;
    if (!cuda_ok(cudaMemsetAsync(h->d_counter, 0, sizeof(int), (cudaStream_t)stream), "queue reset")) return MPCB_E_CUDA;
    e = h->lane.launch(k, h->d_lane_ws, (size_t)grid * h->lane.block, grid, (cudaStream_t)stream);
    main_warps = grid * (h->lane.block / 32);
  } else if (small_batch) {
    grid = (B + h->var.lat_warps - 1) / h->var.lat_warps;
    if (grid > h->lat_grid) grid = h->lat_grid;
    // the batch fits the SMs in one wave of the small-batch kernel (one warp per block, all state in shared memory)
    k.slab = nullptr;
    k.counter = h->d_counter;
    if (!cuda_ok(cudaMemsetAsync(h->d_counter, 0, sizeof(int), (cudaStream_t)stream), "queue reset")) return MPCB_E_CUDA;
    e = h->var.lat_launch(k, grid, h->lat_smem, (cudaStream_t)stream);
  } else {
    if (h->persistent_grid) {
      const int need = (B + h->var.warps - 1) / h->var.warps;
      grid = need < h->persistent_grid ? need : h->persistent_grid;
      k.slab = h->d_slab;
      k.counter = h->d_counter;
      if (!cuda_ok(cudaMemsetAsync(h->d_counter, 0, sizeof(int), (cudaStream_t)stream), "queue reset")) return MPCB_E_CUDA;
    }
    e = h->var.launch(k, grid, h->smem, (cudaStream_t)stream);
  }
  if (!cuda_ok(e, "solve_kernel launch")) return MPCB_E_CUDA;
  if (with_resto) {
    // second pass: the listed scenarios again, from their start points, with the restoration-capable kernel.  Its batch
    // size is known only on the device; blocks that find the queue empty retire at once.
    KParams k2 = k;
    k2.restoration = 1;
    k2.slab = h->d_resto_slab;
    k2.order = nullptr;
    k2.resto_consumer = 1;
    k2.main_warps = lane_engine ? main_warps : grid * (small_batch ? h->var.lat_warps : h->var.warps);
    k2.trace = nullptr;
    int g2 = (B + h->var.resto_warps - 1) / h->var.resto_warps;
    if (g2 > h->resto_grid) g2 = h->resto_grid;
    e = h->var.resto_launch(k2, g2, h->resto_smem, (cudaStream_t)stream);
    if (!cuda_ok(e, "restoration kernel launch")) return MPCB_E_CUDA;
  }
  if (!cuda_ok(cudaEventRecord(h->last_done, (cudaStream_t)stream), "cudaEventRecord")) return MPCB_E_CUDA;
  h->has_last = true;
  h->last_stream = (cudaStream_t)stream;
  {
    const int64_t launches = h->info.launches + (with_resto ? 2 : 1);
    h->info = lane_engine ? h->lane_info : (small_batch ? h->lat_info : h->main_info);
    h->info.grid = grid;
    h->info.launches = launches;
  }
  return MPCB_OK;
}

int mpcb_wait(mpcb_handle *h) {
  if (!h) return MPCB_E_ARG;
  return cuda_ok(cudaStreamSynchronize(h->stream), "solve (stream sync)") ? MPCB_OK : MPCB_E_CUDA;
}

int mpcb_solve_batch_host(mpcb_handle *h, int B, const double *x0, const double *xs, const double *obs,
                          const double *z_init, double *u0, double *cost, int32_t *status, int32_t *iters,
                          double *z_out, double *lam_out) {
  int rc = mpcb_submit_batch_host(h, B, x0, xs, obs, z_init, u0, cost, status, iters, z_out, lam_out);
  return rc != MPCB_OK ? rc : mpcb_wait(h);
}

int mpcb_n_g(const mpcb_cfg *c) {
  if (!c) return 0;
  const int nx = mpcb_nx(c), M = c->obs_mode == MPCB_OBS_NONE ? 0 : c->M;
  const int n_obs_st = (c->obs_mode == MPCB_OBS_ELLIPSE || c->obs_mode == MPCB_OBS_DCBF) ? c->N : (c->obs_mode == MPCB_OBS_SQRT ? c->N + 1 : 0);
  return nx * (c->N + 1) + c->n_rate * (c->N - 1) + M * n_obs_st;
}

int mpcb_reserve(mpcb_handle *h, int B) {
  if (!h || B < 0) return MPCB_E_ARG;
  if (B <= h->cap_B) return MPCB_OK;
  const int nx = h->var.nx, N = h->cfg.N;
  const int M = h->cfg.obs_mode == MPCB_OBS_NONE ? 0 : h->cfg.M;
  const size_t nv = 2 * (size_t)N + (size_t)nx * (N + 1);
  const size_t so = (size_t)M * (h->cfg.obs_input != MPCB_OBS_TRAJECTORY ? 1 : (N + 1)) * 6;
  const size_t sxs = h->cfg.ref_mode == MPCB_REF_TRAJECTORY ? (size_t)nx * N : (size_t)nx;
  if (!cuda_ok(cudaStreamSynchronize(h->stream), "stream sync before regrowing buffers")) return MPCB_E_CUDA;
  free_bufs(h);
  size_t b = (size_t)B;
  bool ok = cuda_ok(cudaMalloc(&h->d_x0, b * nx * 8), "cudaMalloc") && cuda_ok(cudaMalloc(&h->d_xs, b * sxs * 8), "cudaMalloc") &&
            cuda_ok(cudaMalloc(&h->d_obs, b * (so ? so : 1) * 8), "cudaMalloc") && cuda_ok(cudaMalloc(&h->d_zin, b * nv * 8), "cudaMalloc") &&
            cuda_ok(cudaMalloc(&h->d_u0, b * 2 * 8), "cudaMalloc") && cuda_ok(cudaMalloc(&h->d_cost, b * 8), "cudaMalloc") &&
            cuda_ok(cudaMalloc(&h->d_z, b * nv * 8), "cudaMalloc") && cuda_ok(cudaMalloc(&h->d_lam, b * nx * (N + 1) * 8), "cudaMalloc") &&
            cuda_ok(cudaMalloc(&h->d_lamg, b * mpcb_n_g(&h->cfg) * 8), "cudaMalloc") && cuda_ok(cudaMalloc(&h->d_lamx, b * nv * 8), "cudaMalloc") &&
            cuda_ok(cudaMalloc(&h->d_status, b * 4), "cudaMalloc") && cuda_ok(cudaMalloc(&h->d_iters, b * 4), "cudaMalloc");
  if (!ok) { free_bufs(h); return MPCB_E_NOMEM; }
  h->cap_B = B;
  if (h->resto_grid > 0 && B > h->resto_cap) {
    cudaFree(h->d_resto_list);
    h->d_resto_list = nullptr;
    h->resto_cap = 0;
    if (!cuda_ok(cudaMalloc(&h->d_resto_list, (size_t)B * sizeof(int32_t)), "cudaMalloc resto list")) return MPCB_E_NOMEM;
    h->resto_cap = B;
  }
  return MPCB_OK;
}

int mpcb_set_dual_outputs(mpcb_handle *h, double *lam_g, double *lam_x, int on_host) {
  if (!h) return MPCB_E_ARG;
  h->lam_g_out = lam_g;
  h->lam_x_out = lam_x;
  h->duals_on_host = on_host ? 1 : 0;
  return MPCB_OK;
}

int mpcb_submit_batch_host(mpcb_handle *h, int B, const double *x0, const double *xs, const double *obs,
                           const double *z_init, double *u0, double *cost, int32_t *status, int32_t *iters,
                           double *z_out, double *lam_out) {
  if (!h || B < 0) return MPCB_E_ARG;
  if (B == 0) return MPCB_OK;
  if (!x0 || !xs || !u0 || !cost || !status || !iters) return MPCB_E_ARG;
  const int nx = h->var.nx, N = h->cfg.N;
  const int M = h->cfg.obs_mode == MPCB_OBS_NONE ? 0 : h->cfg.M;
  const size_t nv = 2 * (size_t)N + (size_t)nx * (N + 1);
  const size_t so = (size_t)M * (h->cfg.obs_input != MPCB_OBS_TRAJECTORY ? 1 : (N + 1)) * 6;
  const size_t sxs = h->cfg.ref_mode == MPCB_REF_TRAJECTORY ? (size_t)nx * N : (size_t)nx;
  if (M > 0 && !obs) return MPCB_E_ARG;
  if (B > h->cap_B) {  // grows the staging buffers (allocation); call mpcb_reserve up front to keep this path allocation-free
    int rc = mpcb_reserve(h, B);
    if (rc != MPCB_OK) return rc;
  }
  cudaStream_t st = h->stream;
  size_t b = (size_t)B;
  bool ok = cuda_ok(cudaMemcpyAsync(h->d_x0, x0, b * nx * 8, cudaMemcpyHostToDevice, st), "H2D x0") &&
            cuda_ok(cudaMemcpyAsync(h->d_xs, xs, b * sxs * 8, cudaMemcpyHostToDevice, st), "H2D xs");
  if (ok && so) ok = cuda_ok(cudaMemcpyAsync(h->d_obs, obs, b * so * 8, cudaMemcpyHostToDevice, st), "H2D obs");
  if (ok && z_init) ok = cuda_ok(cudaMemcpyAsync(h->d_zin, z_init, b * nv * 8, cudaMemcpyHostToDevice, st), "H2D z_init");
  if (!ok) return MPCB_E_CUDA;
  int rc = mpcb_solve_batch(h, B, h->d_x0, h->d_xs, so ? h->d_obs : nullptr, z_init ? h->d_zin : nullptr, h->d_u0, h->d_cost,
                            h->d_status, h->d_iters, z_out ? h->d_z : nullptr, lam_out ? h->d_lam : nullptr, (void *)st);
  if (rc != MPCB_OK) return rc;
  ok = cuda_ok(cudaMemcpyAsync(u0, h->d_u0, b * 2 * 8, cudaMemcpyDeviceToHost, st), "D2H u0") &&
       cuda_ok(cudaMemcpyAsync(cost, h->d_cost, b * 8, cudaMemcpyDeviceToHost, st), "D2H cost") &&
       cuda_ok(cudaMemcpyAsync(status, h->d_status, b * 4, cudaMemcpyDeviceToHost, st), "D2H status") &&
       cuda_ok(cudaMemcpyAsync(iters, h->d_iters, b * 4, cudaMemcpyDeviceToHost, st), "D2H iters");
  if (ok && z_out) ok = cuda_ok(cudaMemcpyAsync(z_out, h->d_z, b * nv * 8, cudaMemcpyDeviceToHost, st), "D2H z");
  if (ok && lam_out) ok = cuda_ok(cudaMemcpyAsync(lam_out, h->d_lam, b * nx * (N + 1) * 8, cudaMemcpyDeviceToHost, st), "D2H lam");
  if (ok && h->duals_on_host && h->lam_g_out) ok = cuda_ok(cudaMemcpyAsync(h->lam_g_out, h->d_lamg, b * mpcb_n_g(&h->cfg) * 8, cudaMemcpyDeviceToHost, st), "D2H lam_g");
  if (ok && h->duals_on_host && h->lam_x_out) ok = cuda_ok(cudaMemcpyAsync(h->lam_x_out, h->d_lamx, b * nv * 8, cudaMemcpyDeviceToHost, st), "D2H lam_x");
  return ok ? MPCB_OK : MPCB_E_CUDA;
}

int mpcb_set_order(mpcb_handle *h, const int32_t *order, int n) {
  if (!h || (order && n <= 0)) return MPCB_E_ARG;
  h->kp.order = order;
  h->order_n = order ? n : 0;
  return MPCB_OK;
}

int mpcb_set_trace_buffer(mpcb_handle *h, double *trace, int rows) {
  if (!h || rows < 0) return MPCB_E_ARG;
  if (h->lane_forced && rows > 0) return MPCB_E_ARG;  // the per-iteration log exists in the warp kernels only
  h->kp.trace = rows > 0 ? trace : nullptr;
  h->kp.trace_rows = rows;
  return MPCB_OK;
}

int mpcb_debug_slot_errors(mpcb_handle *h, int *count) {
  if (!h || !count) return MPCB_E_ARG;
  *count = -1;  // not a debug build
#ifdef MPCB_DEBUG_SLOTS
  if (h->has_last) cudaEventSynchronize(h->last_done);
  if (!cuda_ok(cudaMemcpy(count, h->d_debug_errors, sizeof(int), cudaMemcpyDeviceToHost), "debug counter")) return MPCB_E_CUDA;
#endif
  return MPCB_OK;
}

int mpcb_get_launch_info(mpcb_handle *h, mpcb_launch_info *out) {
  if (!h || !out) return MPCB_E_ARG;
  *out = h->info;
  return MPCB_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------
// closed-loop helper: plant Euler step + warm-start shift (PKG/main_cbf_kin_c_sim.py:16-26)
// ---------------------------------------------------------------------------------------
namespace {

template <int NX>
__global__ void shift_kernel(KParams p, int B, double *x0, double *z) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int N = p.N, nv = 2 * N + NX * (N + 1);
  double *zz = z + (size_t)b * nv;
  double *xx = x0 + (size_t)b * NX;
  // plant step with the first control
  double x[NX], u[2] = {zz[0], zz[1]}, f[NX];
#pragma unroll
  for (int i = 0; i < NX; i++) x[i] = xx[i];
  if constexpr (NX == 4) {
    double s_, c_, t_;
    d_trig(x[2], u[0], &s_, &c_, &t_);
    f[0] = x[3] * c_; f[1] = x[3] * s_; f[2] = x[3] * t_ / p.Veh_l; f[3] = u[1];
  } else {
    dyn_f(x, u, p, f);
  }
  // shifted copies held in registers before anything is overwritten
  constexpr int PER = (2 * MPCB_NMAX + 6 * (MPCB_NMAX + 1) + 31) / 32;
  double v[PER];
#pragma unroll
  for (int r = 0; r < PER; r++) {
    int idx = lane + 32 * r;
    double val = 0;
    if (idx < 2 * N) {
      int src = idx + 2 < 2 * N ? idx + 2 : idx;  // repeat the last control row
      val = zz[src];
    } else if (idx < nv) {
      int j = idx - 2 * N;
      int src = j + NX < NX * (N + 1) ? j + NX : j;  // repeat the last state row
      val = zz[2 * N + src];
    }
    v[r] = val;
  }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < PER; r++) {
    int idx = lane + 32 * r;
    if (idx < nv) zz[idx] = v[r];
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NX; i++) xx[i] = x[i] + p.T * f[i];
  }
}

}  // namespace

extern "C" int mpcb_shift_batch(mpcb_handle *h, int B, double *x0, double *z, void *stream) {
  if (!h || B < 0 || !x0 || !z) return MPCB_E_ARG;
  if (B == 0) return MPCB_OK;
  KParams k = h->kp;
  const int wpb = 4;
  int grid = (B + wpb - 1) / wpb;
  if (h->cfg.model == MPCB_MODEL_KIN) shift_kernel<4><<<grid, 32 * wpb, 0, (cudaStream_t)stream>>>(k, B, x0, z);
  else shift_kernel<6><<<grid, 32 * wpb, 0, (cudaStream_t)stream>>>(k, B, x0, z);
  if (!cuda_ok(cudaGetLastError(), "shift_kernel launch")) return MPCB_E_CUDA;
  h->info.launches++;
  return MPCB_OK;
}

// ---------------------------------------------------------------------------------------
// batched obs_prediction (PKG/Obs_prediction.py:3-40) and the mains' obstacle update
// (PKG/main_cbf_kin_c_sim_pre.py:106: the obstacle becomes row 1 of its own prediction), one thread per obstacle
// ---------------------------------------------------------------------------------------
namespace {

__global__ void obs_predict_kernel(int n, int N, double dt, double *state, double *traj, int advance) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double *o = state + (size_t)i * 6;
  const double th = o[2], v = o[3], l = o[4], w = o[5];
  // x + v cos(theta) dt, step by step, in the reference's operation order (:27-28)
  const double dx = v * cos(th) * dt, dy = v * sin(th) * dt;
  double x = o[0], y = o[1];
  if (traj) {
    double *t = traj + (size_t)i * (N + 1) * 6;
    for (int k = 0; k <= N; k++) {
      t[6 * k + 0] = x; t[6 * k + 1] = y; t[6 * k + 2] = th; t[6 * k + 3] = v; t[6 * k + 4] = l; t[6 * k + 5] = w;
      x = x + dx;
      y = y + dy;
    }
  }
  if (advance) { o[0] = o[0] + dx; o[1] = o[1] + dy; }
}

}  // namespace

extern "C" int mpcb_obs_prediction_batch(mpcb_handle *h, int n_obstacles, double *obs_state, double *traj, int advance, void *stream) {
  if (!h || n_obstacles < 0 || !obs_state || (!traj && !advance)) return MPCB_E_ARG;
  if (n_obstacles == 0) return MPCB_OK;
  obs_predict_kernel<<<(n_obstacles + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n_obstacles, h->cfg.N, h->cfg.T, obs_state, traj, advance);
  if (!cuda_ok(cudaGetLastError(), "obs_predict_kernel launch")) return MPCB_E_CUDA;
  h->info.launches++;
  return MPCB_OK;
}

// ---------------------------------------------------------------------------------------
// batched RefPathGenerator (PKG/RefPathGenerator.py:9-59): one thread per scenario
// ---------------------------------------------------------------------------------------
namespace {

// Every floating-point step below is written with explicit round-to-nearest intrinsics so that
// no multiply-add is contracted: the index arithmetic (nearest-point walk, linspace, truncation
// to int) has to reproduce numpy's results exactly, not approximately.
__global__ void ref_traj_kernel(int B, int N, double T_horizon, const double *x0, const double *xs, const double *path_x0,
                                int32_t *last_idx, double aa, double *ref, double *stage) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double ex = x0[4 * b + 0], ey = x0[4 * b + 1], ev = x0[4 * b + 3];
  const double tx = xs[4 * b + 0], ty = xs[4 * b + 1], tphi = xs[4 * b + 2], tv = xs[4 * b + 3];
  // define_ref_path (:9-24): np.arange(x_start, xs.x +- 1, +-1) -> element i is x_start + i*step
  const double start = path_x0[b];
  const double step = tx > start ? 1.0 : -1.0;
  const double stop = __dadd_rn(tx, step);
  int len = (int)ceil(__ddiv_rn(__dadd_rn(stop, -start), step));
  if (len < 1) len = 1;
  // numpy fills arange as start + i*delta with delta = (start + step) - start, which is not exactly
  // +-1 when start + step rounds
  const double dlt = __dadd_rn(__dadd_rn(start, step), -start);
  // find_ref_traj (:27-59)
  const double pv = __dadd_rn(__dmul_rn(0.5, ev), __dmul_rn(0.5, tv));
  const int pidx = (int)__dmul_rn(pv, T_horizon);  // step_x = 1
  const int last = last_idx[b];
  const int lo = max(0, last - 5), hi = min(len, last + pidx);
  int mi = lo;
  double best = INFINITY;
  for (int i = lo; i < hi; i++) {
    const double dx = __dadd_rn(__dadd_rn(start, __dmul_rn((double)i, dlt)), -ex), dy = __dadd_rn(ty, -ey);
    const double d = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    if (d < best) { best = d; mi = i; } else break;
  }
  last_idx[b] = mi;
  // idx = np.clip(np.linspace(mi, mi + pidx, N+1), 0, len-1).astype(int)
  const double a0 = (double)mi, a1 = (double)(mi + pidx);
  const double ls = __ddiv_rn(__dadd_rn(a1, -a0), (double)N);
  for (int i = 0; i <= N; i++) {
    double v = i == N ? a1 : __dadd_rn(__dmul_rn((double)i, ls), a0);
    v = fmin(fmax(v, 0.0), (double)(len - 1));
    const int id = (int)v;
    const double rx = __dadd_rn(start, __dmul_rn((double)id, dlt));
    if (ref) {
      double *r = ref + ((size_t)b * (N + 1) + i) * 4;
      r[0] = rx; r[1] = ty; r[2] = tphi; r[3] = tv;
    }
    if (stage && i >= 1) {  // ref_X of stage i-1 = aa*ref_state[i] + (1-aa)*xs   (PKG/MPC_CBF_optimize_kin.py:196)
      double *q = stage + ((size_t)b * N + (i - 1)) * 4;
      const double w1 = __dadd_rn(1.0, -aa);
      q[0] = __dadd_rn(__dmul_rn(aa, rx), __dmul_rn(w1, tx));
      q[1] = __dadd_rn(__dmul_rn(aa, ty), __dmul_rn(w1, ty));
      q[2] = __dadd_rn(__dmul_rn(aa, tphi), __dmul_rn(w1, tphi));
      q[3] = __dadd_rn(__dmul_rn(aa, tv), __dmul_rn(w1, tv));
    }
  }
}

}  // namespace

extern "C" int mpcb_ref_traj_batch(mpcb_handle *h, int B, double T_horizon, const double *x0, const double *xs, const double *path_x0,
                                   int32_t *last_idx, double aa, double *ref, double *stage_targets, void *stream) {
  if (!h || B < 0) return MPCB_E_ARG;
  if (B == 0) return MPCB_OK;
  if (!x0 || !xs || !path_x0 || !last_idx || (!ref && !stage_targets) || !(T_horizon > 0)) return MPCB_E_ARG;
  if (h->cfg.model != MPCB_MODEL_KIN) return MPCB_E_ARG;  // the reference builds references for the 4-state mains only
  if ((int)(T_horizon / h->cfg.T) != h->cfg.N) return MPCB_E_ARG;  // N_p = int(T_horizon/dt), RefPathGenerator.py:31
  ref_traj_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(B, h->cfg.N, T_horizon, x0, xs, path_x0, last_idx, aa, ref, stage_targets);
  if (!cuda_ok(cudaGetLastError(), "ref_traj_kernel launch")) return MPCB_E_CUDA;
  h->info.launches++;
  return MPCB_OK;
}

// ---------------------------------------------------------------------------------------
// FP64 FMA-pipe micro-benchmark: the roofline denominator of the solve kernel (SURVEY.md
// section 8d asks for a measured DFMA peak; MEASURED_PEAKS.json only holds HBM and bf16).
// ---------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a, double b) {
  double v[8];
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = (double)(threadIdx.x + i) * 1e-3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] = fma(v[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += v[i];
  if (s == 123.456) out[0] = s;  // never true; keeps the chain alive
}
}  // namespace

extern "C" int mpcb_fp64_peak_tflops(double *tflops) {
  if (!tflops) return MPCB_E_ARG;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return MPCB_E_NODEVICE; }
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double *d = nullptr;
  if (!cuda_ok(cudaMalloc(&d, 8), "cudaMalloc")) return MPCB_E_NOMEM;
  const int iters = 4096, grid = sms * 8, block = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    dfma_kernel<<<grid, block>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    if (!cuda_ok(cudaEventSynchronize(e1), "dfma_kernel")) { cudaFree(d); return MPCB_E_CUDA; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 32.0 * iters * (double)grid * block;
    double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops = best;
  return MPCB_OK;
}
