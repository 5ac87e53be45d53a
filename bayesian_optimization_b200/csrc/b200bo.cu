// b200bo.cu -- host driver + C ABI (include/b200bo.h) of the B200 GP-surrogate / acquisition engine.
// sm_100a only.  No CPU fallback: every entry point fails with B200BO_E_NODEVICE / B200BO_E_CUDA when the
// device is missing.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/b200bo.h"
#include "band_kernels.cuh"
#include "dgemm.cuh"
#include "fast_kernels.cuh"
#include "fast2_kernels.cuh"
#include "fast3_kernels.cuh"
#include "fast4_kernels.cuh"
#include "fast5_kernels.cuh"
#include "fast6_kernels.cuh"
#include "fit_kernels.cuh"
#include "assemble_kernels.cuh"
#include "grad_kernels.cuh"
#include "host_qr.h"
#include "oz_kernels.cuh"
#include "predict_kernels.cuh"
#include "trend_kernels.cuh"

using namespace b2;

static thread_local std::string g_err;
static int set_err(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CU_TRY(expr)                                                                         \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      char _b[512];                                                                          \
      snprintf(_b, sizeof _b, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return set_err(B200BO_E_CUDA, _b);                                                     \
    }                                                                                        \
  } while (0)

#define CHECK_ARG(cond, msg) \
  do {                       \
    if (!(cond)) return set_err(B200BO_E_ARG, msg); \
  } while (0)

namespace {

// (world, 2 q) int64 exchange block of the global arg-max: this rank's row <- [value bits | global index], other rows
// <- 0, so that ONE integer sum over the ranks (ncclAllReduce) delivers every rank's pairs bit for bit
__global__ void best_pairs_kernel(const double* __restrict__ best_val, const long long* __restrict__ best_idx, int q,
                                  long long offset, int rank, int world, long long* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= world * 2 * q) return;
  const int r = e / (2 * q), c = e % (2 * q);
  long long v = 0;
  if (r == rank) {
    if (c < q) v = __double_as_longlong(best_val[c]);
    else v = best_idx[c - q] >= 0 ? best_idx[c - q] + offset : -1;
  }
  out[e] = v;
}

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t reserve(size_t want) {
    if (want <= n) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e == cudaSuccess) n = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

// A replayable stretch of the factorisation (its kernels take pointers only): run eagerly the first time a
// configuration is seen, captured into a CUDA graph the second time, replayed from then on.  The L-BFGS loop calls
// factor() hundreds of times with the same shapes; the graph removes ~200 launches' worth of CPU time and most of the
// inter-kernel gaps of the latency-bound panel chain.
struct GraphSlot {
  cudaGraphExec_t exec = nullptr;
  const void* key[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int seen = 0, launches = 0;
  void drop() {
    if (exec) cudaGraphExecDestroy(exec);
    exec = nullptr;
    seen = 0;
  }
};

struct EvPool {
  std::vector<cudaEvent_t> ev;
  size_t used = 0;
  cudaEvent_t get() {
    if (used == ev.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      ev.push_back(e);
    }
    return ev[used++];
  }
  void reset() { used = 0; }
  void destroy() {
    for (auto e : ev) cudaEventDestroy(e);
    ev.clear();
  }
};

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

using GemmNT = GemmCore<64, 64, 32, 32, false, false, 3>;  // A (MxK), B (NxK)
using GemmNN = GemmCore<64, 64, 32, 32, false, true, 3>;   // A (MxK), B (KxN)
using GemmTN = GemmCore<64, 64, 32, 32, true, true, 3>;    // A (KxM), B (KxN)

}  // namespace

struct FastModel {  // a-priori rounding model of the tensor-core pass (build_fast_model)
  double a_max = 0, s2 = 0, fro = 0, l1_max = 0, gamma_l2 = 0, f_l2 = 0, b_max = 0, sd_r = 0;
  double dy = 0, du = 0;                 // half-widths of yhat and of u = (Ft^T rt - 1) / G
  double ds_abs[2] = {0, 0}, ds_rel[2] = {0, 0};  // mse: ds_abs + ds_rel sqrt(sum rt^2); [0] one product, [1] three
  double det_ds = 0;                     // deterministic worst case (reported only)
};
struct LastFast {  // the candidate set of the last tensor-core pass (b200bo_debug_fast_check)
  const double* xdev = nullptr;
  int64_t M = 0;
  int nprod = 0;
};
static const size_t PIN_BYTES = 32 * 1024;

struct b200bo_ctx {
  int device = 0, num_sms = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  int prec = B200BO_PREC_FP64;
  bool keepR = false;
  // training data
  int N = 0, D = 0, ld = 0;
  DevBuf<double> Xt, y, F, theta;
  // factorisation
  DevBuf<double> A, W, S, Rkeep, Dinv, Yt, Ft, rho, gamma, part, scal;
  DevBuf<int> status;
  bool factored = false;
  int corr = 0, mode = 0, trend = 0, estimate_trend = 1, n_theta = 0;
  double sigma2 = NAN, noise_var = 0, beta = 0, G = NAN, llf = -INFINITY, par_last = NAN;
  // linear / quadratic trend (p > 1): basis F, Ft = L^-1 F, FV = L^-T Ft as (ld, 64) row-major; R factor and beta
  int p = 1;
  DevBuf<double> FB, FtG, FV, Gm, betav, Vt;
  std::vector<double> h_G, h_beta, h_Ft;  // host copies: (p, p) row-major, (p,), (N, p) row-major
  // predict workspace
  int Mc = 0;
  DevBuf<double> Xc, Kst, yhat, sumsq, dotf, mse, params, part_val, best_val, vals;
  DevBuf<long long> part_idx, best_idx;
  // tensor-core (B200BO_PREC_FAST) state, built lazily after factor()
  bool fast_ready = false;
  bool restricted = false;      // the last factor() evaluated log_likelihood_restricted
  bool fvec_ready = false;      // fvec = L^-T Ft of the current factorisation (gradient path)
  DevBuf<double> gRT, gZ, g_ydx, g_mdx, g_val, g_dx;
  bool calibrated[2] = {false, false};  // [0]: one-product first pass, [1]: three-product pass
  int DP = 0, b_scale_log2 = 0;
  double dy_cal[2] = {0, 0}, ds_cal[2] = {0, 0};
  int fast_products = 1;   // products of the first pass (1: fp16 operands; 3: split fp16); set_fast_products
  bool escalate = false;   // the one-product band was too wide for this fit: go straight to three products
  DevBuf<__half> Lh, Ll;
  DevBuf<float> Xs, dbg_w;
  DevBuf<double> rs_part, cscale, fvec, f_yhat, f_sumsq, f_dotf, stage[2], Xband, errout, band_hiB;
  DevBuf<long long> band_list, band_list0, thr_key, trace;
  DevBuf<int> band_count, err_flag;
  CUtensorMap map_hi, map_lo;
  // second-generation kernel (Gram product on the tensor cores): its own operand copies
  bool use_v2 = false;
  DevBuf<__half> Xh2, Xl2;
  DevBuf<float> aux2;
  DevBuf<float2> exch2;
  DevBuf<double> cmean;
  CUtensorMap map2_hi, map2_lo, map2_xh, map2_xl;
  fk3::PairMaps pair_maps;  // third generation: CTA pairs (cta_group::2)
  bool use_pair = false;
  fk4::ReplayMaps replay_maps;  // fourth generation: CTA pairs + r replay from an L2-resident scratch
  bool use_replay = false;
  int last_n_store = 0;
  int replay_max_chunks = -1;   // test knob: cap on the stored chunks per tile (< 0: none)
  DevBuf<__half> r_scratch;
  int replay_mb = 64;           // scratch budget (MB): sized to stay in L2 next to the fp16 L^-1; 0 = recompute (generation 3)
  std::vector<double> xmean;  // per-feature mean of the training set (host copy from set_train)
  int fast_kernel_pref = 6;   // 6: two CTA pairs share a candidate tile (L2-resident scratch); 5: CTA pairs, producers decoupled through the scratch; 4: CTA pairs + r replay; 3: CTA pairs; 2: single-CTA Gram kernel; 1: first generation
  bool use_decoupled = false;
  bool use_shared = false;    // generation 6 applies to this fit (ld % 256 == 0, ld >= 1024, all CTAs co-resident)
  int shared_ok = -1;         // occupancy query of generation 6: -1 not asked yet, 0 no, 1 yes
  DevBuf<uint32_t> share_flags, smid_dbg;
  int rescore_max = 2048;     // B200BO_RESCORE_MAX: a one-product band up to max(this, M / 50) candidates is re-scored directly
  int fast_max_sms = 1 << 20; // B200BO_FAST_MAX_SMS: cap on the CTAs of the generation-6 grid (developer: scratch size vs SM count)
  int gen6_db_chunks = 0;     // B200BO_GEN6_DB_CHUNKS: leading chunks of a tile with two scratch slots (generation 6)
  int gen6_cooperative = 1;   // B200BO_GEN6_COOPERATIVE=0: plain launch (A/B)
  int gen6_min_ld = 4096;     // smallest padded N generation 6 is used for (B200BO_GEN6_MIN_LD)
  int last_gen = 0;           // generation of the fused kernel the last tensor-core launch used (timings[10])
  cudaStream_t copy_stream = nullptr;
  cudaStream_t la_stream = nullptr;  // low-priority helper stream of the Cholesky look-ahead
  cudaStream_t inv_stream = nullptr; // diagonal-block inverses, off the critical path
  std::vector<cudaEvent_t> la_ev;    // 3 events per panel (untimed)
  int lookahead = 2;  // 0: single stream; 1: look-ahead with separate factor / solve / update kernels; 2: one kernel per
                      // panel step where that measured faster (N <= 2048: 0.53 vs 0.59 ms at N = 1024; at N >= 4096 the
                      // fused step is 38 us against ~35 us for the three small kernels, 3.2 vs 3.04 ms), else as 1
  DevBuf<double> Lside, P0side;
  int use_graphs = 1;
  GraphSlot g_chol, g_trtri;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_used[2] = {nullptr, nullptr};
  bool want_dbg_w = false;
  EvPool evs;
  double timings[B200BO_N_TIMINGS] = {0};
  double fit_timings[B200BO_N_TIMINGS] = {0};
  bool gemm_attr[4] = {false, false, false, false};  // MaxDynamicSharedMemorySize set on this handle's device
  bool contract_attr = false;
  // band stage of the tensor-core path: a-priori error model, device pipeline workspaces, pinned staging
  FastModel model;
  bool use_model = true;          // B200BO_BAND_MODEL=0: calibrated half-widths only (round-1 behaviour, A/B runs)
  double widen[2] = {1.0, 1.0};   // widening factor of the half-widths for this fit (x4 whenever the band shows larger errors)
  double cal_err_y[2] = {0, 0}, cal_err_s[2] = {0, 0};  // largest errors of the calibration sample
  double last_err_y = 0, last_err_s = 0, last_ratio = 0;
  fk::BandArgs last_band{};
  LastFast last_fast;
  std::vector<double> last_theta;  // arguments of the last factor(): b200bo_append re-uses them
  double last_noise_arg = 0, last_beta_fixed = NAN;
  DevBuf<double> ap_Tr, ap_Ts, ap_Tu, ap_Cb, ap_Dv;
  int last_q = 0;                 // criteria of the last acquisition call (b200bo_best_pairs_device)
  int dev_chunk_tiles = 8;        // B200BO_DEV_CHUNK_TILES: fused launches of this many tiles per SM on device-resident input
                                  // (0 = one launch; 8 measured 4 % faster at M = 1e7 under the power cap, profiles/r02/device_path_chunking_ab.txt)
  std::vector<double> Xhost;      // training set as given (N, D): the distance-error bound needs max_j ||x_j||^2 per theta
  DevBuf<double> Xall, f_mse, bd_kst, bd_ypart, bd_part, pm_v, pm_t, pm_t2, rowsq, rowl1;
  DevBuf<bd::BandCtl> bd_ctl;
  uint8_t* pin = nullptr;         // pinned host staging of the band pipeline (parameters in, control block + bests out)
  // Cholesky trailing update on the tcgen05 tensor cores (oz_kernels.cuh): int8 digit planes of the current panel
  int chol_tc = 0;                // 0: fp64 DMMA trailing updates; 7 / 8: digit planes of the exact int8 splitting
  int chol_tc_min_rows = 1024;    // smaller trailing matrices stay on the DMMA path (too few tiles to fill the SMs)
  DevBuf<uint8_t> oz_dig;
  DevBuf<double> oz_scA, oz_scB;
  DevBuf<int> oz_err;
  CUtensorMap oz_mapA, oz_mapB;
  const void* oz_map_base = nullptr;
  int oz_map_rcap = 0;
  bool oz_attr = false;
  // TMA-staged kernel-matrix assembly (assemble_kernels.cuh)
  int assemble_tma = 1;           // B200BO_ASSEMBLE_TMA=0: the first version (fit_kernels.cuh)
  CUtensorMap mapXt;
  const void* mapXt_base = nullptr;
  int mapXt_ld = 0, mapXt_D = 0;
};

namespace {

template <typename Core, bool A_KM, bool B_KN>
cudaError_t launch_gemm_on(b200bo_ctx* h, cudaStream_t st, const GemmArgs& g, int M, int N, int batch) {
  auto kern = dgemm_kernel<64, 64, 32, 32, A_KM, B_KN, 3>;
  // the attribute is per DEVICE: remember it per handle (one handle = one device), not per process
  bool& attr_set = h->gemm_attr[(A_KM ? 2 : 0) + (B_KN ? 1 : 0)];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Core::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid(N / 64, M / 64, batch);
  kern<<<grid, Core::NT, Core::SMEM_BYTES, st>>>(g);
  return cudaGetLastError();
}
template <typename Core, bool A_KM, bool B_KN>
cudaError_t launch_gemm(b200bo_ctx* h, const GemmArgs& g, int M, int N, int batch) {
  return launch_gemm_on<Core, A_KM, B_KN>(h, h->stream, g, M, N, batch);
}

template <class F>
static int run_graphed(b200bo_ctx* h, GraphSlot& slot, const void* const (&key)[8], int& launches, F&& body) {
  cudaStream_t st = h->stream;
  // Measured (scripts/fit_time.py): replay wins where the chain is launch-latency bound (N = 1024: Cholesky 0.68 ->
  // 0.61 ms) and loses slightly at N >= 4096, where the captured nodes no longer carry the helper stream's low priority
  // and the big trailing GEMMs get in the way of the panel chain (3.06 -> 3.23 ms): eager above ld = 2048.
  if (h->ld > 2048) {
    int l = 0;
    int rc = body(l);
    launches += l;
    return rc;
  }
  bool same = true;
  for (int i = 0; i < 8; ++i) same = same && slot.key[i] == key[i];
  if (!h->use_graphs || !same) {
    slot.drop();
    for (int i = 0; i < 8; ++i) slot.key[i] = key[i];
  }
  if (h->use_graphs && slot.exec) {
    CU_TRY(cudaGraphLaunch(slot.exec, st));
    launches += slot.launches;
    return 0;
  }
  if (!h->use_graphs || slot.seen++ == 0) {  // first sighting: eager (also warms every function attribute)
    int l = 0;
    int rc = body(l);
    launches += l;
    return rc;
  }
  int l = 0;
  CU_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
  int rc = body(l);
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture(st, &g);
  if (rc || e != cudaSuccess || !g) {
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
    if (rc) return rc;
    h->use_graphs = 0;  // capture is not available here: stay eager
    l = 0;
    rc = body(l);
    launches += l;
    return rc;
  }
  e = cudaGraphInstantiate(&slot.exec, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) {
    slot.exec = nullptr;
    cudaGetLastError();
    h->use_graphs = 0;
    l = 0;
    rc = body(l);
    launches += l;
    return rc;
  }
  slot.launches = l;
  CU_TRY(cudaGraphLaunch(slot.exec, st));
  launches += l;
  return 0;
}

struct PhaseTimer {
  b200bo_ctx* h;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> spans[4];
  void begin(int k) {
    cudaEvent_t e = h->evs.get();
    cudaEventRecord(e, h->stream);
    spans[k].push_back({e, nullptr});
  }
  void end(int k) {
    cudaEvent_t e = h->evs.get();
    cudaEventRecord(e, h->stream);
    spans[k].back().second = e;
  }
  double total(int k) {
    double t = 0;
    for (auto& s : spans[k]) {
      float ms = 0;
      if (s.second && cudaEventElapsedTime(&ms, s.first, s.second) == cudaSuccess) t += ms;
    }
    return t;
  }
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encode_tiled(EncodeTiledFn* out) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CU_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres));
    if (!f || qres != cudaDriverEntryPointSuccess) return set_err(B200BO_E_CUDA, "cuTensorMapEncodeTiled is not available");
    fn = (EncodeTiledFn)f;
  }
  *out = fn;
  return 0;
}

// generic 2-D row-major tensor map: `rows` rows of `inner` elements, row pitch in bytes, boxes of box_rows x box_inner
static int make_map_2d(CUtensorMap* map, CUtensorMapDataType dt, const void* base, uint64_t inner, uint64_t rows,
                       uint64_t pitch_bytes, int box_inner, int box_rows, CUtensorMapSwizzle sw) {
  EncodeTiledFn fn;
  int rc = get_encode_tiled(&fn);
  if (rc) return rc;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dt, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(B200BO_E_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
  return 0;
}

// ---- Cholesky trailing update on the tensor cores (oz_kernels.cuh) ---------------------------------------------------
// C(r, c) -= sum_k P(r, k) P(c, k) over the lower tiles of the rows x rows matrix at C; P = rows x 64 panel (pitch ldp).
// Two launches on `st`: the digit split of the panel, the tcgen05 int8 SYRK.  h->chol_tc = digit planes (7 or 8).
static int launch_oz_syrk(b200bo_ctx* h, cudaStream_t st, const double* P, int ldp, int rows, double* C, int ldc,
                          int& launches) {
  const int S = h->chol_tc == 7 ? 7 : 8;
  const int rows_pad = round_up(rows, oz::TM);
  const int Rcap = std::max(h->ld, rows_pad);
  CHECK_ARG(rows > 0 && rows % 64 == 0, "oz_syrk: the panel extent must be a positive multiple of 64");
  CU_TRY(h->oz_dig.reserve((size_t)oz::NPL * Rcap * 128));
  CU_TRY(h->oz_scA.reserve(Rcap));
  CU_TRY(h->oz_scB.reserve(Rcap));
  if (!h->err_flag.p) {
    CU_TRY(h->err_flag.reserve(1));
    CU_TRY(cudaMemsetAsync(h->err_flag.p, 0, sizeof(int), st));
  }
  int rc;
  if (h->oz_map_base != h->oz_dig.p || h->oz_map_rcap != Rcap) {
    if ((rc = make_map_2d(&h->oz_mapA, CU_TENSOR_MAP_DATA_TYPE_UINT8, h->oz_dig.p, 128, (uint64_t)oz::NPL * Rcap, 128, 128, oz::TM,
                          CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_map_2d(&h->oz_mapB, CU_TENSOR_MAP_DATA_TYPE_UINT8, h->oz_dig.p, 128, (uint64_t)oz::NPL * Rcap, 128, 128, oz::TN,
                          CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    h->oz_map_base = h->oz_dig.p;
    h->oz_map_rcap = Rcap;
  }
  CUtensorMap mapC;
  if ((rc = make_map_2d(&mapC, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, C, (uint64_t)rows, (uint64_t)rows, (uint64_t)ldc * 8, oz::TM, oz::TN,
                        CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
  if (!h->oz_attr) {
    CU_TRY(cudaFuncSetAttribute(oz::oz_syrk_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, oz::SMEM_BYTES));
    CU_TRY(cudaFuncSetAttribute(oz::oz_syrk_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, oz::SMEM_BYTES));
    h->oz_attr = true;
  }
  oz::OzArgs a;
  a.scA = h->oz_scA.p; a.scB = h->oz_scB.p; a.rows = rows; a.Rcap = Rcap; a.err = h->err_flag.p;
  a.dbg = getenv("B200BO_OZ_DEBUG") ? atoi(getenv("B200BO_OZ_DEBUG")) : 0;
  const int ncb = (rows + oz::TM - 1) / oz::TM, nrb = rows / oz::TN;
  long long tiles = 0;
  for (int cb = 0; cb < ncb; ++cb) tiles += nrb - cb * (oz::TM / oz::TN);
  a.G = (int)std::max<long long>(2, std::min<long long>(16, tiles / (2LL * h->num_sms)));
  if (const char* e = getenv("B200BO_OZ_G")) a.G = std::max(1, atoi(e));
  int items = 0;
  for (int cb = 0; cb < ncb; ++cb) items += (nrb - cb * (oz::TM / oz::TN) + a.G - 1) / a.G;
  const int grid = std::min(items, h->num_sms);
  if (S == 7) {
    oz::oz_split_kernel<7><<<rows_pad / 8, 256, 0, st>>>(P, ldp, rows, rows_pad, Rcap, h->oz_dig.p, h->oz_scA.p, h->oz_scB.p);
    CU_TRY(cudaGetLastError());
    oz::oz_syrk_kernel<7><<<grid, oz::NT_OZ, oz::SMEM_BYTES, st>>>(h->oz_mapA, h->oz_mapB, mapC, a);
  } else {
    oz::oz_split_kernel<8><<<rows_pad / 8, 256, 0, st>>>(P, ldp, rows, rows_pad, Rcap, h->oz_dig.p, h->oz_scA.p, h->oz_scB.p);
    CU_TRY(cudaGetLastError());
    oz::oz_syrk_kernel<8><<<grid, oz::NT_OZ, oz::SMEM_BYTES, st>>>(h->oz_mapA, h->oz_mapB, mapC, a);
  }
  CU_TRY(cudaGetLastError());
  launches += 2;
  return 0;
}

int ensure_predict_ws(b200bo_ctx* h, int q, bool need_vals_stage) {
  if (h->Mc == 0) h->Mc = h->num_sms * PC_BM;
  size_t Mc = h->Mc;
  CU_TRY(h->Xc.reserve(Mc * h->D));
  CU_TRY(h->Kst.reserve(Mc * (size_t)h->ld));
  CU_TRY(h->yhat.reserve(Mc));
  CU_TRY(h->sumsq.reserve(Mc));
  CU_TRY(h->dotf.reserve(Mc));
  CU_TRY(h->mse.reserve(Mc));
  int qq = std::max(q, 1);
  CU_TRY(h->params.reserve(qq));
  CU_TRY(h->part_val.reserve((size_t)qq * h->num_sms));
  CU_TRY(h->part_idx.reserve((size_t)qq * h->num_sms));
  CU_TRY(h->best_val.reserve(qq));
  CU_TRY(h->best_idx.reserve(qq));
  if (need_vals_stage) CU_TRY(h->vals.reserve((size_t)qq * Mc));
  return 0;
}

}  // namespace

extern "C" {

const char* b200bo_last_error(void) { return g_err.c_str(); }
int b200bo_version(void) { return 100; }

int b200bo_create(int device, b200bo_handle* out) {
  CHECK_ARG(out != nullptr, "out is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return set_err(B200BO_E_NODEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e));
  CHECK_ARG(device >= 0 && device < n, "device index out of range");
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return set_err(B200BO_E_NODEVICE, "libb200bo is built for sm_100a (B200) only; found sm_" +
                                          std::to_string(prop.major) + std::to_string(prop.minor));
  CU_TRY(cudaSetDevice(device));
  b200bo_ctx* h = new b200bo_ctx();
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  CU_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CU_TRY(h->scal.reserve(16));
  CU_TRY(h->status.reserve(1));
  // developer knobs (A/B runs): first-pass products and the mbarrier suspend hint of the fused kernels
  if (const char* e = getenv("B200BO_FAST_PRODUCTS")) h->fast_products = atoi(e) == 3 ? 3 : 1;
  if (const char* e = getenv("B200BO_FAST_KERNEL")) h->fast_kernel_pref = std::max(1, std::min(6, atoi(e)));
#ifndef B200BO_DEV_KERNELS
  if (h->fast_kernel_pref == 2 || h->fast_kernel_pref == 3) h->fast_kernel_pref = 4;
#endif
  if (const char* e = getenv("B200BO_CHOL_LOOKAHEAD")) h->lookahead = std::max(0, std::min(2, atoi(e)));
  if (const char* e = getenv("B200BO_GRAPHS")) h->use_graphs = atoi(e) != 0;
  if (const char* e = getenv("B200BO_GEN6_MIN_LD")) h->gen6_min_ld = std::max(1024, atoi(e));
  if (const char* e = getenv("B200BO_GEN6_COOPERATIVE")) h->gen6_cooperative = atoi(e) != 0;
  if (const char* e = getenv("B200BO_GEN6_DB_CHUNKS")) h->gen6_db_chunks = std::max(0, atoi(e));
  if (const char* e = getenv("B200BO_FAST_MAX_SMS")) h->fast_max_sms = std::max(4, atoi(e));
  if (const char* e = getenv("B200BO_RESCORE_MAX")) h->rescore_max = std::max(1, atoi(e));
  if (const char* e = getenv("B200BO_ASSEMBLE_TMA")) h->assemble_tma = atoi(e) != 0;
  if (const char* e = getenv("B200BO_CHOL_TC")) { int v = atoi(e); h->chol_tc = (v == 7 || v == 8) ? v : 0; }
  if (const char* e = getenv("B200BO_CHOL_TC_MIN_ROWS")) h->chol_tc_min_rows = std::max(64, atoi(e));
  if (const char* e = getenv("B200BO_REPLAY_MB")) h->replay_mb = std::max(0, std::min(4096, atoi(e)));
  if (const char* e = getenv("B200BO_BAND_MODEL")) h->use_model = atoi(e) != 0;
  if (const char* e = getenv("B200BO_DEV_CHUNK_TILES")) h->dev_chunk_tiles = std::max(0, atoi(e));
  if (const char* e = getenv("B200BO_WAIT_HINT_NS")) {
    unsigned v = (unsigned)atoi(e);
    CU_TRY(cudaMemcpyToSymbol(fk::g_wait_hint_ns, &v, sizeof v));
  }
  *out = h;
  return 0;
}

int b200bo_destroy(b200bo_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  h->Xt.release(); h->y.release(); h->F.release(); h->theta.release();
  h->A.release(); h->W.release(); h->S.release(); h->Rkeep.release(); h->Dinv.release();
  h->Yt.release(); h->Ft.release(); h->rho.release(); h->gamma.release(); h->part.release();
  h->scal.release(); h->status.release();
  h->Xc.release(); h->Kst.release(); h->yhat.release(); h->sumsq.release(); h->dotf.release();
  h->mse.release(); h->params.release(); h->part_val.release(); h->best_val.release(); h->vals.release();
  h->part_idx.release(); h->best_idx.release();
  h->rs_part.release(); h->Xh2.release(); h->Xl2.release(); h->aux2.release(); h->exch2.release(); h->cmean.release(); h->r_scratch.release();
  h->Lh.release(); h->Ll.release(); h->Xs.release(); h->dbg_w.release();
  h->cscale.release(); h->fvec.release(); h->f_yhat.release(); h->f_sumsq.release(); h->f_dotf.release();
  h->stage[0].release(); h->stage[1].release(); h->Xband.release(); h->band_hiB.release();
  h->trace.release(); h->errout.release(); h->band_list.release(); h->band_list0.release(); h->thr_key.release();
  h->band_count.release(); h->err_flag.release();
  h->Xall.release(); h->f_mse.release(); h->bd_kst.release(); h->bd_ypart.release(); h->bd_part.release();
  h->ap_Tr.release(); h->ap_Ts.release(); h->ap_Tu.release(); h->ap_Cb.release(); h->ap_Dv.release();
  h->oz_dig.release(); h->oz_scA.release(); h->oz_scB.release(); h->share_flags.release();
  h->pm_v.release(); h->pm_t.release(); h->pm_t2.release(); h->rowsq.release(); h->rowl1.release(); h->bd_ctl.release();
  if (h->pin) cudaFreeHost(h->pin);
  for (int i = 0; i < 2; ++i) {
    if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]);
    if (h->ev_used[i]) cudaEventDestroy(h->ev_used[i]);
  }
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  h->g_chol.drop();
  h->g_trtri.drop();
  if (h->la_stream) cudaStreamDestroy(h->la_stream);
  if (h->inv_stream) cudaStreamDestroy(h->inv_stream);
  for (cudaEvent_t e : h->la_ev) cudaEventDestroy(e);
  h->evs.destroy();
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

int b200bo_set_stream(b200bo_handle h, void* s) {
  CHECK_ARG(h, "handle is NULL");
  CU_TRY(cudaSetDevice(h->device));
  CU_TRY(cudaStreamSynchronize(h->stream));
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  h->stream = (cudaStream_t)s;
  h->own_stream = false;
  return 0;
}

int b200bo_set_precision(b200bo_handle h, int prec) {
  CHECK_ARG(h, "handle is NULL");
  CHECK_ARG(prec == B200BO_PREC_FP64 || prec == B200BO_PREC_FAST, "unknown precision id");
  h->prec = prec;
  return 0;
}

int b200bo_set_fast_kernel(b200bo_handle h, int generation) {
  CHECK_ARG(h, "handle is NULL");
  CHECK_ARG(generation >= 1 && generation <= 6, "generation is 1 .. 6");
#ifndef B200BO_DEV_KERNELS
  CHECK_ARG(generation != 2 && generation != 3, "generations 2 and 3 are only in developer builds (-DB200BO_DEV_KERNELS)");
#endif
  h->fast_kernel_pref = generation;
  h->fast_ready = false;
  h->fvec_ready = false;
  h->calibrated[0] = h->calibrated[1] = false;
  return 0;
}

int b200bo_set_replay(b200bo_handle h, int budget_mb, int max_chunks) {
  CHECK_ARG(h, "handle is NULL");
  CHECK_ARG(budget_mb >= 0 && budget_mb <= 4096, "budget_mb out of range");
  h->replay_mb = budget_mb;
  h->replay_max_chunks = max_chunks;
  h->fast_ready = false;
  h->fvec_ready = false;
  h->calibrated[0] = h->calibrated[1] = false;
  return 0;
}

int b200bo_set_fast_products(b200bo_handle h, int products) {
  CHECK_ARG(h, "handle is NULL");
  CHECK_ARG(products == 1 || products == 3, "products is 1 or 3");
  h->fast_products = products;
  h->escalate = false;
  return 0;
}

int b200bo_set_keep_R(b200bo_handle h, int keep) {
  CHECK_ARG(h, "handle is NULL");
  h->keepR = keep != 0;
  return 0;
}

int b200bo_set_chol_tc(b200bo_handle h, int digits, int min_rows) {
  CHECK_ARG(h, "handle is NULL");
  CHECK_ARG(digits == 0 || digits == 7 || digits == 8, "digits must be 0 (fp64 DMMA), 7 or 8");
  h->chol_tc = digits;
  if (min_rows > 0) h->chol_tc_min_rows = std::max(64, min_rows);
  return 0;
}

int b200bo_debug_oz_syrk(b200bo_handle h, const double* P_host, int rows, double* C_host, int digits, int reps, double* out_ms) {
  CHECK_ARG(h && P_host && C_host, "NULL argument");
  CHECK_ARG(rows > 0 && rows % 64 == 0, "rows must be a positive multiple of 64");
  CHECK_ARG(digits == 7 || digits == 8 || digits == 0, "digits must be 7 or 8 (0: the fp64 DMMA kernel, for comparison)");
  CU_TRY(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  DevBuf<double> P, Cd;
  CU_TRY(P.reserve((size_t)rows * 64));
  CU_TRY(Cd.reserve((size_t)rows * rows));
  CU_TRY(cudaMemcpyAsync(P.p, P_host, (size_t)rows * 64 * 8, cudaMemcpyHostToDevice, st));
  const int keep = h->chol_tc;
  h->chol_tc = digits;
  cudaEvent_t e0, e1;
  CU_TRY(cudaEventCreate(&e0));
  CU_TRY(cudaEventCreate(&e1));
  float best = 0;
  int rc = 0;
  for (int r = 0; r < std::max(reps, 1) && !rc; ++r) {
    CU_TRY(cudaMemcpyAsync(Cd.p, C_host, (size_t)rows * rows * 8, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaEventRecord(e0, st));
    int launches = 0;
    if (digits) {
      rc = launch_oz_syrk(h, st, P.p, 64, rows, Cd.p, rows, launches);
    } else {
      GemmArgs g{};
      g.A = P.p; g.B = P.p; g.C = Cd.p; g.lda = 64; g.ldb = 64; g.ldc = rows; g.K = 64; g.alpha = -1.0; g.beta = 1.0; g.lower_only = 1;
      cudaError_t e = launch_gemm<GemmNT, false, false>(h, g, rows, rows, 1);
      if (e != cudaSuccess) rc = set_err(B200BO_E_CUDA, cudaGetErrorString(e));
    }
    CU_TRY(cudaEventRecord(e1, st));
    CU_TRY(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r == 0 || ms < best) best = ms;
  }
  h->chol_tc = keep;
  if (!rc) {
    CU_TRY(cudaMemcpyAsync(C_host, Cd.p, (size_t)rows * rows * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    int flag = 0;
    if (digits && h->err_flag.p) CU_TRY(cudaMemcpy(&flag, h->err_flag.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (flag) rc = set_err(B200BO_E_CUDA, "oz_syrk pipeline wait timed out (code " + std::to_string(flag) + ")");
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  P.release();
  Cd.release();
  if (out_ms) *out_ms = best;
  return rc;
}

int b200bo_set_train(b200bo_handle h, const double* X, const double* y, int N, int D) {
  CHECK_ARG(h && X && y, "NULL argument");
  CHECK_ARG(N >= 1 && D >= 1, "N and D must be positive");
  CHECK_ARG(D <= 1024, "D > 1024 is not supported");
  CU_TRY(cudaSetDevice(h->device));
  h->N = N;
  h->D = D;
  h->ld = round_up(N, 128);
  h->factored = false;
  const int ld = h->ld;
  std::vector<double> xt((size_t)D * ld, 0.0), yy(ld, 0.0), ff(ld, 0.0);
  h->Xhost.assign(X, X + (size_t)N * D);
  h->xmean.assign(D, 0.0);
  for (int i = 0; i < N; ++i)
    for (int d = 0; d < D; ++d) h->xmean[d] += X[(size_t)i * D + d];
  for (int d = 0; d < D; ++d) h->xmean[d] /= N;
  for (int i = 0; i < N; ++i) {
    for (int d = 0; d < D; ++d) xt[(size_t)d * ld + i] = X[(size_t)i * D + d];
    yy[i] = y[i];
    ff[i] = 1.0;  // constant trend basis F = ones (trend.py:76-79); zero on padding rows
  }
  CU_TRY(h->Xt.reserve(xt.size()));
  CU_TRY(h->y.reserve(ld));
  CU_TRY(h->F.reserve(ld));
  CU_TRY(h->theta.reserve(D + 1));
  CU_TRY(cudaMemcpyAsync(h->Xt.p, xt.data(), xt.size() * 8, cudaMemcpyHostToDevice, h->stream));
  CU_TRY(cudaMemcpyAsync(h->y.p, yy.data(), ld * 8, cudaMemcpyHostToDevice, h->stream));
  CU_TRY(cudaMemcpyAsync(h->F.p, ff.data(), ld * 8, cudaMemcpyHostToDevice, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  // predict workspaces depend on ld
  h->Kst.release();
  return 0;
}

static int append_front(b200bo_handle h, int N0, int m, int corr, int mode, double par_last, double noise_var, int& launches);

// append_m > 0: rows append_N0 .. append_N0 + append_m - 1 are new; A and W hold the factor of the leading block and steps
// 1-3 (assembly, Cholesky, L^-1) are replaced by the bordering update of append_front
static int factor_impl(b200bo_handle h, int corr, const double* theta, int n_theta, int mode, double par_last,
                       double noise_var, int trend, const double* beta_or_null, double* out_llf,
                       double* out_sigma2, double* out_noise_var, int* out_status, bool restricted,
                       int append_N0 = 0, int append_m = 0) {
  CHECK_ARG(h && theta, "NULL argument");
  CHECK_ARG(h->N > 0, "set_train first");
  CHECK_ARG(corr >= 0 && corr <= 7, "unknown correlation id");
  CHECK_ARG(mode >= 0 && mode <= 2, "unknown estimation mode");
  CHECK_ARG(trend >= B200BO_TREND_CONSTANT && trend <= B200BO_TREND_QUADRATIC, "unknown trend id");
  CHECK_ARG(trend_p(trend, h->D) <= TR_PMAX, "at most 64 trend basis functions are supported");
  CHECK_ARG(!(restricted && trend != B200BO_TREND_CONSTANT), "the restricted likelihood is implemented for the constant trend");
  if (corr_has_extra_param(corr)) {
    CHECK_ARG(n_theta == 2 || n_theta == h->D + 1, "Length of theta must be 2 or D + 1");  // kernel.py:367-370
    --n_theta;  // the last entry is the exponent (generalized_exponential) or nu (general Matern)
    CHECK_ARG(corr != MATERN_NU || (theta[n_theta] > 0 && theta[n_theta] <= 50.0), "nu must be in (0, 50]");
  }
  CHECK_ARG(n_theta == 1 || n_theta == h->D, "Length of theta must be 1 or D");
  // an estimated trend needs N >= p rows for the thin QR of the (N, p) panel (gpr.py:298-309 raises upstream)
  CHECK_ARG(beta_or_null != nullptr || trend_p(trend, h->D) <= h->N,
            "Ordinary least squares problem is undetermined: n_samples must be >= the trend basis size p");
  CU_TRY(cudaSetDevice(h->device));
  const int N = h->N, D = h->D, ld = h->ld, nb = ld / NB;
  const size_t nn = (size_t)ld * ld;
  h->factored = false;
  h->fast_ready = false;
  h->fvec_ready = false;
  h->calibrated[0] = h->calibrated[1] = false;
  h->escalate = false;
  h->widen[0] = h->widen[1] = 1.0;
  h->last_fast = LastFast();
  CU_TRY(h->A.reserve(nn));
  CU_TRY(h->W.reserve(nn));
  CU_TRY(h->S.reserve(nn));
  CU_TRY(h->Dinv.reserve((size_t)nb * NB * NB));
  CU_TRY(h->Yt.reserve(ld));
  CU_TRY(h->Ft.reserve(ld));
  CU_TRY(h->rho.reserve(ld));
  CU_TRY(h->gamma.reserve(ld));
  const int gchunks = (ld + 255) / 256;
  CU_TRY(h->part.reserve((size_t)gchunks * ld));
  if (h->keepR) CU_TRY(h->Rkeep.reserve(nn));

  std::vector<double> th(D + 1, 0.0);
  for (int d = 0; d < D; ++d) th[d] = theta[n_theta == 1 ? 0 : d];
  if (corr_has_extra_param(corr)) th[D] = theta[n_theta];
  cudaStream_t st = h->stream;
  h->evs.reset();
  PhaseTimer pt{h};
  int launches = 0;
  CU_TRY(cudaMemcpyAsync(h->theta.p, th.data(), (D + 1) * 8, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaMemsetAsync(h->status.p, 0, sizeof(int), st));
  cudaEvent_t e0 = h->evs.get(), e1 = h->evs.get();
  CU_TRY(cudaEventRecord(e0, st));

  if (append_m > 0) {
    pt.begin(0);
    int rc = append_front(h, append_N0, append_m, corr, mode, par_last, noise_var, launches);
    if (rc) return rc;
    pt.end(0);
  } else {
  // ---- 1. kernel-matrix assembly -------------------------------------------------------------------
  pt.begin(0);
  {
    AssembleArgs a;
    a.Xt = h->Xt.p; a.R = h->A.p; a.theta = h->theta.p;
    a.N = N; a.D = D; a.ld = ld; a.corr = corr; a.mode = mode;
    a.sigma2 = mode == B200BO_MODE_NOISY ? par_last : 0.0;
    a.noise_var = mode == B200BO_MODE_NOISY ? noise_var : 0.0;
    a.alpha = mode == B200BO_MODE_NOISE_ESTIM ? par_last : 1.0;
    if (h->assemble_tma && D <= KA_DMAX) {
      // TMA-staged version (assemble_kernels.cuh): the (D, 64) slabs of Xt by cp.async.bulk.tensor, 256-bit stores
      if (h->mapXt_base != h->Xt.p || h->mapXt_ld != ld || h->mapXt_D != D) {
        int rc = make_map_2d(&h->mapXt, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, h->Xt.p, (uint64_t)ld, (uint64_t)D, (uint64_t)ld * 8, NB, D,
                             CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc) return rc;
        h->mapXt_base = h->Xt.p; h->mapXt_ld = ld; h->mapXt_D = D;
      }
      const size_t smem = ((size_t)2 * D * NB + D + 2) * sizeof(double);
#define KA_LAUNCH(C)                                                                                                    \
  do {                                                                                                                  \
    if (smem > 48 * 1024) CU_TRY(cudaFuncSetAttribute(kmat_assemble_tma_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kmat_assemble_tma_kernel<C><<<nb * (nb + 1) / 2, 256, smem, st>>>(h->mapXt, a);                                       \
  } while (0)
      switch (corr) {
        case RBF: KA_LAUNCH(RBF); break;
        case MATERN12: KA_LAUNCH(MATERN12); break;
        case MATERN32: KA_LAUNCH(MATERN32); break;
        case MATERN52: KA_LAUNCH(MATERN52); break;
        case ABSEXP: KA_LAUNCH(ABSEXP); break;
        default: KA_LAUNCH(-1); break;  // cubic, generalized_exponential, general-nu Matern
      }
#undef KA_LAUNCH
    } else {
      size_t smem = ((size_t)2 * D * NB + ((D + 1) & ~1) + NB * 66) * sizeof(double);
      CU_TRY(cudaFuncSetAttribute(kmat_assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kmat_assemble_kernel<<<nb * (nb + 1) / 2, 256, smem, st>>>(a);
    }
    CU_TRY(cudaGetLastError());
    ++launches;
    if (h->keepR) CU_TRY(cudaMemcpyAsync(h->Rkeep.p, h->A.p, nn * 8, cudaMemcpyDeviceToDevice, st));
  }
  pt.end(0);

  // ---- 2. blocked right-looking Cholesky (lower), fp64 DMMA trailing updates -------------------------
  pt.begin(1);
  CU_TRY(cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CHOL_DIAG_SMEM));
  // Look-ahead of depth one: the trailing update of panel j is split into (a) the next block column -- all the next
  // diagonal block and panel need -- on the main stream and (b) the rest on a low-priority helper stream, so that
  // chol_diag(j+1) (one CTA, latency bound) and panel(j+1) run underneath (b) of panel j.  (a) of panel j+1 touches the
  // block column (b) of panel j also updates: it waits for it.
  const bool la = h->lookahead && nb > 2;
  if (la) {
    if (!h->la_stream) {
      int lo = 0, hi = 0;
      CU_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CU_TRY(cudaStreamCreateWithPriority(&h->la_stream, cudaStreamNonBlocking, lo));
      CU_TRY(cudaStreamCreateWithPriority(&h->inv_stream, cudaStreamNonBlocking, lo));
      CU_TRY(cudaFuncSetAttribute(chol_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CHOL_FACTOR_SMEM));
      CU_TRY(cudaFuncSetAttribute(chol_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CHOL_DIAG_SMEM));
      CU_TRY(cudaFuncSetAttribute(panel_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PANEL_TRSM_SMEM));
      CU_TRY(cudaFuncSetAttribute(panel_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PANEL_CHAIN_SMEM));
      CU_TRY(cudaFuncSetAttribute(chol_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CHOL_DIAG_SMEM));
    }
    CU_TRY(h->Lside.reserve((size_t)nb * NB * NB));
    CU_TRY(h->P0side.reserve((size_t)nb * NB * NB));
    while ((int)h->la_ev.size() < 3 * nb + 1) {
      cudaEvent_t e;
      CU_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->la_ev.push_back(e);
    }
  }
  const void* const gkey[8] = {h->A.p, h->W.p, h->S.p, h->Dinv.p, h->status.p, (const void*)(intptr_t)ld,
                               (const void*)(intptr_t)((la ? (h->lookahead >= 2 && ld <= 2048 ? 2 : 1) : 0) + 16 * h->chol_tc + 1024 * h->chol_tc_min_rows), (const void*)st};
  auto chol_body = [&](int& launches) -> int {
  int last_b = -1;  // index of the last panel whose (b) part went to the helper stream
  if (la && h->lookahead >= 2 && ld <= 2048) {
    // one kernel per panel step on the main stream (update of block column jb by panel jb - 1, factor, solve); the
    // rest of panel jb's trailing update (b) and the block inverse / write-backs on the helper streams.  Step jb reads
    // block column jb, which (b) of every panel <= jb - 2 has updated: it waits for the last of them.
    for (int jb = 0; jb < nb; ++jb) {
      const int nrows = nb - 1 - jb;
      if (jb >= 2 && last_b >= 0) CU_TRY(cudaStreamWaitEvent(st, h->la_ev[2 * std::min(last_b, jb - 2) + 1], 0));
      panel_chain_kernel<<<std::max(nrows, 1), 256, PANEL_CHAIN_SMEM, st>>>(h->A.p, ld, jb, nrows, h->Lside.p, h->P0side.p, h->status.p);
      CU_TRY(cudaGetLastError());
      CU_TRY(cudaEventRecord(h->la_ev[2 * jb], st));
      CU_TRY(cudaStreamWaitEvent(h->inv_stream, h->la_ev[2 * jb], 0));
      chol_finish_kernel<<<1, 256, CHOL_DIAG_SMEM, h->inv_stream>>>(h->A.p, ld, jb, nrows, h->Lside.p, h->P0side.p, h->Dinv.p);
      CU_TRY(cudaGetLastError());
      launches += 2;
      if (nrows >= 2) {
        CU_TRY(cudaStreamWaitEvent(h->la_stream, h->la_ev[2 * jb], 0));
        double* P1 = h->A.p + (size_t)(jb + 2) * NB * ld + (size_t)jb * NB;  // rows of panel jb below its first block
        GemmArgs bq{};
        bq.A = P1; bq.B = P1; bq.C = h->A.p + (size_t)(jb + 2) * NB * (ld + 1);
        bq.lda = ld; bq.ldb = ld; bq.ldc = ld; bq.K = NB; bq.alpha = -1.0; bq.beta = 1.0; bq.lower_only = 1;
        if (h->chol_tc && (nrows - 1) * NB >= h->chol_tc_min_rows) {
          int rc = launch_oz_syrk(h, h->la_stream, P1, ld, (nrows - 1) * NB, bq.C, ld, launches);
          if (rc) return rc;
        } else {
          CU_TRY((launch_gemm_on<GemmNT, false, false>(h, h->la_stream, bq, (nrows - 1) * NB, (nrows - 1) * NB, 1)));
          ++launches;
        }
        CU_TRY(cudaEventRecord(h->la_ev[2 * jb + 1], h->la_stream));
        last_b = jb;
      }
    }
  } else
  for (int jb = 0; jb < nb; ++jb) {
    double* Ajj = h->A.p + (size_t)jb * NB * (ld + 1);
    double* Dj = h->Dinv.p + (size_t)jb * NB * NB;
    if (la) {
      // factor on the main stream; the block inverse (only the L^-1 stage needs it) on its own stream
      chol_factor_kernel<<<1, 256, CHOL_FACTOR_SMEM, st>>>(Ajj, ld, h->status.p);
      CU_TRY(cudaGetLastError());
      CU_TRY(cudaEventRecord(h->la_ev[2 * nb + jb], st));
      CU_TRY(cudaStreamWaitEvent(h->inv_stream, h->la_ev[2 * nb + jb], 0));
      chol_inverse_kernel<<<1, 256, CHOL_DIAG_SMEM, h->inv_stream>>>(Ajj, ld, Dj);
      CU_TRY(cudaGetLastError());
      ++launches;
    } else {
      chol_diag_kernel<<<1, 256, CHOL_DIAG_SMEM, st>>>(Ajj, ld, Dj, h->status.p);
      CU_TRY(cudaGetLastError());
    }
    ++launches;
    int mrem = ld - (jb + 1) * NB;
    if (mrem > 0) {
      double* P = h->A.p + (size_t)(jb + 1) * NB * ld + (size_t)jb * NB;
      if (la) {  // panel: P <- P * Ljj^-T by row-parallel substitution against the factor itself
        panel_trsm_kernel<<<mrem / 64, 64, PANEL_TRSM_SMEM, st>>>(P, ld, Ajj);
        CU_TRY(cudaGetLastError());
      } else {
        GemmArgs g{};  // panel: P <- P * Ljj^-T   (C(m,n) = sum_k P(m,k) Dinv(n,k))
        g.A = P; g.B = Dj; g.C = P; g.lda = ld; g.ldb = NB; g.ldc = ld; g.K = NB; g.alpha = 1.0; g.beta = 0.0;
        CU_TRY((launch_gemm<GemmNT, false, false>(h, g, mrem, NB, 1)));
      }
      GemmArgs s{};  // trailing update: A22 <- A22 - P P^T, lower tiles only
      s.A = P; s.B = P; s.C = h->A.p + (size_t)(jb + 1) * NB * (ld + 1);
      s.lda = ld; s.ldb = ld; s.ldc = ld; s.K = NB; s.alpha = -1.0; s.beta = 1.0; s.lower_only = 1;
      if (!la) {
        if (h->chol_tc && mrem >= h->chol_tc_min_rows) {
          int rc = launch_oz_syrk(h, st, P, ld, mrem, s.C, ld, launches);
          if (rc) return rc;
          ++launches;
        } else {
          CU_TRY((launch_gemm<GemmNT, false, false>(h, s, mrem, mrem, 1)));
          launches += 2;
        }
        continue;
      }
      if (last_b >= 0) CU_TRY(cudaStreamWaitEvent(st, h->la_ev[2 * last_b + 1], 0));
      CU_TRY((launch_gemm<GemmNT, false, false>(h, s, mrem, NB, 1)));  // (a): block column jb + 1
      launches += 2;
      if (mrem > NB) {
        CU_TRY(cudaEventRecord(h->la_ev[2 * jb], st));
        CU_TRY(cudaStreamWaitEvent(h->la_stream, h->la_ev[2 * jb], 0));
        GemmArgs b = s;  // (b): columns jb + 2 .. of the trailing matrix, from the rows of P below the first block
        b.A = P + (size_t)NB * ld; b.B = b.A; b.C = s.C + (size_t)NB * (ld + 1);
        if (h->chol_tc && mrem - NB >= h->chol_tc_min_rows) {  // exact int8 digit products on tcgen05 (oz_kernels.cuh)
          int rc = launch_oz_syrk(h, h->la_stream, b.A, ld, mrem - NB, b.C, ld, launches);
          if (rc) return rc;
        } else {
          CU_TRY((launch_gemm_on<GemmNT, false, false>(h, h->la_stream, b, mrem - NB, mrem - NB, 1)));
          ++launches;
        }
        CU_TRY(cudaEventRecord(h->la_ev[2 * jb + 1], h->la_stream));
        last_b = jb;
      }
    }
  }
  if (la) {
    if (last_b >= 0) CU_TRY(cudaStreamWaitEvent(st, h->la_ev[2 * last_b + 1], 0));
    CU_TRY(cudaEventRecord(h->la_ev[3 * nb], h->inv_stream));  // every diagonal inverse is in place
    CU_TRY(cudaStreamWaitEvent(st, h->la_ev[3 * nb], 0));
  }
  zero_upper_kernel<<<h->num_sms * 4, 256, 0, st>>>(h->A.p, ld);
  CU_TRY(cudaGetLastError());
  ++launches;
  return 0;
  };
  {
    int rc = run_graphed(h, h->g_chol, gkey, launches, chol_body);
    if (rc) return rc;
  }
  pt.end(1);

  // ---- 3. L^-1 by recursive doubling: [[A,0],[B,C]]^-1 = [[A^-1,0],[-C^-1 B A^-1, C^-1]] -------------
  pt.begin(2);
  auto trtri_body = [&](int& launches) -> int {
  CU_TRY(cudaMemsetAsync(h->W.p, 0, nn * 8, st));
  scatter_dinv_kernel<<<nb, 256, 0, st>>>(h->Dinv.p, h->W.p, ld);
  CU_TRY(cudaGetLastError());
  ++launches;
  for (int s = NB; s < ld; s *= 2) {
    auto merge = [&](int a0, int c, int batch) -> int {
      // T = B * W_A     (c x s) = (c x s)(s x s), W_A lower  -> k >= n0
      GemmArgs g1{};
      g1.A = h->A.p + (size_t)(a0 + s) * ld + a0; g1.lda = ld;
      g1.B = h->W.p + (size_t)a0 * (ld + 1); g1.ldb = ld;
      g1.C = h->S.p + (size_t)(a0 + s) * ld + a0; g1.ldc = ld;
      g1.sA = g1.sB = g1.sC = (long long)2 * s * (ld + 1);
      g1.K = s; g1.alpha = 1.0; g1.beta = 0.0; g1.kb_mode = 1;
      CU_TRY((launch_gemm<GemmNN, false, true>(h, g1, c, s, batch)));
      // W_B = -W_C * T  (c x s) = (c x c)(c x s), W_C lower  -> k < m0 + BM
      GemmArgs g2{};
      g2.A = h->W.p + (size_t)(a0 + s) * (ld + 1); g2.lda = ld;
      g2.B = h->S.p + (size_t)(a0 + s) * ld + a0; g2.ldb = ld;
      g2.C = h->W.p + (size_t)(a0 + s) * ld + a0; g2.ldc = ld;
      g2.sA = g2.sB = g2.sC = (long long)2 * s * (ld + 1);
      g2.K = c; g2.alpha = -1.0; g2.beta = 0.0; g2.ke_mode = 2;
      CU_TRY((launch_gemm<GemmNN, false, true>(h, g2, c, s, batch)));
      launches += 2;
      return 0;
    };
    int gf = ld / (2 * s);
    if (gf > 0) {
      int rc = merge(0, s, gf);
      if (rc) return rc;
    }
    int rem = ld - gf * 2 * s;
    if (rem > s) {
      int rc = merge(gf * 2 * s, rem - s, 1);
      if (rc) return rc;
    }
  }
  return 0;
  };
  {
    int rc = run_graphed(h, h->g_trtri, gkey, launches, trtri_body);
    if (rc) return rc;
  }
  pt.end(2);
  }  // full factorisation

  // ---- 4. solves: Yt, Ft, rho, gamma and the likelihood scalars ---------------------------------------
  pt.begin(3);
  const int est = beta_or_null == nullptr;
  const double beta_fixed = est ? 0.0 : beta_or_null[0];
  tri_gemv2_kernel<<<(ld + 7) / 8, 256, 0, st>>>(h->W.p, ld, ld, h->y.p, h->F.p, h->Yt.p, h->Ft.p);
  CU_TRY(cudaGetLastError());
  fit_scalars_kernel<<<1, 1024, 0, st>>>(h->A.p, ld, ld, h->Ft.p, h->Yt.p, h->scal.p);
  CU_TRY(cudaGetLastError());
  rho_kernel<<<1, 1024, 0, st>>>(h->Yt.p, h->Ft.p, ld, est, beta_fixed, h->rho.p, h->scal.p);
  CU_TRY(cudaGetLastError());
  tri_gemvT_partial_kernel<<<dim3((ld + 255) / 256, gchunks), 256, 0, st>>>(h->W.p, ld, ld, h->rho.p, h->part.p, 256);
  CU_TRY(cudaGetLastError());
  colsum_partials_kernel<<<(ld + 255) / 256, 256, 0, st>>>(h->part.p, ld, gchunks, h->gamma.p);
  CU_TRY(cudaGetLastError());
  launches += 5;
  pt.end(3);
  CU_TRY(cudaEventRecord(e1, st));

  double sc[5];
  double ft0 = 0;
  int flag = 0;
  CU_TRY(cudaMemcpyAsync(sc, h->scal.p, sizeof sc, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaMemcpyAsync(&ft0, h->Ft.p, 8, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaMemcpyAsync(&flag, h->status.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  h->fit_timings[0] = ms;
  for (int k = 0; k < 4; ++k) h->fit_timings[1 + k] = pt.total(k);
  h->fit_timings[5] = launches;

  const double ff = sc[0], logdet = sc[2];
  double rr = sc[3];
  const int p = trend_p(trend, D);
  h->p = p;
  if (trend != B200BO_TREND_CONSTANT && !flag) {
    // ---- general basis (gpr.py:800-808): Ft = L^-1 F on the device, thin QR of the (N, p) panel on the host with
    // LAPACK's Householder conventions (dgeqr2: R_jj = -sign(alpha) ||x||), rho and beta from it, gamma back on device
    CU_TRY(h->FB.reserve((size_t)ld * TR_PMAX));
    CU_TRY(h->FtG.reserve((size_t)ld * TR_PMAX));
    CU_TRY(h->FV.reserve((size_t)ld * TR_PMAX));
    CU_TRY(h->Gm.reserve((size_t)TR_PMAX * TR_PMAX));
    CU_TRY(h->betav.reserve(TR_PMAX));
    trend_basis_kernel<<<(ld * TR_PMAX + 255) / 256, 256, 0, st>>>(h->Xt.p, N, D, ld, trend, p, h->FB.p);
    CU_TRY(cudaGetLastError());
    GemmArgs g{};  // Ft(m, c) = sum_{k <= m} W(m, k) F(k, c)
    g.A = h->W.p; g.lda = ld; g.B = h->FB.p; g.ldb = TR_PMAX; g.C = h->FtG.p; g.ldc = TR_PMAX;
    g.K = ld; g.alpha = 1.0; g.beta = 0.0; g.ke_mode = 2;
    CU_TRY((launch_gemm<GemmNN, false, true>(h, g, ld, TR_PMAX, 1)));
    std::vector<double> ftp((size_t)ld * TR_PMAX), yt(ld);
    CU_TRY(cudaMemcpyAsync(ftp.data(), h->FtG.p, ftp.size() * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(yt.data(), h->Yt.p, (size_t)ld * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    h->h_Ft.assign((size_t)N * p, 0.0);
    for (int i = 0; i < N; ++i)
      for (int c = 0; c < p; ++c) h->h_Ft[(size_t)i * p + c] = ftp[(size_t)i * TR_PMAX + c];
    std::vector<double> rho(ld, 0.0);
    h->h_beta.assign(p, 0.0);
    h->h_G.assign((size_t)p * p, 0.0);
    if (est) {
      // thin QR with LAPACK's Householder conventions, beta and rho from it (csrc/host_qr.h)
      thin_qr_beta_rho(h->h_Ft.data(), yt.data(), N, p, h->h_G.data(), h->h_beta.data(), rho.data());
    } else {
      for (int c = 0; c < p; ++c) h->h_beta[c] = beta_or_null[c];
      for (int i = 0; i < N; ++i) {                                                       // gpr.py:808
        double v = yt[i];
        for (int c = 0; c < p; ++c) v -= h->h_Ft[(size_t)i * p + c] * h->h_beta[c];
        rho[i] = v;
      }
    }
    rr = 0.0;
    for (int i = 0; i < N; ++i) rr += rho[i] * rho[i];
    std::vector<double> gpad((size_t)TR_PMAX * TR_PMAX, 0.0), bpad(TR_PMAX, 0.0);
    for (int i = 0; i < p; ++i) {
      bpad[i] = h->h_beta[i];
      for (int c = 0; c < p; ++c) gpad[(size_t)i * TR_PMAX + c] = h->h_G[(size_t)i * p + c];
    }
    CU_TRY(cudaMemcpyAsync(h->rho.p, rho.data(), (size_t)ld * 8, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(h->Gm.p, gpad.data(), gpad.size() * 8, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(h->betav.p, bpad.data(), bpad.size() * 8, cudaMemcpyHostToDevice, st));
    tri_gemvT_partial_kernel<<<dim3((ld + 255) / 256, gchunks), 256, 0, st>>>(h->W.p, ld, ld, h->rho.p, h->part.p, 256);
    CU_TRY(cudaGetLastError());
    colsum_partials_kernel<<<(ld + 255) / 256, 256, 0, st>>>(h->part.p, ld, gchunks, h->gamma.p);
    CU_TRY(cudaGetLastError());
    GemmArgs v{};  // FV(n, c) = sum_{k >= n} W(k, n) Ft(k, c)
    v.A = h->W.p; v.lda = ld; v.B = h->FtG.p; v.ldb = TR_PMAX; v.C = h->FV.p; v.ldc = TR_PMAX;
    v.K = ld; v.alpha = 1.0; v.beta = 0.0; v.kb_mode = 2;
    CU_TRY((launch_gemm<GemmTN, true, true>(h, v, ld, TR_PMAX, 1)));
    CU_TRY(cudaStreamSynchronize(st));
  }
  double llf, s2, nv;
  const double two_pi = 6.283185307179586;
  if (mode == B200BO_MODE_NOISELESS) {  // gpr.py:932-945
    int k = est ? p : 0;                // rank(Q Q^T) = p for a full-rank basis
    s2 = rr / (N - k);
    nv = 0.0;
    llf = -0.5 * (N * log(two_pi * s2) + 2.0 * logdet + N);
  } else if (mode == B200BO_MODE_NOISE_ESTIM) {  // gpr.py:949-959
    double s2t = rr / N;
    s2 = par_last * s2t;
    nv = (1.0 - par_last) * s2t;
    llf = -0.5 * (N * log(two_pi * s2t) + 2.0 * logdet + N);
  } else {  // gpr.py:963-977
    s2 = par_last;
    nv = noise_var;
    double s2t = s2 + nv;
    llf = -0.5 * (N * log(two_pi * s2t) + 2.0 * logdet + rr / s2t);
    if (restricted) {
      // log_likelihood_restricted, gpr.py:849-870.  p = 1, F = ones: det(F^T F) = N, prod(diag G)^2 = Ft^T Ft.
      // The simple-kriging branch SUBTRACTS the log-determinant term (:866) -- kept as upstream.
      if (est) llf = -0.5 * ((N - 1) * log(two_pi * s2t) - log((double)N) + 2.0 * logdet + log(ff) + rr / s2t);
      else llf = -0.5 * (N * log(two_pi * s2t) - 2.0 * logdet + rr / s2t);
    }
  }
  int status = B200BO_FIT_OK;
  if (flag || llf != llf) status = B200BO_FIT_NOT_SPD;
  else if (llf > 0) status = B200BO_FIT_REJECTED;  // gpr.py:981-982
  h->restricted = restricted;
  h->last_theta.assign(theta, theta + n_theta + (corr_has_extra_param(corr) ? 1 : 0));
  h->last_noise_arg = noise_var;
  h->last_beta_fixed = est ? NAN : beta_fixed;
  h->corr = corr; h->mode = mode; h->trend = trend; h->estimate_trend = est; h->n_theta = n_theta;
  h->par_last = par_last;
  h->sigma2 = s2; h->noise_var = nv;
  h->beta = trend != B200BO_TREND_CONSTANT ? 0.0 : (est ? sc[4] : beta_fixed);  // p > 1: added by trend_mse_kernel
  // LAPACK dgeqrf sign convention for the 1x1 R factor: -sign(Ft[0]) * ||Ft||  (gpr.py:805)
  h->G = (ft0 >= 0 ? -1.0 : 1.0) * sqrt(ff);
  h->llf = status == B200BO_FIT_OK ? llf : -INFINITY;
  h->factored = status == B200BO_FIT_OK;
  if (out_llf) *out_llf = h->llf;
  if (out_sigma2) *out_sigma2 = s2;
  if (out_noise_var) *out_noise_var = nv;
  if (out_status) *out_status = status;
  return 0;
}

// ---- bordering update: [[R11, .], [R21, R22]] = [[L11, 0], [S, L22]] [[L11, 0], [S, L22]]^T with S = R21 L11^-T,
// L22 L22^T = R22 - S S^T, and the new rows of the inverse [-L22^-1 S L11^-1 | L22^-1].  m <= 64 new points: three thin
// DMMA GEMMs against L^-1 (2 x 64 N0^2 flops each) instead of the N^3 / 3 + N^3 / 3 of a fresh factor + inverse.
static int append_front(b200bo_handle h, int N0, int m, int corr, int mode, double par_last, double noise_var, int& launches) {
  cudaStream_t st = h->stream;
  const int ld = h->ld, D = h->D;
  const int N0p = round_up(N0, NB);
  CU_TRY(h->ap_Tr.reserve((size_t)NB * ld));
  CU_TRY(h->ap_Ts.reserve((size_t)NB * ld));
  CU_TRY(h->ap_Tu.reserve((size_t)NB * ld));
  CU_TRY(h->ap_Cb.reserve((size_t)NB * NB));
  CU_TRY(h->ap_Dv.reserve((size_t)NB * NB));
  AppendRowsArgs a;
  a.Xt = h->Xt.p; a.theta = h->theta.p; a.Tr = h->ap_Tr.p; a.Cb = h->ap_Cb.p;
  a.N0 = N0; a.m = m; a.D = D; a.ld = ld; a.corr = corr; a.mode = mode;
  a.sigma2 = mode == B200BO_MODE_NOISY ? par_last : 0.0;
  a.noise_var = mode == B200BO_MODE_NOISY ? noise_var : 0.0;
  a.alpha = mode == B200BO_MODE_NOISE_ESTIM ? par_last : 1.0;
  const size_t smem = ((size_t)NB * D + D + 1) * sizeof(double);
  if (smem > 48 * 1024) CU_TRY(cudaFuncSetAttribute(kmat_append_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kmat_append_rows_kernel<<<(std::max(ld, N0 + NB) + 255) / 256, 256, smem, st>>>(a);
  CU_TRY(cudaGetLastError());
  append_sym_kernel<<<(NB * NB + 255) / 256, 256, 0, st>>>(h->ap_Cb.p, m);
  CU_TRY(cudaGetLastError());
  // S(i, n) = sum_{k <= n} R21(i, k) W(n, k)
  GemmArgs g{};
  g.A = h->ap_Tr.p; g.lda = ld; g.B = h->W.p; g.ldb = ld; g.C = h->ap_Ts.p; g.ldc = ld;
  g.K = N0p; g.alpha = 1.0; g.beta = 0.0; g.ke_mode = 1;
  CU_TRY((launch_gemm<GemmNT, false, false>(h, g, NB, N0p, 1)));
  // C22 <- R22 - S S^T, then its factor and the factor's inverse
  GemmArgs c{};
  c.A = h->ap_Ts.p; c.lda = ld; c.B = h->ap_Ts.p; c.ldb = ld; c.C = h->ap_Cb.p; c.ldc = NB;
  c.K = N0p; c.alpha = -1.0; c.beta = 1.0;
  CU_TRY((launch_gemm<GemmNT, false, false>(h, c, NB, NB, 1)));
  CU_TRY(cudaFuncSetAttribute(chol_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CHOL_FACTOR_SMEM));
  CU_TRY(cudaFuncSetAttribute(chol_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CHOL_DIAG_SMEM));
  chol_factor_kernel<<<1, 256, CHOL_FACTOR_SMEM, st>>>(h->ap_Cb.p, NB, h->status.p);
  CU_TRY(cudaGetLastError());
  chol_inverse_kernel<<<1, 256, CHOL_DIAG_SMEM, st>>>(h->ap_Cb.p, NB, h->ap_Dv.p);
  CU_TRY(cudaGetLastError());
  // U(i, n) = sum_{k >= n} S(i, k) W(k, n);  W21 = -L22^-1 U
  GemmArgs u{};
  u.A = h->ap_Ts.p; u.lda = ld; u.B = h->W.p; u.ldb = ld; u.C = h->ap_Tu.p; u.ldc = ld;
  u.K = N0p; u.alpha = 1.0; u.beta = 0.0; u.kb_mode = 1;
  CU_TRY((launch_gemm<GemmNN, false, true>(h, u, NB, N0p, 1)));
  GemmArgs w{};
  w.A = h->ap_Dv.p; w.lda = NB; w.B = h->ap_Tu.p; w.ldb = ld; w.C = h->ap_Tr.p; w.ldc = ld;
  w.K = NB; w.alpha = -1.0; w.beta = 0.0;
  CU_TRY((launch_gemm<GemmNN, false, true>(h, w, NB, N0p, 1)));
  append_scatter_kernel<<<(ld + 255) / 256, 256, 0, st>>>(h->A.p, h->W.p, ld, N0, m, h->ap_Ts.p, h->ap_Tr.p, h->ap_Cb.p, h->ap_Dv.p);
  CU_TRY(cudaGetLastError());
  launches += 9;
  return 0;
}

int b200bo_append(b200bo_handle h, const double* X_new, int m_total, const double* y_all, double* out_llf, double* out_sigma2,
                  double* out_noise_var, int* out_status) {
  CHECK_ARG(h && X_new && y_all, "NULL argument");
  CHECK_ARG(m_total >= 1, "m must be positive");
  if (!h->factored) return set_err(B200BO_E_STATE, "append before a successful factor()");
  CHECK_ARG(h->trend == B200BO_TREND_CONSTANT, "append is implemented for the constant trend");
  CU_TRY(cudaSetDevice(h->device));
  const int D = h->D;
  const std::vector<double> theta = h->last_theta;
  const int n_theta_arg = (int)theta.size();
  const int corr = h->corr, mode = h->mode, trend = h->trend;
  const double par_last = h->par_last, noise_arg = h->last_noise_arg;
  const bool restricted = h->restricted;
  const bool est = h->estimate_trend != 0;
  const double beta_fixed = h->last_beta_fixed;
  int rc = 0;
  for (int done = 0; done < m_total && !rc; done += NB) {
    const int m = std::min(NB, m_total - done);
    const int N0 = h->N, N1 = N0 + m, ld0 = h->ld, ld1 = round_up(N1, 128);
    cudaStream_t st = h->stream;
    // ---- training set: host copy, transposed device copy (new pitch when ld grows), targets -----------------------
    h->Xhost.insert(h->Xhost.end(), X_new + (size_t)done * D, X_new + (size_t)(done + m) * D);
    std::vector<double> xt((size_t)D * ld1, 0.0), yy(ld1, 0.0), ff(ld1, 0.0);
    h->xmean.assign(D, 0.0);
    for (int i = 0; i < N1; ++i)
      for (int d = 0; d < D; ++d) {
        const double v = h->Xhost[(size_t)i * D + d];
        xt[(size_t)d * ld1 + i] = v;
        h->xmean[d] += v;
      }
    for (int d = 0; d < D; ++d) h->xmean[d] /= N1;
    // the targets of the rows not appended yet are not part of this step's model
    for (int i = 0; i < N1; ++i) {
      yy[i] = y_all[i];
      ff[i] = 1.0;
    }
    CU_TRY(h->Xt.reserve(xt.size()));
    CU_TRY(h->y.reserve(ld1));
    CU_TRY(h->F.reserve(ld1));
    CU_TRY(cudaMemcpyAsync(h->Xt.p, xt.data(), xt.size() * 8, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(h->y.p, yy.data(), (size_t)ld1 * 8, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(h->F.p, ff.data(), (size_t)ld1 * 8, cudaMemcpyHostToDevice, st));
    if (ld1 != ld0) {
      // the pitch changes: move L and L^-1 into larger buffers (identity on the new padding)
      DevBuf<double> nA, nW;
      CU_TRY(nA.reserve((size_t)ld1 * ld1));
      CU_TRY(nW.reserve((size_t)ld1 * ld1));
      grow_matrix_kernel<<<h->num_sms * 8, 256, 0, st>>>(h->A.p, ld0, nA.p, ld1, N0);
      grow_matrix_kernel<<<h->num_sms * 8, 256, 0, st>>>(h->W.p, ld0, nW.p, ld1, N0);
      CU_TRY(cudaGetLastError());
      CU_TRY(cudaStreamSynchronize(st));
      h->A.release(); h->W.release(); h->S.release();
      h->A = nA; h->W = nW;
      h->Kst.release();
      h->g_chol.drop(); h->g_trtri.drop();
    }
    CU_TRY(cudaStreamSynchronize(st));  // xt / yy / ff are stack-owned
    h->N = N1;
    h->ld = ld1;
    rc = factor_impl(h, corr, theta.data(), n_theta_arg, mode, par_last, noise_arg, trend, est ? nullptr : &beta_fixed, out_llf,
                     out_sigma2, out_noise_var, out_status, restricted, N0, m);
    if (!rc && out_status && *out_status != B200BO_FIT_OK) break;
  }
  return rc;
}

int b200bo_factor(b200bo_handle h, int corr, const double* theta, int n_theta, int mode, double par_last,
                  double noise_var, int trend, const double* beta_or_null, double* out_llf,
                  double* out_sigma2, double* out_noise_var, int* out_status) {
  CHECK_ARG(mode >= 0 && mode <= 2, "unknown estimation mode");
  return factor_impl(h, corr, theta, n_theta, mode, par_last, noise_var, trend, beta_or_null, out_llf, out_sigma2,
                     out_noise_var, out_status, false);
}

int b200bo_factor_restricted(b200bo_handle h, int corr, const double* theta, int n_theta, double sigma2, double noise_var,
                             int trend, const double* beta_or_null, double* out_llf, int* out_status) {
  CHECK_ARG(sigma2 > 0 && noise_var >= 0, "sigma2 must be positive and noise_var non-negative");
  // every estimation mode of log_likelihood_restricted builds R = (s2 R0 + tau2 I) / (s2 + tau2)   gpr.py:826-839
  return factor_impl(h, corr, theta, n_theta, B200BO_MODE_NOISY, sigma2, noise_var, trend, beta_or_null, out_llf,
                     nullptr, nullptr, out_status, true);
}

static int ensure_fvec(b200bo_handle h);

int b200bo_llf_grad_restricted(b200bo_handle h, double* out_grad, int n_par) {
  CHECK_ARG(h && out_grad, "NULL argument");
  if (!h->factored || !h->restricted) return set_err(B200BO_E_STATE, "llf_grad_restricted before a successful factor_restricted()");
  const int N = h->N, D = h->D, ld = h->ld, nb = ld / NB;
  CHECK_ARG(n_par >= 1 && n_par <= D + 2, "n_par out of range");
  if (!corr_has_dtheta(h->corr))
    return set_err(B200BO_E_ARG, "the reference leaves this kernel's theta-gradient unimplemented (gpr.py:758-768)");
  CU_TRY(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  int rc;
  if (h->estimate_trend && (rc = ensure_fvec(h))) return rc;
  GemmArgs g{};  // Rinv = L^-T L^-1 (lower tiles), as in b200bo_llf_grad
  g.A = h->W.p; g.B = h->W.p; g.C = h->S.p; g.lda = g.ldb = g.ldc = ld; g.K = ld; g.alpha = 1.0; g.beta = 0.0;
  g.lower_only = 1; g.kb_mode = 2;
  CU_TRY((launch_gemm<GemmTN, true, true>(h, g, ld, ld, 1)));
  const int ntiles = nb * (nb + 1) / 2, S = D + 4;
  CU_TRY(h->part.reserve((size_t)ntiles * S + 2 * S));
  const double tv = h->sigma2 + h->noise_var;
  size_t smem = ((size_t)2 * D * NB + ((D + 1) & ~1) + 2 * NB + 8 * S) * sizeof(double);
  CU_TRY(cudaFuncSetAttribute(llf_grad_traces_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  std::vector<double> r1(S), r2(S, 0.0);
  double* dout = h->part.p + (size_t)ntiles * S;
  // pass 1 (gamma):  slot d = 0.5 [ gamma^T dR0_d gamma / tv - sum(Rinv * dR0_d) ]; T1, T2, tr(Rinv), gamma^T gamma
  // pass 2 (q = L^-T Q = fvec / G, ordinary kriging only):  slot d = 0.5 tv q^T dR0_d q; T2(q), q^T q      gpr.py:880-882
  for (int pass = 0; pass < (h->estimate_trend ? 2 : 1); ++pass) {
    GradArgs a;
    a.Xt = h->Xt.p; a.theta = h->theta.p; a.Rinv = h->S.p; a.partial = h->part.p;
    a.N = N; a.D = D; a.ld = ld; a.corr = h->corr;
    if (pass == 0) { a.gamma = h->gamma.p; a.a = 1.0 / tv; a.b = 1.0; }
    else { a.gamma = h->fvec.p; a.a = tv / (h->G * h->G); a.b = 0.0; }
    llf_grad_traces_kernel<<<ntiles, 256, smem, st>>>(a);
    CU_TRY(cudaGetLastError());
    grad_reduce_kernel<<<S, 1024, 0, st>>>(h->part.p, ntiles, S, dout);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(pass == 0 ? r1.data() : r2.data(), dout, S * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
  }
  const double T1 = r1[D], T2 = r1[D + 1], trRinv = r1[D + 2], gg = r1[D + 3];
  const double G2 = h->G * h->G;
  const double qR0q = h->estimate_trend ? (2.0 * r2[D + 1] + r2[D + 3]) / G2 : 0.0;   // q^T R0 q (T2, gg of pass 2 are in fvec units)
  const double qq = h->estimate_trend ? r2[D + 3] / G2 : 0.0;
  // the (N,N,D+2) tensor of gpr.py:885-891: D theta slices (tv dR0/dtheta_d), R0, I -- indexed by the PARAMETER number
  std::vector<double> v(D + 2);
  for (int d = 0; d < D; ++d) v[d] = r1[d] + r2[d];
  v[D] = -0.5 * ((2.0 * T1 + trRinv) / tv - (2.0 * T2 + gg) / (tv * tv) - qR0q);
  v[D + 1] = -0.5 * (trRinv / tv - gg / (tv * tv) - qq);
  for (int i = 0; i < n_par; ++i) out_grad[i] = v[i];
  return 0;
}

int b200bo_llf_grad(b200bo_handle h, double* out_grad, int n_par) {
  CHECK_ARG(h && out_grad, "NULL argument");
  if (!h->factored) return set_err(B200BO_E_STATE, "llf_grad before a successful factor()");
  if (h->restricted) return set_err(B200BO_E_STATE, "the last factor() was restricted: use b200bo_llf_grad_restricted");
  const int N = h->N, D = h->D, ld = h->ld, nb = ld / NB, nt = h->n_theta;
  CHECK_ARG(n_par == nt + (h->mode == B200BO_MODE_NOISELESS ? 0 : 1), "n_par does not match the estimation mode");
  if (!corr_has_dtheta(h->corr))
    return set_err(B200BO_E_ARG, "the reference leaves this kernel's theta-gradient unimplemented (gpr.py:758-768)");
  CU_TRY(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  // Rinv = L^-T L^-1 (lower tiles): cho_solve((L, True), eye(N)) of gpr.py:997 as one TN GEMM on L^-1
  GemmArgs g{};
  g.A = h->W.p; g.B = h->W.p; g.C = h->S.p; g.lda = g.ldb = g.ldc = ld; g.K = ld; g.alpha = 1.0; g.beta = 0.0;
  g.lower_only = 1; g.kb_mode = 2;  // W(k,i) = 0 for k < i: start at the row-tile offset (i >= j)
  CU_TRY((launch_gemm<GemmTN, true, true>(h, g, ld, ld, 1)));
  const int ntiles = nb * (nb + 1) / 2, S = D + 4;
  CU_TRY(h->part.reserve((size_t)ntiles * S + S));
  double s2t = h->sigma2 + h->noise_var;
  GradArgs a;
  a.Xt = h->Xt.p; a.theta = h->theta.p; a.Rinv = h->S.p; a.gamma = h->gamma.p; a.partial = h->part.p;
  a.N = N; a.D = D; a.ld = ld; a.corr = h->corr;
  if (h->mode == B200BO_MODE_NOISELESS) { a.a = 1.0 / h->sigma2; a.b = 1.0; }              // gpr.py:1002-1010
  else if (h->mode == B200BO_MODE_NOISE_ESTIM) { a.a = h->par_last / s2t; a.b = h->par_last; }  // :1011-1019
  else { a.a = 1.0 / s2t; a.b = 1.0; }                                                       // :1026-1038 (quirk g2)
  size_t smem = ((size_t)2 * D * NB + ((D + 1) & ~1) + 2 * NB + 8 * S) * sizeof(double);
  CU_TRY(cudaFuncSetAttribute(llf_grad_traces_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  llf_grad_traces_kernel<<<ntiles, 256, smem, st>>>(a);
  CU_TRY(cudaGetLastError());
  double* dout = h->part.p + (size_t)ntiles * S;
  grad_reduce_kernel<<<S, 1024, 0, st>>>(h->part.p, ntiles, S, dout);
  CU_TRY(cudaGetLastError());
  std::vector<double> r(S);
  CU_TRY(cudaMemcpyAsync(r.data(), dout, S * 8, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  const double T1 = r[D], T2 = r[D + 1], trRinv = r[D + 2], gg = r[D + 3];
  double last = 0.0;
  if (h->mode == B200BO_MODE_NOISE_ESTIM) {
    // -0.5 (sum(Rinv * (R0 - I)) - gamma^T (R0 - I) gamma / s2t): off-diagonal pairs only   gpr.py:1021-1025
    last = -(T1 - T2 / s2t);
  } else if (h->mode == B200BO_MODE_NOISY) {
    // -0.5 (sum(Cinv * R0) - gamma_^T R0 gamma_), Cinv = Rinv / s2t, gamma_ = gamma / s2t    gpr.py:1027-1037
    last = -0.5 * ((2.0 * T1 + trRinv) / s2t - (2.0 * T2 + gg) / (s2t * s2t));
  }
  if (nt == D) {
    for (int d = 0; d < D; ++d) out_grad[d] = r[d];
    if (n_par > nt) out_grad[nt] = last;
  } else {
    // Isotropic theta with D > 1.  The reference indexes its (N,N,D[+1]) gradient tensor with the PARAMETER
    // index (gpr.py:1004-1005, :1034-1035), so parameter 0 sees only the feature-0 slice, and in "noisy" mode
    // parameter 1 (sigma2) sees the feature-1 slice instead of R0.  Reproduced as is.
    out_grad[0] = r[0];
    if (n_par > 1) out_grad[1] = (h->mode == B200BO_MODE_NOISY && D > 1) ? r[1] : last;
  }
  return 0;
}

int b200bo_get_state(b200bo_handle h, int what, double* out, size_t n_elems) {
  CHECK_ARG(h && out, "NULL argument");
  if (!(h->factored || (what == B200BO_STATE_R && h->keepR && h->Rkeep.p)))
    return set_err(B200BO_E_STATE, "no successful factor() yet");
  CU_TRY(cudaSetDevice(h->device));
  const int N = h->N, ld = h->ld;
  auto copy_mat = [&](const double* src) -> int {
    CHECK_ARG(n_elems == (size_t)N * N, "expected N*N elements");
    CU_TRY(cudaMemcpy2DAsync(out, (size_t)N * 8, src, (size_t)ld * 8, (size_t)N * 8, N, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    return 0;
  };
  auto copy_vec = [&](const double* src) -> int {
    CHECK_ARG(n_elems == (size_t)N, "expected N elements");
    CU_TRY(cudaMemcpyAsync(out, src, (size_t)N * 8, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(cudaStreamSynchronize(h->stream));
    return 0;
  };
  switch (what) {
    case B200BO_STATE_L: return copy_mat(h->A.p);
    case B200BO_STATE_LINV: return copy_mat(h->W.p);
    case B200BO_STATE_R: return copy_mat(h->Rkeep.p);
    case B200BO_STATE_GAMMA: return copy_vec(h->gamma.p);
    case B200BO_STATE_YT: return copy_vec(h->Yt.p);
    case B200BO_STATE_FT:
      if (h->trend != B200BO_TREND_CONSTANT) {
        CHECK_ARG(n_elems == (size_t)N * h->p, "expected N*p elements");
        for (size_t i = 0; i < n_elems; ++i) out[i] = h->h_Ft[i];
        return 0;
      }
      return copy_vec(h->Ft.p);
    case B200BO_STATE_RHO: return copy_vec(h->rho.p);
    case B200BO_STATE_BETA:
      if (h->trend != B200BO_TREND_CONSTANT) {
        CHECK_ARG(n_elems == (size_t)h->p, "expected p elements");
        for (int i = 0; i < h->p; ++i) out[i] = h->h_beta[i];
        return 0;
      }
      CHECK_ARG(n_elems == 1, "expected 1 element");
      out[0] = h->beta;
      return 0;
    case B200BO_STATE_G:
      if (h->trend != B200BO_TREND_CONSTANT) {
        CHECK_ARG(n_elems == (size_t)h->p * h->p, "expected p*p elements");
        for (size_t i = 0; i < n_elems; ++i) out[i] = h->h_G[i];
        return 0;
      }
      CHECK_ARG(n_elems == 1, "expected 1 element");
      out[0] = h->G;
      return 0;
  }
  return set_err(B200BO_E_ARG, "unknown state id");
}

// ------------------------------------------------------------------------------------------------------
// fp64 building block: moments of m <= Mc device-resident candidates -> yh (m), h->sumsq, h->dotf
// ------------------------------------------------------------------------------------------------------
struct LaunchCount {
  int all = 0, contract = 0;
};
static const int RS_SMALL_MAX = 1024;  // candidates up to which the row-parallel contraction is used

static int fp64_moments(b200bo_handle h, const double* xc_dev, int m, double* yh, int eval_mse, PhaseTimer* pt,
                        LaunchCount* lc) {
  cudaStream_t st = h->stream;
  const int D = h->D, ld = h->ld;
  const int mpad = round_up(m, PC_BM);
  if (!h->contract_attr) {  // per handle: the attribute belongs to the handle's device
    CU_TRY(cudaFuncSetAttribute(contract_fp64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PredCore::SMEM_BYTES));
    h->contract_attr = true;
  }
  const size_t ks_smem = ((size_t)KS_ROWS * D + D) * sizeof(double);
  if (ks_smem > 48 * 1024)
    CU_TRY(cudaFuncSetAttribute(kstar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ks_smem));
  if (pt) pt->begin(0);
  KstarArgs k;
  k.Xc = xc_dev; k.Xt = h->Xt.p; k.theta = h->theta.p; k.gamma = h->gamma.p;
  k.Kst = eval_mse ? h->Kst.p : nullptr; k.yhat = yh;
  k.M = m; k.N = h->N; k.D = D; k.ld = ld; k.corr = h->corr; k.beta = h->beta;
  kstar_kernel<<<mpad / KS_ROWS, 256, ks_smem, st>>>(k);
  CU_TRY(cudaGetLastError());
  ++lc->all;
  if (pt) pt->end(0);
  if (eval_mse && m <= RS_SMALL_MAX) {
    // few candidates: one balanced sweep over the rows of L^-1 instead of one CTA per 128-candidate tile
    if (pt) pt->begin(1);
    const int nblk = ld / RS_ROWS, nctas = (nblk + 1) / 2, chunks = (m + RS_BC - 1) / RS_BC;
    CU_TRY(h->rs_part.reserve((size_t)chunks * nctas * 2 * RS_BC));
    RescoreArgs r;
    r.Kst = h->Kst.p; r.Linv = h->W.p; r.Ft = h->Ft.p; r.part = h->rs_part.p; r.ld = ld; r.M = m;
    rs_contract_kernel<<<dim3(nctas, chunks), 256, 0, st>>>(r);
    CU_TRY(cudaGetLastError());
    rs_reduce_kernel<<<(m + 127) / 128, 128, 0, st>>>(h->rs_part.p, nctas, m, h->sumsq.p, h->dotf.p);
    CU_TRY(cudaGetLastError());
    lc->all += 2;
    ++lc->contract;
    if (pt) pt->end(1);
  } else if (eval_mse) {
    if (pt) pt->begin(1);
    ContractArgs c;
    c.Kst = h->Kst.p; c.Linv = h->W.p; c.Ft = h->Ft.p; c.sumsq = h->sumsq.p; c.dotf = h->dotf.p; c.ld = ld;
    contract_fp64_kernel<<<mpad / PC_BM, PredCore::NT, PredCore::SMEM_BYTES, st>>>(c);
    CU_TRY(cudaGetLastError());
    ++lc->all;
    ++lc->contract;
    if (pt) pt->end(1);
  }
  return 0;
}

// acquisition + arg-max of m candidates whose moments are in (yh, h->sumsq, h->dotf); merges into h->best_*
static int acq_stage(b200bo_handle h, const double* yh, int m, long long idx_base, const long long* idx_map,
                     int acq_id, int minimize, double plugin, int q, double* vals, long long vals_ld,
                     long long vals_off, double* mse_out, LaunchCount* lc, const double* mse_in = nullptr) {
  cudaStream_t st = h->stream;
  AcqArgs g{};
  g.yhat = yh; g.sumsq = h->sumsq.p; g.dotf = h->dotf.p; g.mse_in = mse_in; g.mse_out = mse_in ? nullptr : mse_out;
  g.M = m; g.estimate_trend = h->estimate_trend; g.sigma2 = h->sigma2; g.G = h->G;
  g.idx_base = idx_base; g.idx_map = idx_map;
  g.vals = vals; g.vals_ld = vals_ld; g.vals_off = vals_off;
  g.params = h->params.p; g.part_val = h->part_val.p; g.part_idx = h->part_idx.p;
  g.acq = acq_id; g.minimize = minimize; g.q = q; g.plugin = plugin;
  int nblk = std::min(h->num_sms, (m + 255) / 256);
  acq_kernel<<<dim3(nblk, q), 256, 0, st>>>(g);
  CU_TRY(cudaGetLastError());
  argmax_merge_kernel<<<(q + 63) / 64, 64, 0, st>>>(h->part_val.p, h->part_idx.p, nblk, q, h->best_val.p, h->best_idx.p);
  CU_TRY(cudaGetLastError());
  lc->all += 2;
  return 0;
}

// p > 1 trend: yhat += f(x)^T beta, MSE with the p-vector u (gpr.py:490, :496-510); mse_dev may be NULL (mean only)
static int trend_finish(b200bo_handle h, const double* xc_dev, int m, double* yh, double* mse_dev, LaunchCount* lc) {
  cudaStream_t st = h->stream;
  const int ld = h->ld, mpad = round_up(m, PC_BM);
  if (mse_dev && h->estimate_trend) {
    CU_TRY(h->Vt.reserve((size_t)round_up(h->Mc, PC_BM) * TR_PMAX));
    GemmArgs g{};  // Vt(m, c) = sum_k r(m, k) FV(k, c) = (Ft^T rt)_c
    g.A = h->Kst.p; g.lda = ld; g.B = h->FV.p; g.ldb = TR_PMAX; g.C = h->Vt.p; g.ldc = TR_PMAX;
    g.K = ld; g.alpha = 1.0; g.beta = 0.0;
    CU_TRY((launch_gemm<GemmNN, false, true>(h, g, mpad, TR_PMAX, 1)));
    ++lc->all;
  }
  TrendMseArgs a;
  a.Xc = xc_dev; a.Vt = h->Vt.p; a.sumsq = h->sumsq.p; a.Gm = h->Gm.p; a.beta = h->betav.p; a.yhat = yh; a.mse = mse_dev;
  a.M = m; a.D = h->D; a.trend = h->trend; a.p = h->p; a.estimate_trend = h->estimate_trend; a.sigma2 = h->sigma2;
  const size_t smem = ((size_t)h->p * h->p + h->p) * sizeof(double);
  trend_mse_kernel<<<(m + 127) / 128, 128, smem, st>>>(a);
  CU_TRY(cudaGetLastError());
  ++lc->all;
  return 0;
}

static int upload_params_reset_best(b200bo_handle h, const double* params, int q) {
  cudaStream_t st = h->stream;
  std::vector<double> pr(q);
  for (int c = 0; c < q; ++c) pr[c] = params ? params[c] : 0.0;
  CU_TRY(cudaMemcpyAsync(h->params.p, pr.data(), q * 8, cudaMemcpyHostToDevice, st));
  std::vector<long long> bi(q, -1);
  CU_TRY(cudaMemsetAsync(h->best_val.p, 0, q * 8, st));
  CU_TRY(cudaMemcpyAsync(h->best_idx.p, bi.data(), q * 8, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaStreamSynchronize(st));  // pr / bi are stack-owned
  return 0;
}

static int download_best(b200bo_handle h, int q, double* best_val, int64_t* best_idx) {
  cudaStream_t st = h->stream;
  std::vector<long long> bi(q);
  CU_TRY(cudaMemcpyAsync(best_val, h->best_val.p, q * 8, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaMemcpyAsync(bi.data(), h->best_idx.p, q * 8, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  for (int c = 0; c < q; ++c) best_idx[c] = bi[c];
  return 0;
}

// Shared driver of predict / acq on the fp64 path.  acq_id < 0: predict only.
static int run_candidates_fp64(b200bo_handle h, const double* Xc, int64_t M, int loc, int eval_mse, double* yhat_out,
                               double* mse_out, int acq_id, int minimize, double plugin, const double* params, int q,
                               double* vals, double* best_val, int64_t* best_idx) {
  const bool do_acq = acq_id >= 0;
  const bool dev = loc == B200BO_DEVICE;
  int rc = ensure_predict_ws(h, q, do_acq && vals && !dev);
  if (rc) return rc;
  cudaStream_t st = h->stream;
  const int D = h->D, Mc = h->Mc;
  h->evs.reset();
  PhaseTimer pt{h};
  LaunchCount lc;
  if (do_acq && (rc = upload_params_reset_best(h, params, q))) return rc;
  cudaEvent_t e0 = h->evs.get(), e1 = h->evs.get();
  CU_TRY(cudaEventRecord(e0, st));
  for (int64_t a = 0; a < M; a += Mc) {
    const int m = (int)std::min<int64_t>(Mc, M - a);
    const double* xc = Xc + (size_t)a * D;
    if (!dev) {
      CU_TRY(cudaMemcpyAsync(h->Xc.p, xc, (size_t)m * D * 8, cudaMemcpyHostToDevice, st));
      xc = h->Xc.p;
    }
    double* yh = (dev && yhat_out) ? yhat_out + a : h->yhat.p;
    if ((rc = fp64_moments(h, xc, m, yh, eval_mse, &pt, &lc))) return rc;
    const bool gen_trend = h->trend != B200BO_TREND_CONSTANT;
    double* mt = nullptr;  // p > 1 trend: the finished MSE of this chunk
    if (gen_trend) {
      mt = eval_mse ? ((dev && mse_out) ? mse_out + a : h->mse.p) : nullptr;
      if ((rc = trend_finish(h, xc, m, yh, mt, &lc))) return rc;
    }
    if (eval_mse) {
      pt.begin(2);
      double* mo = mse_out ? (dev ? mse_out + a : h->mse.p) : nullptr;
      if (do_acq) {
        double* vp = vals ? (dev ? vals : h->vals.p) : nullptr;
        if ((rc = acq_stage(h, yh, m, a, nullptr, acq_id, minimize, plugin, q, vp, dev ? M : Mc, dev ? a : 0, mo, &lc, mt)))
          return rc;
      } else if (gen_trend) {
        // trend_finish wrote the MSE already
      } else {
        AcqArgs g{};
        g.yhat = yh; g.sumsq = h->sumsq.p; g.dotf = h->dotf.p; g.mse_out = mo;
        g.M = m; g.estimate_trend = h->estimate_trend; g.sigma2 = h->sigma2; g.G = h->G;
        mse_kernel<<<std::min(h->num_sms, (m + 255) / 256), 256, 0, st>>>(g);
        CU_TRY(cudaGetLastError());
        ++lc.all;
      }
      pt.end(2);
    }
    if (!dev) {
      if (yhat_out) CU_TRY(cudaMemcpyAsync(yhat_out + a, h->yhat.p, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
      if (eval_mse && mse_out) CU_TRY(cudaMemcpyAsync(mse_out + a, h->mse.p, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
      if (do_acq && vals)
        CU_TRY(cudaMemcpy2DAsync(vals + a, (size_t)M * 8, h->vals.p, (size_t)Mc * 8, (size_t)m * 8, q, cudaMemcpyDeviceToHost, st));
    }
  }
  CU_TRY(cudaEventRecord(e1, st));
  if (do_acq && (rc = download_best(h, q, best_val, best_idx))) return rc;
  CU_TRY(cudaStreamSynchronize(st));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  h->timings[0] = ms;
  for (int k = 0; k < 3; ++k) h->timings[1 + k] = pt.total(k);
  h->timings[4] = lc.contract;
  h->timings[5] = lc.all;
  h->timings[6] = 0;
  h->timings[7] = 0;
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// tensor-core path (B200BO_PREC_FAST)
// ------------------------------------------------------------------------------------------------------
static int make_f16_map(CUtensorMap* map, const __half* base, int inner, int rows, int box_inner, int box_rows);
static int make_linv_map(CUtensorMap* map, const __half* base, int ld) {
  return make_f16_map(map, base, ld, ld, fk::KC, fk::BN);
}
// 2-D fp16 row-major tensor (rows x inner), 128-byte-swizzled boxes of box_rows x box_inner
static int make_f16_map(CUtensorMap* map, const __half* base, int inner, int rows, int box_inner, int box_rows) {
  return make_map_2d(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, base, (uint64_t)inner, (uint64_t)rows, (uint64_t)inner * sizeof(__half),
                     box_inner, box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
}

static void build_fast_model(b200bo_ctx* h, const std::vector<double>& rowsq, const std::vector<double>& rowl1,
                             double s2_sq, double gamma_l2, double f_l2, double b_max);

static bool fast_supported(const b200bo_ctx* h) { return h->corr != CUBIC && !corr_has_extra_param(h->corr) && h->D <= 64; }

static int ensure_fast_state(b200bo_handle h) {
  if (h->fast_ready) return 0;
  cudaStream_t st = h->stream;
  const int D = h->D, ld = h->ld;
  const size_t nn = (size_t)ld * ld;
  h->DP = D <= 8 ? 8 : D <= 16 ? 16 : D <= 32 ? 32 : 64;
  // per-feature coordinate scale folding theta (and log2 e / 2 nu) into the coordinates
  std::vector<double> th(D), cs(D);
  CU_TRY(cudaMemcpyAsync(th.data(), h->theta.p, D * 8, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  const double LOG2E = 1.4426950408889634;
  for (int d = 0; d < D; ++d) {
    switch (h->corr) {
      case RBF: cs[d] = sqrt(th[d] * LOG2E); break;
      case MATERN12: cs[d] = sqrt(th[d]); break;
      case MATERN32: cs[d] = sqrt(3.0 * th[d]); break;
      case MATERN52: cs[d] = sqrt(5.0 * th[d]); break;
      default: cs[d] = th[d] * LOG2E; break;  // ABSEXP
    }
  }
  CU_TRY(h->cscale.reserve(D));
  CU_TRY(cudaMemcpyAsync(h->cscale.p, cs.data(), D * 8, cudaMemcpyHostToDevice, st));
  // f = L^-T Ft, so that Ft^T (L^-1 r) = r . f  (gpr.py:498 as a dot product against r)
  CU_TRY(h->fvec.reserve(ld));
  const int gchunks = (ld + 255) / 256;
  CU_TRY(h->part.reserve((size_t)gchunks * ld));
  tri_gemvT_partial_kernel<<<dim3((ld + 255) / 256, gchunks), 256, 0, st>>>(h->W.p, ld, ld, h->Ft.p, h->part.p, 256);
  CU_TRY(cudaGetLastError());
  colsum_partials_kernel<<<(ld + 255) / 256, 256, 0, st>>>(h->part.p, ld, gchunks, h->fvec.p);
  CU_TRY(cudaGetLastError());
  // power-of-two scale of L^-1 into the fp16 range, then the (hi, lo) split
  const int nblk = h->num_sms * 4;
  CU_TRY(h->errout.reserve(std::max(nblk, 8)));
  fk::absmax_kernel<<<nblk, 256, 0, st>>>(h->W.p, nn, h->errout.p);
  CU_TRY(cudaGetLastError());
  std::vector<double> pm(nblk);
  CU_TRY(cudaMemcpyAsync(pm.data(), h->errout.p, nblk * 8, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  double mx = 0;
  for (double v : pm) mx = std::max(mx, v);
  if (!(mx > 0) || !std::isfinite(mx)) return set_err(B200BO_E_STATE, "L^-1 has no finite non-zero entry");
  h->b_scale_log2 = 14 - ilogb(mx);
  CU_TRY(h->Lh.reserve(nn));
  CU_TRY(h->Ll.reserve(nn));
  fk::linv_split_kernel<<<h->num_sms * 8, 256, 0, st>>>(h->W.p, nn, ldexp(1.0, h->b_scale_log2), h->Lh.p, h->Ll.p);
  CU_TRY(cudaGetLastError());
  CU_TRY(h->Xs.reserve((size_t)(h->DP + 2) * ld));
  fk::xs_prep_kernel<<<(ld + 127) / 128, 128, 0, st>>>(h->Xt.p, h->cscale.p, h->gamma.p, h->fvec.p, D, h->DP, ld, h->Xs.p);
  CU_TRY(cudaGetLastError());
  int rc;
  if ((rc = make_linv_map(&h->map_hi, h->Lh.p, ld))) return rc;
  if ((rc = make_linv_map(&h->map_lo, h->Ll.p, ld))) return rc;
  h->use_v2 = h->fast_kernel_pref >= 2 && h->corr != ABSEXP;
#ifndef B200BO_DEV_KERNELS
  if (h->num_sms % 2) h->use_v2 = false;  // the single-CTA Gram kernel (generation 2) is a developer-build kernel: generation 1 instead
#endif
  if (h->use_v2) {
    CU_TRY(h->cmean.reserve(D));
    CU_TRY(cudaMemcpyAsync(h->cmean.p, h->xmean.data(), D * 8, cudaMemcpyHostToDevice, st));
    CU_TRY(h->Xh2.reserve((size_t)ld * 64));
    CU_TRY(h->Xl2.reserve((size_t)ld * 64));
    CU_TRY(h->aux2.reserve((size_t)(ld / fk::KC) * 3 * fk::KC));
    CU_TRY(h->exch2.reserve((size_t)h->num_sms * 3 * fk::BM));
    fk2::xs2_prep_kernel<<<(ld + 127) / 128, 128, 0, st>>>(h->Xt.p, h->cscale.p, h->cmean.p, h->gamma.p, h->fvec.p,
                                                           h->N, D, ld, h->Xh2.p, h->Xl2.p, h->aux2.p);
    CU_TRY(cudaGetLastError());
    if ((rc = make_f16_map(&h->map2_hi, h->Lh.p, ld, ld, fk::KC, fk2::NB))) return rc;
    if ((rc = make_f16_map(&h->map2_lo, h->Ll.p, ld, ld, fk::KC, fk2::NB))) return rc;
    if ((rc = make_f16_map(&h->map2_xh, h->Xh2.p, 64, ld, 64, fk::KC))) return rc;
    if ((rc = make_f16_map(&h->map2_xl, h->Xl2.p, 64, ld, 64, fk::KC))) return rc;
    h->use_pair = h->fast_kernel_pref >= 3 && (h->num_sms % 2 == 0);
    if (h->use_pair) {
      h->pair_maps.hi128 = h->map2_hi;
      h->pair_maps.lo128 = h->map2_lo;
      if ((rc = make_f16_map(&h->pair_maps.hi64, h->Lh.p, ld, ld, fk::KC, 64))) return rc;
      if ((rc = make_f16_map(&h->pair_maps.lo64, h->Ll.p, ld, ld, fk::KC, 64))) return rc;
      if ((rc = make_f16_map(&h->pair_maps.xh32, h->Xh2.p, 64, ld, 64, 32))) return rc;
      if ((rc = make_f16_map(&h->pair_maps.xl32, h->Xl2.p, 64, ld, 64, 32))) return rc;
      // r scratch of the replay kernel: (CTA, chunk, plane) blocks of 128 rows x 64 fp16, capped by the budget
      const size_t blk = (size_t)fk::BM * fk::KC;  // halves per block (16 KB)
      const size_t want = (size_t)h->num_sms * (ld / fk::KC) * 2 * blk;
      const size_t cap = (size_t)h->replay_mb * (1u << 20) / sizeof(__half) / blk * blk;
      // generation 5 keeps every chunk of a tile (both planes): no budget, the scratch may spill out of L2
      h->use_decoupled = h->fast_kernel_pref >= 5 && ld >= 512 && want * sizeof(__half) <= ((size_t)4 << 30);
      const size_t n_halves = h->use_decoupled ? want : std::min(want, cap);
      h->use_replay = h->fast_kernel_pref >= 4 && n_halves >= (size_t)h->num_sms * 2 * blk;
      h->use_shared = false;
      // generation 6 pays where a tile is long: N = 4096 + 4 - 6 %, N <= 2048 slower than generation 5 (the hand-over between
      // the two sides costs a fixed ~3 us per chunk class and tile; profiles/r02/gen6_vs_gen5_by_N.txt)
      if (h->fast_kernel_pref >= 6 && h->use_decoupled && ld % 256 == 0 && ld >= h->gen6_min_ld && h->num_sms % 4 == 0) {
        if (h->shared_ok < 0) {
          // the partner pairs of generation 6 poll each other: every CTA of the grid has to be resident at once
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(h->num_sms); cfg.blockDim = dim3(fk6::NT6); cfg.dynamicSmemBytes = fk6::SMEM6_BYTES;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeClusterDimension;
          at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
          cfg.attrs = at; cfg.numAttrs = 1;
          auto kern = fk6::predict_fused_shared_kernel<MATERN52, 1, false>;
          CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, fk6::SMEM6_BYTES));
          int ncl = 0;
          cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg);
          h->shared_ok = (e == cudaSuccess && ncl * 2 >= h->num_sms) ? 1 : 0;
          if (e != cudaSuccess) cudaGetLastError();
        }
        h->use_shared = h->shared_ok == 1;
      }
      if (h->use_replay) {
        CU_TRY(h->r_scratch.reserve(n_halves));
        h->replay_maps.pm = h->pair_maps;
        if ((rc = make_f16_map(&h->replay_maps.scr, h->r_scratch.p, fk::KC, (int)(h->r_scratch.n / fk::KC), fk::KC, fk::BM))) return rc;
      }
    }
  }
  CU_TRY(h->err_flag.reserve(1));
  CU_TRY(cudaMemsetAsync(h->err_flag.p, 0, sizeof(int), st));
  {
    // ---- statistics of L^-1, gamma, f for the a-priori error model of the pass (build_fast_model) ----------------
    CU_TRY(h->rowsq.reserve(ld));
    CU_TRY(h->rowl1.reserve(ld));
    CU_TRY(h->pm_v.reserve(ld));
    CU_TRY(h->pm_t.reserve(ld));
    CU_TRY(h->pm_t2.reserve(ld));
    bd::linv_rowstats_kernel<<<(ld + 7) / 8, 256, 0, st>>>(h->W.p, ld, ld, h->rowsq.p, h->rowl1.p);
    CU_TRY(cudaGetLastError());
    // ||L^-1||_2^2 = largest eigenvalue of W^T W: power iteration, unnormalised (||W||_2 >= 1 because R has a unit
    // diagonal, so the iterates cannot underflow; ten steps stay far below the double range for any R that factored)
    const int PM_ITERS = 10;
    bd::pm_init_kernel<<<(ld + 255) / 256, 256, 0, st>>>(h->pm_v.p, h->N, ld);
    CU_TRY(cudaGetLastError());
    for (int it = 0; it < PM_ITERS; ++it) {
      if (it == PM_ITERS - 1) bd::sumsq_kernel<<<1, 1024, 0, st>>>(h->pm_v.p, ld, h->errout.p, 0);
      tri_gemv2_kernel<<<(ld + 7) / 8, 256, 0, st>>>(h->W.p, ld, ld, h->pm_v.p, h->pm_v.p, h->pm_t.p, h->pm_t2.p);
      tri_gemvT_partial_kernel<<<dim3((ld + 255) / 256, gchunks), 256, 0, st>>>(h->W.p, ld, ld, h->pm_t.p, h->part.p, 256);
      colsum_partials_kernel<<<(ld + 255) / 256, 256, 0, st>>>(h->part.p, ld, gchunks, h->pm_v.p);
    }
    bd::sumsq_kernel<<<1, 1024, 0, st>>>(h->pm_v.p, ld, h->errout.p, 1);
    bd::sumsq_kernel<<<1, 1024, 0, st>>>(h->gamma.p, ld, h->errout.p, 2);
    bd::sumsq_kernel<<<1, 1024, 0, st>>>(h->fvec.p, ld, h->errout.p, 3);
    CU_TRY(cudaGetLastError());
    std::vector<double> rsq(ld), rl1(ld);
    double e4[4];
    CU_TRY(cudaMemcpyAsync(rsq.data(), h->rowsq.p, (size_t)ld * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(rl1.data(), h->rowl1.p, (size_t)ld * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(e4, h->errout.p, sizeof e4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    double s2_sq = (e4[0] > 0 && std::isfinite(e4[0]) && std::isfinite(e4[1])) ? sqrt(e4[1] / e4[0]) : INFINITY;
    // largest squared norm of a centred training point in the kernel's units (cs folds theta and the kernel's constant)
    double b_max = 0.0;
    for (int i = 0; i < h->N; ++i) {
      double b = 0.0;
      for (int d = 0; d < D; ++d) {
        const double v = (h->Xhost[(size_t)i * D + d] - h->xmean[d]) * cs[d];
        b += h->corr == ABSEXP ? fabs(v) : v * v;
      }
      b_max = std::max(b_max, b);
    }
    build_fast_model(h, rsq, rl1, s2_sq, sqrt(e4[2]), sqrt(e4[3]), b_max);
  }
  if (!h->copy_stream) {
    CU_TRY(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      CU_TRY(cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming));
      CU_TRY(cudaEventCreateWithFlags(&h->ev_used[i], cudaEventDisableTiming));
    }
  }
  CU_TRY(cudaStreamSynchronize(st));
  h->fast_ready = true;
  h->calibrated[0] = h->calibrated[1] = false;
  return 0;
}

// one launch of the fused tensor-core kernel over m device-resident candidates; outputs at out_off
static int launch_fused(b200bo_handle h, const double* xc_dev, long long m, size_t out_off, int nprod = 3) {
  if (h->use_v2) {
    fk2::Fused2Args a;
    a.Xc = xc_dev; a.cscale = h->cscale.p; a.cmean = h->cmean.p; a.aux = h->aux2.p;
    a.yhat = h->f_yhat.p + out_off; a.sumsq = h->f_sumsq.p + out_off; a.dotf = h->f_dotf.p + out_off;
    a.exch = h->exch2.p;
    a.dbg_w = h->want_dbg_w ? h->dbg_w.p : nullptr;
    a.trace = nullptr;
    if (getenv("B200BO_TRACE")) {
      CU_TRY(h->trace.reserve((size_t)fk5::TRACE5_CHUNKS * 8));
      CU_TRY(cudaMemsetAsync(h->trace.p, 0, (size_t)fk5::TRACE5_CHUNKS * 64, h->stream));
      a.trace = h->trace.p;
    }
    a.err = h->err_flag.p;
    a.M = m; a.N = h->N; a.D = h->D; a.ld = h->ld; a.corr = h->corr; a.dk_steps = (h->D + 15) / 16; a.beta = h->beta;
    a.out_scale = (float)ldexp(1.0, -(fk::A_SCALE_LOG2 + h->b_scale_log2));
    const long long tiles = (m + fk::BM - 1) / fk::BM;
    const int grid = h->use_pair ? (int)std::min<long long>(h->num_sms, 2 * ((tiles + 1) / 2))
                                 : (int)std::min<long long>(h->num_sms, tiles);
#ifdef B200BO_DEV_KERNELS  /* generations 2 and 3 are superseded: developer builds only (A/B runs) */
#define FK3_LAUNCH(C)                                                                                               \
  do {                                                                                                             \
    auto kern = nprod == 1 ? fk3::predict_fused_pair_kernel<C, 1> : fk3::predict_fused_pair_kernel<C, 3>;           \
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, fk3::SMEM_BYTES));              \
    kern<<<grid, fk2::NT2, fk3::SMEM_BYTES, h->stream>>>(h->pair_maps, a);                                          \
  } while (0)
#else
#define FK3_LAUNCH(C) return set_err(B200BO_E_STATE, "generation 3 of the fused kernel is only in developer builds (-DB200BO_DEV_KERNELS)")
#endif
#define FK4_LAUNCH(C)                                                                                               \
  do {                                                                                                             \
    auto kern = nprod == 1 ? fk4::predict_fused_replay_kernel<C, 1> : fk4::predict_fused_replay_kernel<C, 3>;       \
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, fk3::SMEM_BYTES));              \
    kern<<<grid, fk2::NT2, fk3::SMEM_BYTES, h->stream>>>(h->replay_maps, a, ra);                                    \
  } while (0)
#define FK5_LAUNCH(C)                                                                                               \
  do {                                                                                                             \
    auto kern = nprod == 1 ? fk5::predict_fused_decoupled_kernel<C, 1, false> : fk5::predict_fused_decoupled_kernel<C, 3, false>; \
    if (C == MATERN52 && nprod == 1 && a.trace) kern = fk5::predict_fused_decoupled_kernel<MATERN52, 1, true>; /* developer timeline */ \
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, fk3::SMEM_BYTES));              \
    kern<<<grid, fk2::NT2, fk3::SMEM_BYTES, h->stream>>>(h->replay_maps, a, ra);                                    \
  } while (0)
#ifdef B200BO_DEV_KERNELS
#define FK2_LAUNCH(C)                                                                                               \
  do {                                                                                                             \
    auto kern = nprod == 1 ? fk2::predict_fused_tc2_kernel<C, 1> : fk2::predict_fused_tc2_kernel<C, 3>;             \
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, fk2::SMEM_BYTES));              \
    kern<<<grid, fk2::NT2, fk2::SMEM_BYTES, h->stream>>>(h->map2_hi, h->map2_lo, h->map2_xh, h->map2_xl, a);        \
  } while (0)
#else
#define FK2_LAUNCH(C) return set_err(B200BO_E_STATE, "generation 2 of the fused kernel is only in developer builds (-DB200BO_DEV_KERNELS)")
#endif
    if (h->use_pair && h->use_replay && h->use_decoupled && h->use_shared) {
      h->last_gen = 6;
      fk4::ReplayArgs ra;
      ra.scratch = h->r_scratch.p;
      const int kdb6 = std::max(0, std::min(h->gen6_db_chunks, h->ld / fk::KC)) & ~3;
      ra.n_store = h->ld / fk::KC + kdb6;   // slots per 128-candidate half (the scratch is sized for generation 5: 2 x nch per pair)
      h->last_n_store = ra.n_store;
      // deal the accumulator super-tiles to the two sides of a group: largest first, to the side that carries less
      const int n_super = (h->ld + fk2::WC - 1) / fk2::WC;
      uint32_t mask = 0;
      long long load[2] = {0, 0};
      for (int s = n_super - 1; s >= 0; --s) {
        const int uses = std::min(h->ld, fk2::WC * (s + 1)) / fk::KC;
        const int side = load[1] < load[0] ? 1 : 0;
        load[side] += uses;
        if (side) mask |= 1u << s;
      }
      fk6::ShareArgs sa;
      const long long ptiles = (tiles + 1) / 2;
      const int groups = (int)std::min<long long>(std::min(h->num_sms, h->fast_max_sms) / 4, ptiles);
      const int grid6 = 4 * groups;
      const size_t nflags = (size_t)groups * 2 * 2 * 3 * 8;
      CU_TRY(h->share_flags.reserve((size_t)(h->num_sms / 4) * 2 * 2 * 3 * 8));
      CU_TRY(cudaMemsetAsync(h->share_flags.p, 0, nflags * sizeof(uint32_t), h->stream));
      // the two sides ADD their parts into the outputs
      const size_t mp = (size_t)tiles * fk::BM;
      CU_TRY(cudaMemsetAsync(a.yhat, 0, mp * 8, h->stream));
      CU_TRY(cudaMemsetAsync(a.sumsq, 0, mp * 8, h->stream));
      CU_TRY(cudaMemsetAsync(a.dotf, 0, mp * 8, h->stream));
      sa.flags = h->share_flags.p;
      sa.side_mask = mask;
      sa.dead_hint = getenv("B200BO_GEN6_DEAD_HINT") ? atoi(getenv("B200BO_GEN6_DEAD_HINT")) : 0;
      sa.kdb = kdb6;
      sa.trace_cta = getenv("B200BO_TRACE_CTA") ? atoi(getenv("B200BO_TRACE_CTA")) : 0;
      sa.smid_out = nullptr;
      if (getenv("B200BO_SMID_DUMP")) {  // developer: where did the CTAs land?
        CU_TRY(h->smid_dbg.reserve(h->num_sms));
        sa.smid_out = h->smid_dbg.p;
      }
#define FK6_LAUNCH(C)                                                                                               \
  do {                                                                                                             \
    auto kern = nprod == 1 ? fk6::predict_fused_shared_kernel<C, 1, false> : fk6::predict_fused_shared_kernel<C, 3, false>; \
    if (C == MATERN52 && nprod == 1 && a.trace) kern = fk6::predict_fused_shared_kernel<MATERN52, 1, true>; /* developer timeline */ \
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, fk6::SMEM6_BYTES));             \
    /* the sides of a group wait for each other: a COOPERATIVE launch, so that the driver refuses a grid it cannot   \
       make co-resident instead of letting it deadlock into the kernel's timeout traps */                            \
    cudaLaunchConfig_t cfg6 = {};                                                                                  \
    cfg6.gridDim = dim3(grid6); cfg6.blockDim = dim3(fk6::NT6); cfg6.dynamicSmemBytes = fk6::SMEM6_BYTES;          \
    cfg6.stream = h->stream;                                                                                       \
    cudaLaunchAttribute at6[1];                                                                                    \
    at6[0].id = cudaLaunchAttributeCooperative;                                                                    \
    at6[0].val.cooperative = h->gen6_cooperative;                                                                  \
    cfg6.attrs = at6; cfg6.numAttrs = 1;                                                                           \
    CU_TRY(cudaLaunchKernelEx(&cfg6, kern, h->replay_maps, a, ra, sa));                                            \
  } while (0)
      switch (h->corr) {
        case RBF: FK6_LAUNCH(RBF); break;
        case MATERN12: FK6_LAUNCH(MATERN12); break;
        case MATERN32: FK6_LAUNCH(MATERN32); break;
        default: FK6_LAUNCH(MATERN52); break;
      }
#undef FK6_LAUNCH
      if (sa.smid_out) {
        std::vector<uint32_t> sm(grid6);
        CU_TRY(cudaMemcpyAsync(sm.data(), sa.smid_out, grid6 * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
        CU_TRY(cudaStreamSynchronize(h->stream));
        if (FILE* f = fopen(getenv("B200BO_SMID_DUMP"), "w")) {
          for (int i = 0; i < grid6; ++i) fprintf(f, "%d %u\n", i, sm[i]);
          fclose(f);
        }
      }
    } else if (h->use_pair && h->use_replay && h->use_decoupled) {
      h->last_gen = 5;
      fk4::ReplayArgs ra;
      ra.scratch = h->r_scratch.p;
      ra.n_store = h->ld / fk::KC;
      h->last_n_store = ra.n_store;
      switch (h->corr) {
        case RBF: FK5_LAUNCH(RBF); break;
        case MATERN12: FK5_LAUNCH(MATERN12); break;
        case MATERN32: FK5_LAUNCH(MATERN32); break;
        default: FK5_LAUNCH(MATERN52); break;
      }
    } else if (h->use_pair && h->use_replay) {
      h->last_gen = 4;
      fk4::ReplayArgs ra;
      ra.scratch = h->r_scratch.p;
      const size_t blk = (size_t)fk::BM * fk::KC;
      const int planes = nprod == 1 ? 1 : 2;
      ra.n_store = (int)std::min<size_t>((size_t)(h->ld / fk::KC), h->r_scratch.n / ((size_t)h->num_sms * planes * blk)) & ~1;
      if (h->replay_max_chunks >= 0) ra.n_store = std::min(ra.n_store, h->replay_max_chunks & ~1);
      h->last_n_store = ra.n_store;
      switch (h->corr) {
        case RBF: FK4_LAUNCH(RBF); break;
        case MATERN12: FK4_LAUNCH(MATERN12); break;
        case MATERN32: FK4_LAUNCH(MATERN32); break;
        default: FK4_LAUNCH(MATERN52); break;
      }
    } else if (h->use_pair) {
      h->last_gen = 3;
      switch (h->corr) {
        case RBF: FK3_LAUNCH(RBF); break;
        case MATERN12: FK3_LAUNCH(MATERN12); break;
        case MATERN32: FK3_LAUNCH(MATERN32); break;
        default: FK3_LAUNCH(MATERN52); break;
      }
    } else {
      h->last_gen = 2;
      switch (h->corr) {
        case RBF: FK2_LAUNCH(RBF); break;
        case MATERN12: FK2_LAUNCH(MATERN12); break;
        case MATERN32: FK2_LAUNCH(MATERN32); break;
        default: FK2_LAUNCH(MATERN52); break;
      }
    }
#undef FK3_LAUNCH
#undef FK4_LAUNCH
#undef FK5_LAUNCH
#undef FK2_LAUNCH
    CU_TRY(cudaGetLastError());
    if (a.trace) {  // developer timeline: dump the stamps of CTA 0 relative to its first one
      std::vector<long long> t((size_t)fk5::TRACE5_CHUNKS * 8);
      CU_TRY(cudaMemcpyAsync(t.data(), a.trace, t.size() * 8, cudaMemcpyDeviceToHost, h->stream));
      CU_TRY(cudaStreamSynchronize(h->stream));
      if (FILE* f = fopen(getenv("B200BO_TRACE"), "w")) {
        long long t0 = 0;
        for (long long v : t) if (v && (!t0 || v < t0)) t0 = v;
        fprintf(f, "# chunk  mma:wait_A  mma:A_ready  mma:issued | prod:start  prod:G_ready  prod:A_free  prod:stored  prod:after_bar   (cycles, nprod=%d)\n", nprod);
        for (int c = 0; c < (h->use_decoupled ? fk5::TRACE5_CHUNKS : fk2::TRACE_CHUNKS); ++c) {
          fprintf(f, "%4d", c);
          for (int k = 0; k < 8; ++k) fprintf(f, " %9lld", t[c * 8 + k] ? t[c * 8 + k] - t0 : -1);
          fprintf(f, "\n");
        }
        fclose(f);
      }
    }
    return 0;
  }
  h->last_gen = 1;
  fk::FusedArgs a;
  a.Xc = xc_dev; a.Xs = h->Xs.p; a.cscale = h->cscale.p;
  a.yhat = h->f_yhat.p + out_off; a.sumsq = h->f_sumsq.p + out_off; a.dotf = h->f_dotf.p + out_off;
  a.dbg_w = h->want_dbg_w ? h->dbg_w.p : nullptr;
  a.err = h->err_flag.p;
  a.M = m; a.N = h->N; a.D = h->D; a.ld = h->ld; a.corr = h->corr; a.beta = h->beta;
  a.out_scale = (float)ldexp(1.0, -(fk::A_SCALE_LOG2 + h->b_scale_log2));
  const long long tiles = (m + fk::BM - 1) / fk::BM;
  const int grid = (int)std::min<long long>(h->num_sms, tiles);
  const int smem = fk::smem_total(h->DP);
  const bool abs_ = h->corr == ABSEXP;
#define FK_LAUNCH(DPV, ABSV)                                                                                     \
  do {                                                                                                           \
    auto kern = fk::predict_fused_tc_kernel<DPV, ABSV>;                                                          \
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, fk::smem_total(DPV)));        \
    kern<<<grid, fk::NT, smem, h->stream>>>(h->map_hi, h->map_lo, a);                                            \
  } while (0)
  switch (h->DP) {
    case 8: if (abs_) FK_LAUNCH(8, true); else FK_LAUNCH(8, false); break;
    case 16: if (abs_) FK_LAUNCH(16, true); else FK_LAUNCH(16, false); break;
    case 32: if (abs_) FK_LAUNCH(32, true); else FK_LAUNCH(32, false); break;
    default: if (abs_) FK_LAUNCH(64, true); else FK_LAUNCH(64, false); break;
  }
#undef FK_LAUNCH
  CU_TRY(cudaGetLastError());
  return 0;
}

static int check_fast_err(b200bo_handle h) {
  int flag = 0;
  CU_TRY(cudaMemcpyAsync(&flag, h->err_flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  if (flag) return set_err(B200BO_E_CUDA, "tensor-core pipeline wait timed out (code " + std::to_string(flag) + ")");
  return 0;
}

static const int FAST_CHUNK_TILES = 8;        // candidate tiles per SM per fused launch when streaming from the host (at N = 4096)
// launches are sized in time, not tiles: (4096 / ld)^2 more tiles per launch below N = 4096
static inline int64_t chunk_scale(int ld) {
  if (ld >= 4096) return 1;
  const int64_t r = (4096 + ld - 1) / ld;
  return r * r;
}
static const int LIST0_CAP = 1 << 16;         // scan survivors (~ subsample stride x q when the criterion is not flat)
static const int THR_STRIDE = 32;
static const double MODEL_LAMBDA = 8.0;       // the a-priori half-widths sit MODEL_LAMBDA modelled standard deviations out

// ---- a-priori rounding model of the tensor-core pass ------------------------------------------
// The fp16 roundings of the operands are modelled as independent, mean-zero, uniform within their unit round-off (the
// probabilistic rounding model of Higham & Mary 2019); the error of sum rt^2 is then a sum of ~N^2 such terms:
//   sd(d ss)  <=  2 u sqrt(ss) sqrt(a_max^2 + ||L^-1||_2^2) r_max        (L^-1 part: row norms; r part: through R^-1 r)
//   E(d ss)   <=  2 u^2 ||L^-1||_F^2 r_max^2                              (second-order term: E sum (d rt)^2)
// with u^2 = 2 (2^-11)^2 / 3 for one fp16 product per MAC and u = 2^-21 for the split-fp16 three-product form, plus
// the fp32 accumulation in TMEM (N products per output against partial sums bounded by the row 1-norms).  yhat and
// Ft^T rt are dot products of the fp32 r: their half-widths come from the fp32 distance / MUFU error per element.
// The deterministic worst-case bound (every rounding at its maximum, all aligned) is kept for the record only: it is
// 2-3 orders of magnitude above any error the pass can produce and would put every candidate in the band.
static void build_fast_model(b200bo_ctx* h, const std::vector<double>& rowsq, const std::vector<double>& rowl1,
                             double s2_sq, double gamma_l2, double f_l2, double b_max) {
  FastModel& m = h->model;
  const int N = h->N;
  double amax2 = 0, fro2 = 0, l1max = 0, l1sq = 0;
  for (int i = 0; i < N; ++i) {
    amax2 = std::max(amax2, rowsq[i]);
    fro2 += rowsq[i];
    l1max = std::max(l1max, rowl1[i]);
    l1sq += rowl1[i] * rowl1[i];
  }
  // ||L^-1||_2^2 lies between the largest squared row norm and the squared Frobenius norm; the power iteration
  // approaches it from below: 25 % on top, clamped into the rigorous interval
  s2_sq = std::min(std::max(1.25 * s2_sq, amax2), fro2);
  m.a_max = sqrt(amax2); m.s2 = sqrt(s2_sq); m.fro = sqrt(fro2); m.l1_max = l1max;
  m.gamma_l2 = gamma_l2; m.f_l2 = f_l2; m.b_max = b_max;
  // per-element error of the fp32 cross-correlation: slope of the kernel in its argument x error of the fp32 squared
  // distance a_m + b_j - 2 <x, X_j> (three split-fp16 Gram products ~2^-22 each, fp32 adds: one fp32 unit round-off,
  // 2^-24, of a_m + b_j as the standard deviation; the product |phi'(acc)| (a_m + b_j) is bounded over ALL candidates
  // by K_phi max(5 b_max, 8), because far candidates have r ~ 0).  Measured on 25 000 float64-checked candidates per
  // BASELINE workload the resulting dy sits 17-60 x above the largest yhat error (profiles/r02/band_probe.json).
  const double e_acc = ldexp(1.0, -24) * std::max(5.0 * b_max, 8.0);
  double kphi = 0.6931471805599453;                       // RBF / abs-exp: d 2^-acc / d acc
  if (h->corr == MATERN32) kphi = 0.5;
  if (h->corr == MATERN52) kphi = 1.0 / 6.0;
  double sd_r = kphi * e_acc + ldexp(1.0, -22);            // + ex2.approx / sqrt.approx / fp32 polynomial
  if (h->corr == MATERN12) sd_r = sqrt(e_acc) + ldexp(1.0, -22);  // exp(-sqrt(acc)): the cusp at acc = 0
  m.sd_r = sd_r;
  const double acc32 = ldexp(1.0, -24);
  m.dy = MODEL_LAMBDA * (sd_r + 4.0 * acc32) * gamma_l2 + 1e-13;
  m.du = h->estimate_trend ? MODEL_LAMBDA * (sd_r + 4.0 * acc32) * f_l2 / fabs(h->G) : 0.0;
  for (int k = 0; k < 2; ++k) {
    const int nprod = k == 0 ? 1 : 3;
    const double u = k == 0 ? ldexp(1.0, -11) * sqrt(2.0 / 3.0) : ldexp(1.0, -21);
    const double sd_op = 2.0 * u * sqrt(amax2 + s2_sq);
    const double sd_acc = 2.0 * sqrt((double)nprod * N) * acc32 * l1max;
    m.ds_rel[k] = MODEL_LAMBDA * sqrt(sd_op * sd_op + sd_acc * sd_acc) * h->sigma2;
    m.ds_abs[k] = (2.0 * u * u * fro2 + ldexp(1.0, -21)) * h->sigma2 + 1e-13 * h->sigma2;
  }
  m.det_ds = 2.0 * ldexp(1.0, -10) * sqrt(2.0) * sqrt(l1sq) * h->sigma2;
}

// fast vs fp64 moments on n candidates whose exact moments are in (h->yhat, h->sumsq, h->dotf): calibration only
static int fast_errors(b200bo_handle h, const long long* list_dev, long long idx_base, int n, double* ey, double* es) {
  fk::band_err_kernel<<<1, 256, 0, h->stream>>>(h->f_yhat.p, h->f_sumsq.p, h->f_dotf.p, h->yhat.p, h->sumsq.p,
                                                 h->dotf.p, list_dev, idx_base, n, h->estimate_trend, h->G,
                                                 h->sigma2, h->errout.p);
  CU_TRY(cudaGetLastError());
  double e[2];
  CU_TRY(cudaMemcpyAsync(e, h->errout.p, 16, cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  *ey = e[0];
  *es = e[1];
  return 0;
}

// phase I: the fused tensor-core pass over all candidates.  Host candidates are copied chunk by chunk into ONE device
// buffer (copies on their own stream, overlapping the fused launches), so that the band stage gathers its rows on the
// device.  *xdev = device address of the whole candidate set.
static int fast_pass(b200bo_handle h, const double* Xc, int64_t M, bool dev, int nprod, const double** xdev, int* launches) {
  cudaStream_t st = h->stream;
  const int D = h->D;
  int rc;
  *launches = 0;
  if (dev) {
    *xdev = Xc;
    // tiles per SM and launch: the knob is quoted at N = 4096 (8 tiles ~ 2.2 ms per launch) and scaled so that a launch
    // stays that long at smaller N -- a fixed 8 tiles cost C2 (N = 1024) a factor two in seven 0.5 ms launches
    const int64_t chunk = h->dev_chunk_tiles > 0 ? (int64_t)h->num_sms * fk::BM * h->dev_chunk_tiles * chunk_scale(h->ld) : M;
    for (int64_t a = 0; a < M; a += chunk) {
      if ((rc = launch_fused(h, Xc + (size_t)a * D, std::min<int64_t>(chunk, M - a), (size_t)a, nprod))) return rc;
      ++*launches;
    }
    return 0;
  }
  CU_TRY(h->Xall.reserve((size_t)M * D));
  *xdev = h->Xall.p;
  const int64_t FMc = (int64_t)h->num_sms * fk::BM * FAST_CHUNK_TILES * chunk_scale(h->ld);
  // the buffer may still be read by earlier work of the compute stream
  CU_TRY(cudaEventRecord(h->ev_used[0], st));
  CU_TRY(cudaStreamWaitEvent(h->copy_stream, h->ev_used[0], 0));
  int64_t i = 0;
  for (int64_t a = 0; a < M; a += FMc, ++i) {
    const int64_t m = std::min<int64_t>(FMc, M - a);
    const int b = (int)(i & 1);
    CU_TRY(cudaMemcpyAsync(h->Xall.p + (size_t)a * D, Xc + (size_t)a * D, (size_t)m * D * 8, cudaMemcpyHostToDevice, h->copy_stream));
    CU_TRY(cudaEventRecord(h->ev_copied[b], h->copy_stream));
    CU_TRY(cudaStreamWaitEvent(st, h->ev_copied[b], 0));
    if ((rc = launch_fused(h, h->Xall.p + (size_t)a * D, m, (size_t)a, nprod))) return rc;
    ++*launches;
  }
  return 0;
}

// once per factorisation and product count: fast vs fp64 moments on a strided sample of the candidate set
static int calibrate_fast(b200bo_handle h, const double* xdev, int64_t M, int ci, LaunchCount* lc) {
  cudaStream_t st = h->stream;
  const int n = (int)std::min<int64_t>(M, std::min(h->Mc, 2048));
  const int64_t stride = std::max<int64_t>(1, M / n);
  CU_TRY(h->band_list.reserve(LIST0_CAP));
  CU_TRY(h->Xband.reserve((size_t)h->Mc * h->D));
  fk::iota_stride_kernel<<<(n + 255) / 256, 256, 0, st>>>(h->band_list.p, n, stride);
  CU_TRY(cudaGetLastError());
  fk::band_gather_kernel<<<(n * h->D + 255) / 256, 256, 0, st>>>(xdev, h->band_list.p, n, 0, h->D, h->Xband.p);
  CU_TRY(cudaGetLastError());
  int rc;
  if ((rc = fp64_moments(h, h->Xband.p, n, h->yhat.p, 1, nullptr, lc))) return rc;
  double ey, es;
  if ((rc = fast_errors(h, h->band_list.p, 0, n, &ey, &es))) return rc;
  h->cal_err_y[ci] = ey;
  h->cal_err_s[ci] = es;
  h->dy_cal[ci] = 8.0 * ey + 1e-13;
  h->ds_cal[ci] = 8.0 * es + 1e-13 * h->sigma2;
  h->calibrated[ci] = true;
  return 0;
}

static int run_candidates_fast(b200bo_handle h, const double* Xc, int64_t M, int loc, int eval_mse, double* yhat_out,
                               double* mse_out, int acq_id, int minimize, double plugin, const double* params, int q,
                               double* best_val, int64_t* best_idx, bool* fell_back) {
  *fell_back = false;
  const bool do_acq = acq_id >= 0;
  const bool dev = loc == B200BO_DEVICE;
  int rc;
  if ((rc = ensure_fast_state(h))) return rc;
  // first pass: one fp16 product when only the arg-max is wanted (its band is re-scored exactly anyway); three
  // products for predict() (documented ~1e-6 moments) or once a fit's one-product band proved too wide
  const int nprod = (do_acq && h->use_v2 && h->fast_products == 1 && !h->escalate) ? 1 : 3;
  const int ci = nprod == 1 ? 0 : 1;
  if ((rc = ensure_predict_ws(h, q, false))) return rc;
  cudaStream_t st = h->stream;
  const int D = h->D, Mc = h->Mc, ld = h->ld;
  CHECK_ARG(M < INT32_MAX - 256, "M too large for one call");
  const size_t Mpad = (size_t)round_up((int)M, fk::BM);
  CU_TRY(h->f_yhat.reserve(Mpad + fk::BM));
  CU_TRY(h->f_sumsq.reserve(Mpad + fk::BM));
  CU_TRY(h->f_dotf.reserve(Mpad + fk::BM));
  if (h->want_dbg_w) CU_TRY(h->dbg_w.reserve(Mpad * (size_t)ld));
  h->evs.reset();
  PhaseTimer pt{h};
  LaunchCount lc;
  int fused_launches = 0;
  cudaEvent_t e0 = h->evs.get(), e1 = h->evs.get();
  CU_TRY(cudaEventRecord(e0, st));

  // ---- phase I -----------------------------------------------------------------------------------------
  pt.begin(1);
  const double* xdev = nullptr;
  if ((rc = fast_pass(h, Xc, M, dev, nprod, &xdev, &fused_launches))) return rc;
  pt.end(1);
  lc.all += fused_launches;
  h->last_fast = {xdev, M, nprod};

  if (!do_acq) {
    // predict(): moments of the fast pass, MSE elementwise (approximate; see include/b200bo.h)
    AcqArgs g{};
    g.yhat = h->f_yhat.p; g.sumsq = h->f_sumsq.p; g.dotf = h->f_dotf.p;
    g.M = (int)M; g.estimate_trend = h->estimate_trend; g.sigma2 = h->sigma2; g.G = h->G;
    double* mo = nullptr;
    if (eval_mse) {
      // host output: the MSE goes to its own device buffer (the fast sums stay intact for b200bo_debug_fast_check)
      if (dev) mo = mse_out;
      else { CU_TRY(h->f_mse.reserve(Mpad + fk::BM)); mo = h->f_mse.p; }
      g.mse_out = mo;
      mse_kernel<<<std::min(h->num_sms * 4, (int)((M + 255) / 256)), 256, 0, st>>>(g);
      CU_TRY(cudaGetLastError());
      ++lc.all;
    }
    const cudaMemcpyKind kind = dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (yhat_out) CU_TRY(cudaMemcpyAsync(yhat_out, h->f_yhat.p, (size_t)M * 8, kind, st));
    if (eval_mse && !dev) CU_TRY(cudaMemcpyAsync(mse_out, mo, (size_t)M * 8, kind, st));
    CU_TRY(cudaEventRecord(e1, st));
    if ((rc = check_fast_err(h))) return rc;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    h->timings[0] = ms; h->timings[1] = 0; h->timings[2] = pt.total(1); h->timings[3] = 0;
    h->timings[4] = fused_launches; h->timings[5] = lc.all; h->timings[6] = 0; h->timings[7] = 0;
    h->timings[8] = 3;
    h->timings[10] = h->last_gen;
    return 0;
  }

  // ---- calibration (once per factor()): observed errors on a strided sample of this candidate set ------------
  pt.begin(2);
  if (!h->calibrated[ci]) {
    if ((rc = check_fast_err(h))) return rc;
    if ((rc = calibrate_fast(h, xdev, M, ci, &lc))) return rc;
  }

  // ---- phase II / III: band selection + exact re-score, one stream-ordered pipeline per pass --------------------
  // on-device re-score up to BD band members: the row-sweep kernel costs ~ count x ld^2 / 2 fp64 MACs out of L2 / shared
  // memory (fine for tens of candidates at N = 8192, for a thousand at N = 1024); wider bands take the tile kernels
  const int BD = (int)std::max<long long>(32, std::min<long long>(bd::BAND_DEV_MAX, (5LL << 28) / ((long long)ld * ld) / 8 * 8));
  const int nslices = ld / bd::KSS_COLS + (ld % bd::KSS_COLS ? 1 : 0);
  const int rd_blocks = ld / bd::RD_WARPS, rd_grid = std::min(rd_blocks, h->num_sms * 2);
  CU_TRY(h->thr_key.reserve(2 * (size_t)q));
  // capacity of the scan's survivor list: the 1/32-subsample threshold lets through ~ stride x q candidates when the
  // criterion is peaked, but several per cent of M when it is not (C2, EI: 78 672 of 1e6, of which 244 form the band);
  // 2^18 entries where q is small (the exact-bound table is cap x q doubles), 2^16 at q = 32
  const int cap0 = (int)std::max<long long>(LIST0_CAP, std::min<long long>(1 << 18, (1LL << 21) / std::max(q, 1)));
  CU_TRY(h->band_list0.reserve(cap0));
  CU_TRY(h->band_list.reserve(cap0));
  CU_TRY(h->band_count.reserve(2));
  CU_TRY(h->band_hiB.reserve((size_t)cap0 * q));
  CU_TRY(h->Xband.reserve((size_t)Mc * D));
  CU_TRY(h->bd_kst.reserve((size_t)BD * ld));
  CU_TRY(h->bd_ypart.reserve((size_t)nslices * BD));
  CU_TRY(h->bd_part.reserve((size_t)rd_blocks * BD * 2));
  CU_TRY(h->bd_ctl.reserve(1));
  if (!h->pin) CU_TRY(cudaHostAlloc(&h->pin, PIN_BYTES, cudaHostAllocDefault));
  // pinned staging: [params q][BandCtl][best_val q][best_idx q]
  double* pin_params = (double*)h->pin;
  bd::BandCtl* pin_ctl = (bd::BandCtl*)(h->pin + (size_t)fk::BAND_MAX_Q * 8);
  double* pin_bv = (double*)(pin_ctl + 1);
  long long* pin_bi = (long long*)(pin_bv + fk::BAND_MAX_Q);
  for (int c = 0; c < q; ++c) pin_params[c] = params ? params[c] : 0.0;
  CU_TRY(cudaMemcpyAsync(h->params.p, pin_params, (size_t)q * 8, cudaMemcpyHostToDevice, st));
  const size_t kss_smem = ((size_t)bd::KSS_ROWS * D + D + 1) * sizeof(double);
  if (kss_smem > 48 * 1024)
    CU_TRY(cudaFuncSetAttribute(bd::kstar_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kss_smem));

  int rescored = 0, passes = 0;
  double widen = h->widen[ci];
  for (;;) {
    ++passes;
    fk::BandArgs b;
    b.yhat = h->f_yhat.p; b.sumsq = h->f_sumsq.p; b.dotf = h->f_dotf.p; b.params = h->params.p;
    b.M = M; b.acq = acq_id; b.minimize = minimize; b.estimate_trend = h->estimate_trend; b.q = q;
    b.sigma2 = h->sigma2; b.plugin = plugin; b.G = h->G;
    // half-widths: the larger of the a-priori model and the observed (calibrated) errors, times the widening factor
    b.dy = widen * std::max(h->dy_cal[ci], h->use_model ? h->model.dy : 0.0);
    b.ds = widen * h->ds_cal[ci];
    b.ds_abs = h->use_model ? widen * h->model.ds_abs[ci] : 0.0;
    b.ds_rel = h->use_model ? widen * h->model.ds_rel[ci] : 0.0;
    b.du = h->use_model ? widen * h->model.du : 0.0;
    h->last_band = b;
    bd::band_reset_kernel<<<(std::max(2 * q, 2) + 255) / 256, 256, 0, st>>>(h->thr_key.p, 2 * q, fk::ord_key(-INFINITY), h->band_count.p,
                                                                          h->best_val.p, h->best_idx.p, q, h->bd_ctl.p);
    CU_TRY(cudaGetLastError());
    const long long ns = (M + THR_STRIDE - 1) / THR_STRIDE;
    fk::band_thr0_kernel<<<(int)std::min<long long>(h->num_sms * 4, (ns + 255) / 256), 256, 0, st>>>(b, THR_STRIDE, h->thr_key.p);
    CU_TRY(cudaGetLastError());
    fk::band_scan_kernel<<<(int)std::min<long long>(h->num_sms * 8, (M + 255) / 256), 256, 0, st>>>(
        b, h->thr_key.p, h->band_list0.p, cap0, h->band_count.p);
    CU_TRY(cudaGetLastError());
    fk::band_refine_kernel<<<h->num_sms, 256, 0, st>>>(b, h->band_list0.p, h->band_count.p, cap0, h->band_hiB.p,
                                                       h->thr_key.p + q);
    CU_TRY(cudaGetLastError());
    fk::band_filter_kernel<<<h->num_sms, 256, 0, st>>>(h->band_list0.p, h->band_count.p, cap0, h->band_hiB.p,
                                                       h->thr_key.p + q, q, h->band_list.p, h->band_count.p + 1);
    CU_TRY(cudaGetLastError());
    // exact float64 re-score of up to BAND_DEV_MAX band members, sizes read on the device
    bd::band_gather_dev_kernel<<<(BD * D + 255) / 256, 256, 0, st>>>(xdev, h->band_list.p, h->band_count.p + 1, BD, D, h->Xband.p);
    CU_TRY(cudaGetLastError());
    bd::KstarSmallArgs ks;
    ks.Xb = h->Xband.p; ks.Xt = h->Xt.p; ks.theta = h->theta.p; ks.gamma = h->gamma.p; ks.Kst = h->bd_kst.p;
    ks.ypart = h->bd_ypart.p; ks.count = h->band_count.p + 1; ks.cap = BD; ks.N = h->N; ks.D = D; ks.ld = ld; ks.corr = h->corr;
    bd::kstar_small_kernel<<<dim3(nslices, BD / bd::KSS_ROWS), bd::KSS_COLS, kss_smem, st>>>(ks);
    CU_TRY(cudaGetLastError());
    bd::RowdotArgs rd;
    rd.Kst = h->bd_kst.p; rd.Linv = h->W.p; rd.Ft = h->Ft.p; rd.part = h->bd_part.p; rd.count = h->band_count.p + 1;
    rd.cap = BD; rd.ld = ld;
    bd::rowdot_kernel<<<rd_grid, 32 * bd::RD_WARPS, 0, st>>>(rd);
    CU_TRY(cudaGetLastError());
    bd::band_moments_kernel<<<BD, 128, 0, st>>>(h->bd_ypart.p, nslices, h->bd_part.p, rd_blocks, h->band_count.p + 1, BD, h->beta,
                                                h->yhat.p, h->sumsq.p, h->dotf.p);
    CU_TRY(cudaGetLastError());
    {
      AcqArgs g{};
      g.yhat = h->yhat.p; g.sumsq = h->sumsq.p; g.dotf = h->dotf.p;
      g.M = BD; g.M_dev = h->band_count.p + 1; g.estimate_trend = h->estimate_trend; g.sigma2 = h->sigma2; g.G = h->G;
      g.idx_map = h->band_list.p;
      g.params = h->params.p; g.part_val = h->part_val.p; g.part_idx = h->part_idx.p;
      g.acq = acq_id; g.minimize = minimize; g.q = q; g.plugin = plugin;
      acq_kernel<<<dim3(1, q), 256, 0, st>>>(g);
      CU_TRY(cudaGetLastError());
      argmax_merge_kernel<<<(q + 63) / 64, 64, 0, st>>>(h->part_val.p, h->part_idx.p, 1, q, h->best_val.p, h->best_idx.p);
      CU_TRY(cudaGetLastError());
    }
    bd::BandCheckArgs ck;
    ck.y_fast = h->f_yhat.p; ck.ss_fast = h->f_sumsq.p; ck.df_fast = h->f_dotf.p;
    ck.y_ex = h->yhat.p; ck.ss_ex = h->sumsq.p; ck.df_ex = h->dotf.p; ck.list = h->band_list.p;
    ck.counts = h->band_count.p; ck.fused_err = h->err_flag.p; ck.cap = BD; ck.estimate_trend = h->estimate_trend;
    ck.G = h->G; ck.sigma2 = h->sigma2; ck.dy = b.dy; ck.ds = b.ds; ck.ds_abs = b.ds_abs; ck.ds_rel = b.ds_rel; ck.du = b.du;
    ck.ctl = h->bd_ctl.p;
    bd::band_check_kernel<<<1, 256, 0, st>>>(ck);
    CU_TRY(cudaGetLastError());
    lc.all += 12;
    CU_TRY(cudaMemcpyAsync(pin_ctl, h->bd_ctl.p, sizeof(bd::BandCtl), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(pin_bv, h->best_val.p, (size_t)q * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(pin_bi, h->best_idx.p, (size_t)q * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaEventRecord(e1, st));
    CU_TRY(cudaStreamSynchronize(st));   // the one host round trip of the pass
    if (pin_ctl->fused_err)
      return set_err(B200BO_E_CUDA, "tensor-core pipeline wait timed out (code " + std::to_string(pin_ctl->fused_err) + ")");
    const int count0 = pin_ctl->count0, count = pin_ctl->count;
    // Escalate from one product to three only where that is the cheaper way to a narrow band: re-scoring `count`
    // candidates in float64 costs ~count x N^2 flops on the fp64 pipe (~20 TFLOP/s), the two extra products ~2 M N^2 on
    // the tensor pipe (~1000 TFLOP/s) -- break-even at count ~ M / 25.  (A fixed 2048 sent C2, whose a-priori half-widths
    // admit a few thousand of its 1e6 candidates, to the three-product pass: 3.4 ms per step instead of 1.5.)
    const long long rescore_max = std::max<long long>(h->rescore_max, M / 50);
    if (getenv("B200BO_BAND_DEBUG")) fprintf(stderr, "band: nprod %d pass %d count0 %d count %d (direct up to %lld)\n", nprod, passes, count0, count, rescore_max);
    if (nprod == 1 && (count0 > cap0 || count > rescore_max || passes > 2)) {
      // the one-product pass cannot separate the top of this criterion: three products for the rest of this fit
      h->escalate = true;
      return run_candidates_fast(h, Xc, M, loc, eval_mse, yhat_out, mse_out, acq_id, minimize, plugin, params, q,
                                 best_val, best_idx, fell_back);
    }
    if (count0 > cap0 || passes > 4) {  // degenerate (flat criterion) or unstable error estimate
      *fell_back = true;
      return 0;
    }
    double ratio = pin_ctl->ratio;
    h->last_err_y = pin_ctl->err_y; h->last_err_s = pin_ctl->err_s;
    if (count > BD) {
      // wide band: re-score it Mc rows at a time with the tile kernels of the float64 path (host-driven)
      if ((rc = upload_params_reset_best(h, params, q))) return rc;
      ratio = 0.0;
      for (int o = 0; o < count; o += Mc) {
        const int m = std::min(Mc, count - o);
        fk::band_gather_kernel<<<(m * D + 255) / 256, 256, 0, st>>>(xdev, h->band_list.p + o, m, 0, D, h->Xband.p);
        CU_TRY(cudaGetLastError());
        ++lc.all;
        if ((rc = fp64_moments(h, h->Xband.p, m, h->yhat.p, 1, nullptr, &lc))) return rc;
        if ((rc = acq_stage(h, h->yhat.p, m, 0, h->band_list.p + o, acq_id, minimize, plugin, q, nullptr, 0, 0, nullptr, &lc)))
          return rc;
        bd::BandCheckArgs c2 = ck;
        c2.list = h->band_list.p + o; c2.cap = m; c2.counts = h->band_count.p;  // counts[1] >= m here
        CU_TRY(cudaMemsetAsync(h->bd_ctl.p, 0, sizeof(bd::BandCtl), st));
        bd::band_check_kernel<<<1, 256, 0, st>>>(c2);
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaMemcpyAsync(pin_ctl, h->bd_ctl.p, sizeof(bd::BandCtl), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        ratio = std::max(ratio, pin_ctl->ratio);
        h->last_err_y = std::max(h->last_err_y, pin_ctl->err_y);
        h->last_err_s = std::max(h->last_err_s, pin_ctl->err_s);
      }
      CU_TRY(cudaEventRecord(e1, st));
      if ((rc = download_best(h, q, pin_bv, (int64_t*)pin_bi))) return rc;
    }
    rescored += count;
    h->last_ratio = ratio;
    if (ratio <= 0.5) break;  // every error seen inside the band is within half of what the band allowed for
    widen *= 4.0;             // larger errors than allowed for: widen and select again (no tensor-core work is redone)
    h->widen[ci] = widen;
  }
  pt.end(2);
  for (int c = 0; c < q; ++c) {
    best_val[c] = pin_bv[c];
    best_idx[c] = pin_bi[c];
  }
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  h->timings[0] = ms; h->timings[1] = 0; h->timings[2] = pt.total(1); h->timings[3] = ms - pt.total(1);
  h->timings[4] = fused_launches; h->timings[5] = lc.all; h->timings[6] = rescored; h->timings[7] = passes;
  h->timings[8] = nprod; h->timings[9] = (nprod == 3 && h->escalate) ? 1 : 0;
  h->timings[10] = h->last_gen;
  return 0;
}

static int run_candidates(b200bo_handle h, const double* Xc, int64_t M, int loc, int eval_mse, double* yhat_out,
                          double* mse_out, int acq_id, int minimize, double plugin, const double* params, int q,
                          double* vals, double* best_val, int64_t* best_idx) {
  CHECK_ARG(h, "handle is NULL");
  if (!h->factored) return set_err(B200BO_E_STATE, "predict before a successful factor()");
  CHECK_ARG(M >= 0, "M < 0");
  CHECK_ARG(loc == B200BO_HOST || loc == B200BO_DEVICE, "bad loc");
  CU_TRY(cudaSetDevice(h->device));
  // the tensor-core pass needs the variance (its product is rt) and cannot return all q x M values exactly
  if (h->prec == B200BO_PREC_FAST && fast_supported(h) && h->trend == B200BO_TREND_CONSTANT && eval_mse && !vals && M > 0 &&
      q <= fk::BAND_MAX_Q) {
    bool fell_back = false;
    int rc = run_candidates_fast(h, Xc, M, loc, eval_mse, yhat_out, mse_out, acq_id, minimize, plugin, params, q,
                                 best_val, best_idx, &fell_back);
    if (rc || !fell_back) return rc;
  }
  return run_candidates_fp64(h, Xc, M, loc, eval_mse, yhat_out, mse_out, acq_id, minimize, plugin, params, q, vals,
                             best_val, best_idx);
}

int b200bo_predict(b200bo_handle h, const double* Xc, int64_t M, int loc, int eval_mse, double* yhat, double* mse) {
  CHECK_ARG(Xc || M == 0, "Xc is NULL");
  CHECK_ARG(yhat, "yhat is NULL");
  CHECK_ARG(!eval_mse || mse, "mse is NULL with eval_mse");
  return run_candidates(h, Xc, M, loc, eval_mse, yhat, eval_mse ? mse : nullptr, -1, 1, 0.0, nullptr, 0, nullptr,
                        nullptr, nullptr);
}

int b200bo_acq(b200bo_handle h, const double* Xc, int64_t M, int loc, int acq_id, int minimize, double plugin,
               const double* params, int q, double* vals, double* best_val, int64_t* best_idx) {
  CHECK_ARG(Xc || M == 0, "Xc is NULL");
  CHECK_ARG(acq_id >= 0 && acq_id <= 3, "unknown acquisition id");
  CHECK_ARG(q >= 1 && q <= 4096, "q out of range");
  CHECK_ARG(best_val && best_idx, "best_val / best_idx are NULL");
  CHECK_ARG(params || acq_id == B200BO_ACQ_EI, "params is NULL");
  int rc = run_candidates(h, Xc, M, loc, 1, nullptr, nullptr, acq_id, minimize, plugin, params, q, vals, best_val, best_idx);
  if (!rc) h->last_q = q;
  return rc;
}

int b200bo_best_pairs_device(b200bo_handle h, int64_t index_offset, int rank, int world, void* out_dev) {
  CHECK_ARG(h && out_dev, "NULL argument");
  CHECK_ARG(world >= 1 && rank >= 0 && rank < world, "bad rank / world");
  if (h->last_q <= 0) return set_err(B200BO_E_STATE, "no acquisition call yet");
  CU_TRY(cudaSetDevice(h->device));
  const int n = world * 2 * h->last_q;
  best_pairs_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(h->best_val.p, h->best_idx.p, h->last_q, index_offset, rank, world,
                                                           (long long*)out_dev);
  CU_TRY(cudaGetLastError());
  return 0;
}

int b200bo_acq_from_moments(b200bo_handle h, const double* yhat, const double* mse, int64_t M, int loc, int acq_id,
                            int minimize, double plugin, const double* params, int q, double* vals,
                            double* best_val, int64_t* best_idx) {
  CHECK_ARG(h && yhat && mse, "NULL argument");
  if (!h->factored) return set_err(B200BO_E_STATE, "acquisition before a successful factor()");
  CHECK_ARG(acq_id >= 0 && acq_id <= 3, "unknown acquisition id");
  CHECK_ARG(q >= 1 && q <= 4096, "q out of range");
  CHECK_ARG(best_val && best_idx, "best_val / best_idx are NULL");
  CHECK_ARG(params || acq_id == B200BO_ACQ_EI, "params is NULL");
  CU_TRY(cudaSetDevice(h->device));
  const bool dev = loc == B200BO_DEVICE;
  int rc = ensure_predict_ws(h, q, vals && !dev);
  if (rc) return rc;
  cudaStream_t st = h->stream;
  const int Mc = h->Mc;
  std::vector<double> pr(q);
  for (int c = 0; c < q; ++c) pr[c] = params ? params[c] : 0.0;
  std::vector<long long> bi(q, -1);
  CU_TRY(cudaMemcpyAsync(h->params.p, pr.data(), q * 8, cudaMemcpyHostToDevice, st));
  CU_TRY(cudaMemsetAsync(h->best_val.p, 0, q * 8, st));
  CU_TRY(cudaMemcpyAsync(h->best_idx.p, bi.data(), q * 8, cudaMemcpyHostToDevice, st));
  for (int64_t a = 0; a < M; a += Mc) {
    const int m = (int)std::min<int64_t>(Mc, M - a);
    const double *yh = yhat + a, *ms = mse + a;
    if (!dev) {
      CU_TRY(cudaMemcpyAsync(h->yhat.p, yh, (size_t)m * 8, cudaMemcpyHostToDevice, st));
      CU_TRY(cudaMemcpyAsync(h->mse.p, ms, (size_t)m * 8, cudaMemcpyHostToDevice, st));
      yh = h->yhat.p;
      ms = h->mse.p;
    }
    AcqArgs g{};
    g.yhat = yh; g.mse_in = ms; g.M = m; g.sigma2 = h->sigma2; g.idx_base = a;
    g.vals = vals ? (dev ? vals : h->vals.p) : nullptr;
    g.vals_ld = dev ? M : Mc; g.vals_off = dev ? a : 0;
    g.params = h->params.p; g.part_val = h->part_val.p; g.part_idx = h->part_idx.p;
    g.acq = acq_id; g.minimize = minimize; g.q = q; g.plugin = plugin;
    int nblk = std::min(h->num_sms, (m + 255) / 256);
    acq_kernel<<<dim3(nblk, q), 256, 0, st>>>(g);
    CU_TRY(cudaGetLastError());
    argmax_merge_kernel<<<(q + 63) / 64, 64, 0, st>>>(h->part_val.p, h->part_idx.p, nblk, q, h->best_val.p, h->best_idx.p);
    CU_TRY(cudaGetLastError());
    if (!dev && vals)
      CU_TRY(cudaMemcpy2DAsync(vals + a, (size_t)M * 8, h->vals.p, (size_t)Mc * 8, (size_t)m * 8, q, cudaMemcpyDeviceToHost, st));
  }
  CU_TRY(cudaMemcpyAsync(best_val, h->best_val.p, q * 8, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaMemcpyAsync(bi.data(), h->best_idx.p, q * 8, cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  for (int c = 0; c < q; ++c) best_idx[c] = bi[c];
  return 0;
}

int b200bo_debug_fast_rt(b200bo_handle h, const double* Xc, int64_t M, float* out_rt, double* yhat, double* sumsq,
                         double* dotf) {
  CHECK_ARG(h && Xc && out_rt, "NULL argument");
  if (!h->factored) return set_err(B200BO_E_STATE, "debug_fast_rt before a successful factor()");
  CHECK_ARG(M >= 1 && M <= 65536, "M out of range for the debug hook");
  CHECK_ARG(fast_supported(h), "the tensor-core path does not cover this kernel / feature count");
  CU_TRY(cudaSetDevice(h->device));
  int rc;
  if ((rc = ensure_fast_state(h))) return rc;
  const size_t Mpad = (size_t)round_up((int)M, fk::BM);
  const int D = h->D, ld = h->ld, N = h->N;
  CU_TRY(h->f_yhat.reserve(Mpad + fk::BM));
  CU_TRY(h->f_sumsq.reserve(Mpad + fk::BM));
  CU_TRY(h->f_dotf.reserve(Mpad + fk::BM));
  CU_TRY(h->dbg_w.reserve(Mpad * (size_t)ld));
  CU_TRY(h->stage[0].reserve((size_t)M * D));
  CU_TRY(cudaMemcpyAsync(h->stage[0].p, Xc, (size_t)M * D * 8, cudaMemcpyHostToDevice, h->stream));
  h->want_dbg_w = true;
  rc = launch_fused(h, h->stage[0].p, M, 0);
  h->want_dbg_w = false;
  if (rc) return rc;
  if ((rc = check_fast_err(h))) return rc;
  CU_TRY(cudaMemcpy2DAsync(out_rt, (size_t)N * 4, h->dbg_w.p, (size_t)ld * 4, (size_t)N * 4, (size_t)M,
                           cudaMemcpyDeviceToHost, h->stream));
  if (yhat) CU_TRY(cudaMemcpyAsync(yhat, h->f_yhat.p, (size_t)M * 8, cudaMemcpyDeviceToHost, h->stream));
  if (sumsq) CU_TRY(cudaMemcpyAsync(sumsq, h->f_sumsq.p, (size_t)M * 8, cudaMemcpyDeviceToHost, h->stream));
  if (dotf) CU_TRY(cudaMemcpyAsync(dotf, h->f_dotf.p, (size_t)M * 8, cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}

// fvec = L^-T Ft, so that Ft^T L^-1 v = fvec . v  (gpr.py:571-572 without the second triangular solve)
static int ensure_fvec(b200bo_handle h) {
  if (h->fvec_ready) return 0;
  const int ld = h->ld;
  cudaStream_t st = h->stream;
  CU_TRY(h->fvec.reserve(ld));
  const int gchunks = (ld + 255) / 256;
  CU_TRY(h->part.reserve((size_t)gchunks * ld));
  tri_gemvT_partial_kernel<<<dim3((ld + 255) / 256, gchunks), 256, 0, st>>>(h->W.p, ld, ld, h->Ft.p, h->part.p, 256);
  CU_TRY(cudaGetLastError());
  colsum_partials_kernel<<<(ld + 255) / 256, 256, 0, st>>>(h->part.p, ld, gchunks, h->fvec.p);
  CU_TRY(cudaGetLastError());
  h->fvec_ready = true;
  return 0;
}

// posterior moments + gradients of up to GRAD_CHUNK device-resident candidates into the g_* buffers
constexpr int GRAD_CHUNK = 1024;
static int grad_chunk(b200bo_handle h, const double* xc_dev, int m) {
  cudaStream_t st = h->stream;
  const int D = h->D, ld = h->ld;
  const int mpad = round_up(m, PC_BM);
  int rc;
  if ((rc = ensure_fvec(h))) return rc;
  const size_t ks_smem = ((size_t)KS_ROWS * D + D) * sizeof(double);
  if (ks_smem > 48 * 1024)
    CU_TRY(cudaFuncSetAttribute(kstar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ks_smem));
  KstarArgs k;
  k.Xc = xc_dev; k.Xt = h->Xt.p; k.theta = h->theta.p; k.gamma = h->gamma.p;
  k.Kst = h->Kst.p; k.yhat = h->yhat.p;
  k.M = m; k.N = h->N; k.D = D; k.ld = ld; k.corr = h->corr; k.beta = h->beta;
  kstar_kernel<<<mpad / KS_ROWS, 256, ks_smem, st>>>(k);
  CU_TRY(cudaGetLastError());
  // rt = r L^-T  (k <= n: L^-1 is lower triangular)                                        gpr.py:564
  GemmArgs g{};
  g.A = h->Kst.p; g.lda = ld; g.B = h->W.p; g.ldb = ld; g.C = h->gRT.p; g.ldc = ld;
  g.sA = g.sB = g.sC = 0; g.K = ld; g.alpha = 1.0; g.beta = 0.0; g.lower_only = 0; g.kb_mode = 0; g.ke_mode = 1;
  CU_TRY((launch_gemm<GemmCore<64, 64, 32, 32, false, false, 3>, false, false>(h, g, mpad, ld, 1)));
  // z = rt L^-1 = R^-1 r  (k >= n)
  g.A = h->gRT.p; g.B = h->W.p; g.C = h->gZ.p; g.kb_mode = 1; g.ke_mode = 0;
  CU_TRY((launch_gemm<GemmCore<64, 64, 32, 32, false, true, 3>, false, true>(h, g, mpad, ld, 1)));
  PostGradArgs a;
  a.Xc = xc_dev; a.Xt = h->Xt.p; a.theta = h->theta.p; a.Kst = h->Kst.p; a.RT = h->gRT.p; a.Z = h->gZ.p;
  a.gamma = h->gamma.p; a.fv = h->fvec.p; a.Ft = h->Ft.p;
  a.y_dx = h->g_ydx.p; a.mse_dx = h->g_mdx.p; a.mse = h->mse.p;
  a.M = m; a.N = h->N; a.D = D; a.ld = ld; a.corr = h->corr; a.estimate_trend = h->estimate_trend;
  a.sigma2 = h->sigma2; a.G = h->G;
  const size_t pg_smem = ((size_t)ld + 2 * D) * sizeof(double);
  if (pg_smem > 48 * 1024)
    CU_TRY(cudaFuncSetAttribute(post_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pg_smem));
  post_grad_kernel<<<m, PG_NT, pg_smem, st>>>(a);
  CU_TRY(cudaGetLastError());
  return 0;
}

static int grad_common(b200bo_handle h, const double* Xc, int64_t M, int acq_id, int minimize, double plugin, double par,
                       double* yhat, double* mse, double* y_dx, double* mse_dx, double* val, double* dx) {
  CHECK_ARG(h && (Xc || M == 0), "NULL argument");
  if (!h->factored) return set_err(B200BO_E_STATE, "gradient before a successful factor()");
  CHECK_ARG(h->corr != CUBIC && !corr_has_extra_param(h->corr), "this kernel has no gradient (corr_dx: `pass`, gpr.py:652-655)");
  CHECK_ARG(h->trend == B200BO_TREND_CONSTANT, "the posterior gradient is implemented for the constant trend");
  CU_TRY(cudaSetDevice(h->device));
  int rc;
  if ((rc = ensure_predict_ws(h, 1, false))) return rc;
  const int D = h->D, ld = h->ld;
  const size_t rows = (size_t)round_up(GRAD_CHUNK, PC_BM);
  CU_TRY(h->gRT.reserve(rows * ld));
  CU_TRY(h->gZ.reserve(rows * ld));
  CU_TRY(h->g_ydx.reserve(rows * D));
  CU_TRY(h->g_mdx.reserve(rows * D));
  CU_TRY(h->g_val.reserve(rows));
  CU_TRY(h->g_dx.reserve(rows * D));
  cudaStream_t st = h->stream;
  for (int64_t a = 0; a < M; a += GRAD_CHUNK) {
    const int m = (int)std::min<int64_t>(GRAD_CHUNK, M - a);
    CU_TRY(cudaMemcpyAsync(h->Xc.p, Xc + (size_t)a * D, (size_t)m * D * 8, cudaMemcpyHostToDevice, st));
    if ((rc = grad_chunk(h, h->Xc.p, m))) return rc;
    if (acq_id >= 0) {
      AcqGradArgs g;
      g.yhat = h->yhat.p; g.mse = h->mse.p; g.y_dx = h->g_ydx.p; g.mse_dx = h->g_mdx.p; g.val = h->g_val.p; g.dx = h->g_dx.p;
      g.M = m; g.D = D; g.acq = acq_id; g.minimize = minimize; g.sigma2 = h->sigma2; g.plugin = plugin; g.par = par;
      acq_grad_kernel<<<(m + 127) / 128, 128, 0, st>>>(g);
      CU_TRY(cudaGetLastError());
      if (val) CU_TRY(cudaMemcpyAsync(val + a, h->g_val.p, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
      if (dx) CU_TRY(cudaMemcpyAsync(dx + (size_t)a * D, h->g_dx.p, (size_t)m * D * 8, cudaMemcpyDeviceToHost, st));
    }
    if (yhat) CU_TRY(cudaMemcpyAsync(yhat + a, h->yhat.p, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
    if (mse) CU_TRY(cudaMemcpyAsync(mse + a, h->mse.p, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
    if (y_dx) CU_TRY(cudaMemcpyAsync(y_dx + (size_t)a * D, h->g_ydx.p, (size_t)m * D * 8, cudaMemcpyDeviceToHost, st));
    if (mse_dx) CU_TRY(cudaMemcpyAsync(mse_dx + (size_t)a * D, h->g_mdx.p, (size_t)m * D * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));  // Xc staging buffer is reused by the next chunk
  }
  return 0;
}

int b200bo_gradient(b200bo_handle h, const double* Xc, int64_t M, double* yhat, double* mse, double* y_dx, double* mse_dx) {
  CHECK_ARG(y_dx && mse_dx, "y_dx / mse_dx are NULL");
  return grad_common(h, Xc, M, -1, 1, 0.0, 0.0, yhat, mse, y_dx, mse_dx, nullptr, nullptr);
}

int b200bo_acq_grad(b200bo_handle h, const double* Xc, int64_t M, int acq_id, int minimize, double plugin, double param,
                    double* val, double* dx) {
  CHECK_ARG(acq_id >= 0 && acq_id <= 3, "unknown acquisition id");
  CHECK_ARG(val && dx, "val / dx are NULL");
  return grad_common(h, Xc, M, acq_id, minimize, plugin, param, nullptr, nullptr, nullptr, nullptr, val, dx);
}

int b200bo_debug_fused_time(b200bo_handle h, const double* Xc_host, int64_t M, int products, int reps, double* out_ms) {
  CHECK_ARG(h && Xc_host && out_ms, "NULL argument");
  if (!h->factored) return set_err(B200BO_E_STATE, "debug_fused_time before a successful factor()");
  CHECK_ARG(M >= 1 && reps >= 1 && (products == 1 || products == 3), "bad argument");
  CHECK_ARG(fast_supported(h), "the tensor-core path does not cover this kernel / feature count");
  CU_TRY(cudaSetDevice(h->device));
  int rc;
  if ((rc = ensure_fast_state(h))) return rc;
  const size_t Mpad = (size_t)round_up((int)M, fk::BM);
  CU_TRY(h->f_yhat.reserve(Mpad + fk::BM));
  CU_TRY(h->f_sumsq.reserve(Mpad + fk::BM));
  CU_TRY(h->f_dotf.reserve(Mpad + fk::BM));
  CU_TRY(h->stage[0].reserve((size_t)M * h->D));
  CU_TRY(cudaMemcpyAsync(h->stage[0].p, Xc_host, (size_t)M * h->D * 8, cudaMemcpyHostToDevice, h->stream));
  cudaEvent_t e0, e1;
  CU_TRY(cudaEventCreate(&e0));
  CU_TRY(cudaEventCreate(&e1));
  if ((rc = launch_fused(h, h->stage[0].p, M, 0, products))) return rc;  // warm-up
  CU_TRY(cudaEventRecord(e0, h->stream));
  for (int i = 0; i < reps; ++i)
    if ((rc = launch_fused(h, h->stage[0].p, M, 0, products))) return rc;
  CU_TRY(cudaEventRecord(e1, h->stream));
  CU_TRY(cudaStreamSynchronize(h->stream));
  if ((rc = check_fast_err(h))) return rc;
  float ms = 0.f;
  CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *out_ms = (double)ms / reps;
  return 0;
}

int b200bo_get_band_info(b200bo_handle h, double* out, int n) {
  CHECK_ARG(h && out && n >= 1, "bad argument");
  if (!h->fast_ready) return set_err(B200BO_E_STATE, "no tensor-core state yet (run a B200BO_PREC_FAST call first)");
  const FastModel& m = h->model;
  const fk::BandArgs& b = h->last_band;
  const double v[B200BO_N_BAND_INFO] = {
      m.dy, m.du, m.ds_abs[0], m.ds_rel[0], m.ds_abs[1], m.ds_rel[1], m.det_ds, m.a_max, m.s2, m.fro, m.l1_max, m.gamma_l2,
      m.f_l2, m.b_max, m.sd_r, h->dy_cal[0], h->ds_cal[0], h->dy_cal[1], h->ds_cal[1], h->cal_err_y[0], h->cal_err_s[0],
      h->cal_err_y[1], h->cal_err_s[1], b.dy, b.ds, b.ds_abs, b.ds_rel, b.du, h->last_err_y, h->last_err_s, h->last_ratio,
      std::max(h->widen[0], h->widen[1])};
  for (int i = 0; i < n && i < B200BO_N_BAND_INFO; ++i) out[i] = v[i];
  return 0;
}

int b200bo_debug_fast_check(b200bo_handle h, int64_t stride, int64_t max_samples, double* out, int n) {
  CHECK_ARG(h && out && n >= 8 && stride >= 1 && max_samples >= 1, "bad argument");
  if (!h->factored || !h->fast_ready || !h->last_fast.xdev)
    return set_err(B200BO_E_STATE, "debug_fast_check needs a preceding tensor-core call on this factorisation");
  CU_TRY(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const int64_t M = h->last_fast.M;
  const int D = h->D, Mc = h->Mc;
  const int ci = h->last_fast.nprod == 1 ? 0 : 1;
  const int64_t total = std::min<int64_t>((M + stride - 1) / stride, max_samples);
  int rc;
  if ((rc = ensure_predict_ws(h, 1, false))) return rc;
  CU_TRY(h->band_list.reserve(LIST0_CAP));
  CU_TRY(h->band_count.reserve(2));
  CU_TRY(h->Xband.reserve((size_t)Mc * D));
  CU_TRY(h->bd_ctl.reserve(1));
  LaunchCount lc;
  double ey = 0, es = 0, ra = 0;
  // half-widths to compare with: those of the last band pass, or (after predict()) the model's for this product count
  bd::BandCheckArgs ck;
  ck.y_fast = h->f_yhat.p; ck.ss_fast = h->f_sumsq.p; ck.df_fast = h->f_dotf.p;
  ck.y_ex = h->yhat.p; ck.ss_ex = h->sumsq.p; ck.df_ex = h->dotf.p; ck.list = h->band_list.p;
  ck.counts = h->band_count.p; ck.fused_err = h->err_flag.p; ck.estimate_trend = h->estimate_trend;
  ck.G = h->G; ck.sigma2 = h->sigma2;
  ck.dy = std::max(h->model.dy, h->calibrated[ci] ? h->dy_cal[ci] : 0.0) * h->widen[ci];
  ck.ds = (h->calibrated[ci] ? h->ds_cal[ci] : 0.0) * h->widen[ci];
  ck.ds_abs = h->model.ds_abs[ci] * h->widen[ci]; ck.ds_rel = h->model.ds_rel[ci] * h->widen[ci]; ck.du = h->model.du * h->widen[ci];
  ck.ctl = h->bd_ctl.p;
  const int step = std::min(Mc, LIST0_CAP);
  for (int64_t o = 0; o < total; o += step) {
    const int m = (int)std::min<int64_t>(step, total - o);
    fk::iota_stride_kernel<<<(m + 255) / 256, 256, 0, st>>>(h->band_list.p, m, stride);
    fk::list_offset_kernel<<<(m + 255) / 256, 256, 0, st>>>(h->band_list.p, m, o * stride);
    fk::band_gather_kernel<<<(int)(((size_t)m * D + 255) / 256), 256, 0, st>>>(h->last_fast.xdev, h->band_list.p, m, 0, D, h->Xband.p);
    CU_TRY(cudaGetLastError());
    if ((rc = fp64_moments(h, h->Xband.p, m, h->yhat.p, 1, nullptr, &lc))) return rc;
    const int cnt[2] = {m, m};
    CU_TRY(cudaMemcpyAsync(h->band_count.p, cnt, sizeof cnt, cudaMemcpyHostToDevice, st));
    ck.cap = m;
    bd::band_check_kernel<<<1, 256, 0, st>>>(ck);
    CU_TRY(cudaGetLastError());
    bd::BandCtl c;
    CU_TRY(cudaMemcpyAsync(&c, h->bd_ctl.p, sizeof c, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    ey = std::max(ey, c.err_y);
    es = std::max(es, c.err_s);
    ra = std::max(ra, c.ratio);
  }
  out[0] = (double)total; out[1] = ey; out[2] = es; out[3] = ra;
  out[4] = ck.dy; out[5] = ck.ds; out[6] = ck.ds_abs + ck.ds_rel; out[7] = h->last_fast.nprod;
  return 0;
}

int b200bo_get_timings(b200bo_handle h, double* out, int n) {
  CHECK_ARG(h && out && n >= 1, "bad argument");
  for (int i = 0; i < n && i < B200BO_N_TIMINGS; ++i) out[i] = h->timings[i];
  return 0;
}
int b200bo_get_fit_timings(b200bo_handle h, double* out, int n) {
  CHECK_ARG(h && out && n >= 1, "bad argument");
  for (int i = 0; i < n && i < B200BO_N_TIMINGS; ++i) out[i] = h->fit_timings[i];
  return 0;
}

}  // extern "C"
