// gamc_abi.cu -- C ABI (include/gamc_b200.h) over the sm_100a kernels.  Host-side orchestration only:
// buffer ownership, TMA descriptors, the per-iteration launch sequence of GAMC's `iterate!`
// (reference src/samplers/iterate/GAMC.jl:1-293), NCCL plumbing, profiling events.  No CPU arithmetic
// on the hot path: every target evaluation and every transition runs in the kernels of datapass.cuh,
// metric.cuh, sampler.cuh.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <set>
#include <map>
#include <dlfcn.h>

#include "ctx.h"
#include "common.cuh"
#include "datapass.cuh"
#include "metric.cuh"
#include "metric_plan.h"
#include "metric_i8_plan.h"
#include "sampler.cuh"
#include "synth.cuh"

using namespace gamc;

// ----------------------------------------------------------------------------------------------------
// minimal NCCL surface, resolved lazily with dlopen so that single-GPU use needs no NCCL at all
// ----------------------------------------------------------------------------------------------------
namespace {
struct NcclUniqueId { char internal[128]; };
typedef void* ncclComm_t;
typedef int (*fn_ncclGetUniqueId)(NcclUniqueId*);
typedef int (*fn_ncclCommInitRank)(ncclComm_t*, int, NcclUniqueId, int);
typedef int (*fn_ncclAllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef int (*fn_ncclAllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
typedef int (*fn_ncclCommDestroy)(ncclComm_t);
typedef const char* (*fn_ncclGetErrorString)(int);
struct NcclApi {
  void* lib = nullptr;
  fn_ncclGetUniqueId GetUniqueId = nullptr;
  fn_ncclCommInitRank CommInitRank = nullptr;
  fn_ncclAllReduce AllReduce = nullptr;
  fn_ncclAllGather AllGather = nullptr;
  fn_ncclCommDestroy CommDestroy = nullptr;
  fn_ncclGetErrorString GetErrorString = nullptr;
  bool load(std::string& err) {
    if (lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) { err = std::string("cannot dlopen libnccl: ") + dlerror(); return false; }
    GetUniqueId = (fn_ncclGetUniqueId)dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (fn_ncclCommInitRank)dlsym(lib, "ncclCommInitRank");
    AllReduce = (fn_ncclAllReduce)dlsym(lib, "ncclAllReduce");
    AllGather = (fn_ncclAllGather)dlsym(lib, "ncclAllGather");
    CommDestroy = (fn_ncclCommDestroy)dlsym(lib, "ncclCommDestroy");
    GetErrorString = (fn_ncclGetErrorString)dlsym(lib, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy) { err = "libnccl lacks required symbols"; return false; }
    return true;
  }
};
NcclApi g_nccl;
constexpr int NCCL_DOUBLE = 8;  // ncclFloat64
constexpr int NCCL_SUM = 0;

typedef CUresult (*fn_cuTensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
fn_cuTensorMapEncodeTiled g_encode = nullptr;

constexpr int PROF_POOL = GAMC_PROF_POOL;
}  // namespace

struct gamc_ctx : gamc_ctx_base {
  cudaStream_t side = nullptr; cudaEvent_t ev_main = nullptr, ev_side = nullptr;   // am_mode 2: speculative factor updates overlap the data pass

  // target
  long long N = 0; int p = 0; long long ld = 0; double lambda = 1.0;
  double* dX = nullptr; double* dy = nullptr; bool own_data = false;
  double* dw = nullptr;
  // data-pass configuration
  int dp_rpl = 1, dp_nu = 0, dp_grid = 0, dp_grid_now = 0, part_g_stride = 0;   // dp_grid_now: grid of the current evaluation (am_mode 2 leaves SMs free)
  CUtensorMap mapXA{}, mapXB{}, mapW{}, mapY{};
  bool metric_fused = false; MetricPlan* d_plans_fused = nullptr;   // level-2 evaluation in one pass (k_metric_fused), p <= 256
  // level-2 evaluation as ONE pass over X out of HBM: k_datapass_co and the metric kernel run at the same time (p <= 256)
  bool level2_co = false; unsigned long long* d_co_flags = nullptr; unsigned co_epoch = 0; long long co_lag_rows = 0;
  unsigned long long* d_co_dbg = nullptr;   // GAMC_CO_DEBUG=1: time stamps and CTA placement of the pair, printed by gamc_eval_upto_tensor
  double* part_ll = nullptr; double* part_g = nullptr;
  // metric on the integer tensor cores (metric_i8.cuh): digit planes of diag(sqrt w) X, column exponents, per-CTA partial tiles
  int metric_route = 0, route_slices = 0, route_smax = 0;   // gamc_set_metric_route
  bool metric_i8 = false, i8_fused = false; int i8_k = 0, i8_smax = 0, i8_grid = 0, i8_nitems = 0, i8_ntiles = 0, i8_nslots = 0; long long i8_ldn = 0;
  int8_t* d_i8S = nullptr; double *d_i8cs = nullptr, *d_i8cg = nullptr; unsigned long long* d_i8cmax = nullptr;
  long long* d_i8acc = nullptr; size_t i8_acc_bytes = 0;   // [item counter (16 bytes) | int64 group sums [tile][slot][128 x 128]]
  I8Kind* d_i8kinds = nullptr; I8Item* d_i8items = nullptr; int* d_i8ints = nullptr; CUtensorMap mapS{};
  int *d_i8tile_bi = nullptr, *d_i8tile_bj = nullptr;
  // metric
  MetricLayout layout; MetricPlan* d_plans = nullptr; TaskCoord* d_coords = nullptr;
  int nkinds = 0, ntasks = 0, metric_grid = 0; long long nslabs = 0; double* ws = nullptr; int* d_ints = nullptr;
  int *d_cta_kind = nullptr, *d_cta_split = nullptr, *d_kind_nsplit = nullptr, *d_task_nsplit = nullptr;
  // landing buffer of an evaluation = all-reduce payload [ll, g(p), pad, G(p x p)]
  double* R = nullptr; int goff = 0;
  double* theta_eval = nullptr;
  // sampler
  bool sampler_ready = false;
  gamc_config cfg{}; TuneCfg tune{};
  Bufs B{};
  double* vecs = nullptr; double* mats = nullptr; DevState* d_state = nullptr;
  unsigned char* arena = nullptr; size_t arena_bytes = 0;   // [state | vecs | R | mats]: the L2-persisting working set of the step kernels
  long long h_count = 0; bool h_present = true, h_past = false, spec_ready = false;
  long long h_tot_prop = 0, h_tot_totprop = 0;   // host mirror of the tuner's proposal counters: which steps are burn-in tuning events (am_mode 3)
  double* part_ll3 = nullptr; int part_ll3_stride = 0;   // per-CTA partials of the three-vector log-likelihood pass (am_mode 3)
  // tape
  long long tape_n = 0, tape_pos = 0, tape_cap = 0; int tape_p = 0;
  double* d_tape = nullptr; std::vector<double> h_usched;
  long long tape_gen_total = 0;
  // chain
  long long chain_cap = 0; double* d_chain = nullptr; unsigned char* d_accept = nullptr;
  // comm
  ncclComm_t comm = nullptr;
  // peer mailboxes (NVLink peer memory, CUDA IPC): scalar all-reduce of the AM step without NCCL or k_reduce
  PeerMail* mbox = nullptr; PeerMail** d_peers = nullptr; void* peer_ptr[32] = {nullptr}; bool p2p_ready = false; unsigned long long p2p_seq = 0;
  // peer landing: two reduced-evaluation buffers per rank, mapped by every peer; k_land sums them over NVLink (no NCCL on the step path)
  double* xR = nullptr; size_t xR_stride = 0; int xR_p = 0; double** d_peerR = nullptr; void* peerR_ptr[32] = {nullptr}; bool land_p2p = false;
  unsigned long long land_seq = 0;
  // generic target: the host evaluates the model (gamc_target_callback) or enqueues device work (gamc_target_device_callback)
  long long launches_cb = 0;
  gamc_target_fn cb = nullptr; void* cb_user = nullptr; double* cb_theta = nullptr; double* cb_R = nullptr;
  gamc_target_enqueue_fn dcb = nullptr; void* dcb_user = nullptr;
  // user schedule (gamc_set_schedule_callback)
  gamc_schedule_fn sched_fn = nullptr; void* sched_user = nullptr;
  // peer waits
  unsigned long long peer_timeout_ns = 30000000000ull; double* d_barrier = nullptr;
  // burn-in report records, per-sample log-target
  double* d_tune_log = nullptr; int tune_log_cap = 0; double* d_chain_lt = nullptr; double* d_monitor = nullptr;
  long long chain_saved_seen = 0;           // DevState.chain_saved at the last check_device_error
};

gamc_ctx_base* gamc_base(gamc_handle h) { return h; }
static inline bool generic_target(const gamc_ctx* h) { return h->cb != nullptr || h->dcb != nullptr; }

// Debug aid (GAMC_DEBUG=1): synchronise and report after every launch.
static bool g_debug = getenv("GAMC_DEBUG") != nullptr;
#define DEBUG_POINT(h, what)                                                                      \
  do {                                                                                            \
    if (g_debug) {                                                                                \
      cudaError_t e__ = cudaStreamSynchronize((h)->stream);                                       \
      cudaError_t l__ = cudaGetLastError();                                                       \
      if (e__ != cudaSuccess || l__ != cudaSuccess)                                               \
        fprintf(stderr, "[gamc debug] %s: sync=%s last=%s\n", what, cudaGetErrorString(e__), cudaGetErrorString(l__)); \
    }                                                                                             \
  } while (0)

namespace {

bool ensure_encode(gamc_ctx* h) {
  if (g_encode) return true;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || !fn) { h->err = "cuTensorMapEncodeTiled entry point not available"; return false; }
  g_encode = (fn_cuTensorMapEncodeTiled)fn;
  return true;
}

gamc_status make_map_2d(gamc_ctx* h, CUtensorMap* map, const double* base, long long N, int p, long long ld, int box_rows, int box_cols) {
  if (!ensure_encode(h)) return GAMC_CUDA_ERROR;
  cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)p};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
  cuuint32_t box[2] = {(cuuint32_t)box_rows, (cuuint32_t)box_cols};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { h->err = "cuTensorMapEncodeTiled(2d) failed with code " + std::to_string((int)r); return GAMC_CUDA_ERROR; }
  return GAMC_OK;
}

gamc_status make_map_1d(gamc_ctx* h, CUtensorMap* map, const double* base, long long N, int box) {
  if (!ensure_encode(h)) return GAMC_CUDA_ERROR;
  cuuint64_t dims[1] = {(cuuint64_t)N};
  cuuint64_t strides[1] = {8};  // unused for rank 1
  cuuint32_t bx[1] = {(cuuint32_t)box};
  cuuint32_t es[1] = {1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 1, (void*)base, dims, strides, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { h->err = "cuTensorMapEncodeTiled(1d) failed with code " + std::to_string((int)r); return GAMC_CUDA_ERROR; }
  return GAMC_OK;
}

void free_target(gamc_ctx* h) {
  if (h->own_data) { cudaFree(h->dX); cudaFree(h->dy); }
  h->dX = h->dy = nullptr; h->own_data = false;
  cudaFree(h->dw); h->dw = nullptr;
  cudaFree(h->part_ll); cudaFree(h->part_g); cudaFree(h->part_ll3); h->part_ll = h->part_g = h->part_ll3 = nullptr;
  cudaFree(h->d_plans_fused); h->d_plans_fused = nullptr; h->metric_fused = false;
  cudaFree(h->d_co_flags); h->d_co_flags = nullptr; h->level2_co = false;
  cudaFree(h->d_co_dbg); h->d_co_dbg = nullptr;
  cudaFree(h->d_i8S); cudaFree(h->d_i8cs); cudaFree(h->d_i8cg); cudaFree(h->d_i8acc); cudaFree(h->d_i8cmax); cudaFree(h->d_i8kinds); cudaFree(h->d_i8items); cudaFree(h->d_i8ints);
  h->d_i8S = nullptr; h->d_i8cs = h->d_i8cg = nullptr; h->d_i8acc = nullptr; h->d_i8cmax = nullptr; h->d_i8kinds = nullptr; h->d_i8items = nullptr; h->d_i8ints = nullptr; h->metric_i8 = false;
  cudaFree(h->d_plans); cudaFree(h->d_coords); cudaFree(h->ws); cudaFree(h->d_ints); h->d_plans = nullptr; h->d_coords = nullptr; h->ws = nullptr; h->d_ints = nullptr;
  cudaFree(h->arena); h->arena = nullptr; h->arena_bytes = 0; h->R = nullptr; h->vecs = h->mats = nullptr; h->d_state = nullptr;
  cudaFree(h->theta_eval); h->theta_eval = nullptr;
  cudaFreeHost(h->cb_theta); cudaFreeHost(h->cb_R); h->cb_theta = h->cb_R = nullptr; h->cb = nullptr; h->cb_user = nullptr; h->dcb = nullptr; h->dcb_user = nullptr;
}

void free_peer_landing(gamc_ctx* h) {
  for (int r = 0; r < h->nranks && r < 32; ++r) if (h->peerR_ptr[r] && r != h->rank) cudaIpcCloseMemHandle(h->peerR_ptr[r]);
  for (int r = 0; r < 32; ++r) h->peerR_ptr[r] = nullptr;
  cudaFree(h->xR); cudaFree(h->d_peerR);
  h->xR = nullptr; h->d_peerR = nullptr; h->land_p2p = false; h->xR_p = 0;
}

// Collective (every rank calls it from gamc_sampler_init): allocate this rank's two reduced-evaluation buffers and map every
// peer's pair (CUDA IPC).  Leaves land_p2p = false (NCCL path) unless it worked on every rank.
void setup_peer_landing(gamc_ctx* h) {
  if (!h->comm || !h->p2p_ready || getenv("GAMC_NO_P2P_LAND")) { h->land_p2p = false; return; }
  if (h->land_p2p && h->xR_p == h->p) return;
  free_peer_landing(h);
  const int n = h->nranks, p = h->p;
  h->xR_stride = ((size_t)h->goff + (size_t)p * p + 15) & ~(size_t)15;
  unsigned char* d_handles = nullptr;
  std::vector<cudaIpcMemHandle_t> handles(n);
  bool ok = cudaMalloc(&h->xR, sizeof(double) * 2 * h->xR_stride) == cudaSuccess &&
            cudaMemsetAsync(h->xR, 0, sizeof(double) * 2 * h->xR_stride, h->stream) == cudaSuccess &&
            cudaMalloc(&d_handles, sizeof(cudaIpcMemHandle_t) * n) == cudaSuccess &&
            cudaIpcGetMemHandle(&handles[h->rank], h->xR) == cudaSuccess &&
            cudaMemcpyAsync(d_handles + sizeof(cudaIpcMemHandle_t) * h->rank, &handles[h->rank], sizeof(cudaIpcMemHandle_t),
                            cudaMemcpyHostToDevice, h->stream) == cudaSuccess;
  if (d_handles) {
    const int r = g_nccl.AllGather(d_handles + sizeof(cudaIpcMemHandle_t) * h->rank, d_handles, sizeof(cudaIpcMemHandle_t), 0 /*ncclChar*/,
                                   h->comm, h->stream);
    ok = ok && r == 0 && cudaMemcpyAsync(handles.data(), d_handles, sizeof(cudaIpcMemHandle_t) * n, cudaMemcpyDeviceToHost, h->stream) == cudaSuccess;
  } else {
    ok = false;
  }
  ok = cudaStreamSynchronize(h->stream) == cudaSuccess && ok;
  cudaFree(d_handles);
  std::vector<double*> peers(n, nullptr);
  for (int r = 0; ok && r < n; ++r) {
    if (r == h->rank) { peers[r] = h->xR; h->peerR_ptr[r] = h->xR; continue; }
    void* ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, handles[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; break; }
    h->peerR_ptr[r] = ptr; peers[r] = (double*)ptr;
  }
  ok = ok && cudaMalloc(&h->d_peerR, sizeof(double*) * n) == cudaSuccess &&
       cudaMemcpy(h->d_peerR, peers.data(), sizeof(double*) * n, cudaMemcpyHostToDevice) == cudaSuccess;
  cudaGetLastError();
  double* d_ok = nullptr;
  bool all_ok = false;
  if (cudaMalloc(&d_ok, sizeof(double)) == cudaSuccess) {
    const double v = ok ? 1.0 : 0.0;
    double all = 0.0;
    cudaMemcpyAsync(d_ok, &v, sizeof(double), cudaMemcpyHostToDevice, h->stream);
    const int r = g_nccl.AllReduce(d_ok, d_ok, 1, NCCL_DOUBLE, 3 /*ncclMin*/, h->comm, h->stream);
    cudaMemcpyAsync(&all, d_ok, sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    cudaStreamSynchronize(h->stream);
    cudaFree(d_ok);
    all_ok = (r == 0 && all == 1.0);
  }
  h->land_p2p = all_ok; h->xR_p = p;
}

// Map every rank's mailbox array into this process (CUDA IPC over NVLink peer memory).  The handles travel through one
// ncclAllGather; the gather also orders every rank's zero-fill before any peer write.  GAMC_NO_P2P=1 disables the path.
void setup_peer_mailboxes(gamc_ctx* h) {
  h->p2p_ready = false;
  for (int r = 0; r < 32; ++r) { if (h->peer_ptr[r] && h->peer_ptr[r] != (void*)h->mbox) cudaIpcCloseMemHandle(h->peer_ptr[r]); h->peer_ptr[r] = nullptr; }
  cudaFree(h->mbox); cudaFree(h->d_peers); h->mbox = nullptr; h->d_peers = nullptr;      // a second gamc_comm_init starts afresh
  if (getenv("GAMC_NO_P2P") || h->nranks < 2 || h->nranks > 32 || !g_nccl.AllGather) return;
  const int n = h->nranks;
  unsigned char* d_handles = nullptr;
  std::vector<cudaIpcMemHandle_t> handles(n);
  bool ok = cudaMalloc(&h->mbox, sizeof(PeerMail) * 4 * n) == cudaSuccess &&     // [0,2n): AM scalars, [2n,4n): landing flags
            cudaMemsetAsync(h->mbox, 0, sizeof(PeerMail) * 4 * n, h->stream) == cudaSuccess &&
            cudaMalloc(&d_handles, sizeof(cudaIpcMemHandle_t) * n) == cudaSuccess &&
            cudaIpcGetMemHandle(&handles[h->rank], h->mbox) == cudaSuccess &&
            cudaMemcpyAsync(d_handles + sizeof(cudaIpcMemHandle_t) * h->rank, &handles[h->rank], sizeof(cudaIpcMemHandle_t),
                            cudaMemcpyHostToDevice, h->stream) == cudaSuccess;
  // every rank must take part in the gather, whatever happened above (a rank that failed contributes zeros)
  if (d_handles) {
    const int r = g_nccl.AllGather(d_handles + sizeof(cudaIpcMemHandle_t) * h->rank, d_handles, sizeof(cudaIpcMemHandle_t), 0 /*ncclChar*/,
                                   h->comm, h->stream);
    ok = ok && r == 0 && cudaMemcpyAsync(handles.data(), d_handles, sizeof(cudaIpcMemHandle_t) * n, cudaMemcpyDeviceToHost, h->stream) == cudaSuccess;
  } else {
    ok = false;
  }
  ok = cudaStreamSynchronize(h->stream) == cudaSuccess && ok;
  cudaFree(d_handles);
  std::vector<PeerMail*> peers(n, nullptr);
  for (int r = 0; ok && r < n; ++r) {
    if (r == h->rank) { peers[r] = h->mbox; h->peer_ptr[r] = h->mbox; continue; }
    void* ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, handles[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; break; }
    h->peer_ptr[r] = ptr; peers[r] = (PeerMail*)ptr;
  }
  ok = ok && cudaMalloc(&h->d_peers, sizeof(PeerMail*) * n) == cudaSuccess &&
       cudaMemcpy(h->d_peers, peers.data(), sizeof(PeerMail*) * n, cudaMemcpyHostToDevice) == cudaSuccess;
  cudaGetLastError();
  // all ranks must agree: a one-element all-reduce of the local verdict (min over ranks)
  double* d_ok = nullptr;
  if (cudaMalloc(&d_ok, sizeof(double)) == cudaSuccess) {
    const double v = ok ? 1.0 : 0.0;
    double all = 0.0;
    cudaMemcpyAsync(d_ok, &v, sizeof(double), cudaMemcpyHostToDevice, h->stream);
    const int r = g_nccl.AllReduce(d_ok, d_ok, 1, NCCL_DOUBLE, 3 /*ncclMin*/, h->comm, h->stream);
    cudaMemcpyAsync(&all, d_ok, sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    cudaStreamSynchronize(h->stream);
    cudaFree(d_ok);
    h->p2p_ready = (r == 0 && all == 1.0);
  }
  h->p2p_seq = 0;
}

void free_sampler(gamc_ctx* h) {
  cudaFree(h->d_chain); cudaFree(h->d_accept); cudaFree(h->d_tune_log); cudaFree(h->d_chain_lt);
  h->B = Bufs{}; h->d_chain = nullptr; h->d_accept = nullptr; h->d_tune_log = nullptr; h->d_chain_lt = nullptr; h->tune_log_cap = 0;
  h->sampler_ready = false;
}

// Keep the step kernels' working set (state, vectors, landing buffer, p x p matrices: a few MB) resident in
// the persisting part of L2: every data pass streams the 2 GB design matrix through L2 and would otherwise
// evict it, leaving the single-CTA linear-algebra kernels bound by DRAM latency.
void apply_l2_policy(gamc_ctx* h) {
  if (!h->arena) return;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, h->device) != cudaSuccess) return;
  if (prop.persistingL2CacheMaxSize <= 0 || prop.accessPolicyMaxWindowSize <= 0) return;
  const size_t want = std::min<size_t>(h->arena_bytes, (size_t)prop.accessPolicyMaxWindowSize);
  cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, std::max<size_t>(want * 2, 8u << 20)));
  cudaStreamAttrValue attr{};
  attr.accessPolicyWindow.base_ptr = h->arena;
  attr.accessPolicyWindow.num_bytes = want;
  attr.accessPolicyWindow.hitRatio = 1.0f;
  attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
  cudaGetLastError();   // best effort: a refusal only costs performance
}

// ---- data-pass dispatch -----------------------------------------------------------------------------
template <int NU, int RPL, bool GRAD>
gamc_status launch_dp_t(gamc_ctx* h, const double* theta, const int* skip) {
  const int grid = h->dp_grid_now > 0 ? h->dp_grid_now : h->dp_grid;
  auto kern = k_datapass<NU, RPL, GRAD>;
  { gamc_status s = ensure_attr(h, (const void*)kern, sizeof(DataPassSmem)); if (s != GAMC_OK) return s; }
  ProfScope ps(h, GRAD ? GAMC_K_LOGLIK_GRAD : GAMC_K_LOGLIK);
  DEBUG_POINT(h, "before datapass");
  kern<<<grid, DP_THREADS, sizeof(DataPassSmem), h->stream>>>(h->mapXA, h->dy, theta, h->p, h->N, h->dw, h->part_ll,
                                                                    h->part_g, h->part_g_stride, skip);
  CUDA_TRY(h, cudaGetLastError());
  DEBUG_POINT(h, "after datapass");
  return GAMC_OK;
}

template <int RPL, bool GRAD>
gamc_status launch_dp_nu(gamc_ctx* h, const double* theta, const int* skip) {
  switch (h->dp_nu) {
    case 1: return launch_dp_t<1, RPL, GRAD>(h, theta, skip);
    case 2: return launch_dp_t<2, RPL, GRAD>(h, theta, skip);
    case 3: return launch_dp_t<3, RPL, GRAD>(h, theta, skip);
    case 4: return launch_dp_t<4, RPL, GRAD>(h, theta, skip);
    case 5: return launch_dp_t<5, RPL, GRAD>(h, theta, skip);
    case 6: return launch_dp_t<6, RPL, GRAD>(h, theta, skip);
    case 7: return launch_dp_t<7, RPL, GRAD>(h, theta, skip);
    case 8: return launch_dp_t<8, RPL, GRAD>(h, theta, skip);
    default: h->err = "unsupported number of column units"; return GAMC_UNSUPPORTED_SHAPE;
  }
}

gamc_status launch_datapass(gamc_ctx* h, const double* theta, bool grad, const int* skip = nullptr) {
  if (h->dp_rpl == 2) return grad ? launch_dp_nu<2, true>(h, theta, skip) : launch_dp_nu<2, false>(h, theta, skip);
  return grad ? launch_dp_nu<1, true>(h, theta, skip) : launch_dp_nu<1, false>(h, theta, skip);
}

// log-likelihood at three parameter vectors in one pass (am_mode 3)
template <int NU>
gamc_status launch_ll3_t(gamc_ctx* h, const double* t0, const double* t1, const double* t2) {
  auto kern = k_datapass_ll3<NU>;
  { gamc_status s = ensure_attr(h, (const void*)kern, sizeof(DataPassLL3Smem)); if (s != GAMC_OK) return s; }
  ProfScope ps(h, GAMC_K_LOGLIK);
  kern<<<h->dp_grid, DP_THREADS, sizeof(DataPassLL3Smem), h->stream>>>(h->mapXA, h->dy, t0, t1, t2, h->p, h->N, h->part_ll3, h->part_ll3_stride);
  CUDA_TRY(h, cudaGetLastError());
  return GAMC_OK;
}
gamc_status launch_ll3(gamc_ctx* h, const double* t0, const double* t1, const double* t2) {
  switch (h->dp_nu) {
    case 1: return launch_ll3_t<1>(h, t0, t1, t2);
    case 2: return launch_ll3_t<2>(h, t0, t1, t2);
    case 3: return launch_ll3_t<3>(h, t0, t1, t2);
    case 4: return launch_ll3_t<4>(h, t0, t1, t2);
    case 5: return launch_ll3_t<5>(h, t0, t1, t2);
    case 6: return launch_ll3_t<6>(h, t0, t1, t2);
    case 7: return launch_ll3_t<7>(h, t0, t1, t2);
    case 8: return launch_ll3_t<8>(h, t0, t1, t2);
    default: h->err = "unsupported number of column units"; return GAMC_UNSUPPORTED_SHAPE;
  }
}

gamc_status launch_metric(gamc_ctx* h, const int* skip = nullptr) {
  auto kern = k_metric_t<MT_NSTAGE, false>;
  { gamc_status s = ensure_attr(h, (const void*)kern, sizeof(MetricSmem)); if (s != GAMC_OK) return s; }
  ProfScope ps(h, GAMC_K_METRIC);
  kern<<<h->metric_grid, MT_THREADS, sizeof(MetricSmem), h->stream>>>(h->mapXB, h->mapW, h->d_plans, h->d_cta_kind, h->d_cta_split,
                                                                    h->d_kind_nsplit, h->nslabs, h->ws, h->layout.max_nsplit, skip, MetricCo{});
  CUDA_TRY(h, cudaGetLastError());
  return GAMC_OK;
}

// Level-2 evaluation with ONE pass over X out of HBM: the data pass (log-likelihood, gradient, weights; HBM-bound, ~2 % of the FP64
// pipe) and the metric pass (FP64-tensor-bound, HBM nearly idle) run at the same time, one CTA of each per SM; the metric kernel
// is launched with programmatic stream serialisation so that it starts as soon as every data-pass CTA is resident, takes the
// weights of a slab when the data pass has published them, and re-reads the slab's rows of X from L2 (the data pass is throttled
// to stay co_lag_rows ahead at most).  One profiling scope for the pair (events between the two launches would serialise them).
template <int NU>
gamc_status launch_level2_co_t(gamc_ctx* h, const double* theta, const int* skip) {
  auto kd = k_datapass_co<NU>;
  auto km = k_metric_t<MT_NSTAGE_CO, true>;
  if (h->attr_set.find((const void*)kd) == h->attr_set.end()) {
    // both kernels must fit one SM together: let each ask for the largest shared-memory carve-out
    cudaFuncSetAttribute(kd, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(km, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }
  { gamc_status s = ensure_attr(h, (const void*)kd, sizeof(DataPassCoSmem<NU>)); if (s != GAMC_OK) return s; }
  { gamc_status s = ensure_attr(h, (const void*)km, sizeof(MetricSmemT<MT_NSTAGE_CO>)); if (s != GAMC_OK) return s; }
  ProfScope ps(h, GAMC_K_METRIC);
  h->launches += 1;   // two kernels in this scope
  h->co_epoch += 1;
  unsigned long long* progress = h->d_co_flags + 8;
  unsigned long long* front = h->d_co_flags;
  const int dgrid = h->dp_grid_now > 0 ? h->dp_grid_now : h->dp_grid;
  kd<<<dgrid, DPC_THREADS, sizeof(DataPassCoSmem<NU>), h->stream>>>(h->mapXA, h->dy, theta, h->p, h->N, h->dw, h->part_ll, h->part_g,
      h->part_g_stride, skip, progress, h->co_lag_rows > 0 ? front : nullptr, h->co_epoch, h->co_lag_rows, h->d_co_dbg);
  CUDA_TRY(h, cudaGetLastError());
  MetricCo co{h->dw, progress, front, h->N, dgrid, h->co_epoch, h->d_co_dbg};
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)h->metric_grid); cfg.blockDim = dim3(MT_THREADS);
  cfg.dynamicSmemBytes = sizeof(MetricSmemT<MT_NSTAGE_CO>); cfg.stream = h->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CUDA_TRY(h, cudaLaunchKernelEx(&cfg, km, h->mapXB, h->mapW, (const MetricPlan*)h->d_plans, (const int*)h->d_cta_kind, (const int*)h->d_cta_split,
                                 (const int*)h->d_kind_nsplit, h->nslabs, h->ws, h->layout.max_nsplit, skip, co));
  return GAMC_OK;
}

gamc_status launch_level2_co(gamc_ctx* h, const double* theta, const int* skip) {
  switch (h->dp_nu) {
    case 1: return launch_level2_co_t<1>(h, theta, skip);
    case 2: return launch_level2_co_t<2>(h, theta, skip);
    case 3: return launch_level2_co_t<3>(h, theta, skip);
    case 4: return launch_level2_co_t<4>(h, theta, skip);
    default: h->err = "co-scheduled level-2 evaluation needs p <= 256"; return GAMC_UNSUPPORTED_SHAPE;
  }
}

// Level-2 evaluation in one pass over X (k_metric_fused): log-likelihood and gradient partials per metric CTA, metric tiles.
gamc_status launch_metric_fused(gamc_ctx* h, const double* theta, const int* skip = nullptr) {
  { gamc_status s = ensure_attr(h, (const void*)k_metric_fused, sizeof(MetricSmemF)); if (s != GAMC_OK) return s; }
  ProfScope ps(h, GAMC_K_METRIC);
  k_metric_fused<<<h->metric_grid, MTF_THREADS, sizeof(MetricSmemF), h->stream>>>(h->mapXB, h->mapY, h->d_plans_fused, h->d_cta_kind, h->d_cta_split,
      h->d_kind_nsplit, h->nslabs, h->ws, h->layout.max_nsplit, theta, h->p, h->N, h->part_ll, h->part_g, h->part_g_stride, h->nkinds, skip);
  CUDA_TRY(h, cudaGetLastError());
  return GAMC_OK;
}

// Metric on the integer tensor cores: digit planes of diag(sqrt w) X (k_i8_slice, after the gradient pass has written w), then the
// int8 GEMMs (k_metric_i8: TMA -> tcgen05.mma kind::i8 -> TMEM), partial tiles summed by k_reduce_i8 inside launch_reduce.
static double i8_bias(int k) { double b = 0.0; for (int i = 0; i + 1 < k; ++i) b += ldexp(128.0, 8 * i); return b; }

// gradient pass that also writes the digit planes (one read of X for the whole level-2 evaluation of the integer route)
template <int NU>
gamc_status launch_dp_i8_t(gamc_ctx* h, const double* theta, const int* skip) {
  const int grid = h->dp_grid_now > 0 ? h->dp_grid_now : h->dp_grid;
  auto kern = k_datapass_i8<NU>;
  { gamc_status s = ensure_attr(h, (const void*)kern, sizeof(DataPassI8Smem)); if (s != GAMC_OK) return s; }
  ProfScope ps(h, GAMC_K_LOGLIK_GRAD);
  kern<<<grid, DPI_THREADS, sizeof(DataPassI8Smem), h->stream>>>(h->mapXA, h->dy, theta, h->p, h->N, h->dw, h->part_ll, h->part_g, h->part_g_stride,
                                                                   h->d_i8cs, h->i8_k, h->d_i8S, i8_bias(h->i8_k), skip);
  CUDA_TRY(h, cudaGetLastError());
  return GAMC_OK;
}
gamc_status launch_dp_i8(gamc_ctx* h, const double* theta, const int* skip) {
  switch (h->dp_nu) {
    case 1: return launch_dp_i8_t<1>(h, theta, skip);
    case 2: return launch_dp_i8_t<2>(h, theta, skip);
    case 3: return launch_dp_i8_t<3>(h, theta, skip);
    case 4: return launch_dp_i8_t<4>(h, theta, skip);
    case 5: return launch_dp_i8_t<5>(h, theta, skip);
    case 6: return launch_dp_i8_t<6>(h, theta, skip);
    case 7: return launch_dp_i8_t<7>(h, theta, skip);
    case 8: return launch_dp_i8_t<8>(h, theta, skip);
    default: h->err = "unsupported number of column units"; return GAMC_UNSUPPORTED_SHAPE;
  }
}

gamc_status launch_metric_i8(gamc_ctx* h, const int* skip = nullptr) {
  { gamc_status s = ensure_attr(h, (const void*)k_metric_i8, I8_SMEM_BYTES); if (s != GAMC_OK) return s; }
  ProfScope ps(h, GAMC_K_METRIC);
  if (!h->i8_fused) {
    h->launches += 1;   // two kernels in this scope
    const dim3 sgrid((unsigned)((h->N + I8_SLICE_ROWS - 1) / I8_SLICE_ROWS), (unsigned)((h->p + I8_SLICE_COLS - 1) / I8_SLICE_COLS));
    k_i8_slice<<<sgrid, 256, 0, h->stream>>>(h->dX, h->ld, h->dw, h->N, h->p, h->d_i8cs, h->i8_k, h->d_i8S, i8_bias(h->i8_k), skip);
    CUDA_TRY(h, cudaGetLastError());
  }
  CUDA_TRY(h, cudaMemsetAsync(h->d_i8acc, 0, h->i8_acc_bytes, h->stream));   // item counter and int64 group sums
  k_metric_i8<<<h->i8_grid, I8_THREADS, I8_SMEM_BYTES, h->stream>>>(h->mapS, h->d_i8kinds, h->d_i8items, h->i8_nitems, reinterpret_cast<int*>(h->d_i8acc),
                                                                    h->d_i8acc + 2, h->i8_nslots, skip);
  CUDA_TRY(h, cudaGetLastError());
  return GAMC_OK;
}

// level: 0 = ll, 1 = ll + grad, 2 = ll + grad + metric
gamc_status launch_reduce(gamc_ctx* h, int level, double* E = nullptr, const int* skip = nullptr) {
  ProfScope ps(h, GAMC_K_REDUCE);
  const int nblk = (level >= 2 ? h->ntasks * 4 : 0) + (level >= 1 ? (h->p + 31) / 32 : 0) + 1;   // tensor tiles | gradient column blocks | scalar
  const int nparts = (level >= 2 && h->metric_fused) ? h->metric_grid : (h->dp_grid_now > 0 ? h->dp_grid_now : h->dp_grid);
  if (level >= 2 && h->metric_i8) {
    const int nb1 = (h->p + 31) / 32 + 1;
    k_reduce<<<nb1, 256, 0, h->stream>>>(E ? E : h->R, h->p, h->goff, h->part_ll, h->part_g, h->part_g_stride, nparts, 1,
                                         nullptr, h->d_coords, h->ntasks, h->d_task_nsplit, h->layout.max_nsplit, skip);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    k_reduce_i8<<<h->i8_ntiles * (I8_TILE * I8_TILE / 256), 256, 0, h->stream>>>(E ? E : h->R, h->p, h->goff, h->d_i8acc + 2, h->i8_smax, h->d_i8tile_bi,
        h->d_i8tile_bj, h->d_i8cg, skip);
    CUDA_TRY(h, cudaGetLastError());
    return GAMC_OK;
  }
  k_reduce<<<nblk, 256, 0, h->stream>>>(E ? E : h->R, h->p, h->goff, h->part_ll, h->part_g, h->part_g_stride, nparts, level >= 1,
                                        level >= 2 ? h->ws : nullptr, h->d_coords, h->ntasks, h->d_task_nsplit, h->layout.max_nsplit, skip);
  CUDA_TRY(h, cudaGetLastError());
  return GAMC_OK;
}

// Likelihood partial sums at a device-resident theta into the landing buffer R, then (multi-GPU) all-reduce.
// Generic target: fetch theta, let the host evaluate, upload [lt, g, G] into the landing buffer (one round trip).
gamc_status eval_callback(gamc_ctx* h, const double* d_theta, int level) {
  const int p = h->p;
  CUDA_TRY(h, cudaMemcpyAsync(h->cb_theta, d_theta, sizeof(double) * p, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  double lt = NAN;
  const int rc = h->cb(h->cb_theta, p, level, &lt, h->cb_R + 1, h->cb_R + h->goff, h->cb_user);
  if (rc != 0) { h->err = "target callback returned " + std::to_string(rc); return GAMC_INVALID_ARG; }
  h->cb_R[0] = lt;
  const size_t count = level == 0 ? 1 : (level == 1 ? (size_t)(1 + p) : (size_t)h->goff + (size_t)p * p);
  CUDA_TRY(h, cudaMemcpyAsync(h->R, h->cb_R, sizeof(double) * count, cudaMemcpyHostToDevice, h->stream));
  h->launches_cb += 1;
  return GAMC_OK;
}

// The kernels of a level-2 evaluation before the reduction: one fused pass when every CTA can hold complete rows (p <= 256),
// otherwise the gradient pass followed by the metric pass.
gamc_status launch_level2(gamc_ctx* h, const double* d_theta, const int* skip = nullptr) {
  if (h->metric_i8) {
    gamc_status s8 = h->i8_fused ? launch_dp_i8(h, d_theta, skip) : launch_datapass(h, d_theta, true, skip);
    if (s8 != GAMC_OK) return s8;
    return launch_metric_i8(h, skip);
  }
  if (h->metric_fused) return launch_metric_fused(h, d_theta, skip);
  if (h->level2_co) return launch_level2_co(h, d_theta, skip);
  gamc_status s = launch_datapass(h, d_theta, true, skip);
  if (s != GAMC_OK) return s;
  return launch_metric(h, skip);
}

// skip_reduce (level 0, single GPU only): leave the per-CTA partials for k_am_post to sum.
gamc_status eval_device(gamc_ctx* h, const double* d_theta, int level, bool allreduce, bool skip_reduce = false) {
  if (h->cb) return eval_callback(h, d_theta, level);
  if (h->dcb) {
    // device-side generic target: the user's work is enqueued on the handle's stream straight into the landing buffer
    const int rc = h->dcb(d_theta, h->p, level, h->R, (void*)h->stream, h->dcb_user);
    if (rc != 0) { h->err = "device target callback returned " + std::to_string(rc); return GAMC_INVALID_ARG; }
    h->launches_cb += 1;
    return GAMC_OK;
  }
  gamc_status s = level >= 2 ? launch_level2(h, d_theta) : launch_datapass(h, d_theta, level >= 1);
  if (s != GAMC_OK) return s;
  if (skip_reduce && level == 0 && (!h->comm || h->p2p_ready)) return GAMC_OK;
  s = launch_reduce(h, level);
  if (s != GAMC_OK) return s;
  if (allreduce && h->comm) {
    ProfScope ps(h, GAMC_K_ALLREDUCE);
    const size_t count = level == 0 ? 1 : (level == 1 ? (size_t)(1 + h->p) : (size_t)h->goff + (size_t)h->p * h->p);
    int r = g_nccl.AllReduce(h->R, h->R, count, NCCL_DOUBLE, NCCL_SUM, h->comm, h->stream);
    if (r != 0) { h->err = std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"); return GAMC_NCCL_ERROR; }
  }
  return GAMC_OK;
}

// prior terms for the standalone evaluation entry points
__global__ void __launch_bounds__(256) k_eval_finish(double* R, int goff, const double* theta, int p, double lambda, int level) {
  __shared__ double red[40];
  if (level == 0) {
    const double lp = log_prior(theta, p, lambda, red);
    if (threadIdx.x == 0) R[0] += lp;
  } else if (level == 1) {
    const double lp = log_prior(theta, p, lambda, red);
    for (int i = threadIdx.x; i < p; i += blockDim.x) R[1 + i] -= theta[i] / lambda;
    if (threadIdx.x == 0) R[0] += lp;
  } else {
    const double lt = add_prior(R, goff, theta, p, lambda, red);
    if (threadIdx.x == 0) R[0] = lt;
  }
}

// Working set of the step kernels: [state | vectors | landing buffer R | p x p matrices], plus the evaluation point.
gamc_status alloc_arena(gamc_ctx* h) {
  const int p = h->p;
  h->goff = (1 + p + 1) & ~1;  // G starts 16-byte aligned
  {
    const size_t pp = (size_t)p * p;
    const size_t n_state = 64, n_vecs = ((size_t)22 * p + 16 + 15) & ~(size_t)15, n_R = ((size_t)h->goff + pp + 15) & ~(size_t)15, n_mats = 9 * pp + 2 * la_invd_doubles(p);
    h->arena_bytes = sizeof(double) * (n_state + n_vecs + n_R + n_mats);
    CUDA_TRY(h, cudaMalloc(&h->arena, h->arena_bytes));
    CUDA_TRY(h, cudaMemsetAsync(h->arena, 0, h->arena_bytes, h->stream));
    double* base = reinterpret_cast<double*>(h->arena);
    static_assert(sizeof(DevState) <= 64 * sizeof(double), "DevState must fit its arena slot");
    h->d_state = reinterpret_cast<DevState*>(base);
    h->vecs = base + n_state;
    h->R = h->vecs + n_vecs;
    h->mats = h->R + n_R;
    apply_l2_policy(h);
  }
  CUDA_TRY(h, cudaMalloc(&h->theta_eval, sizeof(double) * p));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return GAMC_OK;
}

gamc_status setup_metric_i8(gamc_ctx* h, int k, int smax);

gamc_status setup_target(gamc_ctx* h) {
  const int p = h->p;
  REQUIRE(h, p >= 1 && p <= 512, GAMC_UNSUPPORTED_SHAPE, "p must be in 1..512");
  REQUIRE(h, (h->ld % 2) == 0 && ((uintptr_t)h->dX % 16) == 0, GAMC_INVALID_ARG, "device X must be 16-byte aligned with an even leading dimension");
  const char* env = getenv("GAMC_DP_RPL");
  h->dp_rpl = (env && atoi(env) == 2 && p <= 256) ? 2 : 1;
  const int CU = 64 / h->dp_rpl, RB = 32 * h->dp_rpl;
  h->dp_nu = (p + CU - 1) / CU;
  h->part_g_stride = h->dp_nu * CU;
  const long long nrb = (h->N + RB - 1) / RB;
  h->dp_grid = (int)std::min<long long>(nrb, h->num_sms);
  if (h->dp_grid < 1) h->dp_grid = 1;
  gamc_status s = make_map_2d(h, &h->mapXA, h->dX, h->N, p, h->ld, RB, CU);
  if (s != GAMC_OK) return s;
  s = make_map_2d(h, &h->mapXB, h->dX, h->N, p, h->ld, MT_SLAB, MT_BLK);
  if (s != GAMC_OK) return s;
  CUDA_TRY(h, cudaMalloc(&h->dw, sizeof(double) * (size_t)(h->N + 64)));
  CUDA_TRY(h, cudaMemsetAsync(h->dw, 0, sizeof(double) * (size_t)(h->N + 64), h->stream));
  s = make_map_1d(h, &h->mapW, h->dw, h->N, MT_SLAB);
  if (s != GAMC_OK) return s;
  {
    const int rows = std::max(h->dp_grid, h->num_sms + 8);   // the fused level-2 kernel writes one row per metric CTA
    CUDA_TRY(h, cudaMalloc(&h->part_ll, sizeof(double) * rows));
    CUDA_TRY(h, cudaMalloc(&h->part_ll3, sizeof(double) * 3 * rows));
    h->part_ll3_stride = rows;
    CUDA_TRY(h, cudaMalloc(&h->part_g, sizeof(double) * (size_t)rows * h->part_g_stride));
  }
  // metric layout
  h->layout = build_metric_layout(p);
  h->nkinds = (int)h->layout.plans.size();
  h->ntasks = (int)h->layout.coords.size();
  h->nslabs = (h->N + MT_SLAB - 1) / MT_SLAB;
  assign_splits(h->layout, std::max(h->num_sms, h->nkinds), h->nslabs);
  h->metric_grid = (int)h->layout.cta_kind.size();
  {
    std::vector<int> ints;
    auto push = [&](const std::vector<int>& v) { size_t off = ints.size(); ints.insert(ints.end(), v.begin(), v.end()); return off; };
    const size_t o1 = push(h->layout.cta_kind), o2 = push(h->layout.cta_split), o3 = push(h->layout.kind_nsplit), o4 = push(h->layout.task_nsplit);
    CUDA_TRY(h, cudaMalloc(&h->d_ints, sizeof(int) * ints.size()));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_ints, ints.data(), sizeof(int) * ints.size(), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->d_cta_kind = h->d_ints + o1; h->d_cta_split = h->d_ints + o2; h->d_kind_nsplit = h->d_ints + o3; h->d_task_nsplit = h->d_ints + o4;
  }
  CUDA_TRY(h, cudaMalloc(&h->d_plans, sizeof(MetricPlan) * h->nkinds));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_plans, h->layout.plans.data(), sizeof(MetricPlan) * h->nkinds, cudaMemcpyHostToDevice, h->stream));
  {
    // fused level-2 evaluation: every kind keeps ALL column blocks resident (the ones its tasks do not use are appended), so
    // each CTA can form eta = X theta for its slab itself.  Opt-in (GAMC_FUSED_METRIC=1) until it beats the two-pass route.
    const int nblk = (p + MT_BLK - 1) / MT_BLK;
    h->metric_fused = nblk <= MT_MAXS && h->metric_grid <= h->num_sms + 8 && h->part_g_stride >= nblk * MT_BLK && getenv("GAMC_FUSED_METRIC");
    if (h->metric_fused) {
      std::vector<MetricPlan> fp = h->layout.plans;
      for (auto& pl : fp)
        for (int b = 0; b < nblk; ++b) {
          bool have = false;
          for (int sl = 0; sl < pl.nslots; ++sl) have = have || pl.slot_block[sl] == b;
          if (!have) pl.slot_block[pl.nslots++] = b;
        }
      CUDA_TRY(h, cudaMalloc(&h->d_plans_fused, sizeof(MetricPlan) * h->nkinds));
      CUDA_TRY(h, cudaMemcpyAsync(h->d_plans_fused, fp.data(), sizeof(MetricPlan) * h->nkinds, cudaMemcpyHostToDevice, h->stream));
      CUDA_TRY(h, cudaStreamSynchronize(h->stream));
      gamc_status sy = make_map_1d(h, &h->mapY, h->dy, h->N, MT_SLAB);
      if (sy != GAMC_OK) return sy;
    }
  }
  {
    // co-scheduled level-2 evaluation (default for p <= 256; GAMC_LEVEL2_CO=0 keeps the two passes one after the other).
    // GAMC_CO_LAG_MB: how far (in bytes of X) the data pass may run ahead of the metric sweep; 0 = no throttle.
    const char* e = getenv("GAMC_LEVEL2_CO");
    h->level2_co = !h->metric_fused && h->dp_rpl == 1 && h->dp_nu <= 4 && !(e && atoi(e) == 0);
    if (h->level2_co) {
      const char* lg = getenv("GAMC_CO_LAG_MB");
      const double lag_mb = lg ? atof(lg) : 48.0;
      h->co_lag_rows = (long long)(lag_mb * 1048576.0 / (8.0 * p));
      // a metric CTA sweeps every nsplit-th slab, so the frontier moves in steps of nsplit slabs: a smaller lag would make the two kernels wait for each other
      if (h->co_lag_rows > 0) h->co_lag_rows = std::max<long long>(h->co_lag_rows, (long long)MT_SLAB * (h->layout.max_nsplit + 1) + 64);
      CUDA_TRY(h, cudaMalloc(&h->d_co_flags, sizeof(unsigned long long) * (8 + h->num_sms)));
      CUDA_TRY(h, cudaMemsetAsync(h->d_co_flags, 0, sizeof(unsigned long long) * (8 + h->num_sms), h->stream));
      h->co_epoch = 0;
      if (getenv("GAMC_CO_DEBUG")) {
        CUDA_TRY(h, cudaMalloc(&h->d_co_dbg, sizeof(unsigned long long) * 1024));
        std::vector<unsigned long long> z(1024, 0ull);
        z[0] = ~0ull; z[2] = ~0ull;
        CUDA_TRY(h, cudaMemcpy(h->d_co_dbg, z.data(), sizeof(unsigned long long) * 1024, cudaMemcpyHostToDevice));
      }
    }
  }
  CUDA_TRY(h, cudaMalloc(&h->d_coords, sizeof(TaskCoord) * h->ntasks));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_coords, h->layout.coords.data(), sizeof(TaskCoord) * h->ntasks, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaMalloc(&h->ws, sizeof(double) * (size_t)h->layout.max_nsplit * h->ntasks * MT_BLK * MT_BLK));
  {
    // metric route (gamc_set_metric_route; GAMC_METRIC_I8 = "0" | "1" | "slices,smax" overrides it)
    int route = h->metric_route, k = h->route_slices > 0 ? h->route_slices : 6, smax = h->route_smax > 0 ? h->route_smax : k + 1;
    if (const char* e8 = getenv("GAMC_METRIC_I8")) {
      int ek = 0, es = 0;
      if (sscanf(e8, "%d,%d", &ek, &es) == 2) { route = GAMC_METRIC_I8; k = ek; smax = es; }
      else route = atoi(e8) != 0 ? GAMC_METRIC_I8 : GAMC_METRIC_F64;
    }
    const bool automatic = route == GAMC_METRIC_AUTO;
    if (automatic) route = ((double)h->N * (double)p >= 4194304.0) ? GAMC_METRIC_I8 : GAMC_METRIC_F64;
    if (route == GAMC_METRIC_I8) {
      gamc_status s8 = setup_metric_i8(h, k, smax);
      if (s8 != GAMC_OK) {
        if (!automatic) return s8;
        // automatic route: the digit planes (slices / 8 of the size of X) did not fit beside X -- evaluate the metric on the FP64 route
        cudaGetLastError();
        cudaFree(h->d_i8S); cudaFree(h->d_i8cs); cudaFree(h->d_i8cg); cudaFree(h->d_i8acc); cudaFree(h->d_i8cmax); cudaFree(h->d_i8kinds); cudaFree(h->d_i8items); cudaFree(h->d_i8ints);
        h->d_i8S = nullptr; h->d_i8cs = h->d_i8cg = nullptr; h->d_i8acc = nullptr; h->d_i8cmax = nullptr; h->d_i8kinds = nullptr; h->d_i8items = nullptr; h->d_i8ints = nullptr;
        h->metric_i8 = false; h->err.clear();
      }
    }
  }
  return alloc_arena(h);
}


// Integer-tensor-core metric: column exponents of X (once), digit-plane buffer, tensor map, work plan.
gamc_status setup_metric_i8(gamc_ctx* h, int k, int smax) {
  const int p = h->p;
  REQUIRE(h, k >= 2 && k <= I8_MAXK && smax >= 2 && smax <= 2 * k, GAMC_INVALID_ARG, "metric_i8: 2 <= slices <= 7, 2 <= smax <= 2 slices");
  REQUIRE(h, h->N < (1ll << 31) - 256, GAMC_UNSUPPORTED_SHAPE, "metric_i8: too many rows for one rank");
  h->i8_k = k; h->i8_smax = smax;
  h->i8_ldn = (h->N + I8_SLICE_ROWS - 1) / I8_SLICE_ROWS * I8_SLICE_ROWS;
  const long long nchunks = h->i8_ldn / I8_KC;
  CUDA_TRY(h, cudaMalloc(&h->d_i8S, (size_t)k * p * h->i8_ldn));
  CUDA_TRY(h, cudaMemsetAsync(h->d_i8S, 0, (size_t)k * p * h->i8_ldn, h->stream));   // rows between N and the end of the last chunk stay zero
  { const char* ef = getenv("GAMC_I8_FUSED"); h->i8_fused = h->dp_rpl == 1 && !(ef && atoi(ef) == 0); }
  CUDA_TRY(h, cudaMalloc(&h->d_i8cs, sizeof(double) * p));
  CUDA_TRY(h, cudaMalloc(&h->d_i8cg, sizeof(double) * p));
  CUDA_TRY(h, cudaMalloc(&h->d_i8cmax, sizeof(unsigned long long) * p));
  CUDA_TRY(h, cudaMemsetAsync(h->d_i8cmax, 0, sizeof(unsigned long long) * p, h->stream));
  {
    const int ny = (int)std::max<long long>(1, std::min<long long>(64, (h->N + 4095) / 4096));
    k_i8_colmax<<<dim3((unsigned)p, (unsigned)ny), 256, 0, h->stream>>>(h->dX, h->ld, h->N, p, h->d_i8cmax);
    CUDA_TRY(h, cudaGetLastError());
    k_i8_scales<<<(p + 127) / 128, 128, 0, h->stream>>>(h->d_i8cmax, p, 8 * k - 2, h->d_i8cs, h->d_i8cg);
    CUDA_TRY(h, cudaGetLastError());
  }
  {
    if (!ensure_encode(h)) return GAMC_CUDA_ERROR;
    cuuint64_t dims[4] = {(cuuint64_t)I8_KC, (cuuint64_t)p, (cuuint64_t)k, (cuuint64_t)nchunks};
    cuuint64_t strides[3] = {(cuuint64_t)I8_KC, (cuuint64_t)I8_KC * (cuuint64_t)p, (cuuint64_t)I8_KC * (cuuint64_t)p * (cuuint64_t)k};
    cuuint32_t box[4] = {(cuuint32_t)I8_KC, (cuuint32_t)I8_TILE, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = g_encode(&h->mapS, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, (void*)h->d_i8S, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          I8_KC == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { h->err = "cuTensorMapEncodeTiled(digit planes) failed with code " + std::to_string((int)r); return GAMC_CUDA_ERROR; }
  }
  int block_chunks = 0;
  if (const char* eb = getenv("GAMC_I8_BLOCK")) { const int v = atoi(eb); if (v >= 1 && v <= 4096) block_chunks = v; }
  const I8Layout L = build_i8_layout(p, k, smax, h->num_sms, nchunks, block_chunks);
  h->i8_nitems = (int)L.items.size();
  h->i8_grid = std::min(h->num_sms, h->i8_nitems);
  h->i8_ntiles = (int)L.tile_bi.size();
  h->i8_nslots = L.nslots;
  {
    std::vector<int> ints;
    auto push = [&](const std::vector<int>& v) { size_t off = ints.size(); ints.insert(ints.end(), v.begin(), v.end()); return off; };
    const size_t o4 = push(L.tile_bi), o5 = push(L.tile_bj);
    CUDA_TRY(h, cudaMalloc(&h->d_i8ints, sizeof(int) * ints.size()));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_i8ints, ints.data(), sizeof(int) * ints.size(), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMalloc(&h->d_i8kinds, sizeof(I8Kind) * L.kinds.size()));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_i8kinds, L.kinds.data(), sizeof(I8Kind) * L.kinds.size(), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMalloc(&h->d_i8items, sizeof(I8Item) * L.items.size()));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_i8items, L.items.data(), sizeof(I8Item) * L.items.size(), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->d_i8tile_bi = h->d_i8ints + o4; h->d_i8tile_bj = h->d_i8ints + o5;
  }
  h->i8_acc_bytes = 16 + sizeof(long long) * (size_t)h->i8_ntiles * h->i8_nslots * I8_TILE * I8_TILE;
  CUDA_TRY(h, cudaMalloc(&h->d_i8acc, h->i8_acc_bytes));
  h->metric_i8 = true;
  return GAMC_OK;
}

// hand-over error word of the co-scheduled level-2 evaluation (after a stream synchronisation)
gamc_status check_co_error(gamc_ctx* h) {
  if (!h->level2_co || !h->d_co_flags) return GAMC_OK;
  unsigned long long w = 0;
  CUDA_TRY(h, cudaMemcpyAsync(&w, h->d_co_flags + 1, sizeof(w), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  if (w) { h->err = "co-scheduled level-2 evaluation: the metric kernel timed out waiting for the data pass"; return GAMC_CUDA_ERROR; }
  return GAMC_OK;
}

// GAMC_CO_DEBUG=1: when and where the CTAs of the last co-scheduled pair ran (stderr), then re-arm the stamps
void co_debug_report(gamc_ctx* h) {
  std::vector<unsigned long long> d(1024);
  cudaMemcpy(d.data(), h->d_co_dbg, sizeof(unsigned long long) * 1024, cudaMemcpyDeviceToHost);
  if (d[0] != ~0ull && d[2] != ~0ull) {
    int max_dp = 0, max_mt = 0, both = 0, sms = 0;
    for (int i = 0; i < 512; ++i) {
      const int ndp = (int)(d[8 + i] & 0xffffffffull), nmt = (int)(d[8 + i] >> 32);
      max_dp = std::max(max_dp, ndp); max_mt = std::max(max_mt, nmt); both += (ndp > 0 && nmt > 0); sms += (ndp > 0 || nmt > 0);
    }
    int occ_dp = -1, occ_mt = -1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_dp, k_datapass_co<4>, DPC_THREADS, sizeof(DataPassCoSmem<4>));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_mt, k_metric_t<MT_NSTAGE_CO, true>, MT_THREADS, sizeof(MetricSmemT<MT_NSTAGE_CO>));
    fprintf(stderr, "[co] data pass runs %.1f us; first metric CTA starts %+.1f us after the first data-pass CTA; SMs used %d, with both kernels %d, "
                    "max CTAs per SM: data pass %d, metric %d; occupancy alone: %d / %d\n",
            (d[1] - d[0]) * 1e-3, ((double)d[2] - (double)d[0]) * 1e-3, sms, both, max_dp, max_mt, occ_dp, occ_mt);
    if (d[4] && d[5] && d[6])
      fprintf(stderr, "[co] data-pass CTA 0 reaches 25 / 50 / 75 %% of its row blocks at +%.0f / +%.0f / +%.0f us\n",
              ((double)d[4] - (double)d[0]) * 1e-3, ((double)d[5] - (double)d[0]) * 1e-3, ((double)d[6] - (double)d[0]) * 1e-3);
  }
  std::vector<unsigned long long> z(1024, 0ull);
  z[0] = ~0ull; z[2] = ~0ull;
  cudaMemcpy(h->d_co_dbg, z.data(), sizeof(unsigned long long) * 1024, cudaMemcpyHostToDevice);
}

gamc_status check_device_error(gamc_ctx* h) {
  if (!h->B.st) return check_co_error(h);
  DevState st;
  unsigned long long co_word = 0;
  if (h->level2_co && h->d_co_flags) CUDA_TRY(h, cudaMemcpyAsync(&co_word, h->d_co_flags + 1, sizeof(co_word), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(&st, h->B.st, sizeof(DevState), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  if (co_word) { h->err = "co-scheduled level-2 evaluation: the metric kernel timed out waiting for the data pass"; return GAMC_CUDA_ERROR; }
  h->chain_saved_seen = st.chain_saved;
  if (st.error_flag == GAMC_NONFINITE_INIT) {
    h->err = "Log-target, gradient or tensor of log-target not finite: initial values out of support";
    return GAMC_NONFINITE_INIT;
  }
  if (st.error_flag == GAMC_NOT_POSDEF) {
    h->err = "matrix is not positive definite; Cholesky factorization failed at pivot " + std::to_string(st.error_pivot) +
             " (iteration " + std::to_string(st.error_iter) + ")";
    return GAMC_NOT_POSDEF;
  }
  if (st.error_flag == GAMC_CUDA_ERROR) {
    h->err = "internal: a paired AM step (am_mode 3) found a stale candidate proposal (iteration " + std::to_string(st.error_iter) + ")";
    return GAMC_CUDA_ERROR;
  }
  if (st.error_flag == GAMC_NCCL_ERROR) {
    h->err = "peer all-reduce timed out waiting for rank " + std::to_string(st.error_pivot) + " (iteration " + std::to_string(st.error_iter) + ")";
    return GAMC_NCCL_ERROR;
  }
  return GAMC_OK;
}

// `rand(Bernoulli(prob))` decisions of the nine schedules (src/samplers/GAMC.jl:264-337, src/stats/schedule.jl:1-7)
bool schedule_decision(const gamc_config& c, long long i, long long tot, double u) {
  auto par = [&](int k, double dflt) { return std::isnan(c.sched_params[k]) ? dflt : c.sched_params[k]; };
  const double di = (double)i, dt = (double)tot;
  switch (c.schedule) {
    case GAMC_SCHED_AM_ONLY: return false;
    case GAMC_SCHED_SMMALA_ONLY: return true;
    case GAMC_SCHED_MOD: { long long n = (long long)par(0, 10.0); return n > 0 && (i % n) == 0; }
    case GAMC_SCHED_RAND: return u <= par(0, 0.5);
    case GAMC_SCHED_COS: { double a = par(0, 10.0), b = par(1, 0.2), cc = par(2, 0.2); return u <= b * cos(a * di * M_PI / dt) + cc; }
    case GAMC_SCHED_RAND_EXP_DECAY: { double a = par(0, 10.0), b = par(1, 0.0); return u <= (1 - b) * exp(-a * di / dt) + b; }
    case GAMC_SCHED_RAND_LIN_DECAY: { double a = par(0, 1e-2), b = par(1, 0.0); return u <= (1 - b) * (1 / (1 + a * di)) + b; }
    case GAMC_SCHED_RAND_POW_DECAY: { double a = par(0, 1e-3), b = par(1, 0.0); return u <= (1 - b) * pow(a, di / dt) + b; }
    case GAMC_SCHED_RAND_QUAD_DECAY: { double a = par(0, 1e-5), b = par(1, 0.0); return u <= (1 - b) * (1 / (1 + a * (di * di))) + b; }
    default: return true;
  }
}

// the branch of iteration t: the user's schedule function if one is set (GAMC.update!::Function), else the descriptor
bool branch_decision(const gamc_ctx* h, long long t, double u) {
  if (h->sched_fn) return h->sched_fn((int64_t)t, (int64_t)h->cfg.nsteps, u, h->sched_user) != 0;
  return schedule_decision(h->cfg, t, h->cfg.nsteps, u);
}

// Stream-ordered barrier over the ranks of a multi-GPU job (one-element NCCL all-reduce): no rank's later kernels start before
// every rank's earlier work is enqueued and reached, so the in-kernel peer waits only ever cover kernel-time skew.
gamc_status comm_stream_barrier(gamc_ctx* h) {
  if (!h->comm) return GAMC_OK;
  if (!h->d_barrier) { CUDA_TRY(h, cudaMalloc(&h->d_barrier, sizeof(double))); CUDA_TRY(h, cudaMemsetAsync(h->d_barrier, 0, sizeof(double), h->stream)); }
  ProfScope ps(h, GAMC_K_ALLREDUCE);
  const int r = g_nccl.AllReduce(h->d_barrier, h->d_barrier, 1, NCCL_DOUBLE, NCCL_SUM, h->comm, h->stream);
  if (r != 0) { h->err = std::string("ncclAllReduce (barrier): ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"); return GAMC_NCCL_ERROR; }
  return GAMC_OK;
}

// move a landed level-2 evaluation into its slot (mode 0 initial point, 1 current point, 2 proposal)
gamc_status launch_land(gamc_ctx* h, int mode) {
  ProfScope ps(h, GAMC_K_LAND);
  k_land<<<h->p, LAND_THREADS, 0, h->stream>>>(h->B, h->tune, mode, LandPeer{});
  CUDA_TRY(h, cudaGetLastError());
  return GAMC_OK;
}

// Level-2 evaluation at a device-resident theta, landed into its slot.  With peer landing the reduced buffers stay where
// they are and k_land adds them up over NVLink; otherwise NCCL all-reduces R first.
gamc_status eval_and_land(gamc_ctx* h, const double* d_theta, int mode) {
  if (mode == 1 && h->tune.reuse_tensor && !h->comm && !generic_target(h)) {
    // gamc_config.reuse_tensor: every kernel of the re-evaluation returns at once if no AM proposal was accepted since the
    // slot was filled (device flag; the host never synchronises)
    const int* skip = &h->d_state->tensor_valid;
    gamc_status s = launch_level2(h, d_theta, skip);
    if (s != GAMC_OK) return s;
    s = launch_reduce(h, 2, nullptr, skip);
    if (s != GAMC_OK) return s;
    return launch_land(h, mode);
  }
  if (!h->land_p2p) {
    gamc_status s = eval_device(h, d_theta, 2, true);
    if (s != GAMC_OK) return s;
    return launch_land(h, mode);
  }
  gamc_status s = launch_level2(h, d_theta);
  if (s != GAMC_OK) return s;
  const unsigned long long seq = ++h->land_seq;
  const size_t roff = (size_t)(seq & 1ull) * h->xR_stride;
  s = launch_reduce(h, 2, h->xR + roff);
  if (s != GAMC_OK) return s;
  {
    ProfScope ps(h, GAMC_K_ALLREDUCE);
    k_peer_flag<<<1, 32, 0, h->stream>>>(h->d_peers, h->nranks, h->rank, seq);
    CUDA_TRY(h, cudaGetLastError());
  }
  ProfScope ps(h, GAMC_K_LAND);
  LandPeer lp{h->d_peerR, h->d_peers, h->nranks, h->rank, seq, roff, h->peer_timeout_ns};
  k_land<<<h->p, LAND_THREADS, 0, h->stream>>>(h->B, h->tune, mode, lp);
  CUDA_TRY(h, cudaGetLastError());
  return GAMC_OK;
}

template <typename K, typename... Args>
gamc_status launch_step(gamc_ctx* h, int family, K kern, Args... args) {
  const size_t smem = step_smem_bytes(h->p);
  ProfScope ps(h, family);
  kern<<<1, LA_THREADS, smem, h->stream>>>(args...);
  CUDA_TRY(h, cudaGetLastError());
  return GAMC_OK;
}

__global__ void k_state_factor(Bufs B, TuneCfg T);   // defined with the other small kernels of the round-2 entry points below

gamc_status set_step_attrs(gamc_ctx* h) {
  const int smem = (int)step_smem_bytes(512);   // sized for the largest supported p so the attribute is set once
  const void* ks[] = {(const void*)k_sampler_init, (const void*)k_smmala_pre, (const void*)k_smmala_post, (const void*)k_am_pre,
                      (const void*)k_am_post, (const void*)k_state_factor};
  for (const void* k : ks) { gamc_status s = ensure_attr(h, k, smem); if (s != GAMC_OK) return s; }
  {
    // the request depends on p through the number of panel buffers (two while they fit 200 KB: p <= 352), so the limit is the
    // maximum over every supported p, not the value at p = 512
    size_t mx = 0;
    for (int q = 1; q <= 512; ++q) mx = std::max(mx, am_r1_smem_bytes(q, am_r1_nbuf(q)));
    gamc_status s = ensure_attr(h, (const void*)k_am_pre_r1, mx); if (s != GAMC_OK) return s;
  }
  { gamc_status s = ensure_attr(h, (const void*)k_triinv, triinv_smem_bytes(512)); if (s != GAMC_OK) return s; }
  return GAMC_OK;
}

// sstate.oldinvtensor = inv(G) of the current point, into B.C (and W = Lm^-1 into B.W)
gamc_status launch_inverse(gamc_ctx* h, bool write_C = true) {
  const int p = h->p;
  {
    ProfScope ps(h, GAMC_K_HANDOVER);
    const size_t smem = triinv_smem_bytes(p);
    k_triinv<<<(p + 31) / 32, 256, smem, h->stream>>>(h->B.st, h->B.U, h->B.Ustride, h->B.invU, h->B.invUstride, p, h->B.W);
    CUDA_TRY(h, cudaGetLastError());
  }
  if (write_C) {
    ProfScope ps(h, GAMC_K_HANDOVER);
    dim3 grid((p + 15) / 16, (p + 15) / 16);
    k_wtw<<<grid, 256, 0, h->stream>>>(h->B.W, p, h->B.C);
    CUDA_TRY(h, cudaGetLastError());
  }
  return GAMC_OK;
}

// am_mode 1 hand-over: B.Lc = chol(inv(G)) = U^-T of the current point
gamc_status launch_seed_factor(gamc_ctx* h) {
  gamc_status s = launch_inverse(h, false);
  if (s != GAMC_OK) return s;
  ProfScope ps(h, GAMC_K_HANDOVER);
  const int p = h->p;
  k_w_to_l<<<std::min(1024, (p * p + 255) / 256), 256, 0, h->stream>>>(h->B.st, h->B.W, p, h->B.Lc, h->B.rdiagL);
  CUDA_TRY(h, cudaGetLastError());
  return GAMC_OK;
}

gamc_status ensure_tape(gamc_ctx* h, long long n) {
  if (n <= h->tape_cap && h->tape_p == h->p) return GAMC_OK;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  cudaFree(h->d_tape);
  h->d_tape = nullptr; h->tape_cap = 0;
  CUDA_TRY(h, cudaMalloc(&h->d_tape, sizeof(double) * (size_t)n * (3 + h->p)));
  h->tape_cap = n; h->tape_p = h->p;
  return GAMC_OK;
}

void bind_tape(gamc_ctx* h) {
  const long long n = h->tape_cap;
  h->B.tape_usched = h->d_tape;
  h->B.tape_umix = h->d_tape + n;
  h->B.tape_uacc = h->d_tape + 2 * n;
  h->B.tape_z = h->d_tape + 3 * n;
}


// ---- small kernels of the new entry points -----------------------------------------------------------
// gamc_set_state: B.Lc[0] = chol(C) of a host-supplied AM covariance (am_mode >= 1 keeps the covariance in factored form)
__global__ void __launch_bounds__(LA_THREADS, 1) k_state_factor(Bufs B, TuneCfg T) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  StepSmem S = step_smem_carve(smem_raw, T.p);
  const int p = T.p;
  const size_t n = (size_t)p * p;
  for (size_t e = threadIdx.x; e < n; e += blockDim.x) { const int i = (int)(e % p), j = (int)(e / p); B.Lc[e] = (i >= j) ? B.C[e] : 0.0; }
  __syncthreads();
  MatView<false> V{B.Lc, p, p};
  int bad = 0;
  chol_blocked<false>(V, B.rdiagL, nullptr, S.la, &bad);
  if (bad) { raise_error(B.st, 3, bad); return; }
  if (threadIdx.x == 0) B.st->lcur = 0;
}

// gamc_get_monitor / monitored arrays: out = [loglikelihood, logprior | gradloglikelihood(p) | gradlogprior(p)] of the current point
__global__ void __launch_bounds__(256) k_monitor(Bufs B, TuneCfg T, double* out) {
  __shared__ double red[40];
  const int p = T.p;
  const double lp = log_prior(B.theta, p, T.lambda, red);
  const int cur = B.st->cur;
  for (int i = threadIdx.x; i < p; i += blockDim.x) {
    const double glp = T.lambda > 0.0 ? -B.theta[i] / T.lambda : 0.0;
    out[2 + i] = B.gs[(size_t)cur * p + i] - glp;
    out[2 + p + i] = glp;
  }
  if (threadIdx.x == 0) { out[0] = B.st->logtarget - lp; out[1] = lp; }
}

// FP64 yard-sticks (gamc_probe_fp64_peak): register-only loops, 16 independent accumulators per warp
__global__ void __launch_bounds__(256) k_probe_dmma(double* out, int iters, double seed) {
  double c[16][2];
  const double a = seed + threadIdx.x * 1e-9, b = seed - threadIdx.x * 1e-9;
#pragma unroll
  for (int t = 0; t < 16; ++t) { c[t][0] = 0.0; c[t][1] = 0.0; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int t = 0; t < 16; ++t)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[t][0]), "+d"(c[t][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < 16; ++t) s += c[t][0] + c[t][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_probe_dfma(double* out, int iters, double seed) {
  double c[16];
  const double a = seed + threadIdx.x * 1e-9, b = 1.0 - 1e-12;
#pragma unroll
  for (int t = 0; t < 16; ++t) c[t] = t;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int t = 0; t < 16; ++t) c[t] = fma(c[t], b, a);
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < 16; ++t) s += c[t];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

// ====================================================================================================
// C ABI
// ====================================================================================================
extern "C" {

static gamc_status run_monitor(gamc_ctx* h);

const char* gamc_version(void) { return "gamc_b200 0.2.0 sm_100a"; }

gamc_status gamc_create(gamc_handle* out, int32_t device) {
  if (!out) return GAMC_INVALID_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return GAMC_CUDA_ERROR;  // no CPU fallback: fail loudly
  if (device < 0 || device >= ndev) return GAMC_INVALID_ARG;
  gamc_ctx* h = new gamc_ctx();
  h->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete h; return GAMC_CUDA_ERROR; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete h; return GAMC_CUDA_ERROR; }
  if (prop.major != 10) { delete h; return GAMC_CUDA_ERROR; }  // sm_100a code only
  h->num_sms = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return GAMC_CUDA_ERROR; }
  if (cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_side, cudaEventDisableTiming) != cudaSuccess) { delete h; return GAMC_CUDA_ERROR; }
  h->ev0.resize(PROF_POOL); h->ev1.resize(PROF_POOL); h->ev_family.resize(PROF_POOL);
  if (const char* ts = getenv("GAMC_PEER_TIMEOUT_S")) { const double v = atof(ts); if (v > 0.0 && v < 1e6) h->peer_timeout_ns = (unsigned long long)(v * 1e9); }
  *out = h;
  return GAMC_OK;
}

gamc_status gamc_destroy(gamc_handle h) {
  if (!h) return GAMC_INVALID_ARG;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  for (int r = 0; r < h->nranks && r < 32; ++r) if (h->peer_ptr[r] && r != h->rank) cudaIpcCloseMemHandle(h->peer_ptr[r]);
  cudaFree(h->mbox); cudaFree(h->d_peers);
  free_peer_landing(h);
  free_sampler(h);
  free_target(h);
  cudaFree(h->d_barrier); cudaFree(h->d_monitor);
  for (int i = 0; i < GAMC_EXT_COUNT; ++i) if (h->ext_free[i]) h->ext_free[i](h);
  cudaFree(h->d_tape);
  if (h->prof || h->ev_used) for (int i = 0; i < PROF_POOL; ++i) { if (h->ev0[i]) cudaEventDestroy(h->ev0[i]); if (h->ev1[i]) cudaEventDestroy(h->ev1[i]); }
  if (h->own_stream) cudaStreamDestroy(h->stream);
  cudaStreamSynchronize(h->side); cudaStreamDestroy(h->side); cudaEventDestroy(h->ev_main); cudaEventDestroy(h->ev_side);
  delete h;
  return GAMC_OK;
}

const char* gamc_last_error_string(gamc_handle h) { return h ? h->err.c_str() : "null handle"; }

gamc_status gamc_set_stream(gamc_handle h, void* cuda_stream) {
  if (!h) return GAMC_INVALID_ARG;
  cudaSetDevice(h->device);
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  if (h->own_stream) cudaStreamDestroy(h->stream);
  h->stream = (cudaStream_t)cuda_stream;
  h->own_stream = false;
  apply_l2_policy(h);
  return GAMC_OK;
}

gamc_status gamc_sync(gamc_handle h) {
  if (!h) return GAMC_INVALID_ARG;
  cudaSetDevice(h->device);
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  if (h->B.st) return check_device_error(h);
  return GAMC_OK;
}

// ---- targets ----------------------------------------------------------------------------------------
gamc_status gamc_target_logistic(gamc_handle h, const double* X, int64_t ldX, const double* y, int64_t N, int32_t p, double lambda) {
  if (!h) return GAMC_INVALID_ARG;
  REQUIRE(h, X && y && N >= 1 && p >= 1 && ldX >= N && lambda > 0, GAMC_INVALID_ARG, "gamc_target_logistic: bad argument");
  REQUIRE(h, p <= 512, GAMC_UNSUPPORTED_SHAPE, "p must be in 1..512");
  cudaSetDevice(h->device);
  free_sampler(h); free_target(h);
  h->N = N; h->p = p; h->lambda = lambda;
  h->ld = (N + 15) & ~15LL;   // re-pitched: 128-byte aligned columns
  CUDA_TRY(h, cudaMalloc(&h->dX, sizeof(double) * (size_t)h->ld * p));
  CUDA_TRY(h, cudaMalloc(&h->dy, sizeof(double) * (size_t)h->ld));
  h->own_data = true;
  CUDA_TRY(h, cudaMemcpy2DAsync(h->dX, sizeof(double) * h->ld, X, sizeof(double) * ldX, sizeof(double) * N, p, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->dy, y, sizeof(double) * N, cudaMemcpyHostToDevice, h->stream));
  return setup_target(h);
}

gamc_status gamc_target_logistic_device(gamc_handle h, const double* dX, int64_t ldX, const double* dy, int64_t N, int32_t p, double lambda) {
  if (!h) return GAMC_INVALID_ARG;
  REQUIRE(h, dX && dy && N >= 1 && p >= 1 && ldX >= N && lambda > 0, GAMC_INVALID_ARG, "gamc_target_logistic_device: bad argument");
  cudaSetDevice(h->device);
  free_sampler(h); free_target(h);
  h->N = N; h->p = p; h->lambda = lambda; h->ld = ldX;
  h->dX = const_cast<double*>(dX); h->dy = const_cast<double*>(dy); h->own_data = false;
  return setup_target(h);
}

gamc_status gamc_target_logistic_synth(gamc_handle h, uint64_t seed, int64_t row0, int64_t nrows, int32_t p, const double* theta_true, double lambda) {
  if (!h) return GAMC_INVALID_ARG;
  REQUIRE(h, theta_true && nrows >= 1 && row0 >= 0 && p >= 1 && lambda > 0, GAMC_INVALID_ARG, "gamc_target_logistic_synth: bad argument");
  REQUIRE(h, p <= 512, GAMC_UNSUPPORTED_SHAPE, "p must be in 1..512");
  cudaSetDevice(h->device);
  free_sampler(h); free_target(h);
  h->N = nrows; h->p = p; h->lambda = lambda;
  h->ld = (nrows + 15) & ~15LL;
  CUDA_TRY(h, cudaMalloc(&h->dX, sizeof(double) * (size_t)h->ld * p));
  CUDA_TRY(h, cudaMalloc(&h->dy, sizeof(double) * (size_t)h->ld));
  h->own_data = true;
  double* d_tt = nullptr;
  CUDA_TRY(h, cudaMalloc(&d_tt, sizeof(double) * p));
  CUDA_TRY(h, cudaMemcpyAsync(d_tt, theta_true, sizeof(double) * p, cudaMemcpyHostToDevice, h->stream));
  {
    dim3 grid((unsigned)((nrows + 255) / 256), (unsigned)std::min(p, 64));
    k_synth_X<<<grid, 256, 0, h->stream>>>(h->dX, h->ld, row0, nrows, p, seed);
    CUDA_TRY(h, cudaGetLastError());
    k_synth_y<<<(unsigned)((nrows + 255) / 256), 256, 0, h->stream>>>(h->dX, h->ld, row0, nrows, p, d_tt, seed, h->dy);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 2;
  }
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  cudaFree(d_tt);
  return setup_target(h);
}

gamc_status gamc_target_callback(gamc_handle h, int32_t p, gamc_target_fn fn, void* user) {
  if (!h) return GAMC_INVALID_ARG;
  REQUIRE(h, fn && p >= 1, GAMC_INVALID_ARG, "gamc_target_callback: bad argument");
  REQUIRE(h, p <= 512, GAMC_UNSUPPORTED_SHAPE, "p must be in 1..512");
  cudaSetDevice(h->device);
  free_sampler(h); free_target(h);
  h->N = 0; h->p = p; h->lambda = -1.0;   // lambda <= 0: the step kernels add no prior terms (the host's values are complete)
  h->goff = (1 + p + 1) & ~1;
  CUDA_TRY(h, cudaMallocHost(&h->cb_theta, sizeof(double) * p));
  CUDA_TRY(h, cudaMallocHost(&h->cb_R, sizeof(double) * ((size_t)h->goff + (size_t)p * p)));
  memset(h->cb_R, 0, sizeof(double) * ((size_t)h->goff + (size_t)p * p));
  h->cb = fn; h->cb_user = user;
  return alloc_arena(h);
}

gamc_status gamc_target_read_rows(gamc_handle h, int64_t r0, int64_t n, double* X_out, double* y_out) {
  if (!h) return GAMC_INVALID_ARG;
  REQUIRE(h, h->dX, GAMC_NOT_READY, "no target");
  REQUIRE(h, r0 >= 0 && n >= 1 && r0 + n <= h->N, GAMC_INVALID_ARG, "row range out of bounds");
  cudaSetDevice(h->device);
  if (X_out) CUDA_TRY(h, cudaMemcpy2DAsync(X_out, sizeof(double) * n, h->dX + r0, sizeof(double) * h->ld, sizeof(double) * n, h->p, cudaMemcpyDeviceToHost, h->stream));
  if (y_out) CUDA_TRY(h, cudaMemcpyAsync(y_out, h->dy + r0, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return GAMC_OK;
}

// ---- evaluation -------------------------------------------------------------------------------------
static gamc_status eval_host(gamc_handle h, const double* theta, int level, bool reduce_and_prior, double* lt, double* grad, double* G) {
  if (!h) return GAMC_INVALID_ARG;
  REQUIRE(h, h->dX || generic_target(h), GAMC_NOT_READY, "no target: call gamc_target_* first");
  REQUIRE(h, theta, GAMC_INVALID_ARG, "theta is null");
  cudaSetDevice(h->device);
  const int p = h->p;
  CUDA_TRY(h, cudaMemcpyAsync(h->theta_eval, theta, sizeof(double) * p, cudaMemcpyHostToDevice, h->stream));
  gamc_status s = eval_device(h, h->theta_eval, level, reduce_and_prior);
  if (s != GAMC_OK) return s;
  if (reduce_and_prior && !generic_target(h)) {
    ProfScope ps(h, GAMC_K_LINALG);
    k_eval_finish<<<1, 256, 0, h->stream>>>(h->R, h->goff, h->theta_eval, p, h->lambda, level);
    CUDA_TRY(h, cudaGetLastError());
  }
  if (lt) CUDA_TRY(h, cudaMemcpyAsync(lt, h->R, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (grad && level >= 1) CUDA_TRY(h, cudaMemcpyAsync(grad, h->R + 1, sizeof(double) * p, cudaMemcpyDeviceToHost, h->stream));
  if (G && level >= 2) CUDA_TRY(h, cudaMemcpyAsync(G, h->R + h->goff, sizeof(double) * (size_t)p * p, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  if (level >= 2 && h->d_co_dbg) co_debug_report(h);
  return level >= 2 ? check_co_error(h) : GAMC_OK;
}

gamc_status gamc_eval_logtarget(gamc_handle h, const double* theta, double* logtarget) {
  return eval_host(h, theta, 0, true, logtarget, nullptr, nullptr);
}
gamc_status gamc_eval_upto_tensor(gamc_handle h, const double* theta, double* logtarget, double* grad, double* G) {
  return eval_host(h, theta, 2, true, logtarget, grad, G);
}
gamc_status gamc_eval_partials(gamc_handle h, const double* theta, double* ll, double* g_lik, double* G_lik) {
  return eval_host(h, theta, G_lik ? 2 : (g_lik ? 1 : 0), false, ll, g_lik, G_lik);
}

// ---- communicator -----------------------------------------------------------------------------------
gamc_status gamc_comm_unique_id(uint8_t id_out[128]) {
  std::string err;
  if (!id_out || !g_nccl.load(err)) return GAMC_NCCL_ERROR;
  NcclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != 0) return GAMC_NCCL_ERROR;
  memcpy(id_out, id.internal, 128);
  return GAMC_OK;
}

gamc_status gamc_comm_init(gamc_handle h, const uint8_t id[128], int32_t nranks, int32_t rank) {
  if (!h) return GAMC_INVALID_ARG;
  REQUIRE(h, id && nranks >= 1 && rank >= 0 && rank < nranks, GAMC_INVALID_ARG, "gamc_comm_init: bad argument");
  if (!g_nccl.load(h->err)) return GAMC_NCCL_ERROR;
  cudaSetDevice(h->device);
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  if (h->comm) {                     // re-initialisation: drop the previous communicator and its peer mappings
    free_peer_landing(h);
    if (g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    h->comm = nullptr;
  }
  NcclUniqueId uid;
  memcpy(uid.internal, id, 128);
  int r = g_nccl.CommInitRank(&h->comm, nranks, uid, rank);
  if (r != 0) { h->err = std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"); h->comm = nullptr; return GAMC_NCCL_ERROR; }
  h->nranks = nranks; h->rank = rank;
  setup_peer_mailboxes(h);   // optional fast path; falls back to NCCL for the scalar too when it cannot be set up
  return GAMC_OK;
}

gamc_status gamc_comm_peer_path(gamc_handle h, int32_t* active) {
  if (!h || !active) return GAMC_INVALID_ARG;
  *active = (h->p2p_ready ? 1 : 0) | (h->land_p2p ? 2 : 0);   // bit 0: AM-step scalar, bit 1: gradient/metric landing
  return GAMC_OK;
}

// ---- sampler ----------------------------------------------------------------------------------------
gamc_status gamc_sampler_init(gamc_handle h, const gamc_config* cfg, const double* theta0) {
  if (!h) return GAMC_INVALID_ARG;
  REQUIRE(h, h->dX || generic_target(h), GAMC_NOT_READY, "no target: call gamc_target_* first");
  REQUIRE(h, cfg && theta0, GAMC_INVALID_ARG, "null argument");
  // GAMC inner constructor asserts (src/samplers/GAMC.jl:135-138)
  REQUIRE(h, cfg->driftstep > 0, GAMC_INVALID_ARG, "Drift step is not positive");
  REQUIRE(h, cfg->minorscale > 0, GAMC_INVALID_ARG, "Constant minorscale must be positive");
  REQUIRE(h, cfg->c >= 0, GAMC_INVALID_ARG, "Constant c must be non-negative");
  REQUIRE(h, cfg->t0 > 0, GAMC_INVALID_ARG, "t0 is not positive");
  REQUIRE(h, cfg->schedule >= 0 && cfg->schedule <= GAMC_SCHED_RAND_QUAD_DECAY, GAMC_INVALID_ARG, "unknown schedule");
  REQUIRE(h, cfg->period >= 1 && cfg->nsteps >= 1 && cfg->burnin >= 0, GAMC_INVALID_ARG, "bad period / nsteps / burnin");
  REQUIRE(h, cfg->am_mode >= 0 && cfg->am_mode <= 3, GAMC_INVALID_ARG, "am_mode must be 0, 1, 2 or 3");
  REQUIRE(h, cfg->sampler_kind == 0 || cfg->sampler_kind == 1, GAMC_INVALID_ARG, "sampler_kind must be 0 (GAMC) or 1 (MALA)");
  cudaSetDevice(h->device);
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  DEBUG_POINT(h, "sampler_init entry");
  free_sampler(h);
  DEBUG_POINT(h, "sampler_init after free_sampler");
  const int p = h->p;
  h->cfg = *cfg;
  TuneCfg& T = h->tune;
  T = TuneCfg{};
  T.is_accrate = cfg->total_tuner_kind == 1;
  const bool tuner_verbose = cfg->verbose_total || cfg->verbose_smmala || cfg->verbose_am;   // GAMCTuner.jl:22-23
  T.tuning = tuner_verbose || T.is_accrate;
  T.count_total = cfg->verbose_total || T.is_accrate;
  T.verbose_sm = cfg->verbose_smmala; T.verbose_am = cfg->verbose_am;
  T.period = cfg->period; T.burnin = cfg->burnin;
  T.targetrate = cfg->targetrate; T.score_k = cfg->score_k;
  T.minorscale = cfg->minorscale; T.sqrtminorscale = sqrt(cfg->minorscale); T.c = cfg->c;
  T.lambda = h->lambda; T.p = p; T.am_mode = cfg->am_mode;
  T.reuse_tensor = (cfg->reuse_tensor && !h->comm && !generic_target(h)) ? 1 : 0;

  Bufs& B = h->B;
  B = Bufs{};
  B.st = h->d_state;
  const size_t pp = (size_t)p * p;
  CUDA_TRY(h, cudaMemsetAsync(h->mats, 0, sizeof(double) * (pp * 9 + 2 * la_invd_doubles(p)), h->stream));
  B.Gs = h->mats; B.Gstride = pp;
  B.U = h->mats + 2 * pp; B.Ustride = pp;
  B.invU = h->mats + 9 * pp; B.invUstride = la_invd_doubles(p);
  B.C = h->mats + 4 * pp; B.W = h->mats + 5 * pp; B.Lc = h->mats + 6 * pp; B.Lstride = pp;   // Lc: three buffers (am_mode 2 rotates them)
  CUDA_TRY(h, cudaMemsetAsync(h->vecs, 0, sizeof(double) * ((size_t)p * 22 + 16), h->stream));
  double* v = h->vecs;
  B.gs = v; v += 2 * p; B.rdiagU = v; v += 2 * p; B.fterm = v; v += 2 * p;
  B.theta = v; v += p; B.theta_prop = v; v += p; B.mu = v; v += p; B.lastmean = v; v += p; B.secondlastmean = v; v += p; B.rdiagL = v; v += 3 * p; B.accbuf = v; v += 2 * p; B.cand = v; v += 2 * p; B.land = v;
  B.R = h->R; B.goff = h->goff;
  h->chain_cap = std::max<long long>(0, cfg->nsteps - cfg->burnin);
  if (h->chain_cap > 0) {
    CUDA_TRY(h, cudaMalloc(&h->d_chain, sizeof(double) * (size_t)h->chain_cap * p));
    CUDA_TRY(h, cudaMalloc(&h->d_accept, (size_t)h->chain_cap));
    CUDA_TRY(h, cudaMalloc(&h->d_chain_lt, sizeof(double) * (size_t)h->chain_cap));
  }
  B.chain_value = h->d_chain; B.chain_accept = h->d_accept; B.chain_cap = h->chain_cap; B.chain_logtarget = h->d_chain_lt;
  if (T.tuning) {   // one record per tuning event: at most burnin / period + 2 of them (iterate/GAMC.jl:201)
    h->tune_log_cap = (int)std::min<long long>(cfg->burnin / cfg->period + 2, 1 << 20);
    CUDA_TRY(h, cudaMalloc(&h->d_tune_log, sizeof(double) * 16 * (size_t)h->tune_log_cap));
    CUDA_TRY(h, cudaMemsetAsync(h->d_tune_log, 0, sizeof(double) * 16 * (size_t)h->tune_log_cap, h->stream));
  }
  B.tune_log = h->d_tune_log; B.tune_log_cap = h->tune_log_cap;
  if (h->d_tape && h->tape_p == p) bind_tape(h);

  // tuner_state (src/samplers/GAMC.jl:195-200): BasicMCTune(step, 0, 0, period) seeds totproposed with the period
  DevState st{};
  st.step = cfg->driftstep; st.sqrtstep = sqrt(cfg->driftstep);
  st.tot_totprop = cfg->period; st.sm_totprop = cfg->period; st.am_totprop = cfg->period;
  st.tot_rate = st.sm_rate = st.am_rate = NAN; st.smfreq = st.amfreq = NAN; st.ratio = NAN;
  CUDA_TRY(h, cudaMemcpyAsync(B.st, &st, sizeof(DevState), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(B.theta, theta0, sizeof(double) * p, cudaMemcpyHostToDevice, h->stream));
  DEBUG_POINT(h, "sampler_init after allocs");
  gamc_status s = set_step_attrs(h);
  if (s != GAMC_OK) return s;
  DEBUG_POINT(h, "sampler_init after attrs");
  if (cfg->sampler_kind == 1) {
    // MALA: log-target + gradient only
    s = eval_device(h, B.theta, 1, true);
    if (s != GAMC_OK) return s;
    ProfScope ps(h, GAMC_K_LINALG);
    k_mala_init<<<1, 256, 0, h->stream>>>(B, T);
    CUDA_TRY(h, cudaGetLastError());
  } else {
    // initialize! (:157-176): first evaluation of target + gradient + metric at theta0
    setup_peer_landing(h);
    s = comm_stream_barrier(h); if (s != GAMC_OK) return s;
    s = eval_and_land(h, B.theta, 0);
    if (s != GAMC_OK) return s;
    s = launch_step(h, GAMC_K_LINALG, k_sampler_init, B, T);
    if (s != GAMC_OK) return s;
  }
  h->h_count = 0; h->h_present = true; h->h_past = false; h->spec_ready = false;   // MuvGAMCState outer ctor (:117-118)
  h->h_tot_prop = 0; h->h_tot_totprop = cfg->period;
  h->tape_n = 0; h->tape_pos = 0;
  h->sampler_ready = true;
  return check_device_error(h);
}

gamc_status gamc_tape_set(gamc_handle h, int64_t n, const double* u_sched, const double* u_mix, const double* z, const double* u_acc) {
  if (!h) return GAMC_INVALID_ARG;
  REQUIRE(h, h->dX || generic_target(h), GAMC_NOT_READY, "no target");
  REQUIRE(h, n >= 1 && u_sched && u_mix && z && u_acc, GAMC_INVALID_ARG, "gamc_tape_set: bad argument");
  cudaSetDevice(h->device);
  gamc_status s = ensure_tape(h, n);
  if (s != GAMC_OK) return s;
  bind_tape(h);
  const long long cap = h->tape_cap;
  CUDA_TRY(h, cudaMemcpyAsync(h->d_tape, u_sched, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_tape + cap, u_mix, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_tape + 2 * cap, u_acc, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_tape + 3 * cap, z, sizeof(double) * (size_t)n * h->p, cudaMemcpyHostToDevice, h->stream));
  h->h_usched.assign(u_sched, u_sched + n);
  h->tape_n = n; h->tape_pos = 0;
  return GAMC_OK;
}

gamc_status gamc_tape_generate(gamc_handle h, int64_t n, uint64_t seed) {
  if (!h) return GAMC_INVALID_ARG;
  REQUIRE(h, h->dX || generic_target(h), GAMC_NOT_READY, "no target");
  REQUIRE(h, n >= 1, GAMC_INVALID_ARG, "gamc_tape_generate: n < 1");
  cudaSetDevice(h->device);
  gamc_status s = ensure_tape(h, n);
  if (s != GAMC_OK) return s;
  bind_tape(h);
  const long long cap = h->tape_cap;
  const int pairs = (h->p + 1) / 2;
  const long long total = n * (long long)(pairs + 1);
  k_tape_generate<<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>(n, h->p, seed, h->tape_gen_total, h->d_tape, h->d_tape + cap,
                                                                       h->d_tape + 2 * cap, h->d_tape + 3 * cap);
  CUDA_TRY(h, cudaGetLastError());
  h->launches += 1;
  h->tape_gen_total += n;
  h->h_usched.resize(n);
  CUDA_TRY(h, cudaMemcpyAsync(h->h_usched.data(), h->d_tape, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->tape_n = n; h->tape_pos = 0;
  return GAMC_OK;
}

gamc_status gamc_tape_get(gamc_handle h, int64_t first, int64_t n, double* u_sched, double* u_mix, double* z, double* u_acc) {
  if (!h) return GAMC_INVALID_ARG;
  REQUIRE(h, h->d_tape && h->tape_n > 0, GAMC_NOT_READY, "no tape: call gamc_tape_set / gamc_tape_generate");
  REQUIRE(h, first >= 0 && n >= 0 && first + n <= h->tape_n, GAMC_INVALID_ARG, "gamc_tape_get: range out of bounds");
  cudaSetDevice(h->device);
  const long long cap = h->tape_cap;
  if (n == 0) return GAMC_OK;
  if (u_sched) CUDA_TRY(h, cudaMemcpyAsync(u_sched, h->d_tape + first, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  if (u_mix) CUDA_TRY(h, cudaMemcpyAsync(u_mix, h->d_tape + cap + first, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  if (u_acc) CUDA_TRY(h, cudaMemcpyAsync(u_acc, h->d_tape + 2 * cap + first, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  if (z) CUDA_TRY(h, cudaMemcpyAsync(z, h->d_tape + 3 * cap + (size_t)first * h->tape_p, sizeof(double) * (size_t)n * h->tape_p, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return GAMC_OK;
}

gamc_status gamc_peek_kinds(gamc_handle h, int64_t n, uint8_t* kinds_out) {
  if (!h) return GAMC_INVALID_ARG;
  REQUIRE(h, h->sampler_ready, GAMC_NOT_READY, "sampler not initialised");
  REQUIRE(h, kinds_out && n >= 0 && h->tape_pos + n <= h->tape_n, GAMC_INVALID_ARG, "gamc_peek_kinds: tape too short");
  if (h->cfg.sampler_kind == 1) { for (long long k = 0; k < n; ++k) kinds_out[k] = 2; return GAMC_OK; }   // MALA steps
  bool present = h->h_present;
  for (long long k = 0; k < n; ++k) {
    const long long t = h->h_count + k + 1;
    if (t > h->cfg.t0) present = branch_decision(h, t, h->h_usched[h->tape_pos + k]);
    kinds_out[k] = present ? 1 : 0;
  }
  return GAMC_OK;
}

gamc_status gamc_run(gamc_handle h, int64_t n) {
  if (!h) return GAMC_INVALID_ARG;
  REQUIRE(h, h->sampler_ready, GAMC_NOT_READY, "sampler not initialised");
  REQUIRE(h, n >= 0 && h->tape_pos + n <= h->tape_n, GAMC_NOT_READY, "random tape exhausted: call gamc_tape_set / gamc_tape_generate");
  cudaSetDevice(h->device);
  Bufs& B = h->B;
  const TuneCfg& T = h->tune;
  gamc_status s;
  if (h->cfg.sampler_kind == 1) {   // MALA: pre -> fused log-lik + gradient pass -> post
    for (long long k = 0; k < n; ++k) {
      ++h->h_count;
      const long long ti = h->tape_pos++;
      { ProfScope ps(h, GAMC_K_LINALG); k_mala_pre<<<1, 256, 0, h->stream>>>(B, T, ti); CUDA_TRY(h, cudaGetLastError()); }
      s = eval_device(h, B.theta_prop, 1, true); if (s != GAMC_OK) return s;
      { ProfScope ps(h, GAMC_K_LINALG); k_mala_post<<<1, 256, 0, h->stream>>>(B, T, ti); CUDA_TRY(h, cudaGetLastError()); }
    }
    return GAMC_OK;
  }
  if (h->comm && (h->p2p_ready || h->land_p2p) && n > 0) { s = comm_stream_barrier(h); if (s != GAMC_OK) return s; }
  // branch sequence of these n iterations (the schedule coin depends only on (i, nsteps, u_sched))
  std::vector<unsigned char> kinds(n), events(n);
  {
    bool pr = h->h_present;
    long long tp = h->h_tot_prop, ttp = h->h_tot_totprop;
    for (long long k = 0; k < n; ++k) {
      const long long t = h->h_count + k + 1;
      if (t > h->cfg.t0) pr = branch_decision(h, t, h->h_usched[h->tape_pos + k]);
      kinds[k] = pr ? 1 : 0;
      // burn-in tuning events (step_tail: tot_prop counts every transition while the tuner is on and restarts at every event)
      events[k] = 0;
      if (T.tuning) {
        ++tp;
        if (ttp <= T.burnin && (tp % T.period) == 0) { events[k] = 1; ttp += tp; tp = 0; }
      }
    }
    h->h_tot_prop = tp; h->h_tot_totprop = ttp;
  }
  // am_mode 3: two consecutive AM steps share one data pass (needs the partial sums to be added on this rank's device)
  // Pairing puts both factor updates (~2 x 58 us at p = 256) on the critical path and saves one data pass: it pays when the pass over
  // this rank's shard takes longer than ~100 us (0.65 GB); below that (small shards of a many-GPU job) am_mode 3 runs as am_mode 2.
  double pair_min_bytes = 6.5e8;
  if (const char* e = getenv("GAMC_AM_PAIR_MIN_BYTES")) pair_min_bytes = atof(e);   // tests: 0 pairs at any size
  const bool pass_is_long = 8.0 * (double)h->N * (double)h->p >= pair_min_bytes;
  const bool can_pair = T.am_mode == 3 && pass_is_long && h->dp_rpl == 1 && !generic_target(h) && (!h->comm || h->p2p_ready);
  const bool spec_mode = T.am_mode == 2 || (T.am_mode == 3 && !can_pair);
  for (long long k = 0; k < n; ++k) {
    ++h->h_count;                                                                        // iterate/GAMC.jl:2
    const long long ti = h->tape_pos++;
    h->h_present = kinds[k] != 0;                                                        // :8-10
    if (h->h_present) {                                                                  // :12
      h->spec_ready = false;
      const int refactor = h->h_past ? 0 : 1;
      if (refactor) {                                                                    // :20
        s = eval_and_land(h, B.theta, 1); if (s != GAMC_OK) return s;
      }
      s = launch_step(h, GAMC_K_SMMALA_PRE, k_smmala_pre, B, T, ti, refactor); if (s != GAMC_OK) return s;
      s = eval_and_land(h, B.theta_prop, 2); if (s != GAMC_OK) return s;                 // :37
      s = launch_step(h, GAMC_K_SMMALA_POST, k_smmala_post, B, T, ti); if (s != GAMC_OK) return s;
    } else {                                                                             // :128
      const bool next_am = spec_mode && (k + 1 < n) && !kinds[k + 1];                  // next step is AM too: speculate
      const bool pair = can_pair && (k + 1 < n) && !kinds[k + 1] && !events[k];          // ... and share this step's data pass
      if (pair) {
        // ---- am_mode 3: steps k and k + 1 in one data pass.  This step's factor and proposal; both candidate factors and proposals of
        //      the next step (one per outcome of this one); the log-likelihood at all three; the two accept/reject kernels.
        if (!h->spec_ready) {
          if (h->h_past) { s = launch_seed_factor(h); if (s != GAMC_OK) return s; }
          const int nbuf = am_r1_nbuf(h->p);
          ProfScope ps(h, GAMC_K_AM_PRE);
          k_am_pre_r1<<<1, LA_THREADS, am_r1_smem_bytes(h->p, nbuf), h->stream>>>(B, T, ti, nbuf, 0);
          CUDA_TRY(h, cudaGetLastError());
        }
        {
          const int nbuf = am_r1_nbuf(h->p);
          ProfScope ps(h, GAMC_K_AM_PRE);
          k_am_pre_r1<<<2, LA_THREADS, am_r1_smem_bytes(h->p, nbuf), h->stream>>>(B, T, ti + 1, nbuf, 2);
          CUDA_TRY(h, cudaGetLastError());
        }
        s = launch_ll3(h, B.theta_prop, B.cand, B.cand + h->p); if (s != GAMC_OK) return s;    // :148 of both steps
        const size_t smem = step_smem_bytes(h->p);
        for (int half = 0; half < 2; ++half) {
          PeerArgs pa{};
          if (h->comm && h->p2p_ready) { pa.peers = h->d_peers; pa.nranks = h->nranks; pa.rank = h->rank; pa.seq = ++h->p2p_seq; pa.timeout_ns = h->peer_timeout_ns; }
          ProfScope ps(h, GAMC_K_AM_POST);
          if (half == 0)
            k_am_post<<<1, LA_THREADS, smem, h->stream>>>(B, T, ti, 1, h->part_ll3, h->dp_grid, pa, nullptr, 0);
          else
            k_am_post<<<1, LA_THREADS, smem, h->stream>>>(B, T, ti + 1, 0, h->part_ll3 + h->part_ll3_stride, h->dp_grid, pa,
                                                          h->part_ll3 + 2 * (size_t)h->part_ll3_stride, 1);
          CUDA_TRY(h, cudaGetLastError());
        }
        // host bookkeeping of the second step (iterate/GAMC.jl:2,8-10,197)
        ++h->h_count; ++h->tape_pos; ++k;
        h->spec_ready = false;
        h->h_past = false;
        continue;
      }
      if (T.am_mode >= 1) {
        if (!h->spec_ready) {
          if (h->h_past) { s = launch_seed_factor(h); if (s != GAMC_OK) return s; }      // chol(oldinvtensor) = U^-T seeds the update
          const int nbuf = am_r1_nbuf(h->p);
          ProfScope ps(h, GAMC_K_AM_PRE);
          k_am_pre_r1<<<1, LA_THREADS, am_r1_smem_bytes(h->p, nbuf), h->stream>>>(B, T, ti, nbuf, 0);
          CUDA_TRY(h, cudaGetLastError());
        }   // else: the previous k_am_post selected the speculative factor and wrote the proposal
      } else {
        if (h->h_past) { s = launch_inverse(h); if (s != GAMC_OK) return s; }            // oldinvtensor = inv(G) seeds the recursion
        s = launch_step(h, GAMC_K_AM_PRE, k_am_pre, B, T, ti); if (s != GAMC_OK) return s;
      }
      if (next_am) {
        // both outcomes of this step's accept/reject -> next step's factor, on the side stream, while the data pass runs
        CUDA_TRY(h, cudaEventRecord(h->ev_main, h->stream));
        CUDA_TRY(h, cudaStreamWaitEvent(h->side, h->ev_main, 0));
        const int nbuf = am_r1_nbuf(h->p);
        h->launches += 1;
        k_am_pre_r1<<<2, LA_THREADS, am_r1_smem_bytes(h->p, nbuf), h->side>>>(B, T, ti + 1, nbuf, 1);
        CUDA_TRY(h, cudaGetLastError());
        CUDA_TRY(h, cudaEventRecord(h->ev_side, h->side));
        h->dp_grid_now = std::max(1, h->dp_grid - 2);                                    // leave two SMs to the speculative CTAs
      }
      const bool local_sum = (!h->comm || h->p2p_ready) && !generic_target(h);
      const int nparts_post = local_sum ? (h->dp_grid_now > 0 ? h->dp_grid_now : h->dp_grid) : 0;
      PeerArgs pa{};
      if (h->comm && h->p2p_ready) { pa.peers = h->d_peers; pa.nranks = h->nranks; pa.rank = h->rank; pa.seq = ++h->p2p_seq; pa.timeout_ns = h->peer_timeout_ns; }
      s = eval_device(h, B.theta_prop, 0, true, true);                                   // :148
      h->dp_grid_now = 0;
      if (s != GAMC_OK) return s;
      if (next_am) CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_side, 0));
      {
        const size_t smem = step_smem_bytes(h->p);
        ProfScope ps(h, GAMC_K_AM_POST);
        k_am_post<<<1, LA_THREADS, smem, h->stream>>>(B, T, ti, next_am ? 1 : 0, h->part_ll, nparts_post, pa, nullptr, 0);
        CUDA_TRY(h, cudaGetLastError());
      }
      h->spec_ready = next_am;
    }
    h->h_past = h->h_present;                                                            // :197
  }
  return GAMC_OK;
}

gamc_status gamc_chain_get(gamc_handle h, int64_t first, int64_t n, double* value, uint8_t* accept) {
  if (!h) return GAMC_INVALID_ARG;
  REQUIRE(h, h->sampler_ready, GAMC_NOT_READY, "sampler not initialised");
  cudaSetDevice(h->device);
  gamc_status s = check_device_error(h);   // synchronises
  if (s != GAMC_OK) return s;
  REQUIRE(h, first >= 0 && n >= 0 && first + n <= h->chain_cap, GAMC_INVALID_ARG, "chain range out of bounds");
  REQUIRE(h, first + n <= h->chain_saved_seen, GAMC_NOT_READY, "chain range not sampled yet: " + std::to_string(h->chain_saved_seen) +
          " post-burn-in samples saved so far");
  if (n == 0) return GAMC_OK;
  if (value) CUDA_TRY(h, cudaMemcpyAsync(value, h->d_chain + (size_t)first * h->p, sizeof(double) * (size_t)n * h->p, cudaMemcpyDeviceToHost, h->stream));
  if (accept) CUDA_TRY(h, cudaMemcpyAsync(accept, h->d_accept + first, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return GAMC_OK;
}

gamc_status gamc_iterate(gamc_handle h, int64_t n, const double* u_sched, const double* u_mix, const double* z, const double* u_acc,
                         double* value, uint8_t* accept, int64_t* n_saved_out) {
  if (!h) return GAMC_INVALID_ARG;
  REQUIRE(h, h->sampler_ready, GAMC_NOT_READY, "sampler not initialised");
  gamc_status s = gamc_tape_set(h, n, u_sched, u_mix, z, u_acc);
  if (s != GAMC_OK) return s;
  const long long before = std::max<long long>(0, std::min<long long>(h->h_count - h->cfg.burnin, h->chain_cap));
  s = gamc_run(h, n);
  if (s != GAMC_OK) return s;
  const long long after = std::max<long long>(0, std::min<long long>(h->h_count - h->cfg.burnin, h->chain_cap));
  s = gamc_chain_get(h, before, after - before, value, accept);
  if (n_saved_out) *n_saved_out = after - before;
  return s;
}

gamc_status gamc_get_state(gamc_handle h, gamc_state_scalars* out) {
  if (!h || !out) return GAMC_INVALID_ARG;
  REQUIRE(h, h->B.st, GAMC_NOT_READY, "sampler not initialised");
  cudaSetDevice(h->device);
  DevState st;
  CUDA_TRY(h, cudaMemcpyAsync(&st, h->B.st, sizeof(DevState), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  out->count = st.count; out->updatetensorcount = st.updatetensorcount;
  out->step = st.step; out->sqrttunestep = st.sqrtstep;
  out->total_accepted = st.tot_acc; out->total_proposed = st.tot_prop; out->total_totproposed = st.tot_totprop; out->total_rate = st.tot_rate;
  out->smmala_accepted = st.sm_acc; out->smmala_proposed = st.sm_prop; out->smmala_totproposed = st.sm_totprop;
  out->am_accepted = st.am_acc; out->am_proposed = st.am_prop; out->am_totproposed = st.am_totprop;
  out->smmalafrequency = st.smfreq; out->amfrequency = st.amfreq;
  out->logtarget = st.logtarget; out->ratio = st.ratio;
  out->presentupdatetensor = h->h_present; out->pastupdatetensor = h->h_past;
  out->error_flag = st.error_flag; out->error_pivot = st.error_pivot; out->error_iteration = st.error_iter;
  out->chain_saved = st.chain_saved;
  return GAMC_OK;
}

gamc_status gamc_get_array(gamc_handle h, int32_t array_id, double* out) {
  if (!h || !out) return GAMC_INVALID_ARG;
  REQUIRE(h, h->sampler_ready, GAMC_NOT_READY, "sampler not initialised");
  cudaSetDevice(h->device);
  const int p = h->p; const size_t pp = (size_t)p * p;
  Bufs& B = h->B;
  DevState st;
  CUDA_TRY(h, cudaMemcpyAsync(&st, B.st, sizeof(DevState), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  const int cur = st.cur;
  const double* src = nullptr; size_t n = p;
  switch (array_id) {
    case GAMC_ARR_VALUE: src = B.theta; break;
    case GAMC_ARR_GRADLOGTARGET: src = B.gs + (size_t)cur * p; break;
    case GAMC_ARR_TENSORLOGTARGET: src = B.Gs + (size_t)cur * pp; n = pp; break;
    case GAMC_ARR_LASTMEAN: src = B.lastmean; break;
    case GAMC_ARR_SECONDLASTMEAN: src = B.secondlastmean; break;
    case GAMC_ARR_OLDFIRSTTERM: src = B.fterm + (size_t)cur * p; break;
    case GAMC_ARR_PROPOSED_VALUE: src = B.theta_prop; break;
    case GAMC_ARR_OLDINVTENSOR:
      if (h->h_past || h->h_count == 0) { gamc_status s = launch_inverse(h); if (s != GAMC_OK) return s; }
      else if (h->tune.am_mode >= 1) {   // AM covariance is kept in factored form: C = L L'
        k_llt<<<std::min(1024, (p * p + 255) / 256), 256, 0, h->stream>>>(B.Lc + (size_t)st.lcur * B.Lstride, p, B.C);
        CUDA_TRY(h, cudaGetLastError());
      }
      src = B.C; n = pp; break;
    case GAMC_ARR_CHOLINVTENSOR: {
      // L = U^-T: L(a,b) = W(p-1-b, p-1-a) with W = Lm^-1 (M space); data movement only on the host
      gamc_status s = launch_inverse(h, false); if (s != GAMC_OK) return s;
      std::vector<double> W(pp);
      CUDA_TRY(h, cudaMemcpyAsync(W.data(), B.W, sizeof(double) * pp, cudaMemcpyDeviceToHost, h->stream));
      CUDA_TRY(h, cudaStreamSynchronize(h->stream));
      for (int b = 0; b < p; ++b) for (int a = 0; a < p; ++a) out[a + (size_t)b * p] = (a >= b) ? W[(size_t)(p - 1 - b) + (size_t)(p - 1 - a) * p] : 0.0;
      return GAMC_OK;
    }
    case GAMC_ARR_GRADLOGLIKELIHOOD: case GAMC_ARR_GRADLOGPRIOR: {
      REQUIRE(h, h->lambda > 0, GAMC_NOT_READY, "a generic target does not separate likelihood and prior");
      gamc_status s = run_monitor(h); if (s != GAMC_OK) return s;
      src = h->d_monitor + 2 + (array_id == GAMC_ARR_GRADLOGPRIOR ? p : 0); break;
    }
    case GAMC_ARR_TENSORLOGLIKELIHOOD: case GAMC_ARR_TENSORLOGPRIOR: {
      // tensor = X' W X + I/lambda (examples/logistic_regression/src/GAMC/simulations.jl:34-37): the prior part is I/lambda
      REQUIRE(h, h->lambda > 0, GAMC_NOT_READY, "a generic target does not separate likelihood and prior");
      std::vector<double> G(pp);
      CUDA_TRY(h, cudaMemcpyAsync(G.data(), B.Gs + (size_t)cur * pp, sizeof(double) * pp, cudaMemcpyDeviceToHost, h->stream));
      CUDA_TRY(h, cudaStreamSynchronize(h->stream));
      const bool prior = array_id == GAMC_ARR_TENSORLOGPRIOR;
      for (int j = 0; j < p; ++j) for (int i = 0; i < p; ++i) {
        const double pr = (i == j) ? 1.0 / h->lambda : 0.0;
        out[i + (size_t)j * p] = prior ? pr : G[i + (size_t)j * p] - pr;
      }
      return GAMC_OK;
    }
    default: h->err = "unknown array id"; return GAMC_INVALID_ARG;
  }
  CUDA_TRY(h, cudaMemcpyAsync(out, src, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return GAMC_OK;
}

// ---- measurement ------------------------------------------------------------------------------------

// ---- round-2 additions: device-side generic target, user schedule, restart, monitors, burn-in report, probes -------
gamc_status gamc_target_device_callback(gamc_handle h, int32_t p, gamc_target_enqueue_fn enqueue, void* user) {
  if (!h) return GAMC_INVALID_ARG;
  REQUIRE(h, enqueue && p >= 1, GAMC_INVALID_ARG, "gamc_target_device_callback: bad argument");
  REQUIRE(h, p <= 512, GAMC_UNSUPPORTED_SHAPE, "p must be in 1..512");
  cudaSetDevice(h->device);
  free_sampler(h); free_target(h);
  h->N = 0; h->p = p; h->lambda = -1.0;   // lambda <= 0: the step kernels add no prior terms (the user's values are complete)
  h->goff = (1 + p + 1) & ~1;
  h->dcb = enqueue; h->dcb_user = user;
  return alloc_arena(h);
}

gamc_status gamc_set_metric_route(gamc_handle h, int32_t route, int32_t slices, int32_t smax) {
  if (!h) return GAMC_INVALID_ARG;
  REQUIRE(h, route >= GAMC_METRIC_AUTO && route <= GAMC_METRIC_I8, GAMC_INVALID_ARG, "metric route must be GAMC_METRIC_AUTO, _F64 or _I8");
  REQUIRE(h, slices == 0 || (slices >= 2 && slices <= I8_MAXK), GAMC_INVALID_ARG, "slices must be 0 (default) or 2..7");
  REQUIRE(h, smax == 0 || (smax >= 2 && smax <= 2 * (slices ? slices : 6)), GAMC_INVALID_ARG, "smax must be 0 (default) or 2..2*slices");
  h->metric_route = route; h->route_slices = slices; h->route_smax = smax;
  return GAMC_OK;
}

gamc_status gamc_get_metric_route(gamc_handle h, int32_t* route, int32_t* slices, int32_t* smax) {
  if (!h) return GAMC_INVALID_ARG;
  if (route) *route = h->metric_i8 ? GAMC_METRIC_I8 : GAMC_METRIC_F64;
  if (slices) *slices = h->metric_i8 ? h->i8_k : 0;
  if (smax) *smax = h->metric_i8 ? h->i8_smax : 0;
  return GAMC_OK;
}

gamc_status gamc_comm_set_timeout(gamc_handle h, double seconds) {
  if (!h) return GAMC_INVALID_ARG;
  REQUIRE(h, seconds > 0 && seconds < 1e6, GAMC_INVALID_ARG, "timeout must be positive");
  h->peer_timeout_ns = (unsigned long long)(seconds * 1e9);
  return GAMC_OK;
}

gamc_status gamc_set_schedule_callback(gamc_handle h, gamc_schedule_fn fn, void* user) {
  if (!h) return GAMC_INVALID_ARG;
  h->sched_fn = fn; h->sched_user = user;
  return GAMC_OK;
}

gamc_status gamc_set_state(gamc_handle h, const gamc_state_scalars* st_in, const double* theta, const double* lastmean,
                           const double* secondlastmean, const double* oldinvtensor) {
  if (!h) return GAMC_INVALID_ARG;
  REQUIRE(h, h->sampler_ready, GAMC_NOT_READY, "sampler not initialised: call gamc_sampler_init first");
  REQUIRE(h, h->cfg.sampler_kind == 0, GAMC_INVALID_ARG, "gamc_set_state restarts GAMC chains");
  REQUIRE(h, st_in && theta && lastmean && secondlastmean, GAMC_INVALID_ARG, "gamc_set_state: null argument");
  REQUIRE(h, st_in->count >= 0 && st_in->step > 0, GAMC_INVALID_ARG, "gamc_set_state: bad count / step");
  REQUIRE(h, st_in->pastupdatetensor || st_in->count == 0 || oldinvtensor, GAMC_INVALID_ARG,
          "gamc_set_state: the last transition was an AM step, so sstate.oldinvtensor (the AM covariance) is required");
  cudaSetDevice(h->device);
  Bufs& B = h->B;
  const TuneCfg& T = h->tune;
  const int p = h->p;
  // value, then initialize! + sampler_state at it: log-target, gradient, metric, factor, G^-1 grad of slot 0
  CUDA_TRY(h, cudaMemcpyAsync(B.theta, theta, sizeof(double) * p, cudaMemcpyHostToDevice, h->stream));
  gamc_status s = comm_stream_barrier(h); if (s != GAMC_OK) return s;
  s = eval_and_land(h, B.theta, 0); if (s != GAMC_OK) return s;
  s = launch_step(h, GAMC_K_LINALG, k_sampler_init, B, T); if (s != GAMC_OK) return s;
  s = check_device_error(h); if (s != GAMC_OK) return s;
  CUDA_TRY(h, cudaMemcpyAsync(B.lastmean, lastmean, sizeof(double) * p, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(B.secondlastmean, secondlastmean, sizeof(double) * p, cudaMemcpyHostToDevice, h->stream));
  DevState st;
  CUDA_TRY(h, cudaMemcpyAsync(&st, B.st, sizeof(DevState), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  st.count = st_in->count; st.updatetensorcount = st_in->updatetensorcount;
  st.step = st_in->step; st.sqrtstep = sqrt(st_in->step);
  st.tot_acc = st_in->total_accepted; st.tot_prop = st_in->total_proposed; st.tot_totprop = st_in->total_totproposed; st.tot_rate = st_in->total_rate;
  st.sm_acc = st_in->smmala_accepted; st.sm_prop = st_in->smmala_proposed; st.sm_totprop = st_in->smmala_totproposed;
  st.am_acc = st_in->am_accepted; st.am_prop = st_in->am_proposed; st.am_totprop = st_in->am_totproposed;
  st.smfreq = st_in->smmalafrequency; st.amfreq = st_in->amfrequency; st.ratio = st_in->ratio;
  st.chain_saved = 0; st.tune_events = 0; st.lcur = 0;
  CUDA_TRY(h, cudaMemcpyAsync(B.st, &st, sizeof(DevState), cudaMemcpyHostToDevice, h->stream));
  h->h_count = st_in->count;
  h->h_tot_prop = st_in->total_proposed; h->h_tot_totprop = st_in->total_totproposed;
  h->h_present = st_in->count == 0 ? true : st_in->presentupdatetensor != 0;
  h->h_past = st_in->count == 0 ? false : st_in->pastupdatetensor != 0;
  h->spec_ready = false; h->tape_n = 0; h->tape_pos = 0;
  if (!h->h_past && st_in->count > 0) {
    CUDA_TRY(h, cudaMemcpyAsync(B.C, oldinvtensor, sizeof(double) * (size_t)p * p, cudaMemcpyHostToDevice, h->stream));
    if (T.am_mode >= 1) { s = launch_step(h, GAMC_K_LINALG, k_state_factor, B, T); if (s != GAMC_OK) return s; }
  }
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return check_device_error(h);
}

static gamc_status run_monitor(gamc_ctx* h) {
  if (!h->d_monitor) CUDA_TRY(h, cudaMalloc(&h->d_monitor, sizeof(double) * (2 + 2 * 512)));
  ProfScope ps(h, GAMC_K_LINALG);
  k_monitor<<<1, 256, 0, h->stream>>>(h->B, h->tune, h->d_monitor);
  CUDA_TRY(h, cudaGetLastError());
  return GAMC_OK;
}

gamc_status gamc_get_monitor(gamc_handle h, double* loglikelihood, double* logprior) {
  if (!h) return GAMC_INVALID_ARG;
  REQUIRE(h, h->sampler_ready, GAMC_NOT_READY, "sampler not initialised");
  REQUIRE(h, h->lambda > 0, GAMC_NOT_READY, "a generic target does not separate likelihood and prior");
  cudaSetDevice(h->device);
  gamc_status s = run_monitor(h); if (s != GAMC_OK) return s;
  double v[2];
  CUDA_TRY(h, cudaMemcpyAsync(v, h->d_monitor, sizeof(v), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  if (loglikelihood) *loglikelihood = v[0];
  if (logprior) *logprior = v[1];
  return GAMC_OK;
}

gamc_status gamc_chain_get_logtarget(gamc_handle h, int64_t first, int64_t n, double* logtarget) {
  if (!h) return GAMC_INVALID_ARG;
  REQUIRE(h, h->sampler_ready, GAMC_NOT_READY, "sampler not initialised");
  cudaSetDevice(h->device);
  gamc_status s = check_device_error(h);   // synchronises
  if (s != GAMC_OK) return s;
  REQUIRE(h, first >= 0 && n >= 0 && first + n <= h->chain_cap, GAMC_INVALID_ARG, "chain range out of bounds");
  REQUIRE(h, first + n <= h->chain_saved_seen, GAMC_NOT_READY, "chain range not sampled yet");
  if (n == 0 || !logtarget) return GAMC_OK;
  CUDA_TRY(h, cudaMemcpyAsync(logtarget, h->d_chain_lt + first, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return GAMC_OK;
}

gamc_status gamc_tune_log_get(gamc_handle h, double* out, int64_t max_records, int64_t* n_records) {
  if (!h) return GAMC_INVALID_ARG;
  REQUIRE(h, h->sampler_ready, GAMC_NOT_READY, "sampler not initialised");
  cudaSetDevice(h->device);
  DevState st;
  CUDA_TRY(h, cudaMemcpyAsync(&st, h->B.st, sizeof(DevState), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  const long long have = std::min<long long>(st.tune_events, h->tune_log_cap);
  if (n_records) *n_records = have;
  const long long take = std::min<long long>(have, std::max<long long>(0, max_records));
  if (out && take > 0) {
    CUDA_TRY(h, cudaMemcpyAsync(out, h->d_tune_log, sizeof(double) * 16 * (size_t)take, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  }
  return GAMC_OK;
}

gamc_status gamc_probe_fp64_peak(gamc_handle h, double* dmma_tflops, double* dfma_tflops) {
  if (!h) return GAMC_INVALID_ARG;
  cudaSetDevice(h->device);
  const int grid = h->num_sms * 4, iters = 4000;      // 32 warps per SM: the occupancy at which both loops saturate the FP64 pipe
  double* out = nullptr;
  CUDA_TRY(h, cudaMalloc(&out, sizeof(double) * (size_t)grid * 256));
  cudaEvent_t e0, e1;
  CUDA_TRY(h, cudaEventCreate(&e0)); CUDA_TRY(h, cudaEventCreate(&e1));
  double best[2] = {0.0, 0.0};
  for (int which = 0; which < 2; ++which) {
    for (int rep = 0; rep < 4; ++rep) {                // first repetition is the warm-up
      cudaEventRecord(e0, h->stream);
      if (which == 0) k_probe_dmma<<<grid, 256, 0, h->stream>>>(out, iters, 1.0);
      else k_probe_dfma<<<grid, 256, 0, h->stream>>>(out, iters, 1.0);
      cudaEventRecord(e1, h->stream);
      h->launches += 1;
      if (cudaEventSynchronize(e1) != cudaSuccess) break;
      float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
      const double flops = which == 0 ? (double)grid * 8 * iters * 16.0 * 512.0 : (double)grid * 256 * iters * 16.0 * 2.0;
      if (rep > 0 && ms > 0.f) best[which] = std::max(best[which], flops / ms * 1e-9);
    }
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  CUDA_TRY(h, cudaGetLastError());
  if (dmma_tflops) *dmma_tflops = best[0];
  if (dfma_tflops) *dfma_tflops = best[1];
  return GAMC_OK;
}

gamc_status gamc_profile_enable(gamc_handle h, int32_t on) {
  if (!h) return GAMC_INVALID_ARG;
  cudaSetDevice(h->device);
  if (on && !h->ev0[0]) {
    for (int i = 0; i < PROF_POOL; ++i) { CUDA_TRY(h, cudaEventCreate(&h->ev0[i])); CUDA_TRY(h, cudaEventCreate(&h->ev1[i])); }
  }
  h->prof = on != 0;
  return GAMC_OK;
}

gamc_status gamc_profile_reset(gamc_handle h) {
  if (!h) return GAMC_INVALID_ARG;
  cudaSetDevice(h->device);
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->ev_used = 0;
  return GAMC_OK;
}

gamc_status gamc_profile_get(gamc_handle h, int32_t family, int64_t* launches, double* total_ms) {
  if (!h) return GAMC_INVALID_ARG;
  REQUIRE(h, family >= 0 && family < GAMC_K_COUNT, GAMC_INVALID_ARG, "bad kernel family");
  cudaSetDevice(h->device);
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  long long cnt = 0; double ms = 0.0;
  for (int i = 0; i < h->ev_used; ++i) {
    if (h->ev_family[i] != family) continue;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, h->ev0[i], h->ev1[i]) == cudaSuccess) { ms += t; cnt += 1; }
  }
  if (launches) *launches = cnt;
  if (total_ms) *total_ms = ms;
  return GAMC_OK;
}

gamc_status gamc_launch_count(gamc_handle h, int64_t* out) {
  if (!h || !out) return GAMC_INVALID_ARG;
  *out = h->launches;
  return GAMC_OK;
}

#ifdef GAMC_GUARD
}  // extern "C"
#undef cudaMalloc
#undef cudaFree
namespace {
constexpr size_t GUARD_BYTES = 4096;
constexpr unsigned char GUARD_PATTERN = 0xA5;
struct GuardRec { unsigned char* base; size_t bytes; };
std::vector<GuardRec> g_guard;
__global__ void k_guard_scan(const unsigned char* z, size_t n, unsigned long long* bad) {
  unsigned long long c = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) c += z[i] != GUARD_PATTERN;
  if (c) atomicAdd(bad, c);
}
}  // namespace
cudaError_t gamc_guard_malloc(void** p, size_t bytes) {
  unsigned char* base = nullptr;
  const size_t padded = (bytes + 255) & ~(size_t)255;
  cudaError_t e = cudaMalloc(&base, padded + 2 * GUARD_BYTES);
  if (e != cudaSuccess) return e;
  cudaMemset(base, GUARD_PATTERN, GUARD_BYTES);
  cudaMemset(base + GUARD_BYTES + bytes, GUARD_PATTERN, padded - bytes + GUARD_BYTES);
  g_guard.push_back(GuardRec{base, bytes});
  *p = base + GUARD_BYTES;
  return cudaSuccess;
}
cudaError_t gamc_guard_free(void* p) {
  if (!p) return cudaSuccess;
  unsigned char* base = static_cast<unsigned char*>(p) - GUARD_BYTES;
  for (size_t i = 0; i < g_guard.size(); ++i)
    if (g_guard[i].base == base) { g_guard.erase(g_guard.begin() + i); return cudaFree(base); }
  return cudaFree(p);     // not one of ours (e.g. an IPC mapping)
}
extern "C" {
// Number of guard bytes overwritten next to any live allocation (0 = no out-of-bounds write), and the number of live allocations.
long long gamc_debug_guard_check(long long* n_allocations) {
  cudaDeviceSynchronize();
  unsigned long long* d_bad = nullptr;
  cudaMalloc(&d_bad, sizeof(unsigned long long));
  cudaMemset(d_bad, 0, sizeof(unsigned long long));
  for (const GuardRec& r : g_guard) {
    const size_t padded = (r.bytes + 255) & ~(size_t)255;
    k_guard_scan<<<8, 256>>>(r.base, GUARD_BYTES, d_bad);
    k_guard_scan<<<8, 256>>>(r.base + GUARD_BYTES + r.bytes, padded - r.bytes + GUARD_BYTES, d_bad);
  }
  unsigned long long bad = 0;
  cudaMemcpy(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost);
  cudaFree(d_bad);
  if (n_allocations) *n_allocations = (long long)g_guard.size();
  return (long long)bad;
}
#endif

// Debug read-out of the integer-tensor-core metric (tests/test_gpu_i8.py): what = 0: digit planes [chunk][k][p][I8_KC] int8 (as written by the last
// level-2 evaluation), 1: column scales cg[p] (double), 2: {k, smax, ldn, grid} as four int64.  Not part of the public header.
int gamc_debug_i8_read(gamc_handle h, int what, void* out, long long bytes) {
  if (!h || !h->metric_i8) return 1;
  cudaStreamSynchronize(h->stream);
  if (what == 0) {
    const long long need = (long long)h->i8_k * h->p * h->i8_ldn;
    if (bytes < need) return 2;
    return cudaMemcpy(out, h->d_i8S, (size_t)need, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 3;
  }
  if (what == 1) {
    if (bytes < (long long)sizeof(double) * h->p) return 2;
    return cudaMemcpy(out, h->d_i8cg, sizeof(double) * h->p, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 3;
  }
  if (what == 2) {
    if (bytes < 32) return 2;
    long long* o = reinterpret_cast<long long*>(out);
    o[0] = h->i8_k; o[1] = h->i8_smax; o[2] = h->i8_ldn; o[3] = I8_KC;
    return 0;
  }
  return 4;
}

#ifdef GAMC_PHASE_CLOCKS
// Debug build only (make dbg): accumulated clock64 deltas / hit counts of the PHASE() marks; reset = 1 clears them.
int gamc_debug_phases(long long* acc, long long* cnt, int reset) {
  cudaDeviceSynchronize();
  if (acc) cudaMemcpyFromSymbol(acc, gamc::g_phase_acc, sizeof(long long) * 64);
  if (cnt) cudaMemcpyFromSymbol(cnt, gamc::g_phase_cnt, sizeof(long long) * 64);
  if (reset) {
    long long z[64] = {0};
    cudaMemcpyToSymbol(gamc::g_phase_acc, z, sizeof(z));
    cudaMemcpyToSymbol(gamc::g_phase_cnt, z, sizeof(z));
  }
  return 0;
}
#endif

}  // extern "C"
