// vicon2gt_b200: batch handle, host staging, device memory, the C ABI of include/vicon2gt_b200.h.
//
// Host-side counterpart of the reference's ViconGraphSolver / Propagator / Interpolator stores:
//   * v2gt_batch_set_trial   = Propagator::feed_imu x M, Interpolator::feed_pose x V (std::set
//                              semantics: sorted by time, duplicate keys dropped, first wins),
//                              timestamp_cameras
//   * v2gt_batch_upload      = timestamp cleaning of ViconGraphSolver::build_and_solve
//                              (.cpp:114-159; has_bounding_imu is t_imu[0] <= t <= t_imu[M-1] for the
//                              time-ordered store) + one H2D copy per array from pinned memory
//   * build / optimize       = kernel launches on the handle's stream; nothing here computes
// There is no CPU fallback: every entry point that needs the GPU fails with V2GT_ERR_NO_DEVICE /
// V2GT_ERR_CUDA when CUDA is unavailable.
#include "v2gt_common.cuh"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <fstream>
#include <iomanip>
#include <limits>
#include <map>
#include <mutex>
#include <numeric>
#include <sstream>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <vector>

namespace v2 {
int launch_init_states(const DevPtrs &d, int n_max, int init_states, cudaStream_t st);
int launch_preintegrate(const DevPtrs &d, int n_max, int keep_cov, cudaStream_t st);
int launch_linearize(const DevPtrs &d, int n_max, int force_all, cudaStream_t st);
int launch_cost(const DevPtrs &d, int n_max, int which, int force_all, cudaStream_t st);
int launch_solve(const DevPtrs &d, int force_all, cudaStream_t st);
int solve_configure();
int elim_warps_per_cta();
int l2_segments(int P, int l2_max);
int launch_error_stats(const DevPtrs &d, const double *d_truth, double *d_out, cudaStream_t st);
int elim_ctas_per_sm();
int launch_eliminate(const DevPtrs &d, int force_all, cudaStream_t st);
int launch_reduced(const DevPtrs &d, int force_all, cudaStream_t st);
int launch_backsub(const DevPtrs &d, int force_all, cudaStream_t st);
int launch_lm_begin(const DevPtrs &d, cudaStream_t st);
int launch_lm_freeze(const DevPtrs &d, int pass, cudaStream_t st);
int launch_lm_pre_try(const DevPtrs &d, cudaStream_t st);
int launch_lm_decide(const DevPtrs &d, int count_active, cudaStream_t st);
int launch_chain_sum(const DevPtrs &d, int with_halo, cudaStream_t st);
int launch_chain_reduce(const DevPtrs &d, int with_halo, cudaStream_t st);
int launch_chain_unpack(const DevPtrs &d, cudaStream_t st);
int launch_chain_pack(const DevPtrs &d, int stride, cudaStream_t st);
int asm_chunks(int n_max);
int simbatch_view(v2gt_simbatch *s, int trial, int *device, int64_t *n_imu, const double **imu_t, const double **d_imu_s, int64_t *n_vic,
                  const double **vic_t, const double **d_vic_s, int64_t *n_times, const double **state_t);
} // namespace v2

using namespace v2;

// One trial's staged inputs: raw caller-layout arrays, either owned copies (v2gt_batch_set_trial) or borrowed views
// (v2gt_batch_set_trial_view).  The conversion to the device layout happens in v2gt_batch_upload, straight into the
// pinned staging buffers, in parallel over the trials.
struct HostTrial {
  v2gt_params prm;
  std::vector<double> own; // owned copies of the arrays below (empty for a borrowed view)
  const double *imu_t = nullptr, *imu_w = nullptr, *imu_a = nullptr;
  const double *vic_t_in = nullptr, *vic_q = nullptr, *vic_p = nullptr, *vic_Rq = nullptr, *vic_Rp = nullptr;
  const double *state_t_raw = nullptr;
  // trial staged from a simulated batch (v2gt_batch_set_trial_sim): measurement rows already on this device, in
  // the device layout; only the stamps above are host arrays (borrowed from the simulator handle)
  const double *dev_imu_s = nullptr, *dev_vic_s = nullptr;
  int64_t n_imu = 0, n_vic_in = 0, n_times = 0;
  double vic_Rconst[18];
  bool cov_const = true;
  bool imu_unsorted = false; // derived in upload: IMU fed out of time order
  // derived in upload
  std::vector<int64_t> vic_order;  // empty: the fed poses are already strictly increasing in time (kept as they are)
  int64_t n_vic = 0;               // after de-duplication
  std::vector<double> state_t;     // after cleaning
  double vic_tmin = std::numeric_limits<double>::infinity(), vic_tmax = -std::numeric_limits<double>::infinity();
  v2gt_trial_info info{};
  bool set = false;
};

struct v2gt_batch {
  int device = 0;
  int sm_count = 148;
  int T = 0;
  cudaStream_t stream = nullptr;
  std::vector<HostTrial> trials;
  std::vector<std::vector<double>> truth; // per trial: [n][10] truth states (v2gt_batch_set_truth)
  DevPtrs d{};
  std::vector<void *> dev_allocs;
  std::vector<void *> pinned_allocs;
  // allocation reuse: a re-upload of a batch of the same shape hands the existing buffers out again, in order
  std::vector<size_t> dev_sizes, pin_sizes;
  std::vector<long long> alloc_key;
  size_t dev_cur = 0, pin_cur = 0;
  bool reuse = false;
  int n_max = 0; // max raw states over trials
  bool uploaded = false, built = false, values_init = false;
  // chain-sharded mode (v2gt_chain_configure)
  bool chain = false;
  int c_rank = 0, c_world = 1, c_p = 0, c_j0 = 0, c_j1 = 0, c_lo = 0, c_hi = 0;
  long long c_n = 0, c_goff = 0;
  void *c_comm = nullptr; // ncclComm_t of the in-library exchange (v2gt_chain_comm_init)
  // one damped solve captured as a CUDA graph, keyed by the segment count of the launch (v2gt_chain_optimize; small
  // batches in v2gt_batch_optimize): valid for the device buffers of the current upload
  std::map<int, cudaGraphExec_t> try_graphs;
  int opt_graph = -1; // -1 auto (small batches), 0 never, 1 always
  // options
  int opt_segments = 0;       // 0 = auto (adaptive); > 0: fixed number of segments per trial
  int opt_target_warps = 148 * 64; // segment warps the adaptive choice aims for per elimination launch
  int opt_keep_cov = 0;
  int opt_tries_per_sync = 8;
  // pinned readback
  int *h_active = nullptr;
  // timing
  cudaEvent_t ev[16];
  std::vector<cudaEvent_t> ev_pool;
  int64_t tries_enqueued = 0;
  bool ev_ok = false;
  double ms[8] = {0};
  int64_t launches[8] = {0};
  std::vector<LmState> h_lm;
  std::vector<int> h_nstates, h_removed;
  std::vector<long long> h_st_off; // first state slot of every trial (host copy of TrialDev::st_off)
  double *h_stage = nullptr;       // the largest pinned staging buffer of the upload (free after it): bounce buffer of
  size_t h_stage_doubles = 0;      // v2gt_batch_get_all_states
  double *h_rb = nullptr;          // own pinned buffer when the staging buffer is too small
  size_t h_rb_doubles = 0;
  bool results_cached = false;
};

namespace {

template <typename Tp> int dev_alloc(v2gt_batch *b, Tp **p, size_t count) {
  *p = nullptr;
  if (count == 0)
    count = 1;
  if (b->reuse) {
    if (b->dev_cur >= b->dev_allocs.size() || b->dev_sizes[b->dev_cur] != count * sizeof(Tp))
      return V2GT_ERR_CUDA;
    *p = (Tp *)b->dev_allocs[b->dev_cur++];
    return V2GT_OK;
  }
  void *q = nullptr;
  cudaError_t e = cudaMalloc(&q, count * sizeof(Tp));
  if (e != cudaSuccess) {
    fprintf(stderr, "[vicon2gt_b200] cudaMalloc of %zu bytes failed: %s\n", count * sizeof(Tp), cudaGetErrorString(e));
    return V2GT_ERR_CUDA;
  }
  b->dev_allocs.push_back(q);
  b->dev_sizes.push_back(count * sizeof(Tp));
  *p = (Tp *)q;
  return V2GT_OK;
}
template <typename Tp> int pin_alloc(v2gt_batch *b, Tp **p, size_t count) {
  if (count == 0)
    count = 1;
  if (b->reuse) {
    if (b->pin_cur >= b->pinned_allocs.size() || b->pin_sizes[b->pin_cur] != count * sizeof(Tp))
      return V2GT_ERR_CUDA;
    *p = (Tp *)b->pinned_allocs[b->pin_cur++];
    return V2GT_OK;
  }
  void *q = nullptr;
  if (cudaMallocHost(&q, count * sizeof(Tp)) != cudaSuccess)
    return V2GT_ERR_CUDA;
  b->pinned_allocs.push_back(q);
  b->pin_sizes.push_back(count * sizeof(Tp));
  *p = (Tp *)q;
  return V2GT_OK;
}
void drop_graphs(v2gt_batch *b) {
  for (auto &kv : b->try_graphs)
    cudaGraphExecDestroy(kv.second);
  b->try_graphs.clear();
}
void free_device(v2gt_batch *b) {
  drop_graphs(b);
  for (void *p : b->dev_allocs)
    cudaFree(p);
  b->dev_allocs.clear();
  for (void *p : b->pinned_allocs)
    cudaFreeHost(p);
  b->pinned_allocs.clear();
  b->dev_sizes.clear();
  b->pin_sizes.clear();
  b->alloc_key.clear();
  b->reuse = false;
  b->h_active = nullptr;
  b->h_stage = nullptr;
  b->h_stage_doubles = 0;
  if (b->h_rb)
    cudaFreeHost(b->h_rb);
  b->h_rb = nullptr;
  b->h_rb_doubles = 0;
  b->uploaded = false;
  b->built = false;
}

#define V2_TRY(expr)                                                                                                             \
  do {                                                                                                                           \
    int _s = (expr);                                                                                                             \
    if (_s != V2GT_OK)                                                                                                           \
      return _s;                                                                                                                 \
  } while (0)

int fetch_results(v2gt_batch *b) {
  if (b->results_cached)
    return V2GT_OK;
  V2_CUDA_CHECK(cudaStreamSynchronize(b->stream));
  b->h_lm.resize(b->T);
  b->h_nstates.resize(b->T);
  b->h_removed.resize(b->T);
  V2_CUDA_CHECK(cudaMemcpy(b->h_lm.data(), b->d.lm, sizeof(LmState) * b->T, cudaMemcpyDeviceToHost));
  V2_CUDA_CHECK(cudaMemcpy(b->h_nstates.data(), b->d.n_states, sizeof(int) * b->T, cudaMemcpyDeviceToHost));
  V2_CUDA_CHECK(cudaMemcpy(b->h_removed.data(), b->d.removed_no_vicon, sizeof(int) * b->T, cudaMemcpyDeviceToHost));
  b->results_cached = true;
  return V2GT_OK;
}

// Eigen-style dense print with the stream's precision and aligned columns (info file)
std::string eigen_print(const double *m, int R, int C, int precision = 6) {
  std::vector<std::string> cells(R * C);
  size_t width = 0;
  for (int i = 0; i < R * C; i++) {
    std::ostringstream o;
    o << std::setprecision(precision) << m[i];
    cells[i] = o.str();
    width = std::max(width, cells[i].size());
  }
  std::ostringstream out;
  for (int r = 0; r < R; r++) {
    if (r)
      out << "\n";
    for (int c = 0; c < C; c++) {
      if (c)
        out << " ";
      out << std::setw((int)width) << cells[r * C + c];
    }
  }
  return out.str();
}

void make_dirs_for(const std::string &path) {
  for (size_t i = 1; i < path.size(); i++)
    if (path[i] == '/') {
      std::string sub = path.substr(0, i);
      mkdir(sub.c_str(), 0755);
    }
}

// run f(t) for t in [0, T) on a few host threads (static interleaved partition)
template <typename F> void parallel_trials(int t0, int t1, F f) {
  const int nthr = std::max(1, std::min({t1 - t0, (int)std::thread::hardware_concurrency(), 16}));
  if (nthr <= 1) {
    for (int t = t0; t < t1; t++)
      f(t);
    return;
  }
  std::vector<std::thread> th;
  for (int k = 0; k < nthr; k++)
    th.emplace_back([=]() {
      for (int t = t0 + k; t < t1; t += nthr)
        f(t);
    });
  for (auto &x : th)
    x.join();
}
template <typename F> void parallel_trials(int T, F f) { parallel_trials(0, T, f); }

// ---- NCCL, loaded at run time (libnccl.so.2: the copy already mapped into the process, e.g. PyTorch's, else the
// system's).  Only the five entry points the chain exchange needs; declarations follow nccl.h 2.x.
struct NcclApi {
  typedef struct { char internal[V2GT_NCCL_ID_BYTES]; } UniqueId;
  int (*GetUniqueId)(UniqueId *) = nullptr;
  int (*CommInitRank)(void **, int, UniqueId, int) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool ok = false;
};
NcclApi &nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h)
      h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h)
      return;
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather;
  });
  return api;
}
constexpr int NCCL_FLOAT64 = 8; // ncclDouble
int nccl_check(int rc, const char *what) {
  if (rc == 0)
    return V2GT_OK;
  NcclApi &n = nccl_api();
  fprintf(stderr, "[vicon2gt_b200] %s failed: %s\n", what, n.GetErrorString ? n.GetErrorString(rc) : "NCCL error");
  return V2GT_ERR_NCCL;
}


} // namespace

namespace v2 {
// accessors used by eval.cu (the handle layout is private to this file)
const DevPtrs *batch_devptrs(v2gt_batch *b) { return &b->d; }
cudaStream_t batch_stream(v2gt_batch *b) { return b->stream; }
int batch_device(v2gt_batch *b) { return b->device; }
int batch_nmax(v2gt_batch *b) { return b->n_max; }
bool batch_built(v2gt_batch *b) { return b->built; }
bool batch_uploaded(v2gt_batch *b) { return b->uploaded; }
void batch_invalidate(v2gt_batch *b) { b->results_cached = false; }
} // namespace v2

extern "C" {

const char *v2gt_version(void) { return "vicon2gt_b200 0.1 (sm_100a, fp64)"; }

void v2gt_params_default(v2gt_params *p) {
  std::memset(p, 0, sizeof(*p));
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  std::memcpy(p->R_GtoV, I, sizeof(I));
  std::memcpy(p->R_BtoI, I, sizeof(I));
  p->gravity_magnitude = 9.81;
  p->estimate_toff_vicon_to_imu = 1;
  p->estimate_ori_vicon_to_imu = 1;
  p->estimate_pos_vicon_to_imu = 1;
  p->num_loop_relin = 0;
  p->sigma_w = 1.6968e-04;
  p->sigma_wb = 1.9393e-05;
  p->sigma_a = 2.0000e-3;
  p->sigma_ab = 3.0000e-03;
  p->lm_max_iterations = 30;
  p->lm_max_tries = 0;
  p->lm_lambda_initial = 1e-5;
  p->lm_lambda_factor = 10.0;
  p->lm_lambda_upper_bound = 1e20;
  p->lm_lambda_lower_bound = 0.0;
  p->lm_min_model_fidelity = 1e-3;
  p->lm_relative_error_tol = 1e-30;
  p->lm_absolute_error_tol = 1e-30;
  p->huber_k = 1.345;
  p->time_offset_window = 0.25;
  p->interp_gap_init = 5.0;
  p->interp_gap_factor = 0.1;
}

int v2gt_abi_struct_sizes(int32_t *params_bytes, int32_t *trial_info_bytes) {
  if (params_bytes)
    *params_bytes = (int32_t)sizeof(v2gt_params);
  if (trial_info_bytes)
    *trial_info_bytes = (int32_t)sizeof(v2gt_trial_info);
  return V2GT_OK;
}

int v2gt_device_count(int32_t *count) {
  if (!count)
    return V2GT_ERR_INVALID_ARG;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *count = 0;
    cudaGetLastError();
    return V2GT_ERR_NO_DEVICE;
  }
  *count = n;
  return V2GT_OK;
}

int v2gt_batch_create(int32_t device, int32_t n_trials, v2gt_batch **out) {
  if (!out || n_trials <= 0)
    return V2GT_ERR_INVALID_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    fprintf(stderr, "[vicon2gt_b200] no CUDA device: this library has no CPU path\n");
    return V2GT_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n)
    return V2GT_ERR_INVALID_ARG;
  V2_CUDA_CHECK(cudaSetDevice(device));
  v2gt_batch *b = new v2gt_batch();
  b->device = device;
  cudaDeviceGetAttribute(&b->sm_count, cudaDevAttrMultiProcessorCount, device);
  b->T = n_trials;
  b->trials.resize(n_trials);
  bool ok = cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) == cudaSuccess;
  int n_ev = 0;
  for (; ok && n_ev < 16; n_ev++)
    ok = cudaEventCreate(&b->ev[n_ev]) == cudaSuccess;
  if (!ok) { // nothing half-built is handed out or left behind
    for (int i = 0; i + 1 < n_ev; i++)
      cudaEventDestroy(b->ev[i]);
    if (b->stream)
      cudaStreamDestroy(b->stream);
    delete b;
    cudaGetLastError();
    return V2GT_ERR_CUDA;
  }
  b->ev_ok = true;
  *out = b;
  return V2GT_OK;
}

int v2gt_batch_destroy(v2gt_batch *b) {
  if (!b)
    return V2GT_OK;
  cudaSetDevice(b->device);
  if (b->stream)
    cudaStreamSynchronize(b->stream);
  free_device(b);
  if (b->c_comm && nccl_api().ok)
    nccl_api().CommDestroy(b->c_comm);
  b->c_comm = nullptr;
  if (b->ev_ok)
    for (int i = 0; i < 16; i++)
      cudaEventDestroy(b->ev[i]);
  for (cudaEvent_t e : b->ev_pool)
    cudaEventDestroy(e);
  if (b->stream)
    cudaStreamDestroy(b->stream);
  delete b;
  return V2GT_OK;
}

int v2gt_batch_set_option(v2gt_batch *b, const char *name, double value) {
  if (!b || !name)
    return V2GT_ERR_INVALID_ARG;
  const std::string n(name);
  if (n == "segments")
    b->opt_segments = (int)value;
  else if (n == "target_warps")
    b->opt_target_warps = std::max(1, (int)value);
  else if (n == "keep_cov")
    b->opt_keep_cov = (int)value;
  else if (n == "tries_per_sync")
    b->opt_tries_per_sync = std::max(1, (int)value);
  else if (n == "graph")
    b->opt_graph = (int)value;
  else
    return V2GT_ERR_INVALID_ARG;
  return V2GT_OK;
}

static int set_trial_impl(v2gt_batch *b, int32_t trial, const v2gt_params *params, int64_t n_imu, const double *imu_t,
                          const double *imu_w, const double *imu_a, int64_t n_vicon, const double *vic_t, const double *vic_q,
                          const double *vic_p, const double *vic_Rq, const double *vic_Rp, const double *vic_R_const,
                          int64_t n_times, const double *state_t, bool copy) {
  if (!b || trial < 0 || trial >= b->T || !params || n_imu < 0 || n_vicon < 0 || n_times < 0)
    return V2GT_ERR_INVALID_ARG;
  if ((n_imu && (!imu_t || !imu_w || !imu_a)) || (n_vicon && (!vic_t || !vic_q || !vic_p)) || (n_times && !state_t))
    return V2GT_ERR_INVALID_ARG;
  if (n_vicon && !(vic_Rq && vic_Rp) && !vic_R_const)
    return V2GT_ERR_INVALID_ARG;
  HostTrial &h = b->trials[trial];
  h = HostTrial();
  h.prm = *params;
  h.n_imu = n_imu;
  h.n_vic_in = n_vicon;
  h.n_times = n_times;
  h.cov_const = !(vic_Rq && vic_Rp);
  if (h.cov_const && vic_R_const)
    std::memcpy(h.vic_Rconst, vic_R_const, sizeof(h.vic_Rconst));
  else
    std::memset(h.vic_Rconst, 0, sizeof(h.vic_Rconst));
  if (copy) {
    const size_t total = (size_t)n_imu * 7 + (size_t)n_vicon * (8 + (h.cov_const ? 0 : 18)) + (size_t)n_times;
    h.own.resize(std::max<size_t>(1, total));
    double *o = h.own.data();
    auto take = [&](const double *src, size_t n) {
      const double *r = o;
      if (n)
        std::memcpy(o, src, sizeof(double) * n);
      o += n;
      return r;
    };
    h.imu_t = take(imu_t, (size_t)n_imu);
    h.imu_w = take(imu_w, (size_t)n_imu * 3);
    h.imu_a = take(imu_a, (size_t)n_imu * 3);
    h.vic_t_in = take(vic_t, (size_t)n_vicon);
    h.vic_q = take(vic_q, (size_t)n_vicon * 4);
    h.vic_p = take(vic_p, (size_t)n_vicon * 3);
    if (!h.cov_const) {
      h.vic_Rq = take(vic_Rq, (size_t)n_vicon * 9);
      h.vic_Rp = take(vic_Rp, (size_t)n_vicon * 9);
    }
    h.state_t_raw = take(state_t, (size_t)n_times);
  } else {
    h.imu_t = imu_t; h.imu_w = imu_w; h.imu_a = imu_a;
    h.vic_t_in = vic_t; h.vic_q = vic_q; h.vic_p = vic_p;
    h.vic_Rq = h.cov_const ? nullptr : vic_Rq;
    h.vic_Rp = h.cov_const ? nullptr : vic_Rp;
    h.state_t_raw = state_t;
  }
  h.set = true;
  b->uploaded = false;
  b->built = false;
  b->results_cached = false;
  return V2GT_OK;
}

int v2gt_batch_set_trial(v2gt_batch *b, int32_t trial, const v2gt_params *params, int64_t n_imu, const double *imu_t,
                         const double *imu_w, const double *imu_a, int64_t n_vicon, const double *vic_t, const double *vic_q,
                         const double *vic_p, const double *vic_Rq, const double *vic_Rp, const double *vic_R_const,
                         int64_t n_times, const double *state_t) {
  return set_trial_impl(b, trial, params, n_imu, imu_t, imu_w, imu_a, n_vicon, vic_t, vic_q, vic_p, vic_Rq, vic_Rp, vic_R_const,
                        n_times, state_t, true);
}

int v2gt_batch_set_trial_view(v2gt_batch *b, int32_t trial, const v2gt_params *params, int64_t n_imu, const double *imu_t,
                              const double *imu_w, const double *imu_a, int64_t n_vicon, const double *vic_t, const double *vic_q,
                              const double *vic_p, const double *vic_Rq, const double *vic_Rp, const double *vic_R_const,
                              int64_t n_times, const double *state_t) {
  return set_trial_impl(b, trial, params, n_imu, imu_t, imu_w, imu_a, n_vicon, vic_t, vic_q, vic_p, vic_Rq, vic_Rp, vic_R_const,
                        n_times, state_t, false);
}

int v2gt_batch_set_trial_sim(v2gt_batch *b, int32_t trial, const v2gt_params *params, v2gt_simbatch *sb, int32_t sim_trial,
                             const double *vic_R_const) {
  if (!b || trial < 0 || trial >= b->T || !params || !sb || !vic_R_const)
    return V2GT_ERR_INVALID_ARG;
  int dev = -1;
  int64_t n_imu = 0, n_vic = 0, n_times = 0;
  const double *imu_t, *d_imu_s, *vic_t, *d_vic_s, *state_t;
  const int rc = v2::simbatch_view(sb, sim_trial, &dev, &n_imu, &imu_t, &d_imu_s, &n_vic, &vic_t, &d_vic_s, &n_times, &state_t);
  if (rc != V2GT_OK)
    return rc;
  if (dev != b->device)
    return V2GT_ERR_INVALID_ARG;
  HostTrial &h = b->trials[trial];
  h = HostTrial();
  h.prm = *params;
  h.n_imu = n_imu;
  h.n_vic_in = n_vic;
  h.n_times = n_times;
  h.cov_const = true;
  std::memcpy(h.vic_Rconst, vic_R_const, sizeof(h.vic_Rconst));
  h.imu_t = imu_t;
  h.vic_t_in = vic_t;
  h.state_t_raw = state_t;
  h.dev_imu_s = d_imu_s;
  h.dev_vic_s = d_vic_s;
  h.set = true;
  b->uploaded = false;
  b->built = false;
  b->results_cached = false;
  return V2GT_OK;
}

int v2gt_batch_upload(v2gt_batch *b) {
  if (!b)
    return V2GT_ERR_INVALID_ARG;
  V2_CUDA_CHECK(cudaSetDevice(b->device));
  V2_CUDA_CHECK(cudaStreamSynchronize(b->stream));
  b->results_cached = false;
  b->uploaded = false;
  b->built = false;
  const int T = b->T;
  // ---- per-trial guards and timestamp cleaning (ViconGraphSolver.cpp:99-150)
  std::vector<TrialDev> td(T);
  std::vector<LmState> lm(T);
  long long S = 0, M = 0, V = 0, C = 0;
  b->n_max = 0;
  for (int t = 0; t < T; t++)
    if (!b->trials[t].set)
      return V2GT_ERR_INVALID_ARG;
  // pass 1 (parallel over trials): vicon order / de-duplication, timestamp cleaning
  parallel_trials(T, [&](int t) {
    HostTrial &h = b->trials[t];
    h.info = v2gt_trial_info{};
    std::memset(&lm[t], 0, sizeof(LmState));
    h.state_t.clear();
    // ---- vicon: std::set<POSEDATA> keyed by timestamp -> stable sort + keep the first of equal keys
    h.vic_order.clear();
    h.vic_tmin = std::numeric_limits<double>::infinity();
    h.vic_tmax = -std::numeric_limits<double>::infinity();
    const int64_t nv = h.n_vic_in;
    bool increasing = true;
    for (int64_t i = 0; i + 1 < nv; i++)
      if (!(h.vic_t_in[i] < h.vic_t_in[i + 1])) {
        increasing = false;
        break;
      }
    if (increasing) {
      h.n_vic = nv;
      if (nv) {
        h.vic_tmin = h.vic_t_in[0];
        h.vic_tmax = h.vic_t_in[nv - 1];
      }
    } else if (h.dev_vic_s) {
      h.info.status = V2GT_ERR_INVALID_ARG; // simulated stamps are strictly increasing by construction
      lm[t].status = V2GT_ERR_INVALID_ARG;
      return;
    } else {
      std::vector<int64_t> order((size_t)nv);
      std::iota(order.begin(), order.end(), (int64_t)0);
      std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) { return h.vic_t_in[x] < h.vic_t_in[y]; });
      h.vic_order.reserve((size_t)nv);
      for (int64_t k = 0; k < nv; k++) {
        const int64_t i = order[(size_t)k];
        // time_min / time_max are updated for every fed pose, duplicates included (Interpolator.cpp:35-37)
        h.vic_tmin = std::min(h.vic_tmin, h.vic_t_in[i]);
        h.vic_tmax = std::max(h.vic_tmax, h.vic_t_in[i]);
        if (!h.vic_order.empty() && h.vic_t_in[h.vic_order.back()] == h.vic_t_in[i])
          continue;
        h.vic_order.push_back(i);
      }
      h.n_vic = (int64_t)h.vic_order.size();
    }
    int status = V2GT_OK;
    if (h.n_times == 0)
      status = V2GT_ERR_NO_TIMESTAMPS;
    else if (h.n_vic < 2)
      status = V2GT_ERR_NO_VICON;
    else {
      const double first = h.vic_tmin + h.prm.toff_imu_to_vicon + 2 * h.prm.time_offset_window;
      const double last = h.vic_tmax + h.prm.toff_imu_to_vicon - 2 * h.prm.time_offset_window;
      // has_bounding_imu (Propagator.cpp:165-193): some sample <= t and some sample >= t seen before the
      // scan breaks; for the time-ordered store this is t_imu[0] <= t <= t_imu[M-1].  The general
      // (unordered) case is reproduced with running prefix extrema.
      const size_t Mi = (size_t)h.n_imu;
      bool imu_sorted = true;
      for (size_t i = 0; i + 1 < Mi; i++)
        if (h.imu_t[i + 1] < h.imu_t[i]) {
          imu_sorted = false;
          break;
        }
      h.imu_unsorted = !imu_sorted;
      h.state_t.reserve((size_t)h.n_times);
      for (int64_t si = 0; si < h.n_times; si++) {
        const double ts = h.state_t_raw[si];
        bool bounded = false;
        if (Mi > 0) {
          if (imu_sorted) {
            bounded = (h.imu_t[0] <= ts) && (h.imu_t[Mi - 1] >= ts);
          } else {
            bool lo = false, hi = false;
            for (size_t i = 0; i < Mi; i++) {
              if (h.imu_t[i] <= ts)
                lo = true;
              if (h.imu_t[i] >= ts)
                hi = true;
              if (lo && hi)
                break;
            }
            bounded = lo && hi;
          }
        }
        if (!bounded)
          h.info.removed_no_imu++;
        else if (ts < first)
          h.info.removed_before++;
        else if (ts > last)
          h.info.removed_after++;
        else
          h.state_t.push_back(ts);
      }
      if (h.state_t.empty())
        status = V2GT_ERR_ALL_OUT_OF_RANGE;
    }
    h.info.status = status;
    lm[t].status = status;
  });
  for (int t = 0; t < T; t++) {
    HostTrial &h = b->trials[t];
    const int status = h.info.status;
    TrialDev &d = td[t];
    std::memset(&d, 0, sizeof(d));
    d.imu_off = M;
    d.n_imu = (long long)h.n_imu;
    d.vic_off = V;
    d.n_vic = (long long)h.n_vic;
    d.st_off = S;
    d.n_raw = (status == V2GT_OK) ? (long long)h.state_t.size() : 0;
    d.vic_cov_const = h.cov_const ? 1 : 0;
    d.imu_unsorted = h.imu_unsorted ? 1 : 0;
    std::memcpy(d.vic_Rconst, h.vic_Rconst, sizeof(d.vic_Rconst));
    d.prm = h.prm;
    M += d.n_imu;
    V += d.n_vic;
    S += d.n_raw;
    if (!h.cov_const)
      C += d.n_vic;
    b->n_max = std::max<int>(b->n_max, (int)d.n_raw);
  }
  b->h_st_off.resize((size_t)T);
  for (int t = 0; t < T; t++)
    b->h_st_off[(size_t)t] = td[t].st_off;
  // ---- segments per trial.  The buffers are sized for the most segments a launch may use (~sqrt(n) per chain);
  // every round of the optimiser then picks its own count from the number of trials still running, so that the
  // chains of the remaining trials are cut finer as the batch drains (v2gt_batch_optimize)
  int P = b->opt_segments;
  if (P <= 0) // up to 4 sqrt(n): the separator chain is itself eliminated in segments (nested level 2), so a small
              // batch can cut its chains much finer than sqrt(n) before the reduced solve dominates
    P = std::max(1, 4 * (int)std::ceil(std::sqrt((double)std::max(1, b->n_max))));
  if (b->opt_segments <= 0)
    P = std::min(P, std::max(1, b->n_max / 32)); // segments shorter than ~32 states only add separator work
  P = std::max(1, std::min(P, std::max(1, b->n_max / 2)));
  if (b->chain) {
    if (T != 1)
      return V2GT_ERR_INVALID_ARG;
    P = b->c_p; // global segment count of the sharded chain
  }
  drop_graphs(b);
  DevPtrs &d = b->d;
  std::memset(&d, 0, sizeof(d));
  d.T = T;
  d.S = S;
  d.M = M;
  d.V = V;
  d.P = P;
  d.nchunk = std::max(1, (b->n_max + V2GT_CHUNK - 1) / V2GT_CHUNK);
  const int nasm = std::max(1, asm_chunks(b->n_max));
  {
    const std::vector<long long> key = {T, S, M, V, C, P, d.nchunk, nasm, b->opt_keep_cov, b->chain ? b->c_world : 0};
    if (key == b->alloc_key && !b->dev_allocs.empty()) {
      b->reuse = true; // same shape as the previous upload: keep every buffer
      b->dev_cur = b->pin_cur = 0;
    } else {
      free_device(b);
      b->alloc_key = key;
    }
  }

  // ---- pinned staging + device arrays
  double *h_imu_t, *h_imu_s, *h_vic_t, *h_vic_s, *h_vic_c, *h_st_t;
  int *h_trial_of, *h_nst;
  long long *h_coff;
  V2_TRY(pin_alloc(b, &h_imu_t, (size_t)M));
  V2_TRY(pin_alloc(b, &h_imu_s, (size_t)M * 8));
  b->h_stage = h_imu_s;
  b->h_stage_doubles = (size_t)M * 8;
  V2_TRY(pin_alloc(b, &h_vic_t, (size_t)V));
  V2_TRY(pin_alloc(b, &h_vic_s, (size_t)V * 8));
  V2_TRY(pin_alloc(b, &h_vic_c, (size_t)C * 18));
  V2_TRY(pin_alloc(b, &h_st_t, (size_t)S));
  V2_TRY(pin_alloc(b, &h_trial_of, (size_t)S));
  V2_TRY(pin_alloc(b, &h_nst, (size_t)T));
  V2_TRY(pin_alloc(b, &h_coff, (size_t)T));
  V2_TRY(pin_alloc(b, &b->h_active, 1));
  {
    long long coff = 0;
    for (int t = 0; t < T; t++) {
      h_coff[t] = b->trials[t].cov_const ? -1 : coff * 18;
      if (!b->trials[t].cov_const)
        coff += td[t].n_vic;
    }
  }
  // pass 2 (parallel over trials, run below in chunks that overlap with the H2D copies of the chunk before): caller
  // layout -> device layout, written straight into the pinned buffers
  auto fill_trial = [&](int t) {
    const HostTrial &h = b->trials[t];
    const TrialDev &q = td[t];
    if (q.n_imu) {
      // IMU: kept in feed order (Propagator::feed_imu appends); one 64-byte row (t, w, a, pad) per sample
      std::memcpy(h_imu_t + q.imu_off, h.imu_t, sizeof(double) * q.n_imu);
      double *o = h_imu_s + 8 * q.imu_off;
      for (long long i = 0; i < (h.dev_imu_s ? 0 : q.n_imu); i++, o += 8) {
        o[0] = h.imu_t[i];
        o[1] = h.imu_w[3 * i];
        o[2] = h.imu_w[3 * i + 1];
        o[3] = h.imu_w[3 * i + 2];
        o[4] = h.imu_a[3 * i];
        o[5] = h.imu_a[3 * i + 1];
        o[6] = h.imu_a[3 * i + 2];
        o[7] = 0.0;
      }
    }
    const bool reorder = !h.vic_order.empty();
    double *vt = h_vic_t + q.vic_off, *vs = h_vic_s + 8 * q.vic_off;
    double *vc = h.cov_const ? nullptr : h_vic_c + h_coff[t];
    for (long long k = 0; k < q.n_vic; k++) {
      const int64_t i = reorder ? h.vic_order[(size_t)k] : k;
      vt[k] = h.vic_t_in[i];
      if (h.dev_vic_s)
        continue;
      double *o = vs + 8 * k;
      o[0] = h.vic_q[4 * i];
      o[1] = h.vic_q[4 * i + 1];
      o[2] = h.vic_q[4 * i + 2];
      o[3] = h.vic_q[4 * i + 3];
      o[4] = h.vic_p[3 * i];
      o[5] = h.vic_p[3 * i + 1];
      o[6] = h.vic_p[3 * i + 2];
      o[7] = 0.0;
      if (vc) {
        std::memcpy(vc + 18 * k, h.vic_Rq + 9 * i, sizeof(double) * 9);
        std::memcpy(vc + 18 * k + 9, h.vic_Rp + 9 * i, sizeof(double) * 9);
      }
    }
    for (long long i = 0; i < q.n_raw; i++) {
      h_st_t[q.st_off + i] = h.state_t[(size_t)i];
      h_trial_of[q.st_off + i] = t;
    }
    h_nst[t] = (int)q.n_raw;
  };
  double *p_imu_t, *p_imu_s, *p_vic_t, *p_vic_s, *p_vic_c;
  long long *p_coff;
  TrialDev *p_trials;
  V2_TRY(dev_alloc(b, &p_imu_t, (size_t)M));
  V2_TRY(dev_alloc(b, &p_imu_s, (size_t)M * 8));
  V2_TRY(dev_alloc(b, &p_vic_t, (size_t)V));
  V2_TRY(dev_alloc(b, &p_vic_s, (size_t)V * 8));
  V2_TRY(dev_alloc(b, &p_vic_c, (size_t)C * 18));
  V2_TRY(dev_alloc(b, &p_coff, (size_t)T));
  V2_TRY(dev_alloc(b, &p_trials, (size_t)T));
  V2_TRY(dev_alloc(b, &d.n_states, (size_t)T));
  V2_TRY(dev_alloc(b, &d.trial_of, (size_t)S));
  V2_TRY(dev_alloc(b, &d.st_t, (size_t)S));
  V2_TRY(dev_alloc(b, &d.st_t_tmp, (size_t)S));
  V2_TRY(dev_alloc(b, &d.x, (size_t)2 * S * X_STRIDE));
  V2_TRY(dev_alloc(b, &d.glob, (size_t)2 * T * GLOB_STRIDE));
  V2_TRY(dev_alloc(b, &d.valid, (size_t)S));
  V2_TRY(dev_alloc(b, &d.pre, (size_t)soa_alloc(S, PRE_STRIDE)));
  V2_TRY(dev_alloc(b, &d.info, (size_t)soa_alloc(S, INFO_STRIDE)));
  if (b->opt_keep_cov)
    V2_TRY(dev_alloc(b, &d.cov, (size_t)S * 225));
  V2_TRY(dev_alloc(b, &d.fpi, (size_t)soa_alloc(S, FPI_STRIDE)));
  V2_TRY(dev_alloc(b, &d.fpv, (size_t)soa_alloc(S, FPV_STRIDE)));
  V2_TRY(dev_alloc(b, &d.ne, (size_t)S * NE_STRIDE));
  V2_TRY(dev_alloc(b, &d.grad, (size_t)S * 16));
  V2_TRY(dev_alloc(b, &d.fac, (size_t)S * FAC_STRIDE));
  V2_TRY(dev_alloc(b, &d.seg, (size_t)T * P * SEG_STRIDE));
  V2_TRY(dev_alloc(b, &d.red_ne, ((size_t)T * (P + 1) + 2) * NE_STRIDE));
  V2_TRY(dev_alloc(b, &d.red_fac, ((size_t)T * (P + 1) + 2) * FAC_STRIDE));
  {
    // level 2 of the nested elimination (separator chain of level 1 cut into ~sqrt(P) segments)
    const int P2 = std::max(2, (int)std::ceil(std::sqrt((double)P)) + 1);
    d.l2_P = P2;
    V2_TRY(dev_alloc(b, &d.l2_trials, (size_t)T));
    V2_TRY(dev_alloc(b, &d.l2_nstates, (size_t)T));
    V2_TRY(dev_alloc(b, &d.l2_seg, (size_t)T * P2 * SEG_STRIDE));
    V2_TRY(dev_alloc(b, &d.l2_red_ne, ((size_t)T * (P2 + 1) + 2) * NE_STRIDE));
    V2_TRY(dev_alloc(b, &d.l2_red_fac, ((size_t)T * (P2 + 1) + 2) * FAC_STRIDE));
    V2_TRY(dev_alloc(b, &d.l2_xsep, (size_t)T * (P2 + 1) * 16));
    V2_TRY(dev_alloc(b, &d.l2_gamma_sub, (size_t)T * 96));
  }
  V2_TRY(dev_alloc(b, &d.red_fb, (size_t)T * 640));
  V2_TRY(dev_alloc(b, &d.xsep, (size_t)T * (P + 1) * 16));
  V2_TRY(dev_alloc(b, &d.delta, (size_t)S * 16));
  V2_TRY(dev_alloc(b, &d.part_asm, (size_t)T * nasm * PART_STRIDE));
  V2_TRY(dev_alloc(b, &d.part_cost, (size_t)T * d.nchunk * 4));
  V2_TRY(dev_alloc(b, &d.gamma, (size_t)T * 96));
  V2_TRY(dev_alloc(b, &d.xg, (size_t)T * 16));
  V2_TRY(dev_alloc(b, &d.lm, (size_t)T));
  V2_TRY(dev_alloc(b, &d.trace, (size_t)T * TRACE_MAX * 5));
  V2_TRY(dev_alloc(b, &d.n_active, 1));
  V2_TRY(dev_alloc(b, &d.removed_no_vicon, (size_t)T));
  if (b->chain) {
    const int ploc = std::max(1, b->c_j1 - b->c_j0);
    V2_TRY(dev_alloc(b, &d.c_sepne, ((size_t)P + 2) * NE_STRIDE));
    V2_TRY(dev_alloc(b, &d.c_sums, 8));
    V2_TRY(dev_alloc(b, &d.c_gamma_loc, 96));
    const size_t per_rank = (size_t)ploc * (SEG_STRIDE + NE_STRIDE) + 96;
    V2_TRY(dev_alloc(b, &d.c_pack_send, per_rank));
    V2_TRY(dev_alloc(b, &d.c_pack_all, per_rank * std::max(1, b->c_world)));
    V2_TRY(dev_alloc(b, &d.c_sums_send, CSUM_N));
    V2_TRY(dev_alloc(b, &d.c_sums_all, (size_t)CSUM_N * std::max(1, b->c_world)));
    d.chain = 1;
    d.c_n = (int)b->c_n;
    d.c_goff = (int)b->c_goff;
    d.c_j0 = b->c_j0;
    d.c_j1 = b->c_j1;
    d.c_lo = b->c_lo;
    d.c_hi = b->c_hi;
    d.c_rank = b->c_rank;
    d.c_world = b->c_world;
  }
  d.trials = p_trials;
  d.imu_t = p_imu_t;
  d.imu_s = p_imu_s;
  d.vic_t = p_vic_t;
  d.vic_s = p_vic_s;
  d.vic_c = p_vic_c;
  d.vic_c_off = p_coff;
  cudaStream_t st = b->stream;
  {
    // chunks of consecutive trials: the pinned fill of chunk k+1 runs while chunk k's ranges (contiguous per array) are
    // on their way to the device
    const int nch = std::max(1, std::min(T, 6));
    for (int ch = 0; ch < nch; ch++) {
      const int t0 = (int)((long long)T * ch / nch), t1 = (int)((long long)T * (ch + 1) / nch);
      if (t1 <= t0)
        continue;
      parallel_trials(t0, t1, fill_trial);
      const TrialDev &a = td[t0], &z = td[t1 - 1];
      const long long m0 = a.imu_off, m1 = z.imu_off + z.n_imu, v0 = a.vic_off, v1 = z.vic_off + z.n_vic;
      const long long s0 = a.st_off, s1 = z.st_off + z.n_raw;
      bool all_dev = true; // every trial of the chunk comes from a simulated batch: no measurement rows to send
      for (int t = t0; t < t1; t++)
        all_dev = all_dev && b->trials[t].dev_imu_s;
      if (m1 > m0) {
        V2_CUDA_CHECK(cudaMemcpyAsync(p_imu_t + m0, h_imu_t + m0, sizeof(double) * (m1 - m0), cudaMemcpyHostToDevice, st));
        if (!all_dev)
          V2_CUDA_CHECK(cudaMemcpyAsync(p_imu_s + 8 * m0, h_imu_s + 8 * m0, sizeof(double) * 8 * (m1 - m0), cudaMemcpyHostToDevice, st));
      }
      if (v1 > v0) {
        V2_CUDA_CHECK(cudaMemcpyAsync(p_vic_t + v0, h_vic_t + v0, sizeof(double) * (v1 - v0), cudaMemcpyHostToDevice, st));
        if (!all_dev)
          V2_CUDA_CHECK(cudaMemcpyAsync(p_vic_s + 8 * v0, h_vic_s + 8 * v0, sizeof(double) * 8 * (v1 - v0), cudaMemcpyHostToDevice, st));
      }
      for (int t = t0; t < t1; t++) {
        const HostTrial &h = b->trials[t];
        if (!h.dev_imu_s)
          continue;
        if (td[t].n_imu)
          V2_CUDA_CHECK(cudaMemcpyAsync(p_imu_s + 8 * td[t].imu_off, h.dev_imu_s, sizeof(double) * 8 * td[t].n_imu, cudaMemcpyDeviceToDevice, st));
        if (td[t].n_vic)
          V2_CUDA_CHECK(cudaMemcpyAsync(p_vic_s + 8 * td[t].vic_off, h.dev_vic_s, sizeof(double) * 8 * td[t].n_vic, cudaMemcpyDeviceToDevice, st));
      }
      long long c0 = -1, c1 = -1; // range of the per-pose covariances of this chunk (doubles)
      for (int t = t0; t < t1; t++)
        if (h_coff[t] >= 0) {
          if (c0 < 0)
            c0 = h_coff[t];
          c1 = h_coff[t] + 18 * td[t].n_vic;
        }
      if (c1 > c0 && c0 >= 0)
        V2_CUDA_CHECK(cudaMemcpyAsync(p_vic_c + c0, h_vic_c + c0, sizeof(double) * (c1 - c0), cudaMemcpyHostToDevice, st));
      if (s1 > s0) {
        V2_CUDA_CHECK(cudaMemcpyAsync(d.st_t + s0, h_st_t + s0, sizeof(double) * (s1 - s0), cudaMemcpyHostToDevice, st));
        V2_CUDA_CHECK(cudaMemcpyAsync(d.trial_of + s0, h_trial_of + s0, sizeof(int) * (s1 - s0), cudaMemcpyHostToDevice, st));
      }
    }
  }
  V2_CUDA_CHECK(cudaMemcpyAsync(p_coff, h_coff, sizeof(long long) * T, cudaMemcpyHostToDevice, st));
  V2_CUDA_CHECK(cudaMemcpyAsync(d.n_states, h_nst, sizeof(int) * T, cudaMemcpyHostToDevice, st));
  V2_CUDA_CHECK(cudaMemcpyAsync(p_trials, td.data(), sizeof(TrialDev) * T, cudaMemcpyHostToDevice, st));
  V2_CUDA_CHECK(cudaMemcpyAsync(d.lm, lm.data(), sizeof(LmState) * T, cudaMemcpyHostToDevice, st));
  V2_CUDA_CHECK(cudaMemsetAsync(d.removed_no_vicon, 0, sizeof(int) * T, st));
  V2_CUDA_CHECK(cudaMemsetAsync(d.x, 0, sizeof(double) * 2 * S * X_STRIDE, st));
  V2_CUDA_CHECK(cudaMemsetAsync(d.delta, 0, sizeof(double) * S * 16, st));
  V2_CUDA_CHECK(cudaMemsetAsync(d.trace, 0, sizeof(double) * T * TRACE_MAX * 5, st));
  V2_CUDA_CHECK(cudaMemsetAsync(d.xg, 0, sizeof(double) * T * 16, st));
  V2_CUDA_CHECK(cudaMemsetAsync(d.l2_trials, 0, sizeof(TrialDev) * T, st));
  V2_CUDA_CHECK(cudaMemsetAsync(d.l2_nstates, 0, sizeof(int) * T, st));
  if (b->chain) {
    V2_CUDA_CHECK(cudaMemsetAsync(d.c_sums, 0, sizeof(double) * 8, st));
    V2_CUDA_CHECK(cudaMemsetAsync(d.c_gamma_loc, 0, sizeof(double) * 96, st));
    V2_CUDA_CHECK(cudaMemsetAsync(d.c_sums_send, 0, sizeof(double) * CSUM_N, st));
    V2_CUDA_CHECK(cudaMemsetAsync(d.c_sums_all, 0, sizeof(double) * CSUM_N * std::max(1, b->c_world), st));
    V2_CUDA_CHECK(cudaMemsetAsync(d.c_sepne, 0, sizeof(double) * ((size_t)P + 2) * NE_STRIDE, st));
    V2_CUDA_CHECK(cudaMemsetAsync(d.seg, 0, sizeof(double) * (size_t)T * P * SEG_STRIDE, st));
  }
  V2_CUDA_CHECK(cudaStreamSynchronize(st)); // td / lm are stack-owned host vectors
  b->reuse = false;
  b->uploaded = true;
  b->built = false;
  b->values_init = false;
  for (int k = 0; k < 8; k++) {
    b->ms[k] = 0;
    b->launches[k] = 0;
  }
  return V2GT_OK;
}

// initial values of the four global nodes (ViconGraphSolver.cpp:400-413)
static int init_globals(v2gt_batch *b) {
  const int T = b->T;
  std::vector<double> g((size_t)2 * T * GLOB_STRIDE, 0.0);
  for (int t = 0; t < T; t++) {
    const v2gt_params &p = b->trials[t].prm;
    double q[4], rpy[3];
    v2::rot_2_quat(p.R_BtoI, q);
    v2::rot2rpy(p.R_GtoV, rpy);
    for (int half = 0; half < 2; half++) {
      double *o = &g[((size_t)half * T + t) * GLOB_STRIDE];
      o[GL_Q] = q[0];
      o[GL_Q + 1] = q[1];
      o[GL_Q + 2] = q[2];
      o[GL_Q + 3] = q[3];
      o[GL_P] = p.p_BinI[0];
      o[GL_P + 1] = p.p_BinI[1];
      o[GL_P + 2] = p.p_BinI[2];
      o[GL_TOFF] = p.toff_imu_to_vicon;
      o[GL_THX] = rpy[0];
      o[GL_THY] = rpy[1];
    }
  }
  V2_CUDA_CHECK(cudaMemcpyAsync(b->d.glob, g.data(), sizeof(double) * g.size(), cudaMemcpyHostToDevice, b->stream));
  V2_CUDA_CHECK(cudaStreamSynchronize(b->stream));
  return V2GT_OK;
}

int v2gt_batch_build(v2gt_batch *b, int32_t init_states) {
  if (!b)
    return V2GT_ERR_INVALID_ARG;
  V2_CUDA_CHECK(cudaSetDevice(b->device));
  if (!b->uploaded)
    V2_TRY(v2gt_batch_upload(b));
  b->results_cached = false;
  if (init_states) {
    V2_TRY(init_globals(b));
    b->values_init = true;
  } else if (!b->values_init) {
    return V2GT_ERR_NOT_BUILT;
  }
  V2_CUDA_CHECK(cudaEventRecord(b->ev[0], b->stream));
  V2_TRY(launch_init_states(b->d, b->n_max, init_states ? 1 : 0, b->stream));
  std::swap(b->d.st_t, b->d.st_t_tmp); // k_compact wrote the surviving times into st_t_tmp
  V2_TRY(launch_preintegrate(b->d, b->n_max, b->opt_keep_cov, b->stream));
  V2_CUDA_CHECK(cudaEventRecord(b->ev[1], b->stream));
  b->launches[0] += 4;
  b->built = true;
  return V2GT_OK;
}

// Segments per trial for an elimination launch over n_running trials: close to target_warps / n_running, but such
// that the grid (n_running x ceil(P / warps per CTA) CTAs of equal work) fills whole waves of the GPU's resident CTA
// slots -- a launch of 5.5 waves costs 6.
static int choose_segments(int target_warps, int n_running, int p_max, int sm_count) {
  const int wpc = elim_warps_per_cta();
  const double slots = (double)std::max(1, sm_count) * elim_ctas_per_sm();
  const int base = std::max(1, (target_warps + n_running - 1) / n_running);
  if (base <= wpc) // very large batches: one CTA per trial is already more than enough warps; do not cut finer than asked
    return std::max(1, std::min(p_max, base));
  // candidates: 0.75 ... 1.25 of the aim for small CTAs; large CTAs (one per SM, whole waves are coarse) look further
  const double lo = wpc >= 8 ? 0.6 : 0.75, hi = wpc >= 8 ? 1.7 : 1.25;
  const int k0 = std::max(1, (int)(lo * base) / wpc), k1 = std::max(k0, (int)(hi * base + wpc - 1) / wpc);
  int best_p = std::max(1, std::min(p_max, base));
  double best_eff = -1.0;
  for (int k = k0; k <= k1; k++) {
    const int P = std::min(p_max, k * wpc);
    const double waves = (double)n_running * ((P + wpc - 1) / wpc) / slots;
    const double eff = waves / std::ceil(waves);
    if (eff > best_eff + 1e-9 || (eff > best_eff - 1e-9 && std::abs(P - base) < std::abs(best_p - base))) {
      best_eff = std::max(eff, best_eff);
      best_p = P;
    }
    if (P == p_max)
      break;
  }
  return std::max(1, best_p);
}

// evaluation-only: the segment geometry the optimiser would pick (host logic, no device needed)
int v2gt_plan_segments(int32_t n_states, int32_t n_running, int32_t target_warps, int32_t sm_count, int32_t *p_level1, int32_t *p_level2) {
  if (n_states < 1 || n_running < 1 || target_warps < 1 || sm_count < 1 || !p_level1 || !p_level2)
    return V2GT_ERR_INVALID_ARG;
  int pmax = std::max(1, 4 * (int)std::ceil(std::sqrt((double)n_states)));
  pmax = std::min(pmax, std::max(1, n_states / 32));
  pmax = std::max(1, std::min(pmax, std::max(1, n_states / 2)));
  const int P = choose_segments(target_warps, n_running, pmax, sm_count);
  const int l2max = std::max(2, (int)std::ceil(std::sqrt((double)pmax)) + 1);
  *p_level1 = P;
  *p_level2 = l2_segments(P, l2max);
  return V2GT_OK;
}

// ---- one damped solve ("try") of every running trial, enqueued on the handle's stream.
// ev: 7 events around the phases (plain launches only), or null.  Chain mode: the two exchanges are part of it.
static int chain_exchange(v2gt_batch *b, int which, cudaStream_t st);
static int seg_stride_host(long long n, int P) {
  long long st = (n + 1 + P - 1) / P;
  return (int)(st < 2 ? 2 : st);
}
static int enqueue_try(v2gt_batch *b, const DevPtrs &d, int count_active, cudaEvent_t *ev) {
  cudaStream_t st = b->stream;
  auto mark = [&](int k) -> int {
    if (ev)
      V2_CUDA_CHECK(cudaEventRecord(ev[k], st));
    return V2GT_OK;
  };
  V2_TRY(launch_lm_pre_try(d, st));
  V2_TRY(mark(0));
  V2_TRY(launch_linearize(d, b->n_max, 0, st));
  V2_TRY(mark(1));
  V2_TRY(launch_eliminate(d, 0, st));
  if (b->chain) {
    V2_TRY(launch_chain_pack(d, seg_stride_host(b->c_n, d.P), st));
    V2_TRY(mark(2));
    V2_TRY(chain_exchange(b, 0, st));
    V2_TRY(launch_chain_unpack(d, st));
  } else {
    V2_TRY(mark(2));
  }
  V2_TRY(launch_reduced(d, 0, st));
  V2_TRY(mark(3));
  V2_TRY(launch_backsub(d, 0, st));
  V2_TRY(mark(4));
  V2_TRY(launch_cost(d, b->n_max, 1, 0, st));
  V2_TRY(mark(5));
  if (b->chain) {
    V2_TRY(launch_chain_sum(d, 1, st));
    V2_TRY(chain_exchange(b, 1, st));
    V2_TRY(launch_chain_reduce(d, 1, st));
  }
  V2_TRY(launch_lm_decide(d, count_active, st));
  V2_TRY(mark(6));
  b->launches[1] += 3;
  b->launches[2] += 1;
  b->launches[3] += (l2_segments(d.P, d.l2_P) ? 7 : 2) + (b->chain ? 2 : 0);
  b->launches[4] += 1;
  b->launches[5] += 1;
  b->launches[6] += 2 + (b->chain ? 2 : 0);
  return V2GT_OK;
}

// The same sequence as an executable CUDA graph (captured once per segment count and upload): one graph launch per
// damped solve instead of ~20 kernel launches -- what a small batch or a single trajectory spends its time on.
static int try_graph(v2gt_batch *b, const DevPtrs &d, cudaGraphExec_t *out) {
  auto it = b->try_graphs.find(d.P);
  if (it != b->try_graphs.end()) {
    *out = it->second;
    return V2GT_OK;
  }
  cudaStream_t st = b->stream;
  int64_t keep[8];
  std::memcpy(keep, b->launches, sizeof(keep));
  V2_CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
  const int rc = enqueue_try(b, d, 1, nullptr);
  cudaGraph_t g = nullptr;
  const cudaError_t e = cudaStreamEndCapture(st, &g);
  std::memcpy(b->launches, keep, sizeof(keep)); // capturing launches nothing
  if (rc != V2GT_OK || e != cudaSuccess || !g) {
    if (g)
      cudaGraphDestroy(g);
    cudaGetLastError();
    return rc != V2GT_OK ? rc : V2GT_ERR_CUDA;
  }
  cudaGraphExec_t ge = nullptr;
  const cudaError_t e2 = cudaGraphInstantiate(&ge, g, 0);
  cudaGraphDestroy(g);
  if (e2 != cudaSuccess) {
    fprintf(stderr, "[vicon2gt_b200] cudaGraphInstantiate failed: %s\n", cudaGetErrorString(e2));
    return V2GT_ERR_CUDA;
  }
  b->try_graphs[d.P] = ge;
  *out = ge;
  return V2GT_OK;
}
static void count_try_launches(v2gt_batch *b, const DevPtrs &d) {
  b->launches[1] += 3;
  b->launches[2] += 1;
  b->launches[3] += (l2_segments(d.P, d.l2_P) ? 7 : 2) + (b->chain ? 2 : 0);
  b->launches[4] += 1;
  b->launches[5] += 1;
  b->launches[6] += 2 + (b->chain ? 2 : 0);
}

// worst case number of tries: every outer iteration may walk lambda from its floor to the upper bound
static int max_rounds_of(const v2gt_batch *b) {
  int max_rounds = 0;
  for (int t = 0; t < b->T; t++) {
    const v2gt_params &p = b->trials[t].prm;
    const double span = std::log(std::max(p.lm_lambda_upper_bound, 1.0) / std::max(1e-300, std::min(p.lm_lambda_initial, 1e-30))) /
                        std::log(std::max(p.lm_lambda_factor, 1.0001));
    int r = (int)((p.lm_max_iterations + 1) * (span + 2));
    if (p.lm_max_tries > 0)
      r = std::min(r, p.lm_max_tries + 1);
    max_rounds = std::max(max_rounds, r);
  }
  return max_rounds;
}

// The optimiser rounds shared by the batch and the chain-sharded solve.  Every `chunk` damped solves the host reads
// the number of trials still running; in between nothing leaves the device.
static int run_rounds(v2gt_batch *b, int max_rounds, bool use_graph) {
  cudaStream_t st = b->stream;
  const int chunk = b->opt_tries_per_sync;
  if (!use_graph)
    while ((int)b->ev_pool.size() < 7 * chunk) { // per-phase device timing: 7 events per try, re-used every chunk
      cudaEvent_t e;
      V2_CUDA_CHECK(cudaEventCreate(&e));
      b->ev_pool.push_back(e);
    }
  int rounds = 0;
  int n_running = b->T;
  const DevPtrs &dmax = b->d;
  while (rounds < max_rounds) {
    // segments per trial for this chunk of tries: fill the GPU with the trials that are still running
    DevPtrs d = dmax;
    if (b->opt_segments <= 0 && !b->chain)
      d.P = choose_segments(b->opt_target_warps, std::max(1, n_running), dmax.P, b->sm_count);
    const int todo = std::min(chunk, max_rounds - rounds);
    cudaGraphExec_t ge = nullptr;
    if (use_graph)
      V2_TRY(try_graph(b, d, &ge));
    for (int k = 0; k < todo; k++) {
      if (k == todo - 1)
        V2_CUDA_CHECK(cudaMemsetAsync(d.n_active, 0, sizeof(int), st)); // only the chunk's last decision counts
      if (use_graph) {
        V2_CUDA_CHECK(cudaGraphLaunch(ge, st));
        count_try_launches(b, d);
      } else {
        V2_TRY(enqueue_try(b, d, k == todo - 1 ? 1 : 0, &b->ev_pool[7 * k]));
      }
    }
    rounds += todo;
    V2_CUDA_CHECK(cudaMemcpyAsync(b->h_active, d.n_active, sizeof(int), cudaMemcpyDeviceToHost, st));
    V2_CUDA_CHECK(cudaStreamSynchronize(st));
    if (!use_graph)
      for (int k = 0; k < todo; k++) {
        cudaEvent_t *e = &b->ev_pool[7 * k];
        for (int ph = 0; ph < 6; ph++) {
          float m = 0;
          cudaEventElapsedTime(&m, e[ph], e[ph + 1]);
          b->ms[1 + ph] += m;
        }
      }
    b->tries_enqueued += todo;
    n_running = *b->h_active;
    if (n_running == 0)
      break;
  }
  return V2GT_OK;
}

int v2gt_batch_optimize(v2gt_batch *b) {
  if (!b)
    return V2GT_ERR_INVALID_ARG;
  if (!b->built)
    return V2GT_ERR_NOT_BUILT;
  if (b->chain)
    return V2GT_ERR_INVALID_ARG; // a sharded chain is optimised by v2gt_chain_optimize / v2gt_chain_phase
  V2_CUDA_CHECK(cudaSetDevice(b->device));
  b->results_cached = false;
  cudaStream_t st = b->stream;
  const DevPtrs &d = b->d;
  V2_CUDA_CHECK(cudaEventRecord(b->ev[2], st));
  V2_TRY(launch_cost(d, b->n_max, 0, 1, st));
  V2_TRY(launch_lm_begin(d, st));
  b->launches[5] += 1;
  b->launches[6] += 1;
  V2_TRY(solve_configure());
  // small problems are bound by the ~20 launches of a damped solve, not by the kernels: replay a captured graph
  const bool use_graph = b->opt_graph > 0 || (b->opt_graph < 0 && (long long)b->T * b->n_max <= 400000);
  V2_TRY(run_rounds(b, max_rounds_of(b), use_graph));
  V2_CUDA_CHECK(cudaEventRecord(b->ev[3], st));
  V2_CUDA_CHECK(cudaStreamSynchronize(st));
  float ms = 0;
  cudaEventElapsedTime(&ms, b->ev[0], b->ev[1]);
  b->ms[0] = ms;
  cudaEventElapsedTime(&ms, b->ev[2], b->ev[3]);
  b->ms[7] = ms;
  return V2GT_OK;
}

int v2gt_batch_build_and_solve(v2gt_batch *b) {
  if (!b)
    return V2GT_ERR_INVALID_ARG;
  if (!b->uploaded)
    V2_TRY(v2gt_batch_upload(b));
  int loops = 0;
  for (int t = 0; t < b->T; t++)
    loops = std::max(loops, (int)b->trials[t].prm.num_loop_relin);
  for (int i = 0; i <= loops; i++) {
    // a trial takes part in passes 0..num_loop_relin of ITS parameters (the reference loops per solver,
    // ViconGraphSolver.cpp:163); afterwards it keeps its result while the rest of the batch goes on
    V2_CUDA_CHECK(cudaSetDevice(b->device));
    V2_TRY(launch_lm_freeze(b->d, i, b->stream));
    V2_TRY(v2gt_batch_build(b, i == 0 ? 1 : 0));
    V2_TRY(v2gt_batch_optimize(b));
  }
  return V2GT_OK;
}

int v2gt_batch_sync(v2gt_batch *b) {
  if (!b)
    return V2GT_ERR_INVALID_ARG;
  V2_CUDA_CHECK(cudaSetDevice(b->device));
  V2_CUDA_CHECK(cudaStreamSynchronize(b->stream));
  return V2GT_OK;
}

int v2gt_batch_get_info(v2gt_batch *b, int32_t trial, v2gt_trial_info *info) {
  if (!b || !info || trial < 0 || trial >= b->T)
    return V2GT_ERR_INVALID_ARG;
  *info = b->trials[trial].info;
  if (!b->uploaded)
    return V2GT_OK;
  V2_CUDA_CHECK(cudaSetDevice(b->device));
  V2_TRY(fetch_results(b));
  const LmState &lm = b->h_lm[trial];
  info->status = lm.status;
  info->n_states = b->h_nstates[trial];
  info->removed_no_vicon = b->h_removed[trial];
  info->iterations = lm.iterations;
  info->tries = lm.tries;
  info->linearizations = lm.linearizations;
  info->lm_state = lm.lm_state;
  info->initial_error = lm.initial_error;
  info->final_error = lm.error;
  info->lambda = lm.lambda;
  return V2GT_OK;
}

int v2gt_batch_get_states(v2gt_batch *b, int32_t trial, int64_t capacity, double *times, double *states, int64_t *n_out) {
  if (!b || trial < 0 || trial >= b->T || !b->built)
    return b && !b->built ? V2GT_ERR_NOT_BUILT : V2GT_ERR_INVALID_ARG;
  V2_CUDA_CHECK(cudaSetDevice(b->device));
  V2_TRY(fetch_results(b));
  const int n = b->h_nstates[trial];
  if (n_out)
    *n_out = n;
  if (capacity < n)
    return (times || states) ? V2GT_ERR_INVALID_ARG : V2GT_OK;
  std::vector<TrialDev> td(1);
  V2_CUDA_CHECK(cudaMemcpy(td.data(), b->d.trials + trial, sizeof(TrialDev), cudaMemcpyDeviceToHost));
  const long long off = td[0].st_off;
  const int cur = b->h_lm[trial].cur;
  if (times && n)
    V2_CUDA_CHECK(cudaMemcpy(times, b->d.st_t + off, sizeof(double) * n, cudaMemcpyDeviceToHost));
  if (states && n)
    V2_CUDA_CHECK(cudaMemcpy(states, b->d.x + ((long long)cur * b->d.S + off) * X_STRIDE, sizeof(double) * n * X_STRIDE,
                             cudaMemcpyDeviceToHost));
  return V2GT_OK;
}

int v2gt_batch_get_all_states(v2gt_batch *b, int64_t capacity, int64_t *offsets, double *times, double *states, double *globals10) {
  if (!b || !b->built || !offsets)
    return b && !b->built ? V2GT_ERR_NOT_BUILT : V2GT_ERR_INVALID_ARG;
  V2_CUDA_CHECK(cudaSetDevice(b->device));
  V2_TRY(fetch_results(b));
  const int T = b->T;
  offsets[0] = 0;
  for (int t = 0; t < T; t++)
    offsets[t + 1] = offsets[t] + b->h_nstates[t];
  const int64_t total = offsets[T];
  if (!times && !states && !globals10)
    return V2GT_OK;
  if (capacity < total)
    return V2GT_ERR_INVALID_ARG;
  // bounce buffer: [states 16 x total | times total | globals 2 x T x 16]
  const size_t need = (size_t)total * 17 + (size_t)2 * T * GLOB_STRIDE;
  double *pin = b->h_stage;
  if (b->h_stage_doubles < need) {
    if (b->h_rb_doubles < need) {
      if (b->h_rb)
        cudaFreeHost(b->h_rb);
      b->h_rb = nullptr;
      b->h_rb_doubles = 0;
      V2_CUDA_CHECK(cudaMallocHost((void **)&b->h_rb, need * sizeof(double)));
      b->h_rb_doubles = need;
    }
    pin = b->h_rb;
  }
  double *p_x = pin, *p_t = pin + (size_t)total * 16, *p_g = pin + (size_t)total * 17;
  cudaStream_t st = b->stream;
  const DevPtrs &d = b->d;
  V2_CUDA_CHECK(cudaMemcpyAsync(p_g, d.glob, sizeof(double) * 2 * T * GLOB_STRIDE, cudaMemcpyDeviceToHost, st));
  // chunks of trials: the copies of chunk k+1 are in flight while chunk k is moved to the caller's arrays on host threads
  const int nch = std::max(1, std::min(T, 8));
  std::vector<cudaEvent_t> ev((size_t)nch);
  for (int ch = 0; ch < nch; ch++) {
    const int t0 = (int)((long long)T * ch / nch), t1 = (int)((long long)T * (ch + 1) / nch);
    for (int t = t0; t < t1; t++) {
      const int n = b->h_nstates[t];
      if (!n)
        continue;
      const long long off = b->h_st_off[(size_t)t];
      if (states)
        V2_CUDA_CHECK(cudaMemcpyAsync(p_x + offsets[t] * 16, d.x + ((long long)b->h_lm[t].cur * d.S + off) * X_STRIDE,
                                      sizeof(double) * n * X_STRIDE, cudaMemcpyDeviceToHost, st));
      if (times)
        V2_CUDA_CHECK(cudaMemcpyAsync(p_t + offsets[t], d.st_t + off, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    }
    V2_CUDA_CHECK(cudaEventCreateWithFlags(&ev[(size_t)ch], cudaEventDisableTiming));
    V2_CUDA_CHECK(cudaEventRecord(ev[(size_t)ch], st));
  }
  int rc = V2GT_OK;
  for (int ch = 0; ch < nch; ch++) {
    const int t0 = (int)((long long)T * ch / nch), t1 = (int)((long long)T * (ch + 1) / nch);
    if (cudaEventSynchronize(ev[(size_t)ch]) != cudaSuccess)
      rc = V2GT_ERR_CUDA;
    cudaEventDestroy(ev[(size_t)ch]);
    if (rc != V2GT_OK || t1 <= t0)
      continue;
    parallel_trials(t0, t1, [&](int t) {
      const size_t n = (size_t)b->h_nstates[t];
      if (states && n)
        std::memcpy(states + offsets[t] * 16, p_x + offsets[t] * 16, sizeof(double) * 16 * n);
      if (times && n)
        std::memcpy(times + offsets[t], p_t + offsets[t], sizeof(double) * n);
    });
  }
  if (rc == V2GT_OK && globals10)
    for (int t = 0; t < T; t++) {
      const double *g = p_g + ((size_t)b->h_lm[t].cur * T + t) * GLOB_STRIDE;
      double *o = globals10 + 10 * t;
      std::memcpy(o, g + GL_Q, 4 * sizeof(double));
      std::memcpy(o + 4, g + GL_P, 3 * sizeof(double));
      o[7] = g[GL_TOFF];
      o[8] = g[GL_THX];
      o[9] = g[GL_THY];
    }
  return rc;
}

int v2gt_batch_get_calibration(v2gt_batch *b, int32_t trial, double *q_BtoI, double *p_BinI, double *t_off, double *theta_xy) {
  if (!b || trial < 0 || trial >= b->T || !b->values_init)
    return b && !b->values_init ? V2GT_ERR_NOT_BUILT : V2GT_ERR_INVALID_ARG;
  V2_CUDA_CHECK(cudaSetDevice(b->device));
  V2_TRY(fetch_results(b));
  const int cur = b->h_lm[trial].cur;
  double g[GLOB_STRIDE];
  V2_CUDA_CHECK(cudaMemcpy(g, b->d.glob + ((long long)cur * b->T + trial) * GLOB_STRIDE, sizeof(g), cudaMemcpyDeviceToHost));
  if (q_BtoI)
    std::memcpy(q_BtoI, g + GL_Q, 4 * sizeof(double));
  if (p_BinI)
    std::memcpy(p_BinI, g + GL_P, 3 * sizeof(double));
  if (t_off)
    *t_off = g[GL_TOFF];
  if (theta_xy) {
    theta_xy[0] = g[GL_THX];
    theta_xy[1] = g[GL_THY];
  }
  return V2GT_OK;
}

// ViconGraphSolver::write_to_file (.cpp:194-248) from host copies of one trial's results.  txt_path (optional): the same
// trajectory as the text file the Monte-Carlo script of the reference feeds to its evaluation tool
// (scripts/run_monte_carlo.sh:66-72): "timestamp(s) tx ty tz qx qy qz qw", one pose per line.
static int write_trial_files(const double *times, const double *X, int64_t n, const double *q_BtoI, const double *p_BinI, double toff,
                             const double *th, double gravity, const char *csv_path, const char *info_path, const char *txt_path) {
  double R_GtoV[9], q_GtoV[4];
  v2::rot_xy(th[0], th[1], R_GtoV);
  v2::rot_2_quat(R_GtoV, q_GtoV);
  if (csv_path) {
    std::remove(csv_path);
    make_dirs_for(csv_path);
  }
  if (txt_path) {
    std::remove(txt_path);
    make_dirs_for(txt_path);
  }
  std::ofstream of_state, of_txt;
  if (csv_path) {
    of_state.open(csv_path, std::ofstream::out | std::ofstream::app);
    if (!of_state)
      return V2GT_ERR_INVALID_ARG;
    of_state << "#time(ns),px,py,pz,qw,qx,qy,qz,vx,vy,vz,bwx,bwy,bwz,bax,bay,baz" << std::endl;
  }
  if (txt_path) {
    of_txt.open(txt_path, std::ofstream::out | std::ofstream::app);
    if (!of_txt)
      return V2GT_ERR_INVALID_ARG;
    of_txt << "# timestamp(s) tx ty tz qx qy qz qw" << std::endl;
  }
  for (int64_t i = 0; i < n; i++) {
    const double *x = &X[(size_t)i * 16];
    double q[4], p[3], v[3];
    v2::quat_mul(x, q_GtoV, q);
    v2::mv3_t(R_GtoV, x + 13, p);
    v2::mv3_t(R_GtoV, x + 7, v);
    if (csv_path)
      of_state << std::setprecision(20) << std::floor(1e9 * times[(size_t)i]) << "," << std::setprecision(6) << p[0] << "," << p[1]
               << "," << p[2] << "," << q[3] << "," << q[0] << "," << q[1] << "," << q[2] << "," << v[0] << "," << v[1] << "," << v[2]
               << "," << x[4] << "," << x[5] << "," << x[6] << "," << x[10] << "," << x[11] << "," << x[12] << "\n";
    if (txt_path) {
      // the time the CSV holds (floor of the nanosecond count), back in seconds
      const double ts = 1e-9 * std::floor(1e9 * times[(size_t)i]);
      of_txt << std::fixed << std::setprecision(9) << ts << std::defaultfloat << std::setprecision(6) << " " << p[0] << " " << p[1] << " "
             << p[2] << " " << q[0] << " " << q[1] << " " << q[2] << " " << q[3] << "\n";
    }
  }
  if (csv_path)
    of_state.close();
  if (txt_path)
    of_txt.close();
  if (!info_path)
    return V2GT_OK;
  std::remove(info_path);
  make_dirs_for(info_path);
  std::ofstream of_info(info_path, std::ofstream::out | std::ofstream::app);
  if (!of_info)
    return V2GT_ERR_INVALID_ARG;
  double R_BtoI[9];
  v2::quat_2_Rot(q_BtoI, R_BtoI);
  of_info << "R_BtoI: " << std::endl << eigen_print(R_BtoI, 3, 3) << std::endl << std::endl;
  of_info << "q_BtoI: " << std::endl << eigen_print(q_BtoI, 4, 1) << std::endl << std::endl;
  of_info << "p_BinI: " << std::endl << eigen_print(p_BinI, 3, 1) << std::endl << std::endl;
  of_info << "R_GtoV: " << std::endl << eigen_print(R_GtoV, 3, 3) << std::endl << std::endl;
  of_info << "R_GtoV (thetax, thetay): " << std::endl;
  of_info << th[0] << " " << th[0] << std::endl << std::endl; // thetax printed twice, as the reference does (.cpp:244)
  of_info << "gravity norm: " << std::endl << gravity << std::endl << std::endl;
  of_info << "t_off_vicon_to_imu: " << std::endl << eigen_print(&toff, 1, 1) << std::endl << std::endl;
  of_info.close();
  return V2GT_OK;
}

int v2gt_batch_write_to_file(v2gt_batch *b, int32_t trial, const char *csv_path, const char *info_path) {
  if (!b || !csv_path || !info_path)
    return V2GT_ERR_INVALID_ARG;
  int64_t n = 0;
  V2_TRY(v2gt_batch_get_states(b, trial, 0, nullptr, nullptr, &n));
  std::vector<double> times((size_t)n), X((size_t)n * 16);
  V2_TRY(v2gt_batch_get_states(b, trial, n, times.data(), X.data(), &n));
  double q_BtoI[4], p_BinI[3], toff, th[2];
  V2_TRY(v2gt_batch_get_calibration(b, trial, q_BtoI, p_BinI, &toff, th));
  return write_trial_files(times.data(), X.data(), n, q_BtoI, p_BinI, toff, th, b->trials[trial].prm.gravity_magnitude, csv_path,
                           info_path, nullptr);
}

// Every solved trial of the batch: <dir>/<trial, 5 digits>_states.csv, _info.txt and (with_txt) _states.txt.  One device
// read-back for the whole batch, then the files are written on several host threads.
int v2gt_batch_write_all(v2gt_batch *b, const char *dir, int32_t with_txt, int32_t *n_written) {
  if (!b || !dir || !b->built)
    return b && !b->built ? V2GT_ERR_NOT_BUILT : V2GT_ERR_INVALID_ARG;
  V2_CUDA_CHECK(cudaSetDevice(b->device));
  V2_TRY(fetch_results(b));
  const int T = b->T;
  const long long S = b->d.S;
  std::vector<double> X((size_t)2 * S * X_STRIDE), times((size_t)S), glob((size_t)2 * T * GLOB_STRIDE);
  std::vector<TrialDev> td((size_t)T);
  V2_CUDA_CHECK(cudaMemcpy(X.data(), b->d.x, sizeof(double) * X.size(), cudaMemcpyDeviceToHost));
  V2_CUDA_CHECK(cudaMemcpy(times.data(), b->d.st_t, sizeof(double) * times.size(), cudaMemcpyDeviceToHost));
  V2_CUDA_CHECK(cudaMemcpy(glob.data(), b->d.glob, sizeof(double) * glob.size(), cudaMemcpyDeviceToHost));
  V2_CUDA_CHECK(cudaMemcpy(td.data(), b->d.trials, sizeof(TrialDev) * T, cudaMemcpyDeviceToHost));
  std::vector<int> rc((size_t)T, V2GT_OK), wrote((size_t)T, 0);
  const std::string base(dir);
  parallel_trials(T, [&](int t) {
    if (b->h_lm[t].status != V2GT_OK)
      return;
    const int cur = b->h_lm[t].cur;
    const long long off = td[t].st_off;
    const double *g = &glob[((size_t)cur * T + t) * GLOB_STRIDE];
    char name[32];
    std::snprintf(name, sizeof(name), "/%05d", t);
    const std::string stem = base + name;
    const std::string csv = stem + "_states.csv", info = stem + "_info.txt", txt = stem + "_states.txt";
    const double th[2] = {g[GL_THX], g[GL_THY]};
    rc[t] = write_trial_files(&times[(size_t)off], &X[((size_t)cur * S + off) * X_STRIDE], b->h_nstates[t], g + GL_Q, g + GL_P,
                              g[GL_TOFF], th, b->trials[t].prm.gravity_magnitude, csv.c_str(), info.c_str(),
                              with_txt ? txt.c_str() : nullptr);
    wrote[t] = (rc[t] == V2GT_OK) ? 1 : 0;
  });
  int total = 0;
  for (int t = 0; t < T; t++) {
    if (rc[t] != V2GT_OK)
      return rc[t];
    total += wrote[t];
  }
  if (n_written)
    *n_written = total;
  return V2GT_OK;
}

// ---- batched trajectory-error statistics (csrc/k_stats.cu)
int v2gt_batch_set_truth(v2gt_batch *b, int32_t trial, int64_t n, const double *truth10) {
  if (!b || trial < 0 || trial >= b->T || n < 0 || (n && !truth10))
    return V2GT_ERR_INVALID_ARG;
  if ((int)b->truth.size() != b->T)
    b->truth.assign((size_t)b->T, std::vector<double>());
  b->truth[trial].assign(truth10, truth10 + (size_t)n * 10);
  return V2GT_OK;
}

int v2gt_batch_error_stats(v2gt_batch *b, double *stats) {
  if (!b || !stats)
    return V2GT_ERR_INVALID_ARG;
  if (!b->built)
    return V2GT_ERR_NOT_BUILT;
  V2_CUDA_CHECK(cudaSetDevice(b->device));
  V2_TRY(fetch_results(b));
  const int T = b->T;
  const long long S = b->d.S;
  if ((int)b->truth.size() != T)
    return V2GT_ERR_INVALID_ARG;
  std::vector<TrialDev> td((size_t)T);
  V2_CUDA_CHECK(cudaMemcpy(td.data(), b->d.trials, sizeof(TrialDev) * T, cudaMemcpyDeviceToHost));
  std::vector<double> h((size_t)std::max<long long>(1, S) * 10, 0.0);
  for (int t = 0; t < T; t++) {
    if (b->h_lm[t].status != V2GT_OK)
      continue;
    const size_t n = (size_t)b->h_nstates[t];
    if (b->truth[t].size() != n * 10)
      return V2GT_ERR_INVALID_ARG; // truth must be given at the surviving state times of the trial
    std::memcpy(&h[(size_t)td[t].st_off * 10], b->truth[t].data(), sizeof(double) * n * 10);
  }
  double *d_truth = nullptr, *d_out = nullptr;
  V2_CUDA_CHECK(cudaMalloc(&d_truth, sizeof(double) * h.size()));
  cudaError_t e = cudaMalloc(&d_out, sizeof(double) * (size_t)T * 24);
  if (e != cudaSuccess) {
    cudaFree(d_truth);
    return V2GT_ERR_CUDA;
  }
  int rc = V2GT_OK;
  if (cudaMemcpyAsync(d_truth, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice, b->stream) != cudaSuccess)
    rc = V2GT_ERR_CUDA;
  if (rc == V2GT_OK)
    rc = launch_error_stats(b->d, d_truth, d_out, b->stream);
  if (rc == V2GT_OK && cudaMemcpy(stats, d_out, sizeof(double) * (size_t)T * 24, cudaMemcpyDeviceToHost) != cudaSuccess)
    rc = V2GT_ERR_CUDA;
  cudaFree(d_truth);
  cudaFree(d_out);
  return rc;
}

int v2gt_chain_configure(v2gt_batch *b, int32_t rank, int32_t world, int64_t n_total, int64_t g_off, int32_t p_total, int32_t j0,
                         int32_t j1, int32_t own_lo, int32_t own_hi) {
  if (!b || b->T != 1 || world < 1 || rank < 0 || rank >= world || n_total < 2 || p_total < 1 || j0 < 0 || j1 <= j0 || j1 > p_total ||
      own_lo < 0 || own_hi <= own_lo || g_off < 0)
    return V2GT_ERR_INVALID_ARG;
  b->chain = true;
  b->c_rank = rank;
  b->c_world = world;
  b->c_n = n_total;
  b->c_goff = g_off;
  b->c_p = p_total;
  b->c_j0 = j0;
  b->c_j1 = j1;
  b->c_lo = own_lo;
  b->c_hi = own_hi;
  b->uploaded = false;
  b->built = false;
  return V2GT_OK;
}

// exchange 0: PACK_SEND -> PACK_ALL, 1: SUMS_SEND -> SUMS_ALL; an all-gather in rank order on the handle's stream
static int chain_exchange(v2gt_batch *b, int which, cudaStream_t st) {
  const DevPtrs &d = b->d;
  const size_t n = which == 0 ? (size_t)(b->c_j1 - b->c_j0) * (SEG_STRIDE + NE_STRIDE) + 96 : (size_t)CSUM_N;
  const double *src = which == 0 ? d.c_pack_send : d.c_sums_send;
  double *dst = which == 0 ? d.c_pack_all : d.c_sums_all;
  if (b->c_world == 1) {
    V2_CUDA_CHECK(cudaMemcpyAsync(dst, src, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    return V2GT_OK;
  }
  if (!b->c_comm)
    return V2GT_ERR_NCCL; // v2gt_chain_comm_init first (or drive the phases and move the buffers yourself)
  return nccl_check(nccl_api().AllGather(src, dst, n, NCCL_FLOAT64, b->c_comm, st), "ncclAllGather");
}

int v2gt_chain_plan(int64_t n_total, int32_t world, int32_t segs_per_rank, int32_t rank, v2gt_chain_plan_t *plan) {
  if (!plan || world < 1 || segs_per_rank < 1 || rank < 0 || rank >= world || n_total < 2)
    return V2GT_ERR_INVALID_ARG;
  const long long p = (long long)world * segs_per_rank;
  if (p > (1 << 24))
    return V2GT_ERR_INVALID_ARG;
  const int stride = seg_stride_host(n_total, (int)p);
  // every segment must exist and hold at least one interior state besides its separator
  if ((p - 1) * stride >= n_total || stride < 3)
    return V2GT_ERR_INVALID_ARG;
  const int j0 = rank * segs_per_rank, j1 = (rank + 1) * segs_per_rank;
  // own states: the separators in front of the own segments and the segments themselves, i.e. the contiguous range
  // [j0 * stride - 1, j1 * stride - 1); rank 0 starts at state 0, the last rank ends at n_total
  const long long start = rank > 0 ? (long long)j0 * stride - 1 : 0;
  const long long end = rank < world - 1 ? (long long)j1 * stride - 1 : n_total;
  const long long g_off = start - (rank > 0 ? 1 : 0);                        // left halo: the previous rank's last own state
  const long long n_local = end + (rank < world - 1 ? 1 : 0) - g_off;        // right halo: the next rank's leading separator
  plan->rank = rank;
  plan->world = world;
  plan->p_total = (int32_t)p;
  plan->stride = stride;
  plan->j0 = j0;
  plan->j1 = j1;
  plan->own_lo = (int32_t)(start - g_off);
  plan->own_hi = (int32_t)(end - g_off);
  plan->n_total = n_total;
  plan->g_off = g_off;
  plan->n_local = n_local;
  return V2GT_OK;
}

int v2gt_chain_fit_segments(int64_t n_total, int32_t world, int32_t target, int32_t *segs_per_rank) {
  if (!segs_per_rank || world < 1)
    return V2GT_ERR_INVALID_ARG;
  *segs_per_rank = 0;
  v2gt_chain_plan_t pl;
  for (int spr = std::max(1, (int)target); spr >= 1; spr--)
    if (v2gt_chain_plan(n_total, world, spr, 0, &pl) == V2GT_OK) {
      *segs_per_rank = spr;
      return V2GT_OK;
    }
  return V2GT_ERR_INVALID_ARG;
}

int v2gt_chain_nccl_unique_id(uint8_t *id_bytes) {
  if (!id_bytes)
    return V2GT_ERR_INVALID_ARG;
  NcclApi &n = nccl_api();
  if (!n.ok) {
    fprintf(stderr, "[vicon2gt_b200] libnccl.so.2 could not be loaded\n");
    return V2GT_ERR_NCCL;
  }
  NcclApi::UniqueId id;
  V2_TRY(nccl_check(n.GetUniqueId(&id), "ncclGetUniqueId"));
  std::memcpy(id_bytes, id.internal, V2GT_NCCL_ID_BYTES);
  return V2GT_OK;
}

int v2gt_chain_comm_init(v2gt_batch *b, const uint8_t *id_bytes) {
  if (!b || !b->chain || !id_bytes)
    return V2GT_ERR_INVALID_ARG;
  NcclApi &n = nccl_api();
  if (!n.ok) {
    fprintf(stderr, "[vicon2gt_b200] libnccl.so.2 could not be loaded\n");
    return V2GT_ERR_NCCL;
  }
  V2_CUDA_CHECK(cudaSetDevice(b->device));
  if (b->c_comm) {
    n.CommDestroy(b->c_comm);
    b->c_comm = nullptr;
    drop_graphs(b);
  }
  NcclApi::UniqueId id;
  std::memcpy(id.internal, id_bytes, V2GT_NCCL_ID_BYTES);
  return nccl_check(n.CommInitRank(&b->c_comm, b->c_world, id, b->c_rank), "ncclCommInitRank");
}

int v2gt_chain_optimize(v2gt_batch *b, int32_t max_tries, int32_t use_graph) {
  if (!b || !b->chain)
    return V2GT_ERR_INVALID_ARG;
  if (!b->built)
    return V2GT_ERR_NOT_BUILT;
  if (b->c_world > 1 && !b->c_comm)
    return V2GT_ERR_NCCL;
  V2_CUDA_CHECK(cudaSetDevice(b->device));
  b->results_cached = false;
  cudaStream_t st = b->stream;
  const DevPtrs &d = b->d;
  V2_CUDA_CHECK(cudaEventRecord(b->ev[2], st));
  V2_TRY(launch_cost(d, b->n_max, 0, 1, st));
  V2_TRY(launch_chain_sum(d, 0, st));
  V2_TRY(chain_exchange(b, 1, st));
  V2_TRY(launch_chain_reduce(d, 0, st));
  V2_TRY(launch_lm_begin(d, st));
  b->launches[5] += 1;
  b->launches[6] += 3;
  V2_TRY(solve_configure());
  int cap = max_rounds_of(b);
  if (max_tries > 0)
    cap = std::min(cap, (int)max_tries);
  V2_TRY(run_rounds(b, cap, use_graph != 0));
  V2_CUDA_CHECK(cudaEventRecord(b->ev[3], st));
  V2_CUDA_CHECK(cudaStreamSynchronize(st));
  float ms = 0;
  cudaEventElapsedTime(&ms, b->ev[0], b->ev[1]);
  b->ms[0] = ms;
  cudaEventElapsedTime(&ms, b->ev[2], b->ev[3]);
  b->ms[7] = ms;
  return V2GT_OK;
}

int v2gt_chain_phase(v2gt_batch *b, int32_t phase) {
  if (!b || !b->chain)
    return V2GT_ERR_INVALID_ARG;
  if (!b->built)
    return V2GT_ERR_NOT_BUILT;
  V2_CUDA_CHECK(cudaSetDevice(b->device));
  b->results_cached = false;
  cudaStream_t st = b->stream;
  const DevPtrs &d = b->d;
  switch (phase) {
  case V2GT_CHAIN_PHASE_COST0:
    V2_TRY(launch_cost(d, b->n_max, 0, 1, st));
    V2_TRY(launch_chain_sum(d, 0, st));
    break;
  case V2GT_CHAIN_PHASE_BEGIN:
    V2_TRY(launch_chain_reduce(d, 0, st));
    V2_TRY(launch_lm_begin(d, st));
    V2_TRY(solve_configure());
    break;
  case V2GT_CHAIN_PHASE_LINEARIZE:
    V2_TRY(launch_lm_pre_try(d, st));
    V2_TRY(launch_linearize(d, b->n_max, 0, st)); // the local part of the global block stays in c_gamma_loc
    break;
  case V2GT_CHAIN_PHASE_ELIMINATE:
    V2_TRY(launch_eliminate(d, 0, st));
    V2_TRY(launch_chain_pack(d, seg_stride_host(b->c_n, d.P), st));
    break;
  case V2GT_CHAIN_PHASE_SOLVE:
    V2_TRY(launch_chain_unpack(d, st));
    V2_TRY(launch_reduced(d, 0, st));
    V2_TRY(launch_backsub(d, 0, st));
    break;
  case V2GT_CHAIN_PHASE_COST:
    V2_TRY(launch_cost(d, b->n_max, 1, 0, st));
    V2_TRY(launch_chain_sum(d, 1, st));
    break;
  case V2GT_CHAIN_PHASE_DECIDE:
    V2_TRY(launch_chain_reduce(d, 1, st));
    V2_TRY(launch_lm_decide(d, 0, st));
    break;
  default:
    return V2GT_ERR_INVALID_ARG;
  }
  return V2GT_OK;
}

int v2gt_chain_buffer(v2gt_batch *b, int32_t which, void **dev_ptr, int64_t *n_doubles) {
  if (!b || !b->chain || !b->uploaded || !dev_ptr || !n_doubles)
    return V2GT_ERR_INVALID_ARG;
  const DevPtrs &d = b->d;
  const int64_t ploc = b->c_j1 - b->c_j0;
  switch (which) {
  case V2GT_CHAIN_BUF_SUMS: *dev_ptr = d.c_sums; *n_doubles = 8; break;
  case V2GT_CHAIN_BUF_GAMMA: *dev_ptr = d.gamma; *n_doubles = 96; break;
  case V2GT_CHAIN_BUF_PACK_SEND: *dev_ptr = d.c_pack_send; *n_doubles = ploc * (SEG_STRIDE + NE_STRIDE) + 96; break;
  case V2GT_CHAIN_BUF_PACK_ALL:
    *dev_ptr = d.c_pack_all;
    *n_doubles = (ploc * (SEG_STRIDE + NE_STRIDE) + 96) * std::max(1, b->c_world);
    break;
  case V2GT_CHAIN_BUF_SEG_ALL: *dev_ptr = d.seg; *n_doubles = (int64_t)d.P * SEG_STRIDE; break;
  case V2GT_CHAIN_BUF_SEPNE_ALL: *dev_ptr = d.c_sepne; *n_doubles = (int64_t)d.P * NE_STRIDE; break;
  case V2GT_CHAIN_BUF_SUMS_SEND: *dev_ptr = d.c_sums_send; *n_doubles = CSUM_N; break;
  case V2GT_CHAIN_BUF_SUMS_ALL: *dev_ptr = d.c_sums_all; *n_doubles = (int64_t)CSUM_N * b->c_world; break;
  default: return V2GT_ERR_INVALID_ARG;
  }
  return V2GT_OK;
}

int v2gt_batch_stream(v2gt_batch *b, void **stream) {
  if (!b || !stream)
    return V2GT_ERR_INVALID_ARG;
  *stream = (void *)b->stream;
  return V2GT_OK;
}

int v2gt_batch_get_timing(v2gt_batch *b, double *ms, int64_t *launches) {
  if (!b)
    return V2GT_ERR_INVALID_ARG;
  for (int k = 0; k < 8; k++) {
    if (ms)
      ms[k] = b->ms[k];
    if (launches)
      launches[k] = b->launches[k];
  }
  return V2GT_OK;
}

int v2gt_batch_get_lm_trace(v2gt_batch *b, int32_t trial, int64_t capacity, double *rows, int64_t *n_out) {
  if (!b || trial < 0 || trial >= b->T || !b->built)
    return V2GT_ERR_INVALID_ARG;
  V2_CUDA_CHECK(cudaSetDevice(b->device));
  V2_TRY(fetch_results(b));
  const int n = b->h_lm[trial].n_trace;
  if (n_out)
    *n_out = n;
  const int m = (int)std::min<int64_t>(n, capacity);
  if (rows && m > 0)
    V2_CUDA_CHECK(cudaMemcpy(rows, b->d.trace + (long long)trial * TRACE_MAX * 5, sizeof(double) * 5 * m, cudaMemcpyDeviceToHost));
  return V2GT_OK;
}

} // extern "C"
