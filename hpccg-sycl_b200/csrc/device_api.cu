// device_api.cu -- hpccg_dev_* entry points of include/hpccg_b200.h: the ELL device mirror, the kernel
// launchers and the device-resident CG loop.  There is NO CPU fallback in this file: every operation is
// a CUDA kernel launch, and every CUDA error is returned to the caller.
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <cmath>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.hpp"
#include "context.hpp"
#include "device_matrix.hpp"
#include "kernels.cuh"
#include "pattern_march.cuh"
#include "persistent_cg.cuh"

namespace hpccg {

// ---- error state -------------------------------------------------------------------------------------
static thread_local std::string t_error;
std::atomic<long long> g_launch_count{0};

void set_error(const std::string &msg) { t_error = msg; }
const char *last_error() { return t_error.c_str(); }

int fail(int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  t_error = buf;
  return code;
}

int fail_cuda(cudaError_t e, const char *what, const char *file, int line) {
  char buf[1024];
  snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  t_error = buf;
  return (int)e;
}

// ---- device info ---------------------------------------------------------------------------------------
// Everything cached here is PER DEVICE (a process may drive several GPUs through hpccg_set_device): the SM count, and
// below the occupancy results and the MaxDynamicSharedMemorySize opt-in, which CUDA keeps per device as well.
constexpr int kMaxDevices = 16;
static Device g_dev[kMaxDevices];
static std::atomic<int> g_dev_known[kMaxDevices];

static int current_device_slot() {
  int id = 0;
  if (cudaGetDevice(&id) != cudaSuccess) {
    cudaGetLastError();
    id = 0;
  }
  return (id >= 0 && id < kMaxDevices) ? id : 0;
}

const Device &device_info() {
  const int id = current_device_slot();
  if (!g_dev_known[id].load(std::memory_order_acquire)) {
    Device d;
    d.id = id;
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, id) == cudaSuccess && sms > 0) d.sm_count = sms;
    else cudaGetLastError();
    g_dev[id] = d;
    g_dev_known[id].store(1, std::memory_order_release);
  }
  return g_dev[id];
}

// one int per device, 0 = not computed yet
struct PerDeviceInt {
  std::atomic<int> v[kMaxDevices];
  int get() { return v[current_device_slot()].load(std::memory_order_relaxed); }
  void set(int x) { v[current_device_slot()].store(x, std::memory_order_relaxed); }
};

int stream_grid(long long work_items, int blocks_per_sm) {
  long long blocks = (work_items + kThreads - 1) / kThreads;
  long long cap = (long long)device_info().sm_count * blocks_per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks > kMaxPartials) blocks = kMaxPartials;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

static inline long long round_up(long long v, long long m) { return (v + m - 1) / m * m; }
static bool aligned16(const void *p) { return (reinterpret_cast<size_t>(p) & 15) == 0; }
static bool aligned32(const void *p) { return (reinterpret_cast<size_t>(p) & 31) == 0; }

// A small per-device workspace for the stand-alone reductions (hpccg_dev_dot, max_abs_diff).
struct GlobalWorkspace {
  double *partials = nullptr;
  unsigned *counter = nullptr;
  std::mutex mu;
};
static GlobalWorkspace g_ws[16];

static int global_workspace(GlobalWorkspace **out) {
  int dev = 0;
  HPCCG_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return fail(HPCCG_ERR_ARG, "device index %d out of range", dev);
  GlobalWorkspace &w = g_ws[dev];
  std::lock_guard<std::mutex> lk(w.mu);
  if (!w.partials) {
    HPCCG_CUDA(cudaMalloc(&w.partials, sizeof(double) * kMaxPartials));
    HPCCG_CUDA(cudaMalloc(&w.counter, sizeof(unsigned)));
    HPCCG_CUDA(cudaMemset(w.counter, 0, sizeof(unsigned)));
  }
  *out = &w;
  return 0;
}

int ensure_solver_workspace(hpccg_dev_matrix *m, int max_iter, int nranks) {
  const long long ncol_pad = round_up(std::max(m->ncol, 2), 512);
  if (!m->r) {
    HPCCG_CUDA(cudaMalloc(&m->r, sizeof(double) * m->npad));
    HPCCG_CUDA(cudaMalloc(&m->Ap, sizeof(double) * m->npad));
    HPCCG_CUDA(cudaMalloc(&m->p, sizeof(double) * ncol_pad));
    HPCCG_CUDA(cudaMemset(m->p, 0, sizeof(double) * ncol_pad));
  }
  if (m->hist_cap < max_iter + 1) {
    if (m->hist) HPCCG_CUDA(cudaFree(m->hist));
    m->hist = nullptr;
    HPCCG_CUDA(cudaMalloc(&m->hist, sizeof(double) * (max_iter + 1)));
    m->hist_cap = max_iter + 1;
  }
  if (m->gathered_cap < nranks) {
    if (m->gathered) HPCCG_CUDA(cudaFree(m->gathered));
    m->gathered = nullptr;
    HPCCG_CUDA(cudaMalloc(&m->gathered, sizeof(double) * std::max(nranks, 8)));
    m->gathered_cap = std::max(nranks, 8);
  }
  if (!m->comm_stream) {
    HPCCG_CUDA(cudaStreamCreateWithFlags(&m->comm_stream, cudaStreamNonBlocking));
    HPCCG_CUDA(cudaEventCreateWithFlags(&m->ev_p_ready, cudaEventDisableTiming));
    HPCCG_CUDA(cudaEventCreateWithFlags(&m->ev_halo_done, cudaEventDisableTiming));
    for (cudaEvent_t &e : m->ev_io) HPCCG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  return 0;
}

int ensure_scratch(hpccg_dev_matrix *m, long long doubles) {
  if (m->scratch_cap >= doubles) return 0;
  if (m->scratch_x) HPCCG_CUDA(cudaFree(m->scratch_x));
  if (m->scratch_y) HPCCG_CUDA(cudaFree(m->scratch_y));
  m->scratch_x = m->scratch_y = nullptr;
  m->scratch_cap = 0;
  HPCCG_CUDA(cudaMalloc(&m->scratch_x, sizeof(double) * doubles));
  HPCCG_CUDA(cudaMalloc(&m->scratch_y, sizeof(double) * doubles));
  m->scratch_cap = doubles;
  return 0;
}

static int alloc_common(hpccg_dev_matrix *m, long long elems = -1) {
  HPCCG_CUDA(cudaGetDevice(&m->device));
  if (elems < 0) elems = (long long)m->slots * m->npad;
  // memory guard: say what is missing instead of failing somewhere inside the repack
  size_t free_b = 0, total_b = 0;
  HPCCG_CUDA(cudaMemGetInfo(&free_b, &total_b));
  const double need = 12.0 * (double)elems + 64.0 * (double)m->npad;  // matrix + the solver's vectors
  if (need > 0.97 * (double)free_b)
    return fail(HPCCG_ERR_ALLOC, "device mirror needs %.1f GB (%lld stored slots of 12 bytes + vectors) but only %.1f GB are free",
                need / 1e9, elems, (double)free_b / 1e9);
  HPCCG_CUDA(cudaMalloc(&m->vals, sizeof(double) * (size_t)std::max<long long>(elems, 1)));
  HPCCG_CUDA(cudaMalloc(&m->cols, sizeof(int) * (size_t)std::max<long long>(elems, 1)));
  HPCCG_CUDA(cudaMalloc(&m->partials, sizeof(double) * kMaxPartials));
  HPCCG_CUDA(cudaMalloc(&m->state, sizeof(CgState)));
  cg_state_init_kernel<<<1, 32>>>(m->state, nullptr);
  count_launch();
  HPCCG_LAUNCH_CHECK();
  return 0;
}

// Buffers a captured solve graph points at are about to change (format switch, new halo plan): drop the graph.
static void invalidate_graph(hpccg_dev_matrix *m) {
  if (m->graph_exec) cudaGraphExecDestroy(m->graph_exec);
  m->graph_exec = nullptr;
  m->graph_b = nullptr;
  m->graph_x = nullptr;
  m->graph_max_iter = 0;
  ++m->generation;
}

// ---- SpMV launch ----------------------------------------------------------------------------------------
template <int SLOTS, int RPT, bool DOT>
static int spmv_occupancy_grid() {
  static PerDeviceInt cache;
  int cached = cache.get();
  if (cached) return cached;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spmv_ell_kernel<SLOTS, RPT, DOT>, kThreads, 0) != cudaSuccess ||
      per_sm < 1)
    per_sm = 2;
  cached = per_sm * device_info().sm_count;
  if (cached > kMaxPartials / 4) cached = kMaxPartials / 4;  // room for interior + two boundary launches
  cache.set(cached);
  return cached;
}

struct SpmvPlan {
  int row_begin, row_end, tiles, grid;
};

template <int RPT>
static SpmvPlan plan_range(int row_begin, int row_end, int max_grid) {
  SpmvPlan p{row_begin, row_end, 0, 0};
  if (row_end <= row_begin) return p;
  const int base = row_begin & ~(RPT - 1);
  const int tile_rows = kThreads * RPT;
  p.tiles = (row_end - base + tile_rows - 1) / tile_rows;
  p.grid = std::min(p.tiles, max_grid);
  return p;
}

template <int SLOTS, int RPT, bool DOT>
static int launch_spmv_t(const hpccg_dev_matrix *m, const double *x, double *y, const SpmvPlan &pl, int partial_offset,
                         int total_partials, const FinishParams &fp, cudaStream_t s) {
  if (pl.grid == 0) return 0;
  spmv_ell_kernel<SLOTS, RPT, DOT><<<pl.grid, kThreads, 0, s>>>(m->vals, m->cols, m->npad, m->slots, x, y, pl.row_begin,
                                                                 pl.row_end, pl.tiles, m->partials, partial_offset,
                                                                 total_partials, &m->state->counter, fp);
  count_launch();
  HPCCG_LAUNCH_CHECK();
  return 0;
}

template <bool DOT>
static int spmv_max_grid(int slots) {
  if (slots == 27) return spmv_occupancy_grid<27, 2, DOT>();
  if (slots == 7) return spmv_occupancy_grid<7, 2, DOT>();
  return spmv_occupancy_grid<0, 2, DOT>();
}

template <bool DOT>
static int launch_spmv_reg(const hpccg_dev_matrix *m, const double *x, double *y, const SpmvPlan &pl, int partial_offset,
                           int total_partials, const FinishParams &fp, cudaStream_t s) {
  if (m->slots == 27) return launch_spmv_t<27, 2, DOT>(m, x, y, pl, partial_offset, total_partials, fp, s);
  if (m->slots == 7) return launch_spmv_t<7, 2, DOT>(m, x, y, pl, partial_offset, total_partials, fp, s);
  return launch_spmv_t<0, 2, DOT>(m, x, y, pl, partial_offset, total_partials, fp, s);
}

// ---- TMA path (27 and 7 slots): stages of SPS slices, persistent CTAs ------------------------------------
// HPCCG_B200_SPMV=reg forces the register path (A/B measurements).
static bool use_tma_path(int slots) {
  static int mode = -1;
  if (mode < 0) {
    const char *e = std::getenv("HPCCG_B200_SPMV");
    mode = (e && std::string(e) == "reg") ? 0 : 1;
  }
  return mode == 1 && (slots == 27 || slots == 7);
}

// L2 prefetch distance of the TMA SpMV in stages (0 = off); HPCCG_B200_L2_AHEAD overrides for A/B runs.
static int tma_l2_ahead() {
  static int v = -1;
  if (v < 0) {
    const char *e = std::getenv("HPCCG_B200_L2_AHEAD");
    v = e ? std::atoi(e) : 0;
    if (v < 0 || v > 64) v = 0;
  }
  return v;
}

template <int SLOTS, int SPS, int NSTAGES, bool DOT>
static int tma_ctas_per_sm() {
  static PerDeviceInt cache;
  int cached = cache.get();
  if (cached) return cached;
  using Cfg = SpmvTmaCfg<SLOTS, SPS, NSTAGES>;
  auto kern = spmv_sell_tma_kernel<SLOTS, SPS, NSTAGES, DOT>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess) return -1;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, Cfg::kRows, Cfg::kSmemBytes) != cudaSuccess || per_sm < 1)
    return -1;
  cached = per_sm;
  cache.set(cached);
  return cached;
}

template <int SLOTS, int SPS, int NSTAGES, bool DOT>
static SpmvPlan plan_tma(int row_begin, int row_end) {
  SpmvPlan p{row_begin, row_end, 0, 0};
  if (row_end <= row_begin) return p;
  constexpr int rows = SPS * kSliceRows;
  const int sb = row_begin / rows, se = (row_end + rows - 1) / rows;
  p.tiles = se - sb;
  const int per_sm = tma_ctas_per_sm<SLOTS, SPS, NSTAGES, DOT>();
  p.grid = std::min(p.tiles, std::max(1, per_sm) * device_info().sm_count);
  return p;
}

template <int SLOTS, int SPS, int NSTAGES, bool DOT>
static int launch_tma_t(const hpccg_dev_matrix *m, const double *x, double *y, const SpmvPlan &pl, int partial_offset,
                        int total_partials, const FinishParams &fp, cudaStream_t s, const SpmvHalo &halo) {
  if (pl.grid == 0) return 0;
  using Cfg = SpmvTmaCfg<SLOTS, SPS, NSTAGES>;
  if (tma_ctas_per_sm<SLOTS, SPS, NSTAGES, DOT>() < 1)
    return fail(HPCCG_ERR_STATE, "TMA SpMV kernel cannot be resident (%d bytes of shared memory)", Cfg::kSmemBytes);
  constexpr int rows = Cfg::kRows;
  const int sb = pl.row_begin / rows;
  spmv_sell_tma_kernel<SLOTS, SPS, NSTAGES, DOT><<<pl.grid, rows, Cfg::kSmemBytes, s>>>(
      m->vals, m->cols, x, y, pl.row_begin, pl.row_end, sb, sb + pl.tiles, m->partials, partial_offset, total_partials,
      &m->state->counter, fp, halo, tma_l2_ahead());
  count_launch();
  HPCCG_LAUNCH_CHECK();
  return 0;
}

// Stage shapes.  27 slots: 2 slices (256 rows, 81 KB) x 2 stages, 1 CTA/SM (variant 0, default);
// 1 slice x 2 stages, 2 CTAs/SM (variant 1); 1 slice x 5 stages, 1 CTA/SM (variant 2) -- HPCCG_B200_TMA_VARIANT
// selects, for the A/B measurements recorded in DESIGN.md.  7 slots: 2 slices (21 KB) x 4 stages, 2 CTAs/SM.
static int tma_variant() {
  static int v = -1;
  if (v < 0) {
    const char *e = std::getenv("HPCCG_B200_TMA_VARIANT");
    v = e ? std::atoi(e) : 0;
    if (v < 0 || v > 2) v = 0;
  }
  return v;
}

#define HPCCG_TMA_DISPATCH(FN, DOT, ...)                                   \
  do {                                                                     \
    if (m->slots == 7) {                                                   \
      switch (tma_variant()) {                                             \
        case 1: return FN<7, 2, 8, DOT>(__VA_ARGS__);                      \
        case 2: return FN<7, 4, 4, DOT>(__VA_ARGS__);                      \
        default: return FN<7, 2, 4, DOT>(__VA_ARGS__);                     \
      }                                                                    \
    }                                                                      \
    switch (tma_variant()) {                                               \
      case 1: return FN<27, 1, 2, DOT>(__VA_ARGS__);                       \
      case 2: return FN<27, 1, 5, DOT>(__VA_ARGS__);                       \
      default: return FN<27, 2, 2, DOT>(__VA_ARGS__);                      \
    }                                                                      \
  } while (0)

// ---- pattern-coded format: tiles of kThreads rows, persistent CTAs ------------------------------------------------
template <int SLOTS, bool DOT>
static int pattern_ctas_per_sm() {
  static PerDeviceInt cache;
  int cached = cache.get();
  if (cached) return cached;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spmv_pattern_kernel<SLOTS, DOT>, kThreads, 0) != cudaSuccess || per_sm < 1)
    per_sm = 2;
  cached = per_sm;
  cache.set(cached);
  return cached;
}

template <int SLOTS, bool DOT>
static SpmvPlan plan_pattern(int row_begin, int row_end) {
  SpmvPlan p{row_begin, row_end, 0, 0};
  if (row_end <= row_begin) return p;
  const int tb = row_begin / kThreads, te = (row_end + kThreads - 1) / kThreads;
  p.tiles = te - tb;
  p.grid = std::min(p.tiles, std::min(pattern_ctas_per_sm<SLOTS, DOT>() * device_info().sm_count, kMaxPartials / 4));
  return p;
}

template <int SLOTS, bool DOT>
static int launch_pattern_t(const hpccg_dev_matrix *m, const double *x, double *y, const SpmvPlan &pl, int partial_offset,
                            int total_partials, const FinishParams &fp, cudaStream_t s, const SpmvHalo &halo) {
  if (pl.grid == 0) return 0;
  const int tb = pl.row_begin / kThreads;
  spmv_pattern_kernel<SLOTS, DOT><<<pl.grid, kThreads, 0, s>>>(m->pat_id, m->pat_val, m->pat_delta, m->pat_len, m->pattern0, x, y,
                                                               m->n, pl.row_begin, pl.row_end, tb, tb + pl.tiles, m->partials,
                                                               partial_offset, total_partials, &m->state->counter, fp, halo);
  count_launch();
  HPCCG_LAUNCH_CHECK();
  return 0;
}

// ---- pattern-coded format, stencil-structured pattern 0: z-marching kernel (pattern_march.cuh) ---------------------
// HPCCG_B200_PATTERN=classic keeps the one-gather-per-entry kernel (A/B).
static bool use_march(const hpccg_dev_matrix *m, const double *x) {
  static int mode = -1;
  if (mode < 0) {
    const char *e = std::getenv("HPCCG_B200_PATTERN");
    mode = (e && std::string(e) == "classic") ? 0 : 1;
  }
  return mode == 1 && m->format == 1 && m->march.ok && aligned32(x);
}

template <class K>
static int march_grid(K kern, long long steps, PerDeviceInt &cache) {
  int per_sm = cache.get();
  if (!per_sm) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    cache.set(per_sm);
  }
  return (int)std::min<long long>(steps, std::min(per_sm * device_info().sm_count, kMaxPartials / 4));
}

template <int SLOTS, bool DOT, bool NEG1>
static int launch_march_t(const hpccg_dev_matrix *m, const double *x, double *y, int partial_offset, const FinishParams &fp,
                          cudaStream_t s, const SpmvHalo &halo) {
  const MarchGeom &g = m->march;
  static PerDeviceInt cache;
  const int grid = march_grid(spmv_pattern_march_kernel<SLOTS, DOT, NEG1>, (long long)g.cols_x * g.cols_y * g.nz, cache);
  spmv_pattern_march_kernel<SLOTS, DOT, NEG1><<<grid, kThreads, 0, s>>>(
      m->pat_id, m->pat_desc, m->pat_val, m->pat_delta, m->pat_len, m->pattern0, g, x, y, m->n, m->interior_begin,
      m->interior_end, m->partials, partial_offset, grid, &m->state->counter, fp, halo);
  count_launch();
  HPCCG_LAUNCH_CHECK();
  return 0;
}

// whole-matrix launch (the marching kernels have no row-range form)
template <bool DOT>
static int launch_march(const hpccg_dev_matrix *m, const double *x, double *y, int partial_offset, const FinishParams &fp,
                        cudaStream_t s, const SpmvHalo &halo) {
  const bool neg1 = m->march.neg1 != 0;
  if (m->slots == 7)
    return neg1 ? launch_march_t<7, DOT, true>(m, x, y, partial_offset, fp, s, halo) : launch_march_t<7, DOT, false>(m, x, y, partial_offset, fp, s, halo);
  return neg1 ? launch_march_t<27, DOT, true>(m, x, y, partial_offset, fp, s, halo) : launch_march_t<27, DOT, false>(m, x, y, partial_offset, fp, s, halo);
}

#define HPCCG_PATTERN_DISPATCH(FN, DOT, ...)             \
  do {                                                   \
    if (m->slots == 7) return FN<7, DOT>(__VA_ARGS__);   \
    return FN<27, DOT>(__VA_ARGS__);                     \
  } while (0)

// ---- ragged SELL-C-sigma (format 2): one thread per row position ----------------------------------------------------------
template <bool DOT>
static SpmvPlan plan_ragged(const hpccg_dev_matrix *m, int row_begin, int row_end) {
  SpmvPlan p{row_begin, row_end, 0, 0};
  if (row_end <= row_begin) return p;
  // with sigma-sorting a row lives anywhere in its window: whole-matrix launches only (the caller never splits then)
  const int base = row_begin & ~(kRaggedRows - 1);
  const int end = m->perm ? (int)m->npad : row_end;
  p.tiles = (end - base + kThreads - 1) / kThreads;
  p.grid = std::min(p.tiles, std::min(8 * device_info().sm_count, kMaxPartials / 4));
  return p;
}

template <bool DOT>
static int launch_ragged(const hpccg_dev_matrix *m, const double *x, double *y, const SpmvPlan &pl, int partial_offset,
                         int total_partials, const FinishParams &fp, cudaStream_t s) {
  if (pl.grid == 0) return 0;
  if (m->perm && (pl.row_begin != 0 || pl.row_end != m->n))
    return fail(HPCCG_ERR_STATE, "a sigma-sorted mirror has no contiguous row ranges");
  const int pos_end = m->perm ? (int)m->npad : pl.row_end;
  spmv_sell_ragged_kernel<DOT><<<pl.grid, kThreads, 0, s>>>(m->vals, m->cols, m->slice_slots, m->slice_off, m->perm, m->n, x, y,
                                                            pl.row_begin, pos_end, m->partials, partial_offset, total_partials,
                                                            &m->state->counter, fp);
  count_launch();
  HPCCG_LAUNCH_CHECK();
  return 0;
}

template <bool DOT>
static SpmvPlan plan_spmv(const hpccg_dev_matrix *m, int row_begin, int row_end) {
  if (m->format == 2) return plan_ragged<DOT>(m, row_begin, row_end);
  if (m->format == 1) HPCCG_PATTERN_DISPATCH(plan_pattern, DOT, row_begin, row_end);
  if (use_tma_path(m->slots)) HPCCG_TMA_DISPATCH(plan_tma, DOT, row_begin, row_end);
  return plan_range<2>(row_begin, row_end, spmv_max_grid<DOT>(m->slots));
}

template <bool DOT>
static int launch_spmv(const hpccg_dev_matrix *m, const double *x, double *y, const SpmvPlan &pl, int partial_offset,
                       int total_partials, const FinishParams &fp, cudaStream_t s, const SpmvHalo &halo = SpmvHalo{}) {
  if (m->format == 2) {
    if (halo.link) return fail(HPCCG_ERR_STATE, "peer-memory halo wait needs the TMA or pattern SpMV path");
    return launch_ragged<DOT>(m, x, y, pl, partial_offset, total_partials, fp, s);
  }
  if (m->format == 1) HPCCG_PATTERN_DISPATCH(launch_pattern_t, DOT, m, x, y, pl, partial_offset, total_partials, fp, s, halo);
  if (use_tma_path(m->slots)) HPCCG_TMA_DISPATCH(launch_tma_t, DOT, m, x, y, pl, partial_offset, total_partials, fp, s, halo);
  if (halo.link) return fail(HPCCG_ERR_STATE, "peer-memory halo wait needs the TMA SpMV path");
  return launch_spmv_reg<DOT>(m, x, y, pl, partial_offset, total_partials, fp, s);
}

// 256-bit accesses in the loop's vector kernels unless HPCCG_B200_VEC=2 (A/B) or a caller's vector is only 16-byte aligned
static bool use_vec4(const void *a, const void *b, const void *c) {
  static int mode = -1;
  if (mode < 0) {
    const char *e = std::getenv("HPCCG_B200_VEC");
    mode = (e && std::atoi(e) == 2) ? 2 : 4;
  }
  return mode == 4 && aligned32(a) && aligned32(b) && aligned32(c);
}

// HPCCG_B200_VEC_TMA="tile,stages" (doubles per tile, ring depth; "0" = plain 256-bit LDG/STG kernels): the loop's two vector
// kernels stream through shared memory with cp.async.bulk (vec_stream_tma_kernel)
struct VecTmaMode {
  int tile = 0, stages = 0;
};
static const VecTmaMode &vec_tma() {
  static VecTmaMode mode;
  static bool init = false;
  if (!init) {
    init = true;
    mode.tile = 2048;
    mode.stages = 3;
    if (const char *e = std::getenv("HPCCG_B200_VEC_TMA")) {
      int t = 0, st = 0;
      const int got = std::sscanf(e, "%d,%d", &t, &st);
      if (got >= 1) {
        mode.tile = t;
        mode.stages = got >= 2 ? st : 3;
      }
    }
  }
  return mode;
}

template <class OP, int TILE, int NSTAGES, bool PUT>
static int launch_vec_tma_t(hpccg_dev_matrix *m, const VecPtrs &vp, const FinishParams &fp, const HaloPut *put, cudaStream_t s) {
  using Cfg = VecTmaCfg<OP, TILE, NSTAGES>;
  auto kern = vec_stream_tma_kernel<OP, TILE, NSTAGES, PUT>;
  static PerDeviceInt cache;
  int per_sm = cache.get();
  if (!per_sm) {
    HPCCG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, Cfg::kSmemBytes) != cudaSuccess || per_sm < 1)
      return fail(HPCCG_ERR_STATE, "TMA vector kernel cannot be resident (%d bytes of shared memory)", Cfg::kSmemBytes);
    cache.set(per_sm);
  }
  const int grid = std::max(1, std::min(per_sm * device_info().sm_count, std::min(m->n / TILE, kMaxPartials)));
  kern<<<grid, kThreads, Cfg::kSmemBytes, s>>>(m->n, m->state, vp, m->partials, grid, &m->state->counter, fp, put ? *put : HaloPut{});
  count_launch();
  HPCCG_LAUNCH_CHECK();
  return 0;
}

template <class OP>
static int launch_vec_tma(hpccg_dev_matrix *m, const VecPtrs &vp, const FinishParams &fp, const HaloPut *put, cudaStream_t s) {
  const VecTmaMode &v = vec_tma();
  // (only the p-producing operation has a put form)
#define HPCCG_VEC_CASE(T, S)                                                                        \
  if (v.tile == T && v.stages == S && VecTmaCfg<OP, T, S>::kSmemBytes <= 227 * 1024) {              \
    if constexpr (OP::kOut == 2) {                                                                  \
      if (put) return launch_vec_tma_t<OP, T, S, true>(m, vp, fp, put, s);                          \
    }                                                                                               \
    return launch_vec_tma_t<OP, T, S, false>(m, vp, fp, nullptr, s);                                \
  }
  HPCCG_VEC_CASE(1024, 2)
  HPCCG_VEC_CASE(1024, 3)
  HPCCG_VEC_CASE(1024, 4)
  HPCCG_VEC_CASE(2048, 2)
  HPCCG_VEC_CASE(2048, 3)
  HPCCG_VEC_CASE(4096, 2)
#undef HPCCG_VEC_CASE
  // a shape this operation has no room for: the nearest that fits
  if constexpr (OP::kOut == 2) {
    if (put) return launch_vec_tma_t<OP, 1024, 3, true>(m, vp, fp, put, s);
  }
  return launch_vec_tma_t<OP, 1024, 3, false>(m, vp, fp, nullptr, s);
}

// Whole-matrix SpMV (+ optional fused x.y) in one launch.
static int spmv_full(const hpccg_dev_matrix *m, const double *x, double *y, bool dot, const FinishParams &fp,
                     cudaStream_t s, const SpmvHalo &halo = SpmvHalo{}) {
  if (!aligned16(x) || !aligned16(y)) return fail(HPCCG_ERR_ARG, "SpMV vectors must be 16-byte aligned");
  if (use_march(m, x) && aligned32(y)) {
    // the 256-bit loads may start up to 3 elements before the last needed one: ncol % 4 == 0 keeps them inside ncol
    // (checked when the geometry was accepted), the solver's own p is padded anyway
    return dot ? launch_march<true>(m, x, y, 0, fp, s, halo) : launch_march<false>(m, x, y, 0, fp, s, halo);
  }
  if (dot) {
    SpmvPlan pl = plan_spmv<true>(m, 0, m->n);
    return launch_spmv<true>(m, x, y, pl, 0, pl.grid, fp, s, halo);
  }
  SpmvPlan pl = plan_spmv<false>(m, 0, m->n);
  return launch_spmv<false>(m, x, y, pl, 0, 0, fp, s, halo);
}

static FinishParams fin_store(double *out) {
  FinishParams fp{};
  fp.mode = FIN_STORE;
  fp.out = out;
  return fp;
}

// ---- ragged SELL-C-sigma mirror (format 2) from assembled host rows ----------------------------------------------------
// HPCCG_B200_RAGGED=0 keeps the uniform-slot layout for every matrix (A/B); HPCCG_B200_SIGMA=<rows> sets the sorting window
// (1 = no sorting).  Default 8192 rows = 256 slices: on a power-law matrix 1.15 x the stored entries (4096: 1.28 x, 1024: 1.56 x).
static int ragged_sigma(bool has_halo) {
  if (has_halo) return 1;  // halo-touching row RANGES must stay contiguous
  int v = 8192;
  if (const char *e = std::getenv("HPCCG_B200_SIGMA")) v = std::atoi(e);
  return std::max(1, std::min(v, 1 << 24));
}

static int build_ragged(hpccg_dev_matrix *m, const int *nnz_in_row, const double *const *ptr_to_vals_in_row,
                        const int *const *ptr_to_inds_in_row, unsigned nthreads) {
  const int n = m->n;
  const long long npos = m->npad, nslices = npos / kRaggedRows;
  m->sigma = ragged_sigma(m->ncol > m->n);
  // row of every position: identity, or -- per window of sigma rows -- rows by decreasing length (stable)
  std::vector<int> perm(npos);
  for (long long i = 0; i < npos; ++i) perm[i] = i < n ? (int)i : -1;
  if (m->sigma > 1) {
    const long long win = std::max<long long>(m->sigma, kRaggedRows);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nthreads; ++t)
      th.emplace_back([&, t] {
        const long long nwin = (n + win - 1) / win;
        for (long long w = nwin * t / nthreads; w < nwin * (t + 1) / nthreads; ++w) {
          const long long lo = w * win, hi = std::min<long long>(n, lo + win);
          std::stable_sort(perm.begin() + lo, perm.begin() + hi, [&](int a, int b) { return nnz_in_row[a] > nnz_in_row[b]; });
        }
      });
    for (auto &t : th) t.join();
  }
  std::vector<int> slots(nslices, 0);
  std::vector<long long> off(nslices + 1, 0);
  for (long long sl = 0; sl < nslices; ++sl) {
    int mx = 0;
    for (int l = 0; l < kRaggedRows; ++l) {
      const int row = perm[sl * kRaggedRows + l];
      if (row >= 0) mx = std::max(mx, nnz_in_row[row]);
    }
    slots[sl] = mx;
    off[sl + 1] = off[sl] + (long long)mx * kRaggedRows;
  }
  m->total_elems = off[nslices];
  HPCCG_TRY(alloc_common(m, m->total_elems));
  HPCCG_CUDA(cudaMalloc(&m->slice_slots, sizeof(int) * nslices));
  HPCCG_CUDA(cudaMalloc(&m->slice_off, sizeof(long long) * (nslices + 1)));
  HPCCG_CUDA(cudaMemcpy(m->slice_slots, slots.data(), sizeof(int) * nslices, cudaMemcpyHostToDevice));
  HPCCG_CUDA(cudaMemcpy(m->slice_off, off.data(), sizeof(long long) * (nslices + 1), cudaMemcpyHostToDevice));
  if (m->sigma > 1) {
    HPCCG_CUDA(cudaMalloc(&m->perm, sizeof(int) * npos));
    HPCCG_CUDA(cudaMemcpy(m->perm, perm.data(), sizeof(int) * npos, cudaMemcpyHostToDevice));
  }
  // repack through one staging buffer per array, runs of whole slices at a time (a run is contiguous on the device)
  long long cap = 1 << 22;
  for (long long sl = 0; sl < nslices; ++sl) cap = std::max(cap, off[sl + 1] - off[sl]);
  std::vector<double> sv(cap);
  std::vector<int> sc(cap);
  long long s0 = 0;
  while (s0 < nslices) {
    long long s1 = s0 + 1;
    while (s1 < nslices && off[s1 + 1] - off[s0] <= cap) ++s1;
    const long long e0 = off[s0], cnt = off[s1] - e0;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nthreads; ++t)
      th.emplace_back([&, t] {
        for (long long sl = s0 + (s1 - s0) * t / nthreads; sl < s0 + (s1 - s0) * (t + 1) / nthreads; ++sl)
          for (int l = 0; l < kRaggedRows; ++l) {
            const int row = perm[sl * kRaggedRows + l];
            const int nnz = row >= 0 ? nnz_in_row[row] : 0;
            const long long base = off[sl] - e0 + l;
            for (int j = 0; j < nnz; ++j) {
              sv[base + (long long)j * kRaggedRows] = ptr_to_vals_in_row[row][j];
              sc[base + (long long)j * kRaggedRows] = ptr_to_inds_in_row[row][j];
            }
            for (int j = nnz; j < slots[sl]; ++j) {
              sv[base + (long long)j * kRaggedRows] = 0.0;
              sc[base + (long long)j * kRaggedRows] = -1;
            }
          }
      });
    for (auto &t : th) t.join();
    if (cnt > 0) {
      HPCCG_CUDA(cudaMemcpy(m->vals + e0, sv.data(), sizeof(double) * cnt, cudaMemcpyHostToDevice));
      HPCCG_CUDA(cudaMemcpy(m->cols + e0, sc.data(), sizeof(int) * cnt, cudaMemcpyHostToDevice));
    }
    s0 = s1;
  }
  m->format = 2;
  return 0;
}

}  // namespace hpccg

using namespace hpccg;

// ================================================================================================
// C-ABI: library / device
// ================================================================================================
extern "C" {

const char *hpccg_last_error(void) { return last_error(); }
int hpccg_version(void) { return 100; }
long long hpccg_launch_count(void) { return g_launch_count.load(); }

int hpccg_device_count(int *count) {
  HPCCG_CUDA(cudaGetDeviceCount(count));
  return 0;
}
int hpccg_set_device(int device) {
  HPCCG_CUDA(cudaSetDevice(device));
  return 0;
}
int hpccg_device_synchronize(void) {
  HPCCG_CUDA(cudaDeviceSynchronize());
  return 0;
}
int hpccg_dev_malloc(void **ptr, long long bytes) {
  HPCCG_CUDA(cudaMalloc(ptr, (size_t)std::max<long long>(bytes, 16)));
  return 0;
}
int hpccg_dev_free(void *ptr) {
  HPCCG_CUDA(cudaFree(ptr));
  return 0;
}
int hpccg_host_malloc_pinned(void **ptr, long long bytes) {
  HPCCG_CUDA(cudaMallocHost(ptr, (size_t)std::max<long long>(bytes, 16)));
  return 0;
}
int hpccg_host_free_pinned(void *ptr) {
  HPCCG_CUDA(cudaFreeHost(ptr));
  return 0;
}
int hpccg_memcpy_h2d(void *dst, const void *src, long long bytes, void *stream) {
  HPCCG_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return 0;
}
int hpccg_memcpy_d2h(void *dst, const void *src, long long bytes, void *stream) {
  HPCCG_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return 0;
}
int hpccg_stream_synchronize(void *stream) {
  HPCCG_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

// ================================================================================================
// Device matrix
// ================================================================================================
int hpccg_dev_matrix_create(int local_nrow, int local_ncol, const int *nnz_in_row,
                            const double *const *ptr_to_vals_in_row, const int *const *ptr_to_inds_in_row,
                            hpccg_dev_matrix **out) {
  if (!out || local_nrow <= 0 || local_ncol < local_nrow || !nnz_in_row || !ptr_to_vals_in_row || !ptr_to_inds_in_row)
    return fail(HPCCG_ERR_ARG, "hpccg_dev_matrix_create: bad arguments");
  const int n = local_nrow;
  // slot count = longest row; first/last rows that touch halo columns bound the interior range
  const unsigned nthreads = std::max(1u, std::min(std::thread::hardware_concurrency(), 16u));
  std::vector<int> tmax(nthreads, 0), tfirst(nthreads, n), tlast(nthreads, -1), tbad(nthreads, 0);
  std::vector<long long> tsum(nthreads, 0);
  {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nthreads; ++t)
      th.emplace_back([&, t] {
        const long long lo = (long long)n * t / nthreads, hi = (long long)n * (t + 1) / nthreads;
        for (long long i = lo; i < hi; ++i) {
          const int nnz = nnz_in_row[i];
          if (nnz > tmax[t]) tmax[t] = nnz;
          tsum[t] += nnz;
          const int *ci = ptr_to_inds_in_row[i];
          for (int j = 0; j < nnz; ++j) {
            if (ci[j] < 0 || ci[j] >= local_ncol) tbad[t] = 1;
            if (ci[j] >= n) {
              if (i < tfirst[t]) tfirst[t] = (int)i;
              if (i > tlast[t]) tlast[t] = (int)i;
            }
          }
        }
      });
    for (auto &t : th) t.join();
  }
  int slots = 1;
  bool bad = false;
  long long stored = 0;
  for (unsigned t = 0; t < nthreads; ++t) {
    slots = std::max(slots, tmax[t]);
    bad = bad || tbad[t];
    stored += tsum[t];
  }
  if (bad) return fail(HPCCG_ERR_ARG, "hpccg_dev_matrix_create: column index outside [0, local_ncol) -- run make_local_matrix first");

  hpccg_dev_matrix *m = new hpccg_dev_matrix();
  m->n = n;
  m->ncol = local_ncol;
  m->slots = slots;
  m->npad = round_up(n, kRowPad);
  // halo-touching rows: leading run [0,a) and trailing run [b,n); a row in the first half extends a,
  // one in the second half lowers b
  int a = 0, b = n;
  {
    // a second, cheap pass over the per-thread extrema is not enough to split the two runs, so rescan
    // only when a halo exists at all
    bool any = false;
    for (unsigned t = 0; t < nthreads; ++t) any = any || tlast[t] >= 0;
    if (any) {
      const int half = n / 2;
      std::vector<int> ta(nthreads, 0), tb(nthreads, n);
      std::vector<std::thread> th;
      for (unsigned t = 0; t < nthreads; ++t)
        th.emplace_back([&, t] {
          const long long lo = (long long)n * t / nthreads, hi = (long long)n * (t + 1) / nthreads;
          for (long long i = lo; i < hi; ++i) {
            const int nnz = nnz_in_row[i];
            const int *ci = ptr_to_inds_in_row[i];
            bool ext = false;
            for (int j = 0; j < nnz; ++j) ext = ext || ci[j] >= n;
            if (!ext) continue;
            if (i < half) ta[t] = std::max(ta[t], (int)i + 1);
            else tb[t] = std::min(tb[t], (int)i);
          }
        });
      for (auto &t : th) t.join();
      for (unsigned t = 0; t < nthreads; ++t) {
        a = std::max(a, ta[t]);
        b = std::min(b, tb[t]);
      }
    }
  }
  m->interior_begin = a;
  m->interior_end = std::max(a, b);

  // Rows of different lengths (file matrices, thin blocks): SELL-C-sigma proper -- per-slice slot counts, rows sorted by
  // length inside windows of sigma rows -- instead of padding every row to the longest one, when that padding would exceed
  // 10 % of the stored entries.  The 27- / 7-slot stencil matrices keep the uniform layout: that is what the TMA kernels and
  // the pattern encoder read, and their padding is 0.6 %.
  {
    const char *e = std::getenv("HPCCG_B200_RAGGED");
    const bool want = e ? e[0] == '1' : (slots != 27 && slots != 7 && 10 * stored < 9 * (long long)slots * n);
    if (want && !(e && e[0] == '0')) {
      int rc = build_ragged(m, nnz_in_row, ptr_to_vals_in_row, ptr_to_inds_in_row, nthreads);
      if (rc) {
        hpccg_dev_matrix_destroy(m);
        return rc;
      }
      *out = m;
      return 0;
    }
  }

  int rc = alloc_common(m);
  if (rc) {
    hpccg_dev_matrix_destroy(m);
    return rc;
  }

  // Repack rows into the SELL-C layout through two pinned staging buffers.  A chunk of whole slices is one
  // contiguous block of the device arrays, so each chunk is two plain 1-D copies.
  const int chunk = (int)std::min<long long>(m->npad, 1 << 18);  // rows per staging chunk (a multiple of kRowPad)
  double *hv[2] = {nullptr, nullptr};
  int *hc[2] = {nullptr, nullptr};
  cudaStream_t cs = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  auto cleanup = [&] {
    for (int i = 0; i < 2; ++i) {
      if (hv[i]) cudaFreeHost(hv[i]);
      if (hc[i]) cudaFreeHost(hc[i]);
      if (ev[i]) cudaEventDestroy(ev[i]);
    }
    if (cs) cudaStreamDestroy(cs);
  };
#define HPCCG_CUDA_CLEAN(call)                                         \
  do {                                                                 \
    cudaError_t e_ = (call);                                           \
    if (e_ != cudaSuccess) {                                           \
      cleanup();                                                       \
      hpccg_dev_matrix_destroy(m);                                     \
      return fail_cuda(e_, #call, __FILE__, __LINE__);                 \
    }                                                                  \
  } while (0)
  HPCCG_CUDA_CLEAN(cudaStreamCreate(&cs));
  for (int i = 0; i < 2; ++i) {
    HPCCG_CUDA_CLEAN(cudaMallocHost(&hv[i], sizeof(double) * (size_t)slots * chunk));
    HPCCG_CUDA_CLEAN(cudaMallocHost(&hc[i], sizeof(int) * (size_t)slots * chunk));
    HPCCG_CUDA_CLEAN(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
  }
  int buf = 0;
  for (long long r0 = 0; r0 < m->npad; r0 += chunk, buf ^= 1) {
    const int rows = (int)std::min<long long>(chunk, m->npad - r0);
    HPCCG_CUDA_CLEAN(cudaEventSynchronize(ev[buf]));  // staging buffer free again
    double *sv = hv[buf];
    int *sc = hc[buf];
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nthreads; ++t)
      th.emplace_back([&, t] {
        const int lo = (int)((long long)rows * t / nthreads), hi = (int)((long long)rows * (t + 1) / nthreads);
        for (int i = lo; i < hi; ++i) {
          const long long row = r0 + i;
          int nnz = 0;
          const double *cv = nullptr;
          const int *ci = nullptr;
          if (row < n) {
            nnz = nnz_in_row[row];
            cv = ptr_to_vals_in_row[row];
            ci = ptr_to_inds_in_row[row];
          }
          for (int j = 0; j < nnz; ++j) {
            const long long o = sell_offset(i, j, slots);  // chunk-relative: r0 is slice-aligned
            sv[o] = cv[j];
            sc[o] = ci[j];
          }
          for (int j = nnz; j < slots; ++j) {
            const long long o = sell_offset(i, j, slots);
            sv[o] = 0.0;
            sc[o] = -1;
          }
        }
      });
    for (auto &t : th) t.join();
    HPCCG_CUDA_CLEAN(cudaMemcpyAsync(m->vals + r0 * slots, sv, sizeof(double) * (size_t)rows * slots, cudaMemcpyHostToDevice, cs));
    HPCCG_CUDA_CLEAN(cudaMemcpyAsync(m->cols + r0 * slots, sc, sizeof(int) * (size_t)rows * slots, cudaMemcpyHostToDevice, cs));
    HPCCG_CUDA_CLEAN(cudaEventRecord(ev[buf], cs));
  }
  HPCCG_CUDA_CLEAN(cudaStreamSynchronize(cs));
#undef HPCCG_CUDA_CLEAN
  cleanup();
  *out = m;
  return 0;
}

int hpccg_dev_matrix_generate(int nx, int ny, int nz, int rank, int size, int stencil, const int *lower_plane_to_local,
                              const int *upper_plane_to_local, int local_ncol, hpccg_dev_matrix **out) {
  if (!out || nx <= 0 || ny <= 0 || nz <= 0 || size <= 0 || rank < 0 || rank >= size || (stencil != 27 && stencil != 7))
    return fail(HPCCG_ERR_ARG, "hpccg_dev_matrix_generate: bad arguments");
  const long long n = (long long)nx * ny * nz;
  if (n > 2147483647LL - 1024) return fail(HPCCG_ERR_ARG, "local_nrow %lld does not fit int32", n);
  const long long plane = (long long)nx * ny;
  const bool has_lower = rank > 0, has_upper = rank < size - 1;
  if ((has_lower && !lower_plane_to_local) || (has_upper && !upper_plane_to_local))
    return fail(HPCCG_ERR_ARG, "hpccg_dev_matrix_generate: neighbour plane tables missing");
  hpccg_dev_matrix *m = new hpccg_dev_matrix();
  m->n = (int)n;
  m->ncol = local_ncol;
  // slot count = longest row of THIS rank's block (27 / 7 for an interior point, fewer for thin blocks)
  auto span = [](int len, bool lo_open, bool hi_open) {
    // max number of in-range offsets {-1,0,1} over positions of a 1-D extent
    int best = 1;
    for (int i = 0; i < len && i < 3; ++i) {
      for (int pos : {i, len - 1 - i}) {
        int c = 1 + ((pos > 0 || lo_open) ? 1 : 0) + ((pos < len - 1 || hi_open) ? 1 : 0);
        best = std::max(best, c);
      }
    }
    return best;
  };
  const int cx = span(nx, false, false), cy = span(ny, false, false), cz = span(nz, has_lower, has_upper);
  m->slots = stencil == 27 ? cx * cy * cz : 1 + (cx - 1) + (cy - 1) + (cz - 1);
  m->npad = round_up(n, kRowPad);
  int a = has_lower ? (int)plane : 0, b = has_upper ? (int)(n - plane) : (int)n;
  if (a > b) a = b = (int)n;
  m->interior_begin = a;
  m->interior_end = b;
  int rc = alloc_common(m);
  if (rc) {
    hpccg_dev_matrix_destroy(m);
    return rc;
  }
  int *d_lower = nullptr, *d_upper = nullptr;
  auto fail_clean = [&](int code) {
    if (d_lower) cudaFree(d_lower);
    if (d_upper) cudaFree(d_upper);
    hpccg_dev_matrix_destroy(m);
    return code;
  };
  if (has_lower) {
    cudaError_t e = cudaMalloc(&d_lower, sizeof(int) * plane);
    if (e == cudaSuccess) e = cudaMemcpy(d_lower, lower_plane_to_local, sizeof(int) * plane, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return fail_clean(fail_cuda(e, "lower plane table", __FILE__, __LINE__));
  }
  if (has_upper) {
    cudaError_t e = cudaMalloc(&d_upper, sizeof(int) * plane);
    if (e == cudaSuccess) e = cudaMemcpy(d_upper, upper_plane_to_local, sizeof(int) * plane, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return fail_clean(fail_cuda(e, "upper plane table", __FILE__, __LINE__));
  }
  generate_ell_kernel<<<stream_grid(m->npad, 16), kThreads>>>(nx, ny, nz, n * rank, n * size, stencil == 7 ? 1 : 0,
                                                               m->slots, m->npad, d_lower, d_upper, m->vals, m->cols);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return fail_clean(fail_cuda(e, "generate_ell_kernel", __FILE__, __LINE__));
  if (d_lower) cudaFree(d_lower);
  if (d_upper) cudaFree(d_upper);
  *out = m;
  return 0;
}

int hpccg_dev_matrix_set_halo(hpccg_dev_matrix *m, int num_neighbors, const int *neighbors, const int *recv_length,
                              const int *send_length, const int *elements_to_send, int total_to_be_sent) {
  if (!m || num_neighbors < 0 || total_to_be_sent < 0) return fail(HPCCG_ERR_ARG, "hpccg_dev_matrix_set_halo: bad arguments");
  long long recv_total = 0, send_total = 0;
  for (int i = 0; i < num_neighbors; ++i) {
    recv_total += recv_length[i];
    send_total += send_length[i];
  }
  if (recv_total != m->ncol - m->n || send_total != total_to_be_sent)
    return fail(HPCCG_ERR_ARG, "halo plan inconsistent: recv %lld vs %d externals, send %lld vs %d", recv_total,
                m->ncol - m->n, send_total, total_to_be_sent);
  m->num_neighbors = num_neighbors;
  m->neighbors.assign(neighbors, neighbors + num_neighbors);
  m->recv_length.assign(recv_length, recv_length + num_neighbors);
  m->send_length.assign(send_length, send_length + num_neighbors);
  m->total_to_be_sent = total_to_be_sent;
  invalidate_graph(m);
  if (m->d_elements_to_send) HPCCG_CUDA(cudaFree(m->d_elements_to_send));
  if (m->d_send_buffer) HPCCG_CUDA(cudaFree(m->d_send_buffer));
  if (m->d_put_inv) HPCCG_CUDA(cudaFree(m->d_put_inv));
  m->d_elements_to_send = nullptr;
  m->d_send_buffer = nullptr;
  m->d_put_inv = nullptr;
  m->put_fusable = 0;
  m->put_plan = HaloPut{};
  // Inverse send maps (row -> position in the neighbour's halo tail), one per neighbour, over the row range the segment
  // draws from: what lets the p-producing kernel do the put itself.  z-slabs send one boundary plane per neighbour (a
  // permutation of a contiguous row range in the 27-pt case, make_local_matrix.cpp:218-230), so the maps are exactly
  // plane-sized; a scattered or repeating send list keeps the stand-alone put kernel.
  if (num_neighbors > 0 && num_neighbors <= kMaxPeerNb) {
    std::vector<int> inv;
    bool ok = true;
    int seg = 0;
    for (int i = 0; i < num_neighbors && ok; ++i) {
      const int len = send_length[i];
      int lo = 0, hi = 0;
      if (len > 0) {
        lo = elements_to_send[seg];
        hi = lo + 1;
        for (int k = 0; k < len; ++k) {
          const int e = elements_to_send[seg + k];
          if (e < 0 || e >= m->n) return fail(HPCCG_ERR_ARG, "elements_to_send[%d] = %d outside local rows", seg + k, e);
          lo = std::min(lo, e);
          hi = std::max(hi, e + 1);
        }
        if ((long long)(hi - lo) > 4LL * len + 4096) ok = false;
      }
      m->put_plan.lo[i] = lo;
      m->put_plan.hi[i] = hi;
      m->put_plan.inv_off[i] = (int)inv.size();
      if (ok && len > 0) {
        const size_t base = inv.size();
        inv.resize(base + (size_t)(hi - lo), -1);
        for (int k = 0; k < len && ok; ++k) {
          int &slot = inv[base + (size_t)(elements_to_send[seg + k] - lo)];
          if (slot >= 0) ok = false;  // the same row twice for one neighbour: not a function
          slot = k;
        }
      }
      seg += len;
    }
    if (ok) {
      HPCCG_CUDA(cudaMalloc(&m->d_put_inv, sizeof(int) * std::max<size_t>(inv.size(), 1)));
      if (!inv.empty()) HPCCG_CUDA(cudaMemcpy(m->d_put_inv, inv.data(), sizeof(int) * inv.size(), cudaMemcpyHostToDevice));
      m->put_plan.inv = m->d_put_inv;
      m->put_plan.nseg = num_neighbors;
      m->put_fusable = 1;
    }
  }
  if (total_to_be_sent > 0) {
    for (int i = 0; i < total_to_be_sent; ++i)
      if (elements_to_send[i] < 0 || elements_to_send[i] >= m->n)
        return fail(HPCCG_ERR_ARG, "elements_to_send[%d] = %d outside local rows", i, elements_to_send[i]);
    HPCCG_CUDA(cudaMalloc(&m->d_elements_to_send, sizeof(int) * total_to_be_sent));
    HPCCG_CUDA(cudaMalloc(&m->d_send_buffer, sizeof(double) * total_to_be_sent));
    HPCCG_CUDA(cudaMemcpy(m->d_elements_to_send, elements_to_send, sizeof(int) * total_to_be_sent, cudaMemcpyHostToDevice));
  }
  return 0;
}

int hpccg_dev_matrix_destroy(hpccg_dev_matrix *m) {
  if (!m) return 0;
  cudaFree(m->vals);
  cudaFree(m->cols);
  cudaFree(m->pat_id);
  cudaFree(m->pat_val);
  cudaFree(m->pat_delta);
  cudaFree(m->pat_len);
  cudaFree(m->pat_desc);
  cudaFree(m->slice_slots);
  cudaFree(m->slice_off);
  cudaFree(m->perm);
  cudaFree(m->persist_win);

  cudaFree(m->d_elements_to_send);
  cudaFree(m->d_send_buffer);
  cudaFree(m->d_put_inv);
  cudaFree(m->partials);
  cudaFree(m->state);
  cudaFree(m->gathered);
  cudaFree(m->r);
  cudaFree(m->p);
  cudaFree(m->Ap);
  cudaFree(m->hist);
  cudaFree(m->scratch_x);
  cudaFree(m->scratch_y);
  peer_link_destroy(m);
  if (m->graph_exec) cudaGraphExecDestroy(m->graph_exec);
  if (m->graph_stream) cudaStreamDestroy(m->graph_stream);
  if (m->comm_stream) cudaStreamDestroy(m->comm_stream);
  if (m->ev_p_ready) cudaEventDestroy(m->ev_p_ready);
  if (m->ev_halo_done) cudaEventDestroy(m->ev_halo_done);
  for (cudaEvent_t e : m->ev_io)
    if (e) cudaEventDestroy(e);
  delete m;
  return 0;
}

int hpccg_dev_matrix_info(const hpccg_dev_matrix *m, int *local_nrow, int *local_ncol, int *slots, long long *padded_rows) {
  if (!m) return fail(HPCCG_ERR_ARG, "null matrix");
  if (local_nrow) *local_nrow = m->n;
  if (local_ncol) *local_ncol = m->ncol;
  if (slots) *slots = m->slots;
  if (padded_rows) *padded_rows = m->npad;
  return 0;
}

int hpccg_dev_matrix_download(const hpccg_dev_matrix *m, double *vals_host, int *cols_host) {
  if (!m) return fail(HPCCG_ERR_ARG, "null matrix");
  // The device arrays are SELL-C (or pattern-coded); the caller receives the canonical column-major
  // [slots][padded_rows] view with the original values and column ids.
  const size_t total = (size_t)m->slots * m->npad;
  if (m->format == 2) {
    const long long nslices = m->npad / kRaggedRows;
    std::vector<double> tv(vals_host ? m->total_elems : 0);
    std::vector<int> tc(m->total_elems), sl(nslices), pm(m->npad);
    std::vector<long long> of(nslices + 1);
    if (vals_host && m->total_elems) HPCCG_CUDA(cudaMemcpy(tv.data(), m->vals, sizeof(double) * m->total_elems, cudaMemcpyDeviceToHost));
    if (m->total_elems) HPCCG_CUDA(cudaMemcpy(tc.data(), m->cols, sizeof(int) * m->total_elems, cudaMemcpyDeviceToHost));
    HPCCG_CUDA(cudaMemcpy(sl.data(), m->slice_slots, sizeof(int) * nslices, cudaMemcpyDeviceToHost));
    HPCCG_CUDA(cudaMemcpy(of.data(), m->slice_off, sizeof(long long) * (nslices + 1), cudaMemcpyDeviceToHost));
    if (m->perm) HPCCG_CUDA(cudaMemcpy(pm.data(), m->perm, sizeof(int) * m->npad, cudaMemcpyDeviceToHost));
    else
      for (long long i = 0; i < m->npad; ++i) pm[i] = (int)i;
    for (size_t i = 0; i < total; ++i) {
      if (vals_host) vals_host[i] = 0.0;
      if (cols_host) cols_host[i] = -1;
    }
    for (long long pos = 0; pos < m->npad; ++pos) {
      const int row = pm[pos];
      if (row < 0 || row >= m->npad) continue;
      const long long s_ = pos / kRaggedRows, l = pos % kRaggedRows;
      for (int j = 0; j < sl[s_]; ++j) {
        const long long o = of[s_] + (long long)j * kRaggedRows + l;
        if (tc[o] < 0) continue;
        if (vals_host) vals_host[(size_t)j * m->npad + row] = tv[o];
        if (cols_host) cols_host[(size_t)j * m->npad + row] = tc[o];
      }
    }
    return 0;
  }
  if (m->format == 0) {
    std::vector<double> tv(vals_host ? total : 0);
    std::vector<int> tc(cols_host ? total : 0);
    if (vals_host) HPCCG_CUDA(cudaMemcpy(tv.data(), m->vals, sizeof(double) * total, cudaMemcpyDeviceToHost));
    if (cols_host) HPCCG_CUDA(cudaMemcpy(tc.data(), m->cols, sizeof(int) * total, cudaMemcpyDeviceToHost));
    for (int j = 0; j < m->slots; ++j)
      for (long long r = 0; r < m->npad; ++r) {
        if (vals_host) vals_host[(size_t)j * m->npad + r] = tv[sell_offset(r, j, m->slots)];
        if (cols_host) cols_host[(size_t)j * m->npad + r] = tc[sell_offset(r, j, m->slots)];
      }
    return 0;
  }
  std::vector<unsigned short> ids(m->npad);
  std::vector<double> pv((size_t)m->npat * m->slots);
  std::vector<int> pd((size_t)m->npat * m->slots), plen(m->npat);
  HPCCG_CUDA(cudaMemcpy(ids.data(), m->pat_id, sizeof(unsigned short) * m->npad, cudaMemcpyDeviceToHost));
  HPCCG_CUDA(cudaMemcpy(pv.data(), m->pat_val, sizeof(double) * pv.size(), cudaMemcpyDeviceToHost));
  HPCCG_CUDA(cudaMemcpy(pd.data(), m->pat_delta, sizeof(int) * pd.size(), cudaMemcpyDeviceToHost));
  HPCCG_CUDA(cudaMemcpy(plen.data(), m->pat_len, sizeof(int) * plen.size(), cudaMemcpyDeviceToHost));
  for (long long r = 0; r < m->npad; ++r) {
    const int id = ids[r];
    const int len = id == 0xFFFF ? 0 : plen[id];
    for (int j = 0; j < m->slots; ++j) {
      if (vals_host) vals_host[(size_t)j * m->npad + r] = j < len ? pv[(size_t)id * m->slots + j] : 0.0;
      if (cols_host) cols_host[(size_t)j * m->npad + r] = j < len ? (int)(r + pd[(size_t)id * m->slots + j]) : -1;
    }
  }
  return 0;
}

int hpccg_dev_matrix_bytes(const hpccg_dev_matrix *m, long long *bytes) {
  if (!m || !bytes) return fail(HPCCG_ERR_ARG, "null argument");
  if (m->format == 0) *bytes = (long long)m->slots * m->npad * 12;
  else if (m->format == 2) *bytes = m->total_elems * 12 + (m->npad / kRaggedRows) * 12 + (m->perm ? 4 * m->npad : 0);
  else *bytes = 2LL * m->npad + (long long)m->npat * (m->slots * 12 + 4);
  return 0;
}

int hpccg_dev_matrix_format(const hpccg_dev_matrix *m, int *format, int *patterns) {
  if (!m) return fail(HPCCG_ERR_ARG, "null matrix");
  if (format) *format = m->format;
  if (patterns) *patterns = m->npat;
  return 0;
}

int hpccg_dev_matrix_comm(const hpccg_dev_matrix *m, int *peer, int *fused_put) {
  if (!m) return fail(HPCCG_ERR_ARG, "null matrix");
  if (peer) *peer = m->peer_link ? 1 : 0;
  if (fused_put) *fused_put = (m->peer_link && m->put_fusable && !std::getenv("HPCCG_B200_SEPARATE_PUT")) ? 1 : 0;
  return 0;
}

// Lossless re-encoding of the SELL arrays: the distinct row patterns -- sequences of (value, column - row)
// pairs -- are collected in a device hash table, numbered, checked entry by entry against every row, and the mirror
// keeps one 16-bit id per row.  A matrix with more than 65535 patterns (or a slot count without a pattern kernel) is
// left in format 0; that is not an error.
int hpccg_dev_matrix_compress(hpccg_dev_matrix *m) {
  if (!m) return fail(HPCCG_ERR_ARG, "null matrix");
  if (m->format != 0 || (m->slots != 27 && m->slots != 7) || m->slots > kPatternSlots) return 0;
  const unsigned table_size = 1u << 21, mask = table_size - 1;
  unsigned long long *keys = nullptr, *freq = nullptr;
  int *ids = nullptr, *scal = nullptr, *rep = nullptr, *pat_delta = nullptr, *pat_len = nullptr;
  unsigned short *pat_id = nullptr;
  double *pat_val = nullptr;
  auto drop = [&] {
    cudaFree(keys);
    cudaFree(freq);
    cudaFree(ids);
    cudaFree(scal);
    cudaFree(rep);
    cudaFree(pat_delta);
    cudaFree(pat_len);
    cudaFree(pat_id);
    cudaFree(pat_val);
  };
#define HPCCG_CUDA_DROP(call)                                            \
  do {                                                                   \
    cudaError_t e_ = (call);                                             \
    if (e_ != cudaSuccess) {                                             \
      drop();                                                            \
      return fail_cuda(e_, #call, __FILE__, __LINE__);                   \
    }                                                                    \
  } while (0)
  HPCCG_CUDA_DROP(cudaMalloc(&keys, sizeof(unsigned long long) * table_size));
  HPCCG_CUDA_DROP(cudaMalloc(&ids, sizeof(int) * table_size));
  HPCCG_CUDA_DROP(cudaMalloc(&scal, sizeof(int) * 4));  // [0] overflow, [1] pattern count, [2] mismatch
  HPCCG_CUDA_DROP(cudaMemset(keys, 0, sizeof(unsigned long long) * table_size));
  HPCCG_CUDA_DROP(cudaMemset(scal, 0, sizeof(int) * 4));
  const int grid = stream_grid(m->npad, 16);
  pattern_insert_kernel<<<grid, kThreads>>>(m->vals, m->cols, m->slots, m->n, keys, mask, scal);
  pattern_number_kernel<<<stream_grid(table_size, 16), kThreads>>>(keys, table_size, ids, scal + 1);
  count_launch(2);
  int h_scal[4] = {0, 0, 0, 0};
  HPCCG_CUDA_DROP(cudaMemcpy(h_scal, scal, sizeof h_scal, cudaMemcpyDeviceToHost));
  const int npat = h_scal[1];
  if (h_scal[0] || npat > kMaxPatterns || npat < 1) {  // does not compress: stay in format 0
    drop();
    return 0;
  }
  HPCCG_CUDA_DROP(cudaMalloc(&pat_id, sizeof(unsigned short) * m->npad));
  HPCCG_CUDA_DROP(cudaMalloc(&rep, sizeof(int) * npat));
  HPCCG_CUDA_DROP(cudaMalloc(&freq, sizeof(unsigned long long) * npat));
  HPCCG_CUDA_DROP(cudaMemset(freq, 0, sizeof(unsigned long long) * npat));
  HPCCG_CUDA_DROP(cudaMalloc(&pat_val, sizeof(double) * (size_t)npat * m->slots));
  HPCCG_CUDA_DROP(cudaMalloc(&pat_delta, sizeof(int) * (size_t)npat * m->slots));
  HPCCG_CUDA_DROP(cudaMalloc(&pat_len, sizeof(int) * npat));
  pattern_assign_kernel<<<grid, kThreads>>>(m->vals, m->cols, m->slots, m->n, m->npad, keys, ids, mask, pat_id, rep, freq);
  pattern_fill_kernel<<<stream_grid(npat, 16), kThreads>>>(m->vals, m->cols, m->slots, npat, rep, pat_val, pat_delta, pat_len);
  count_launch(2);
  // the most frequent pattern becomes id 0 (the constant-bank fast path of the SpMV)
  std::vector<unsigned long long> h_freq(npat);
  std::vector<double> h_val((size_t)npat * m->slots);
  std::vector<int> h_delta((size_t)npat * m->slots), h_len(npat);
  HPCCG_CUDA_DROP(cudaMemcpy(h_freq.data(), freq, sizeof(unsigned long long) * npat, cudaMemcpyDeviceToHost));
  HPCCG_CUDA_DROP(cudaMemcpy(h_val.data(), pat_val, sizeof(double) * h_val.size(), cudaMemcpyDeviceToHost));
  HPCCG_CUDA_DROP(cudaMemcpy(h_delta.data(), pat_delta, sizeof(int) * h_delta.size(), cudaMemcpyDeviceToHost));
  HPCCG_CUDA_DROP(cudaMemcpy(h_len.data(), pat_len, sizeof(int) * npat, cudaMemcpyDeviceToHost));
  int top = 0;
  for (int i = 1; i < npat; ++i)
    if (h_freq[i] > h_freq[top]) top = i;
  if (top != 0) {
    for (int j = 0; j < m->slots; ++j) {
      std::swap(h_val[j], h_val[(size_t)top * m->slots + j]);
      std::swap(h_delta[j], h_delta[(size_t)top * m->slots + j]);
    }
    std::swap(h_len[0], h_len[top]);
    HPCCG_CUDA_DROP(cudaMemcpy(pat_val, h_val.data(), sizeof(double) * h_val.size(), cudaMemcpyHostToDevice));
    HPCCG_CUDA_DROP(cudaMemcpy(pat_delta, h_delta.data(), sizeof(int) * h_delta.size(), cudaMemcpyHostToDevice));
    HPCCG_CUDA_DROP(cudaMemcpy(pat_len, h_len.data(), sizeof(int) * npat, cudaMemcpyHostToDevice));
  }
  pattern_verify_kernel<<<grid, kThreads>>>(m->vals, m->cols, m->slots, m->n, pat_id, 0, top, pat_val, pat_delta, pat_len, scal + 2);
  count_launch();
  HPCCG_CUDA_DROP(cudaGetLastError());
  HPCCG_CUDA_DROP(cudaMemcpy(h_scal, scal, sizeof h_scal, cudaMemcpyDeviceToHost));
  if (h_scal[2]) {  // a hash collision merged two different patterns: keep the uncompressed matrix
    drop();
    return 0;
  }
#undef HPCCG_CUDA_DROP
  // Sub-pattern descriptors: pattern `id` = pattern 0 with some entries missing (same values, same deltas, same order).
  // Boundary rows of a stencil are exactly that -- whole (sy, sz) lines and / or the x-1 / x+1 entries are absent -- and the
  // marching SpMV runs them through the same unrolled code as interior rows: bits 0..8 = lines present, bit 9 = x-1 entries
  // missing, bit 10 = x+1 entries missing.  Anything else (halo rows with remapped columns, perturbed rows, irregular
  // sub-patterns) is marked generic and takes the per-row table path.
  std::vector<unsigned> h_desc(npat, 0xFFFFFFFFu);
  {
    const bool s27 = m->slots == 27;
    const int nruns = s27 ? 9 : 5;
    auto run_first = [&](int k) { return s27 ? 3 * k : (k < 2 ? k : (k == 2 ? 2 : k + 2)); };
    auto run_len = [&](int k) { return s27 ? 3 : (k == 2 ? 3 : 1); };
    for (int id = 0; id < npat && h_len[0] == m->slots; ++id) {
      unsigned mask = 0;
      int j0 = 0;
      bool sub = true;
      for (int e = 0; e < h_len[id] && sub; ++e) {
        const double v = h_val[(size_t)id * m->slots + e];
        const int d = h_delta[(size_t)id * m->slots + e];
        while (j0 < h_len[0] && !(h_delta[j0] == d && std::memcmp(&h_val[j0], &v, sizeof v) == 0)) ++j0;
        if (j0 == h_len[0]) sub = false;
        else mask |= 1u << j0++;
      }
      if (!sub) continue;
      unsigned runs = 0;
      int lm = -1, rm = -1;  // -1: no 3-line seen yet
      bool regular = true;
      for (int k = 0; k < nruns && regular; ++k) {
        const int e0 = run_first(k);
        const unsigned bits = (mask >> e0) & ((1u << run_len(k)) - 1u);
        if (!bits) continue;
        runs |= 1u << k;
        if (run_len(k) == 3) {
          const int l = (bits & 1u) ? 0 : 1, r = (bits & 4u) ? 0 : 1;
          regular = (bits & 2u) != 0 && (lm < 0 || (lm == l && rm == r));
          lm = l;
          rm = r;
        }
      }
      if (regular) h_desc[id] = runs | (lm == 1 ? (1u << 9) : 0u) | (rm == 1 ? (1u << 10) : 0u);
    }
  }
  unsigned *pat_desc = nullptr;
  {
    cudaError_t e_ = cudaMalloc(&pat_desc, sizeof(unsigned) * npat);
    if (e_ == cudaSuccess) e_ = cudaMemcpy(pat_desc, h_desc.data(), sizeof(unsigned) * npat, cudaMemcpyHostToDevice);
    if (e_ != cudaSuccess) {
      cudaFree(pat_desc);
      drop();
      return fail_cuda(e_, "pattern masks", __FILE__, __LINE__);
    }
  }
  // Stencil structure of pattern 0, read off its own deltas: lines of (x-1, x, x+1) around centre deltas sy*nx + sz*nx*ny.
  {
    MarchGeom g{};
    const int *d = h_delta.data();
    const double *v = h_val.data();
    bool ok = h_len[0] == m->slots;
    int nx = 0;
    long long plane = 0;
    if (ok && m->slots == 27) {
      for (int k = 0; k < 9 && ok; ++k) {
        g.base[k] = d[3 * k + 1];
        ok = d[3 * k] == g.base[k] - 1 && d[3 * k + 2] == g.base[k] + 1;
      }
      nx = g.base[5];
      plane = g.base[7];
      for (int sz = -1; sz <= 1 && ok; ++sz)
        for (int sy = -1; sy <= 1; ++sy) ok = ok && g.base[3 * (sz + 1) + (sy + 1)] == sz * plane + sy * nx;
      g.diag = 13;
    } else if (ok && m->slots == 7) {
      ok = d[2] == -1 && d[3] == 0 && d[4] == 1 && d[1] == -d[5] && d[0] == -d[6];
      nx = d[5];
      plane = d[6];
      g.base[0] = d[0];
      g.base[1] = d[1];
      g.base[2] = 0;
      g.base[3] = d[5];
      g.base[4] = d[6];
      g.diag = 3;
    } else {
      ok = false;
    }
    ok = ok && nx >= 16 && nx % 4 == 0 && plane > 0 && plane % nx == 0 && m->n % plane == 0 && m->ncol % 4 == 0;
    if (ok) {
      g.nx = nx;
      g.ny = (int)(plane / nx);
      g.nz = (int)(m->n / plane);
      g.cols_x = (g.nx + kMarchWidth - 1) / kMarchWidth;
      g.cols_y = (g.ny + kMarchLines - 1) / kMarchLines;
      g.neg1 = 1;
      for (int j = 0; j < m->slots; ++j)
        if (j != g.diag && v[j] != -1.0) g.neg1 = 0;
      g.ok = 1;
    }
    m->march = g;
  }
  m->pat_desc = pat_desc;
  std::memset(&m->pattern0, 0, sizeof m->pattern0);
  for (int j = 0; j < m->slots; ++j) {
    m->pattern0.value[j] = h_val[j];
    m->pattern0.delta[j] = h_delta[j];
  }
  m->pattern0.len = h_len[0];
  cudaFree(keys);
  cudaFree(freq);
  cudaFree(ids);
  cudaFree(scal);
  cudaFree(rep);
  invalidate_graph(m);  // a captured solve replays the SELL kernels on the arrays released here
  m->persist_state = -1;  // the single-kernel solve reads the SELL arrays
  cudaFree(m->vals);
  cudaFree(m->cols);
  m->vals = nullptr;
  m->cols = nullptr;
  m->pat_id = pat_id;
  m->pat_val = pat_val;
  m->pat_delta = pat_delta;
  m->pat_len = pat_len;
  m->npat = npat;
  m->format = 1;
  return 0;
}

}  // extern "C"

extern "C" {
// ================================================================================================
// Kernels
// ================================================================================================
int hpccg_dev_spmv(const hpccg_dev_matrix *m, const double *x, double *y, void *stream) {
  if (!m || !x || !y) return fail(HPCCG_ERR_ARG, "hpccg_dev_spmv: null argument");
  FinishParams fp{};
  return spmv_full(m, x, y, false, fp, (cudaStream_t)stream);
}

int hpccg_dev_spmv_dot(const hpccg_dev_matrix *m, const double *x, double *y, double *result_dev, void *stream) {
  if (!m || !x || !y || !result_dev) return fail(HPCCG_ERR_ARG, "hpccg_dev_spmv_dot: null argument");
  return spmv_full(m, x, y, true, fin_store(result_dev), (cudaStream_t)stream);
}

int hpccg_dev_dot(int n, const double *x, const double *y, double *result_dev, void *stream) {
  if (n < 0 || !x || !y || !result_dev) return fail(HPCCG_ERR_ARG, "hpccg_dev_dot: bad argument");
  GlobalWorkspace *w = nullptr;
  HPCCG_TRY(global_workspace(&w));
  std::lock_guard<std::mutex> lk(w->mu);
  const int grid = stream_grid((n + 1) / 2);
  if (x == y)
    dot_kernel<true><<<grid, kThreads, 0, (cudaStream_t)stream>>>(n, x, y, w->partials, grid, w->counter, fin_store(result_dev));
  else
    dot_kernel<false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(n, x, y, w->partials, grid, w->counter, fin_store(result_dev));
  count_launch();
  HPCCG_LAUNCH_CHECK();
  return 0;
}

static int launch_waxpby(int n, double alpha, const double *x, double beta, const double *beta_dev, const double *y,
                         double *w, const CgState *st_check, cudaStream_t s) {
  if (n == 0) return 0;
  const int grid = stream_grid((n + 1) / 2);
  // same branch selection as waxpby.cpp:73,79,85; a device-side beta always takes the alpha==1 form
  if (alpha == 1.0) waxpby_kernel<0><<<grid, kThreads, 0, s>>>(n, alpha, x, beta, beta_dev, y, w, st_check);
  else if (beta == 1.0 && !beta_dev) waxpby_kernel<1><<<grid, kThreads, 0, s>>>(n, alpha, x, beta, beta_dev, y, w, st_check);
  else waxpby_kernel<2><<<grid, kThreads, 0, s>>>(n, alpha, x, beta, beta_dev, y, w, st_check);
  count_launch();
  HPCCG_LAUNCH_CHECK();
  return 0;
}

int hpccg_dev_waxpby(int n, double alpha, const double *x, double beta, const double *y, double *w, void *stream) {
  if (n < 0 || !x || !y || !w) return fail(HPCCG_ERR_ARG, "hpccg_dev_waxpby: bad argument");
  return launch_waxpby(n, alpha, x, beta, nullptr, y, w, nullptr, (cudaStream_t)stream);
}

int hpccg_dev_update_xr_dot(int n, const double *alpha_dev, const double *p, const double *Ap, double *x, double *r,
                            double *rr_dev, void *stream) {
  if (n < 0 || !alpha_dev || !p || !Ap || !x || !r || !rr_dev) return fail(HPCCG_ERR_ARG, "hpccg_dev_update_xr_dot: bad argument");
  if (!aligned16(p) || !aligned16(Ap) || !aligned16(x) || !aligned16(r))
    return fail(HPCCG_ERR_ARG, "hpccg_dev_update_xr_dot: vectors must be 16-byte aligned");
  GlobalWorkspace *w = nullptr;
  HPCCG_TRY(global_workspace(&w));
  std::lock_guard<std::mutex> lk(w->mu);
  const int grid = stream_grid((n + 1) / 2);
  update_xr_dot_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(n, alpha_dev, p, Ap, x, r, w->partials, grid,
                                                                    w->counter, fin_store(rr_dev));
  count_launch();
  HPCCG_LAUNCH_CHECK();
  return 0;
}

int hpccg_dev_p_update(int n, const double *beta_dev, const double *r, double *p, void *stream) {
  if (n < 0 || !beta_dev || !r || !p) return fail(HPCCG_ERR_ARG, "hpccg_dev_p_update: bad argument");
  return launch_waxpby(n, 1.0, r, 0.0, beta_dev, p, p, nullptr, (cudaStream_t)stream);
}

int hpccg_dev_halo_pack(const hpccg_dev_matrix *m, const double *x, double *send_buffer_dev, void *stream) {
  if (!m || !x) return fail(HPCCG_ERR_ARG, "hpccg_dev_halo_pack: null argument");
  if (m->total_to_be_sent == 0) return 0;
  double *dst = send_buffer_dev ? send_buffer_dev : m->d_send_buffer;
  halo_pack_kernel<<<stream_grid(m->total_to_be_sent), kThreads, 0, (cudaStream_t)stream>>>(
      m->total_to_be_sent, m->d_elements_to_send, x, dst, nullptr);
  count_launch();
  HPCCG_LAUNCH_CHECK();
  return 0;
}

int hpccg_dev_max_abs_diff(int n, const double *v1, const double *v2, double *result_dev, void *stream) {
  if (n < 0 || !v1 || !v2 || !result_dev) return fail(HPCCG_ERR_ARG, "hpccg_dev_max_abs_diff: bad argument");
  GlobalWorkspace *w = nullptr;
  HPCCG_TRY(global_workspace(&w));
  std::lock_guard<std::mutex> lk(w->mu);
  const int grid = stream_grid(n);
  max_abs_diff_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(n, v1, v2, w->partials, grid, w->counter, result_dev);
  count_launch();
  HPCCG_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"

// ================================================================================================
// The CG loop (HPCCG.cpp:312-402)
// ================================================================================================
namespace hpccg {

enum TimerCat { T_DDOT = 1, T_WAXPBY = 2, T_SPMV = 3, T_ALLRED = 4, T_EXCH = 5, T_FUSED_SPMV = 8, T_FUSED_UPD = 9, T_PUPD = 10 };

// Pairs of CUDA events around launches, accumulated per category after the solve.
struct EventTimers {
  bool on = false;
  cudaStream_t s = nullptr;
  std::vector<cudaEvent_t> pool;
  std::vector<int> cats;
  size_t used = 0;
  int open_cat = 0;
  cudaEvent_t get() {
    if (used == pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      pool.push_back(e);
    }
    return pool[used++];
  }
  void tick(int cat) {
    if (!on) return;
    open_cat = cat;
    cudaEventRecord(get(), s);
  }
  void tock() {
    if (!on) return;
    cudaEventRecord(get(), s);
    cats.push_back(open_cat);
  }
  // acc[cat] += seconds
  void collect(double *acc, int ncat) {
    for (size_t i = 0; i < cats.size(); ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, pool[2 * i], pool[2 * i + 1]);
      if (cats[i] < ncat) acc[cats[i]] += ms * 1e-3;
    }
  }
  void reset() {
    used = 0;
    cats.clear();
  }
  ~EventTimers() {
    for (cudaEvent_t e : pool) cudaEventDestroy(e);
  }
};

struct SolveRank {
  hpccg_dev_matrix *m;
  const double *b;
  double *x;
  int grank;  // global rank id
};

static FinishParams make_fp(int mode, hpccg_dev_matrix *m, int k, int last, double tol, bool check) {
  FinishParams fp{};
  fp.mode = mode;
  fp.k = k;
  fp.last = last;
  fp.check_active = check ? 1 : 0;
  fp.tol = tol;
  fp.st = m->state;
  fp.hist = m->hist;
  fp.out = nullptr;
  return fp;
}

// The kernel that produces p (kernels.cuh): mode 0 = waxpby copy p = src (HPCCG.cpp:347,362), mode 1 = deferred x update +
// p = r + beta p (:383,:369); put != nullptr folds the peer-memory halo put of exchange `put->exch_idx` into it.
template <int VEC, int MODE>
static int launch_p_update_t(int n, const CgState *st, bool check, const double *src_r, double *p, double *x, const HaloPut *put,
                             cudaStream_t s) {
  const int grid = stream_grid((n + VEC - 1) / VEC);
  if (put) p_update_x_kernel<VEC, MODE, true><<<grid, kThreads, 0, s>>>(n, st, check ? 1 : 0, src_r, p, x, *put);
  else p_update_x_kernel<VEC, MODE, false><<<grid, kThreads, 0, s>>>(n, st, check ? 1 : 0, src_r, p, x, HaloPut{});
  count_launch();
  HPCCG_LAUNCH_CHECK();
  return 0;
}

static int launch_p_update(int mode, int n, const CgState *st, bool check, const double *src_r, double *p, double *x,
                           const HaloPut *put, cudaStream_t s) {
  if (n == 0) return 0;
  const bool v4 = use_vec4(src_r, p, mode == 1 ? (const void *)x : (const void *)p);
  if (mode == 1) return v4 ? launch_p_update_t<4, 1>(n, st, check, src_r, p, x, put, s) : launch_p_update_t<2, 1>(n, st, check, src_r, p, x, put, s);
  return v4 ? launch_p_update_t<4, 0>(n, st, check, src_r, p, x, put, s) : launch_p_update_t<2, 0>(n, st, check, src_r, p, x, put, s);
}

// Halo exchange of vector v (ncol doubles) for all local ranks.  nccl: one local rank, transfers on the
// matrix's comm stream (caller fences with events); otherwise every rank of the world is local and the
// transfers are device copies on `s`.
static int exchange_halo(std::vector<SolveRank> &rk, int R, bool nccl, double *const *v, const CgState *const *chk,
                         cudaStream_t s) {
  if (R == 1) return 0;
  for (size_t q = 0; q < rk.size(); ++q) {
    hpccg_dev_matrix *m = rk[q].m;
    if (m->total_to_be_sent > 0) {
      halo_pack_kernel<<<stream_grid(m->total_to_be_sent), kThreads, 0, s>>>(m->total_to_be_sent, m->d_elements_to_send,
                                                                              v[q], m->d_send_buffer, chk ? chk[q] : nullptr);
      count_launch();
      HPCCG_LAUNCH_CHECK();
    }
  }
  if (nccl) {
    hpccg_dev_matrix *m = rk[0].m;
    return nccl_halo_exchange(m->d_send_buffer, m->send_length.data(), v[0] + m->n, m->recv_length.data(),
                              m->neighbors.data(), m->num_neighbors, s);
  }
  // in-process world: rank q receives from neighbour i the slice that neighbour packed for q
  for (size_t q = 0; q < rk.size(); ++q) {
    hpccg_dev_matrix *m = rk[q].m;
    double *dst = v[q] + m->n;
    for (int i = 0; i < m->num_neighbors; ++i) {
      const int nb = m->neighbors[i];
      if (nb < 0 || nb >= (int)rk.size()) return fail(HPCCG_ERR_STATE, "neighbour %d is not a local rank", nb);
      hpccg_dev_matrix *mb = rk[nb].m;
      const double *src = mb->d_send_buffer;
      int found = -1;
      for (int k = 0; k < mb->num_neighbors; ++k) {
        if (mb->neighbors[k] == rk[q].grank) {
          found = k;
          break;
        }
        src += mb->send_length[k];
      }
      if (found < 0 || mb->send_length[found] != m->recv_length[i])
        return fail(HPCCG_ERR_STATE, "halo plans of ranks %d and %d do not match", rk[q].grank, nb);
      if (m->recv_length[i] > 0)
        HPCCG_CUDA(cudaMemcpyAsync(dst, src, sizeof(double) * m->recv_length[i], cudaMemcpyDeviceToDevice, s));
      dst += m->recv_length[i];
    }
  }
  return 0;
}

// capture_only: the launches are being recorded into a CUDA graph (stream capture): no event timing, no host polling,
// no read-back -- solve_readback() runs after the graph has been launched.
static int solve_readback(hpccg_dev_matrix *m, int max_iter, int *niters_out, double *normr_out, double *hist_host,
                          CgState *hs_out, cudaStream_t s) {
  CgState hs;
  HPCCG_CUDA(cudaMemcpyAsync(&hs, m->state, sizeof(CgState), cudaMemcpyDeviceToHost, s));
  if (hist_host) HPCCG_CUDA(cudaMemcpyAsync(hist_host, m->hist, sizeof(double) * max_iter, cudaMemcpyDeviceToHost, s));
  HPCCG_CUDA(cudaStreamSynchronize(s));
  if (niters_out) *niters_out = hs.niters;
  if (normr_out) *normr_out = hs.normr;
  if (hs_out) *hs_out = hs;
  return 0;
}

static int cg_solve_impl(std::vector<SolveRank> &rk, int R, bool nccl, int max_iter, double tol, int *niters_out,
                         double *normr_out, double *hist_host, double *times, double *loop_ms, int flags, cudaStream_t s,
                         bool capture_only = false, const SolveIO *io = nullptr) {
  const int L = (int)rk.size();
  const bool multi = R > 1;
  const bool unfused = (flags & HPCCG_SOLVE_UNFUSED) != 0;
  const bool defer_x = !unfused && !(flags & HPCCG_SOLVE_EAGER_X);  // x += alpha p rides in the next p-update (kernels.cuh)
  bool overlap = nccl && !(flags & HPCCG_SOLVE_NO_OVERLAP) && !unfused;
  if (max_iter < 1) max_iter = 1;
  for (auto &q : rk) {
    if (!q.m || !q.b || !q.x) return fail(HPCCG_ERR_ARG, "cg_solve: null argument");
    if (!aligned16(q.b) || !aligned16(q.x)) return fail(HPCCG_ERR_ARG, "cg_solve: b and x must be 16-byte aligned");
    if (multi && q.m->ncol > q.m->n && q.m->num_neighbors == 0)
      return fail(HPCCG_ERR_STATE, "cg_solve: matrix has halo columns but no halo plan (call make_local_matrix)");
    // a localised matrix solved as a single rank would read a halo nobody fills (e.g. the rank context of another thread)
    if (!multi && q.m->ncol > q.m->n)
      return fail(HPCCG_ERR_STATE, "cg_solve: matrix has %d halo columns but this thread's rank context is 1 rank "
                  "(hpccg_ctx_set / hpccg_nccl_init on this thread)", q.m->ncol - q.m->n);
    HPCCG_TRY(ensure_solver_workspace(q.m, max_iter, R));
  }
  // all local ranks share rank 0's gather array in the in-process world
  double *gathered = rk[0].m->gathered;
  // Multi-process runs: halos and scalar sums go through peer memory inside the kernels (PeerLink) when every rank
  // could map its peers and the matrix is served by the TMA SpMV; otherwise NCCL send/recv + gathers between kernels.
  // peer_link_create is a COLLECTIVE (it allgathers over the NCCL communicator), so whether it is entered may depend only
  // on things that are equal on every rank (the flags); whether THIS rank can serve the peer path -- its slot count has a
  // TMA kernel, or it is pattern-coded -- travels inside the exchange as an eligibility bit and all ranks take one decision.
  PeerLink *link = nullptr;
  if (nccl && !(flags & HPCCG_SOLVE_NCCL_ONLY)) {
    const bool eligible = rk[0].m->format == 1 || (rk[0].m->format == 0 && use_tma_path(rk[0].m->slots));
    HPCCG_TRY(peer_link_create(rk[0].m, eligible));
    link = rk[0].m->peer_link;
  }
  const bool p2p = link != nullptr;
  if (p2p) overlap = false;  // the exchange is hidden inside the single SpMV launch instead
  static thread_local EventTimers timers;
  timers.reset();
  timers.on = (flags & HPCCG_SOLVE_TIMERS) != 0 && times != nullptr;
  timers.s = s;
  // released on every return path
  struct LoopResources {
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int *h_active = nullptr;
    ~LoopResources() {
      if (ev0) cudaEventDestroy(ev0);
      if (ev1) cudaEventDestroy(ev1);
      if (h_active) cudaFreeHost(h_active);
    }
  } res;
  cudaEvent_t &ev_loop0 = res.ev0, &ev_loop1 = res.ev1;
  if (!capture_only) {
    HPCCG_CUDA(cudaEventCreate(&ev_loop0));
    HPCCG_CUDA(cudaEventCreate(&ev_loop1));
  }
  double t4_host = 0.0;

  std::vector<double *> pv(L);
  std::vector<const CgState *> chk(L);
  for (int q = 0; q < L; ++q) {
    pv[q] = rk[q].m->p;
    chk[q] = rk[q].m->state;
  }

  // -- reduction tail: local sums -> (gather) -> scalar finish on every local rank
  auto finish_multi = [&](int mode, int k, int last, bool check) -> int {
    if (p2p) return 0;  // already summed over the ranks inside the reducing kernel
    timers.tick(T_ALLRED);
    if (nccl) HPCCG_TRY(nccl_allgather_double(gathered, s));
    for (int q = 0; q < L; ++q) {
      cg_scalar_kernel<<<1, 32, 0, s>>>(gathered, R, make_fp(mode, rk[q].m, k, last, tol, check));
      count_launch();
    }
    HPCCG_LAUNCH_CHECK();
    timers.tock();
    return 0;
  };
  auto fp_for = [&](int mode, int q, int k, int last, bool check) {
    if (!multi) return make_fp(mode, rk[q].m, k, last, tol, check);
    if (p2p) {
      FinishParams fp = make_fp(mode, rk[q].m, k, last, tol, check);
      fp.peer = link;
      return fp;
    }
    FinishParams fp = make_fp(FIN_STORE, rk[q].m, k, last, tol, check);
    fp.out = gathered + rk[q].grank;
    return fp;
  };
  // p2p with compact inverse send maps: the put rides in the kernel that produces p (no exchange launch at all)
  // (the literal / eager-x sequences, kept for validation and A/B, use the stand-alone put kernel)
  const bool fused_put = p2p && defer_x && rk[0].m->put_fusable && !std::getenv("HPCCG_B200_SEPARATE_PUT");
  auto put_for = [&](int exch_idx) {
    HaloPut hp = rk[0].m->put_plan;
    hp.exch_idx = exch_idx;
    return hp;
  };
  auto do_exchange = [&](bool check, int exch_idx) -> int {
    if (!multi || fused_put) return 0;
    timers.tick(T_EXCH);
    if (p2p) {
      hpccg_dev_matrix *m = rk[0].m;
      if (m->num_neighbors > 0) {
        // one element per thread: the remote stores are posted writes, so the kernel's length is one gather + one
        // NVLink round trip for the system-scope fence, not a per-thread chain of dependent iterations
        const int grid = std::max(1, std::min(2048, (m->total_to_be_sent + kThreads - 1) / kThreads));
        halo_put_kernel<<<grid, kThreads, 0, s>>>(m->total_to_be_sent, m->d_elements_to_send, m->p, link, check ? m->state : nullptr, exch_idx);
        count_launch();
        HPCCG_LAUNCH_CHECK();
      }
    } else {
      HPCCG_TRY(exchange_halo(rk, R, nccl, pv.data(), check ? chk.data() : nullptr, s));
    }
    timers.tock();
    return 0;
  };
  SpmvHalo halo{};
  if (p2p) halo = SpmvHalo{link, rk[0].m->n, rk[0].m->interior_begin, rk[0].m->interior_end, 0};
  auto spmv_dot_all = [&](int mode, int k, bool check, bool dot) -> int {
    halo.exch_idx = k + 1;  // the exchange that precedes this SpMV (1 = set-up)
    for (int q = 0; q < L; ++q) {
      hpccg_dev_matrix *m = rk[q].m;
      FinishParams fp = dot ? fp_for(mode, q, k, 0, check) : make_fp(FIN_STORE, m, k, 0, tol, check);
      HPCCG_TRY(spmv_full(m, m->p, m->Ap, dot, fp, s, halo));
    }
    return 0;
  };

  for (int q = 0; q < L; ++q) {
    hpccg_dev_matrix *m = rk[q].m;
    cg_state_init_kernel<<<1, 32, 0, s>>>(m->state, p2p ? link : nullptr);
    count_launch();
    HPCCG_CUDA(cudaMemsetAsync(m->hist, 0xFF, sizeof(double) * (max_iter + 1), s));  // NaN = "no iteration ran"
  }
  HPCCG_LAUNCH_CHECK();
  // one stream-ordered rendezvous per solve: every rank's workspace and peer mappings exist before the first remote store
  if (p2p) HPCCG_TRY(nccl_allgather_double(gathered, s));

  // ---- set-up: p = x ; Ap = A p ; r = b - Ap ; rtrans = r.r (HPCCG.cpp:347-354) ----
  timers.tick(T_WAXPBY);
  if (fused_put) {
    const HaloPut hp = put_for(1);
    HPCCG_TRY(launch_p_update(0, rk[0].m->n, rk[0].m->state, false, rk[0].x, rk[0].m->p, nullptr, &hp, s));
  } else {
    for (int q = 0; q < L; ++q) HPCCG_TRY(launch_waxpby(rk[q].m->n, 1.0, rk[q].x, 0.0, nullptr, rk[q].x, rk[q].m->p, nullptr, s));
  }
  timers.tock();
  HPCCG_TRY(do_exchange(false, 1));
  timers.tick(T_SPMV);
  HPCCG_TRY(spmv_dot_all(FIN_STORE, 0, false, false));
  timers.tock();
  // b is first read here: a host b may still be on its way (HPCCG() uploads it behind x, under the set-up SpMV)
  if (io && io->b_ready) HPCCG_CUDA(cudaStreamWaitEvent(s, io->b_ready, 0));
  if (!unfused) {
    timers.tick(T_WAXPBY);
    for (int q = 0; q < L; ++q) {
      hpccg_dev_matrix *m = rk[q].m;
      const int grid = stream_grid(m->n);
      residual_dot_kernel<<<grid, kThreads, 0, s>>>(m->n, rk[q].b, m->Ap, m->r, m->partials, grid, &m->state->counter,
                                                    fp_for(FIN_INIT, q, 0, max_iter <= 1, false));
      count_launch();
    }
    HPCCG_LAUNCH_CHECK();
    timers.tock();
  } else {
    timers.tick(T_WAXPBY);
    for (int q = 0; q < L; ++q) HPCCG_TRY(launch_waxpby(rk[q].m->n, 1.0, rk[q].b, -1.0, nullptr, rk[q].m->Ap, rk[q].m->r, nullptr, s));
    timers.tock();
    timers.tick(T_DDOT);
    for (int q = 0; q < L; ++q) {
      hpccg_dev_matrix *m = rk[q].m;
      const int grid = stream_grid((m->n + 1) / 2);
      dot_kernel<true><<<grid, kThreads, 0, s>>>(m->n, m->r, m->r, m->partials, grid, &m->state->counter,
                                                 fp_for(FIN_INIT, q, 0, max_iter <= 1, false));
      count_launch();
    }
    HPCCG_LAUNCH_CHECK();
    timers.tock();
  }
  if (multi) HPCCG_TRY(finish_multi(FIN_INIT, 0, max_iter <= 1, false));

  // ---- iterations (HPCCG.cpp:358-386) ----
  if (!capture_only) HPCCG_CUDA(cudaEventRecord(ev_loop0, s));
  int *&h_active = res.h_active;
  if (tol > 0.0 && !capture_only) HPCCG_CUDA(cudaMallocHost(&h_active, sizeof(int)));
  for (int k = 1; k < max_iter; ++k) {
    const int last = (k + 1 == max_iter) ? 1 : 0;
    if (unfused && k > 1) {
      // rtrans = r.r ; beta (HPCCG.cpp:366-368)
      timers.tick(T_DDOT);
      for (int q = 0; q < L; ++q) {
        hpccg_dev_matrix *m = rk[q].m;
        const int grid = stream_grid((m->n + 1) / 2);
        // the unfused sequence evaluates the loop condition at the top of iteration k: FIN_RR of k-1
        dot_kernel<true><<<grid, kThreads, 0, s>>>(m->n, m->r, m->r, m->partials, grid, &m->state->counter,
                                                   fp_for(FIN_RR, q, k - 1, 0, true));
        count_launch();
      }
      HPCCG_LAUNCH_CHECK();
      timers.tock();
      if (multi) HPCCG_TRY(finish_multi(FIN_RR, k - 1, 0, true));
    }
    // p = r (k==1, HPCCG.cpp:362) or p = r + beta p (:369)
    timers.tick(T_PUPD);
    for (int q = 0; q < L; ++q) {
      hpccg_dev_matrix *m = rk[q].m;
      const HaloPut hp = fused_put ? put_for(k + 1) : HaloPut{};
      const HaloPut *put = fused_put ? &hp : nullptr;
      if (k == 1) {
        if (put) HPCCG_TRY(launch_p_update(0, m->n, m->state, true, m->r, m->p, nullptr, put, s));
        else HPCCG_TRY(launch_waxpby(m->n, 1.0, m->r, 0.0, nullptr, m->r, m->p, m->state, s));
      } else if (defer_x && vec_tma().tile > 0 && m->n >= (1 << 20) && aligned16(rk[q].x)) {
        VecPtrs vp{};
        vp.in[0] = m->r;
        vp.in[1] = m->p;
        vp.in[2] = rk[q].x;
        vp.out[0] = m->p;
        vp.out[1] = rk[q].x;
        HPCCG_TRY(launch_vec_tma<VecOpPUpdate>(m, vp, FinishParams{}, put, s));
      } else if (defer_x) {
        HPCCG_TRY(launch_p_update(1, m->n, m->state, true, m->r, m->p, rk[q].x, put, s));
      } else {
        HPCCG_TRY(launch_waxpby(m->n, 1.0, m->r, 0.0, &m->state->beta, m->p, m->p, m->state, s));
      }
    }
    timers.tock();

    if (overlap) {
      // halo on the comm stream, interior rows meanwhile, halo-touching rows afterwards
      hpccg_dev_matrix *m = rk[0].m;
      HPCCG_CUDA(cudaEventRecord(m->ev_p_ready, s));
      HPCCG_CUDA(cudaStreamWaitEvent(m->comm_stream, m->ev_p_ready, 0));
      HPCCG_TRY(exchange_halo(rk, R, true, pv.data(), chk.data(), m->comm_stream));
      HPCCG_CUDA(cudaEventRecord(m->ev_halo_done, m->comm_stream));
      SpmvPlan pi = plan_spmv<true>(m, m->interior_begin, m->interior_end);
      SpmvPlan pa = plan_spmv<true>(m, 0, m->interior_begin);
      SpmvPlan pb = plan_spmv<true>(m, m->interior_end, m->n);
      const int total = pi.grid + pa.grid + pb.grid;
      FinishParams fp = fp_for(FIN_PAP, 0, k, 0, true);
      timers.tick(T_FUSED_SPMV);
      HPCCG_TRY(launch_spmv<true>(m, m->p, m->Ap, pi, 0, total, fp, s));
      HPCCG_CUDA(cudaStreamWaitEvent(s, m->ev_halo_done, 0));
      HPCCG_TRY(launch_spmv<true>(m, m->p, m->Ap, pa, pi.grid, total, fp, s));
      HPCCG_TRY(launch_spmv<true>(m, m->p, m->Ap, pb, pi.grid + pa.grid, total, fp, s));
      timers.tock();
    } else {
      HPCCG_TRY(do_exchange(true, k + 1));
      if (!unfused) {
        timers.tick(T_FUSED_SPMV);
        HPCCG_TRY(spmv_dot_all(FIN_PAP, k, true, true));
        timers.tock();
      } else {
        timers.tick(T_SPMV);
        HPCCG_TRY(spmv_dot_all(FIN_STORE, k, true, false));
        timers.tock();
        timers.tick(T_DDOT);
        for (int q = 0; q < L; ++q) {
          hpccg_dev_matrix *m = rk[q].m;
          const int grid = stream_grid((m->n + 1) / 2);
          dot_kernel<false><<<grid, kThreads, 0, s>>>(m->n, m->p, m->Ap, m->partials, grid, &m->state->counter,
                                                      fp_for(FIN_PAP, q, k, 0, true));
          count_launch();
        }
        HPCCG_LAUNCH_CHECK();
        timers.tock();
      }
    }
    if (multi) HPCCG_TRY(finish_multi(FIN_PAP, k, 0, true));

    // x += alpha p ; r -= alpha Ap (HPCCG.cpp:383-384) [+ r.r of the next iteration when fused]
    if (!unfused) {
      timers.tick(T_FUSED_UPD);
      for (int q = 0; q < L; ++q) {
        hpccg_dev_matrix *m = rk[q].m;
        const bool v4 = defer_x && use_vec4(m->Ap, m->r, m->r);
        const int grid = stream_grid(v4 ? (m->n + 3) / 4 : (m->n + 1) / 2);
        if (defer_x && vec_tma().tile > 0 && m->n >= (1 << 20)) {
          VecPtrs vp{};
          vp.in[0] = m->Ap;
          vp.in[1] = m->r;
          vp.out[0] = m->r;
          HPCCG_TRY(launch_vec_tma<VecOpRUpdate>(m, vp, fp_for(FIN_RR, q, k, last, true), nullptr, s));
          continue;  // (counted by the launcher)
        } else if (v4)
          update_r_dot_kernel<4><<<grid, kThreads, 0, s>>>(m->n, &m->state->alpha, m->Ap, m->r, m->partials, grid, &m->state->counter,
                                                            fp_for(FIN_RR, q, k, last, true));
        else if (defer_x)
          update_r_dot_kernel<2><<<grid, kThreads, 0, s>>>(m->n, &m->state->alpha, m->Ap, m->r, m->partials, grid, &m->state->counter,
                                                            fp_for(FIN_RR, q, k, last, true));
        else
          update_xr_dot_kernel<<<grid, kThreads, 0, s>>>(m->n, &m->state->alpha, m->p, m->Ap, rk[q].x, m->r, m->partials, grid,
                                                          &m->state->counter, fp_for(FIN_RR, q, k, last, true));
        count_launch();
      }
      HPCCG_LAUNCH_CHECK();
      timers.tock();
      if (multi) HPCCG_TRY(finish_multi(FIN_RR, k, last, true));
    } else {
      timers.tick(T_WAXPBY);
      for (int q = 0; q < L; ++q) {
        hpccg_dev_matrix *m = rk[q].m;
        HPCCG_TRY(launch_waxpby(m->n, 1.0, rk[q].x, 0.0, &m->state->alpha, m->p, rk[q].x, m->state, s));
        HPCCG_TRY(launch_waxpby(m->n, 1.0, m->r, 0.0, &m->state->neg_alpha, m->Ap, m->r, m->state, s));
      }
      timers.tock();
    }
    if (h_active && (k % 16 == 0)) {
      // a positive tolerance can end the loop early; look every 16 iterations instead of every one
      HPCCG_CUDA(cudaMemcpyAsync(h_active, &rk[0].m->state->active, sizeof(int), cudaMemcpyDeviceToHost, s));
      HPCCG_CUDA(cudaStreamSynchronize(s));
      if (*h_active == 0) break;
    }
  }
  const bool stream_x_out = io && io->x_host && io->copy_stream && L == 1 && !capture_only;
  if (!capture_only && stream_x_out) HPCCG_CUDA(cudaEventRecord(ev_loop1, s));
  if (stream_x_out) {
    // final x to the host in chunks: chunk c is fixed up (the x update of the last executed iteration) on `s` and copied
    // on copy_stream while chunk c+1 is being fixed up
    hpccg_dev_matrix *m = rk[0].m;
    const int chunks = hpccg_dev_matrix::kIoChunks;
    const long long per = ((m->n + chunks - 1) / chunks + 511) / 512 * 512;  // 4 KiB-aligned chunk boundaries
    int c = 0;
    for (long long off = 0; off < m->n; off += per, ++c) {
      const int len = (int)std::min<long long>(per, m->n - off);
      if (defer_x && max_iter > 1) {
        x_fixup_kernel<<<stream_grid((len + 1) / 2), kThreads, 0, s>>>(len, m->state, m->p + off, rk[0].x + off);
        count_launch();
        HPCCG_LAUNCH_CHECK();
      }
      HPCCG_CUDA(cudaEventRecord(m->ev_io[1 + c], s));
      HPCCG_CUDA(cudaStreamWaitEvent(io->copy_stream, m->ev_io[1 + c], 0));
      HPCCG_CUDA(cudaMemcpyAsync(io->x_host + off, rk[0].x + off, sizeof(double) * len, cudaMemcpyDeviceToHost, io->copy_stream));
    }
  } else if (defer_x && max_iter > 1) {  // the x update of the last executed iteration
    for (int q = 0; q < L; ++q) {
      hpccg_dev_matrix *m = rk[q].m;
      x_fixup_kernel<<<stream_grid((m->n + 1) / 2), kThreads, 0, s>>>(m->n, m->state, m->p, rk[q].x);
      count_launch();
    }
    HPCCG_LAUNCH_CHECK();
  }
  if (capture_only) return 0;
  if (!stream_x_out) HPCCG_CUDA(cudaEventRecord(ev_loop1, s));

  // ---- results ----
  CgState hs;
  HPCCG_TRY(solve_readback(rk[0].m, max_iter, niters_out, normr_out, hist_host, &hs, s));
  if (stream_x_out) HPCCG_CUDA(cudaStreamSynchronize(io->copy_stream));
  if (p2p) {
    int perr = 0;
    HPCCG_CUDA(cudaMemcpy(&perr, &link->error, sizeof(int), cudaMemcpyDeviceToHost));
    if (perr) return fail(HPCCG_ERR_COMM, "peer-memory wait timed out (%s): a rank of the job did not arrive",
                          perr == 2 ? "halo" : "scalar reduction");
  }
  float ms = 0.f;
  HPCCG_CUDA(cudaEventElapsedTime(&ms, ev_loop0, ev_loop1));
  if (loop_ms) *loop_ms = ms;
  if (times) {
    double acc[16] = {0};
    timers.collect(acc, 16);
    // Fused kernels are split by the unfused algorithmic byte counts ("times[]"):
    //   SpMV+p.Ap : 340n SpMV / 16n ddot ; update+r.r : 48n waxpby / 8n ddot
    //   deferred-x form: r-update+r.r : 24n waxpby / 8n ddot ; the p-update kernel carries both remaining waxpbys
    const double upd_dot = defer_x ? 8.0 / 32.0 : 8.0 / 56.0;
    times[1] = acc[T_DDOT] + acc[T_FUSED_SPMV] * (16.0 / 356.0) + acc[T_FUSED_UPD] * upd_dot;
    times[2] = acc[T_WAXPBY] + acc[T_FUSED_UPD] * (1.0 - upd_dot);
    times[3] = acc[T_SPMV] + acc[T_FUSED_SPMV] * (340.0 / 356.0);
    times[2] += acc[T_PUPD];
    times[4] = acc[T_ALLRED] + t4_host;
    times[5] = acc[T_EXCH];
    times[7] = acc[T_FUSED_SPMV];
    times[8] = acc[T_FUSED_UPD];
    times[9] = acc[T_PUPD];
    times[10] = hs.niters;
  }
  return 0;
}

}  // namespace hpccg

namespace hpccg {
// ---- the whole solve as one kernel of one thread-block cluster (persistent_cg.cuh) -----------------------------------------
template <int SLOTS>
static int cluster_launch(hpccg_dev_matrix *m, const double *b, double *x, int max_iter, double tol, cudaStream_t s, bool query_only,
                          bool *fits) {
  auto kern = cg_cluster_kernel<SLOTS>;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(m->persist_G);
  cfg.blockDim = dim3(m->persist_threads);
  cfg.dynamicSmemBytes = (size_t)m->persist_smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = m->persist_G;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (query_only) {
    *fits = false;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, m->persist_smem) != cudaSuccess ||
        cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    int clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&clusters, kern, &cfg) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    *fits = clusters >= 1;
    return 0;
  }
  const int *wlo = m->persist_win, *whi = m->persist_win + m->persist_G;
  HPCCG_CUDA(cudaLaunchKernelEx(&cfg, kern, (const double *)m->vals, (const int *)m->cols, m->slots, m->n, m->persist_rows, m->persist_window, wlo, whi, b,
                                x, max_iter, tol, m->state, m->hist));
  count_launch();
  return 0;
}

static int cluster_dispatch(hpccg_dev_matrix *m, const double *b, double *x, int max_iter, double tol, cudaStream_t s, bool query_only,
                            bool *fits) {
  if (m->slots == 27) return cluster_launch<27>(m, b, x, max_iter, tol, s, query_only, fits);
  if (m->slots == 7) return cluster_launch<7>(m, b, x, max_iter, tol, s, query_only, fits);
  return cluster_launch<0>(m, b, x, max_iter, tol, s, query_only, fits);
}

static int persistent_prepare(hpccg_dev_matrix *m) {
  if (m->persist_state != 0) return 0;
  m->persist_state = -1;
  if (m->format != 0 || m->ncol != m->n || !m->vals) return 0;
  int dev = 0, max_smem = 0;
  HPCCG_CUDA(cudaGetDevice(&dev));
  HPCCG_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  // as many CTAs as the cluster allows, one row per thread
  const int G = std::min(kClusterMax, std::max(1, (m->n + 31) / 32));
  const int rows = (m->n + G - 1) / G;
  const int used = (m->n + rows - 1) / rows;  // CTAs that own rows
  const int threads = (int)round_up(rows, 32);
  if (threads > kClusterThreadsMax || (long long)m->slots * rows * 12 > max_smem) return 0;
  int *win = nullptr;
  HPCCG_CUDA(cudaMalloc(&win, sizeof(int) * 2 * used));
  persist_window_kernel<<<used, kThreads>>>(m->cols, m->slots, m->n, rows, win, win + used);
  count_launch();
  std::vector<int> h(2 * used);
  cudaError_t e = cudaMemcpy(h.data(), win, sizeof(int) * 2 * used, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) {
    cudaFree(win);
    return fail_cuda(e, "cluster solve plan", __FILE__, __LINE__);
  }
  int max_w = 0;
  for (int b = 0; b < used; ++b) max_w = std::max(max_w, h[used + b] - h[b]);
  const long long smem = (long long)rows * 16 + (long long)m->slots * rows * 8 + ((long long)m->slots * rows + 2) * 4 +
                         (long long)(max_w + 2) * 16 + 64;
  m->persist_window = (max_w + 1) & ~1;
  m->persist_win = win;
  m->persist_G = used;
  m->persist_rows = rows;
  m->persist_threads = threads;
  m->persist_smem = (int)smem;
  bool fits = false;
  if (smem <= max_smem - 1024) HPCCG_TRY(cluster_dispatch(m, nullptr, nullptr, 0, 0.0, nullptr, true, &fits));
  if (!fits) {  // the column reach of a row block does not fit, or no GPC can host the cluster: normal loop
    cudaFree(win);
    m->persist_win = nullptr;
    return 0;
  }
  m->persist_state = 1;
  return 0;
}

static int cg_solve_persistent(hpccg_dev_matrix *m, const double *b, double *x, int max_iter, double tol, int *niters,
                               double *normr, double *hist_host, double *loop_ms, cudaStream_t s) {
  if (max_iter < 1) max_iter = 1;
  HPCCG_TRY(ensure_solver_workspace(m, max_iter, 1));
  HPCCG_CUDA(cudaMemsetAsync(m->hist, 0xFF, sizeof(double) * (max_iter + 1), s));  // NaN = "no iteration ran"
  struct Ev {
    cudaEvent_t a = nullptr, b = nullptr;
    ~Ev() {
      if (a) cudaEventDestroy(a);
      if (b) cudaEventDestroy(b);
    }
  } ev;
  HPCCG_CUDA(cudaEventCreate(&ev.a));
  HPCCG_CUDA(cudaEventCreate(&ev.b));
  HPCCG_CUDA(cudaEventRecord(ev.a, s));
  HPCCG_TRY(cluster_dispatch(m, b, x, max_iter, tol, s, false, nullptr));
  HPCCG_CUDA(cudaEventRecord(ev.b, s));
  HPCCG_TRY(solve_readback(m, max_iter, niters, normr, hist_host, nullptr, s));
  float ms = 0.f;
  HPCCG_CUDA(cudaEventElapsedTime(&ms, ev.a, ev.b));
  if (loop_ms) *loop_ms = ms;
  return 0;
}

int cg_solve_io(hpccg_dev_matrix *m, const double *b, double *x, int max_iter, double tolerance, int *niters, double *normr,
                double *hist_host, double *times, double *loop_ms, int flags, cudaStream_t stream, const SolveIO *io) {
  if (!m) return fail(HPCCG_ERR_ARG, "cg_solve: null matrix");
  const RankContext &c = ctx();
  std::vector<SolveRank> rk{{m, b, x, c.rank}};
  if (c.size > 1) {
    if (!nccl_ready() || nccl_size() != c.size || nccl_rank() != c.rank)
      return fail(HPCCG_ERR_STATE, "hpccg_dev_cg_solve: rank context is %d/%d but no matching NCCL communicator (hpccg_nccl_init)",
                  c.rank, c.size);
    return cg_solve_impl(rk, c.size, true, max_iter, tolerance, niters, normr, hist_host, times, loop_ms, flags, stream, false, io);
  }
  return cg_solve_impl(rk, 1, false, max_iter, tolerance, niters, normr, hist_host, times, loop_ms, flags, stream, false, io);
}
}  // namespace hpccg

extern "C" {

int hpccg_dev_cg_solve(hpccg_dev_matrix *m, const double *b, double *x, int max_iter, double tolerance, int *niters,
                       double *normr, double *hist_host, double *times, double *loop_ms, int flags, void *stream) {
  if (!m) return fail(HPCCG_ERR_ARG, "hpccg_dev_cg_solve: null matrix");
  const RankContext &c = ctx();
  if ((flags & HPCCG_SOLVE_PERSISTENT) && c.size == 1 && !(flags & (HPCCG_SOLVE_TIMERS | HPCCG_SOLVE_UNFUSED))) {
    HPCCG_TRY(persistent_prepare(m));
    if (m->persist_state == 1) return cg_solve_persistent(m, b, x, max_iter, tolerance, niters, normr, hist_host, loop_ms, (cudaStream_t)stream);
  }
  if (!(flags & HPCCG_SOLVE_GRAPH) || (flags & HPCCG_SOLVE_TIMERS))
    return cg_solve_io(m, b, x, max_iter, tolerance, niters, normr, hist_host, times, loop_ms, flags, (cudaStream_t)stream, nullptr);
  const bool multi = c.size > 1;
  if (multi) {
    // One rank of a multi-GPU job: the peer-memory plane can be captured too -- its exchange stamps and reduction sequence
    // numbers are device state, the one NCCL call of a solve (the rendezvous gather) is capturable -- the NCCL plane is not
    // (send / recv on a side stream).  A rank that replays and a rank that launches directly run the same kernels and the
    // same collective, so ranks need not agree on which of the two they do.
    if (!nccl_ready() || nccl_size() != c.size || nccl_rank() != c.rank)
      return fail(HPCCG_ERR_STATE, "hpccg_dev_cg_solve: rank context is %d/%d but no matching NCCL communicator (hpccg_nccl_init)",
                  c.rank, c.size);
    bool peer = false;
    if (!(flags & (HPCCG_SOLVE_NCCL_ONLY | HPCCG_SOLVE_UNFUSED | HPCCG_SOLVE_EAGER_X))) {
      HPCCG_TRY(ensure_solver_workspace(m, std::max(max_iter, 1), c.size));
      HPCCG_TRY(peer_link_create(m, m->format == 1 || (m->format == 0 && use_tma_path(m->slots))));
      peer = m->peer_link != nullptr;
    }
    if (!peer)
      return cg_solve_io(m, b, x, max_iter, tolerance, niters, normr, hist_host, times, loop_ms, flags, (cudaStream_t)stream, nullptr);
  }
  std::vector<SolveRank> rk{{m, b, x, c.rank}};
  const int R = multi ? c.size : 1;

  // ---- CUDA-graph replay (launch-bound sizes): the launch sequence of a solve depends only on this key ----
  if (max_iter < 1) max_iter = 1;
  const bool same = m->graph_b == b && m->graph_x == x && m->graph_max_iter == max_iter && m->graph_tol == tolerance &&
                    m->graph_flags == flags;
  if (!same) {  // first solve with this key runs directly; a second one is worth a capture
    if (m->graph_exec) cudaGraphExecDestroy(m->graph_exec);
    m->graph_exec = nullptr;
    m->graph_b = b;
    m->graph_x = x;
    m->graph_max_iter = max_iter;
    m->graph_tol = tolerance;
    m->graph_flags = flags;
    return cg_solve_impl(rk, R, multi, max_iter, tolerance, niters, normr, hist_host, times, loop_ms, flags, (cudaStream_t)stream);
  }
  if (!m->graph_stream) HPCCG_CUDA(cudaStreamCreateWithFlags(&m->graph_stream, cudaStreamNonBlocking));
  cudaStream_t gs = m->graph_stream;
  if (!m->graph_exec) {
    HPCCG_TRY(ensure_solver_workspace(m, max_iter, R));
    HPCCG_CUDA(cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal));
    int rc = cg_solve_impl(rk, R, multi, max_iter, tolerance, nullptr, nullptr, nullptr, nullptr, nullptr, flags, gs, true);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(gs, &graph);
    if (rc) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    if (e != cudaSuccess) return fail_cuda(e, "cudaStreamEndCapture", __FILE__, __LINE__);
    e = cudaGraphInstantiate(&m->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
      m->graph_exec = nullptr;
      return fail_cuda(e, "cudaGraphInstantiate", __FILE__, __LINE__);
    }
  }
  // order the replay after whatever the caller enqueued on its stream (b, x uploads), and time it as a whole
  struct ReplayEvents {
    cudaEvent_t in = nullptr, t0 = nullptr, t1 = nullptr;
    ~ReplayEvents() {
      if (in) cudaEventDestroy(in);
      if (t0) cudaEventDestroy(t0);
      if (t1) cudaEventDestroy(t1);
    }
  } ev;
  HPCCG_CUDA(cudaEventCreateWithFlags(&ev.in, cudaEventDisableTiming));
  HPCCG_CUDA(cudaEventCreate(&ev.t0));
  HPCCG_CUDA(cudaEventCreate(&ev.t1));
  HPCCG_CUDA(cudaEventRecord(ev.in, (cudaStream_t)stream));
  HPCCG_CUDA(cudaStreamWaitEvent(gs, ev.in, 0));
  HPCCG_CUDA(cudaEventRecord(ev.t0, gs));
  HPCCG_CUDA(cudaGraphLaunch(m->graph_exec, gs));
  count_launch(3 * (max_iter - 1) + 5);
  HPCCG_CUDA(cudaEventRecord(ev.t1, gs));
  HPCCG_TRY(solve_readback(m, max_iter, niters, normr, hist_host, nullptr, gs));
  if (multi) {
    int perr = 0;
    HPCCG_CUDA(cudaMemcpy(&perr, &m->peer_link->error, sizeof(int), cudaMemcpyDeviceToHost));
    if (perr) return fail(HPCCG_ERR_COMM, "peer-memory wait timed out (%s): a rank of the job did not arrive",
                          perr == 2 ? "halo" : "scalar reduction");
  }
  float ms = 0.f;
  HPCCG_CUDA(cudaEventElapsedTime(&ms, ev.t0, ev.t1));
  if (loop_ms) *loop_ms = ms;
  return 0;
}

int hpccg_dev_cg_solve_group(int nranks, hpccg_dev_matrix *const *m, const double *const *b, double *const *x, int max_iter,
                             double tolerance, int *niters, double *normr, double *hist_host, double *loop_ms, int flags,
                             void *stream) {
  if (nranks < 1 || !m || !b || !x) return fail(HPCCG_ERR_ARG, "hpccg_dev_cg_solve_group: bad argument");
  std::vector<SolveRank> rk;
  for (int q = 0; q < nranks; ++q) rk.push_back({m[q], b[q], x[q], q});
  return cg_solve_impl(rk, nranks, false, max_iter, tolerance, niters, normr, hist_host, nullptr, loop_ms, flags,
                       (cudaStream_t)stream);
}

}  // extern "C"
