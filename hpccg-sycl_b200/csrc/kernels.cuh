// kernels.cuh -- the hand-written sm_100a kernels of the HPCCG hot path.
//
// All kernels are HBM-bandwidth bound (arithmetic intensity ~0.155 flop/B);
// nothing here is a dense contraction, so no tensor cores.  Design rules applied:
//   * SELL-C layout (C = 128 rows per slice, slot-major inside a slice, cg_state.hpp) so that a warp reads
//     32 (or 64) consecutive rows of one slot -- every matrix load is a fully coalesced 128-bit (vals) /
//     64-bit (cols) access -- and a slice is one contiguous block that a TMA bulk copy can move;
//   * the main SpMV streams the matrix with cp.async.bulk (TMA) into a shared-memory ring guarded by
//     mbarriers, so the bytes in flight per SM do not depend on registers or occupancy; the register-path
//     SpMV reads it with ld.global.nc.L1::no_allocate; either way the matrix stream never enters L1, which
//     is left to the gathered vector x (__ldg through L1/L2);
//   * persistent grid-stride tiles, grid = a multiple of the SM count, so reductions have a small,
//     fixed number of block partials and a deterministic two-pass finish;
//   * products and sums use __dmul_rn/__dadd_rn (never contracted into FMA) in the reference's
//     stored-entry order, which makes HPC_sparsemv and waxpby bit-identical to the reference's
//     g++ -O3 x86-64 build (it emits no FMA); only reduction order differs in ddot.
#pragma once

#include <cuda_runtime.h>

#include "cg_state.hpp"

namespace hpccg {

// What the thread that holds a finished reduction does with it.
enum FinishMode : int {
  FIN_STORE = 0,   // *out = sum                     (public ops, multi-rank local sums)
  FIN_INIT = 1,    // rtrans = sum, normr, hist[0]   (HPCCG.cpp:353-356)
  FIN_PAP = 2,     // alpha = rtrans / sum, niters=k (HPCCG.cpp:381-382,385)
  FIN_RR = 3,      // next iteration's rtrans, beta  (HPCCG.cpp:366-368,371) + loop condition (:358)
};

struct FinishParams {
  int mode;
  int k;            // iteration the reduction belongs to
  int last;         // FIN_RR: k+1 == max_iter, i.e. no further iteration will be enqueued
  int check_active; // kernel returns at once when st->active == 0
  double tol;
  CgState *st;
  double *out;      // FIN_STORE target
  double *hist;     // residual history (device), may be null
  PeerLink *peer;   // non-null: sum the block total over all ranks through peer mailboxes before the finish action
};

__device__ __forceinline__ void cg_finish(const FinishParams &fp, double sum) {
  CgState *st = fp.st;
  switch (fp.mode) {
    case FIN_STORE:
      *fp.out = sum;
      break;
    case FIN_INIT:
      st->rtrans = sum;
      st->normr = sqrt(sum);
      if (fp.hist) fp.hist[0] = st->normr;
      st->niters = 0;
      // condition of the first iteration: k=1 < max_iter (fp.last == 0) && normr > tolerance
      st->active = (!fp.last && st->normr > fp.tol) ? 1 : 0;
      if (st->active && fp.hist) fp.hist[1] = st->normr;  // iteration 1 prints the same normr (HPCCG.cpp:362,371)
      break;
    case FIN_PAP:
      st->pAp = sum;
      st->alpha = st->rtrans / sum;
      st->neg_alpha = -st->alpha;
      st->niters = fp.k;
      break;
    case FIN_RR: {
      // `sum` is r.r after iteration k's update = the rtrans iteration k+1 would compute first
      // (HPCCG.cpp:367).  Iteration k+1 runs iff k+1 < max_iter and normr_k > tolerance (:358).
      const bool next = !fp.last && (st->normr > fp.tol);
      if (next) {
        st->oldrtrans = st->rtrans;
        st->rtrans = sum;
        st->beta = sum / st->oldrtrans;
        st->normr = sqrt(sum);
        if (fp.hist) fp.hist[fp.k + 1] = st->normr;
      } else {
        st->active = 0;
      }
      break;
    }
  }
}

// ---- streaming loads -------------------------------------------------------------------------------
__device__ __forceinline__ double2 ld_stream_f64x2(const double *p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ int2 ld_stream_s32x2(const int *p) {
  int2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_stream_f64(const double *p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream_s32(const int *p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// ---- peer-memory primitives (system-scope release / acquire over NVLink) ----------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
constexpr unsigned long long kPeerTimeoutNs = 60ull * 1000ull * 1000ull * 1000ull;  // a dead peer must not hang the GPU

// Spins until *flag >= target.  Returns false on time-out (the caller records the error and carries on with NaN).
__device__ __forceinline__ bool peer_wait_ge(const unsigned long long *flag, unsigned long long target) {
  if (ld_acquire_sys(flag) >= target) return true;
  const unsigned long long t0 = global_ns();
  while (ld_acquire_sys(flag) < target) {
    __nanosleep(64);
    if (global_ns() - t0 > kPeerTimeoutNs) return false;
  }
  return true;
}

// All-reduce (sum in RANK ORDER, so every rank gets the same bits) of one double per rank, executed by the block that
// finished the local reduction: thread r stores this rank's value + stamp into rank r's mailbox over NVLink and then
// waits for rank r's value in the local mailbox.  Called by every thread of the block; result valid in thread 0.
__device__ __forceinline__ double peer_allreduce(PeerLink *pl, double local, double *smem_vals /* >= kMaxRanks */) {
  __shared__ unsigned long long s_seq;
  __shared__ double s_local;
  if (threadIdx.x == 0) {
    s_seq = ++pl->reduce_seq;
    s_local = local;
  }
  __syncthreads();
  const int t = threadIdx.x;
  if (t < pl->size) {
    const unsigned long long seq = s_seq;
    const int slot = (int)(seq % kMailSlots);
    Mailbox *dst = pl->box[t];
    *reinterpret_cast<volatile double *>(&dst->value[slot][pl->rank]) = s_local;
    st_release_sys(&dst->seq[slot][pl->rank], seq);  // release: the value store above is ordered before the stamp
    Mailbox *own = pl->box[pl->rank];
    double v;
    // once a wait has timed out the job is lost: later waits give up at once, so a dead peer costs one time-out, not one per kernel
    if (*reinterpret_cast<volatile int *>(&pl->error) == 0 && peer_wait_ge(&own->seq[slot][t], seq)) {
      v = *reinterpret_cast<volatile double *>(&own->value[slot][t]);
    } else {
      pl->error = 1;
      v = __longlong_as_double(0x7ff8000000000000LL);
    }
    smem_vals[t] = v;
  }
  __syncthreads();
  double g = 0.0;
  if (t == 0) {
    g = smem_vals[0];
    for (int r = 1; r < pl->size; ++r) g = __dadd_rn(g, smem_vals[r]);
  }
  return g;
}

// ---- deterministic block reduction + last-block finish ---------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Sum over the block in a fixed order; result valid in thread 0.
__device__ __forceinline__ double block_sum(double v, double *smem /* kThreads/32 doubles */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (warp == 0) {
    r = (lane < (kThreads >> 5)) ? smem[lane] : 0.0;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

// Publishes this block's partial; the block that takes the last ticket sums all partials in index
// order (fixed tree, independent of which block happens to be last) and runs the finish action.
// NT = threads of the calling block; smem holds NT/32 doubles.  The order in which the partials are combined
// depends only on (total_partials, NT), never on block scheduling.
template <int NT>
__device__ __forceinline__ void publish_and_finish_n(double block_total, double *partials, int partial_index,
                                                     int total_partials, unsigned *counter, const FinishParams &fp,
                                                     double *smem) {
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    partials[partial_index] = block_total;
    __threadfence();
    const unsigned ticket = atomicInc(counter, (unsigned)(total_partials - 1));  // wraps to 0 for the next launch
    s_last = (ticket == (unsigned)(total_partials - 1));
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double acc = 0.0;
  for (int i = threadIdx.x; i < total_partials; i += NT) acc = __dadd_rn(acc, __ldcg(partials + i));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  acc = warp_sum(acc);
  if (lane == 0) smem[warp] = acc;
  __syncthreads();
  double r = 0.0;
  if (warp == 0) {
    r = (lane < NT / 32) ? smem[lane] : 0.0;
    r = warp_sum(r);
  }
  if (fp.peer) {  // multi-GPU: the global sum, over peer memory, inside this kernel
    __shared__ double s_ranks[kMaxRanks];
    r = peer_allreduce(fp.peer, r, s_ranks);
  }
  if (threadIdx.x == 0) cg_finish(fp, r);
}

__device__ __forceinline__ void publish_and_finish(double block_total, double *partials, int partial_index,
                                                   int total_partials, unsigned *counter, const FinishParams &fp,
                                                   double *smem) {
  publish_and_finish_n<kThreads>(block_total, partials, partial_index, total_partials, counter, fp, smem);
}

// ---- HPC_sparsemv.cpp:68-89, register path (any slot count, any row range) -----------------------------
// SLOTS > 0: compile-time slot count (27, 7), fully unrolled so all matrix loads of a row pair are in
// flight before the first gather.  SLOTS == 0: run-time slot count.  RPT rows per thread (2 = 128-bit
// value loads).  Rows [row_begin,row_end) are processed; tiles start at the even row below row_begin.
template <int SLOTS, int RPT, bool DOT>
__global__ void __launch_bounds__(kThreads)
spmv_ell_kernel(const double *__restrict__ vals, const int *__restrict__ cols, long long npad, int slots_rt,
                const double *__restrict__ x, double *__restrict__ y, int row_begin, int row_end, int tiles,
                double *partials, int partial_offset, int total_partials, unsigned *counter, FinishParams fp) {
  __shared__ double smem[kThreads / 32];
  if (fp.check_active && fp.st->active == 0) return;
  const int slots = SLOTS > 0 ? SLOTS : slots_rt;
  constexpr int kTileRows = kThreads * RPT;
  const int base = row_begin & ~(RPT - 1);
  double dot = 0.0;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int r0 = base + tile * kTileRows + threadIdx.x * RPT;
    if (r0 >= row_end) continue;
    // SELL-C addressing: both rows of a pair are in the same slice (r0 and C are even)
    const long long off0 = (long long)(r0 / kSliceRows) * slots * kSliceRows + (r0 % kSliceRows);
    if (RPT == 2) {
      double s0 = 0.0, s1 = 0.0;
      const double *vp = vals + off0;
      const int *cp = cols + off0;
#pragma unroll
      for (int j = 0; j < slots; ++j) {
        const double2 v = ld_stream_f64x2(vp + j * kSliceRows);
        const int2 c = ld_stream_s32x2(cp + j * kSliceRows);
        if (c.x >= 0) s0 = __dadd_rn(s0, __dmul_rn(v.x, __ldg(x + c.x)));
        if (c.y >= 0) s1 = __dadd_rn(s1, __dmul_rn(v.y, __ldg(x + c.y)));
      }
      const bool in0 = r0 >= row_begin, in1 = r0 + 1 < row_end;
      if (in0 && in1) {
        *reinterpret_cast<double2 *>(y + r0) = make_double2(s0, s1);
      } else {
        if (in0) y[r0] = s0;
        if (in1) y[r0 + 1] = s1;
      }
      if (DOT) {
        if (in0) dot = __dadd_rn(dot, __dmul_rn(__ldg(x + r0), s0));
        if (in1) dot = __dadd_rn(dot, __dmul_rn(__ldg(x + r0 + 1), s1));
      }
    } else {
      double s0 = 0.0;
      const double *vp = vals + off0;
      const int *cp = cols + off0;
#pragma unroll
      for (int j = 0; j < slots; ++j) {
        const double v = ld_stream_f64(vp + j * kSliceRows);
        const int c = ld_stream_s32(cp + j * kSliceRows);
        if (c >= 0) s0 = __dadd_rn(s0, __dmul_rn(v, __ldg(x + c)));
      }
      y[r0] = s0;
      if (DOT) dot = __dadd_rn(dot, __dmul_rn(__ldg(x + r0), s0));
    }
  }
  if (DOT) {
    const double total = block_sum(dot, smem);
    publish_and_finish(total, partials, partial_offset + blockIdx.x, total_partials, counter, fp, smem);
  }
}

// ---- HPC_sparsemv.cpp:68-89 on the ragged SELL-C-sigma mirror (format 2) --------------------------------------------------
// One thread per row position, one warp per slice: a slice (kRaggedRows = 32 positions) has its own slot count and element
// offset, so a matrix with a few long rows pads only the slices those rows fall into (C = 32 instead of the uniform layout's
// 128 because a power-law tail needs it: 1.3 x the stored entries instead of 3.5 x on the test matrix).  Entries are summed in stored order (padding slots carry column -1
// and are skipped), i.e. the result of a row is bit-identical to every other path; with sigma-sorting the row of a position
// is perm[position].  Positions [pos_begin, pos_end) are processed (contiguous row ranges need sigma = 1).
template <bool DOT>
__global__ void __launch_bounds__(kThreads)
spmv_sell_ragged_kernel(const double *__restrict__ vals, const int *__restrict__ cols, const int *__restrict__ slice_slots,
                        const long long *__restrict__ slice_off, const int *__restrict__ perm, int n,
                        const double *__restrict__ x, double *__restrict__ y, int pos_begin, int pos_end, double *partials,
                        int partial_offset, int total_partials, unsigned *counter, FinishParams fp) {
  __shared__ double smem[kThreads / 32];
  if (fp.check_active && fp.st->active == 0) return;
  double dot = 0.0;
  const int base = pos_begin & ~(kRaggedRows - 1);
  for (long long pos = base + (long long)blockIdx.x * kThreads + threadIdx.x; pos < pos_end; pos += (long long)gridDim.x * kThreads) {
    if (pos < pos_begin) continue;
    const int slice = (int)(pos / kRaggedRows), l = (int)(pos % kRaggedRows);
    const int ns = __ldg(slice_slots + slice);
    const long long off = __ldg(slice_off + slice) + l;
    double sum = 0.0;
    int j = 0;
    for (; j + 4 <= ns; j += 4) {  // four entries' loads in flight, summed in order
      int c[4];
      double v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        c[u] = ld_stream_s32(cols + off + (long long)(j + u) * kRaggedRows);
        v[u] = ld_stream_f64(vals + off + (long long)(j + u) * kRaggedRows);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (c[u] >= 0) sum = __dadd_rn(sum, __dmul_rn(v[u], __ldg(x + c[u])));
    }
    for (; j < ns; ++j) {
      const int c = ld_stream_s32(cols + off + (long long)j * kRaggedRows);
      const double v = ld_stream_f64(vals + off + (long long)j * kRaggedRows);
      if (c >= 0) sum = __dadd_rn(sum, __dmul_rn(v, __ldg(x + c)));
    }
    const int row = perm ? __ldg(perm + pos) : (int)pos;
    if (row >= 0 && row < n) {
      y[row] = sum;
      if (DOT) dot = __dadd_rn(dot, __dmul_rn(__ldg(x + row), sum));
    }
  }
  if (DOT) {
    const double total = block_sum(dot, smem);
    publish_and_finish(total, partials, partial_offset + blockIdx.x, total_partials, counter, fp, smem);
  }
}

// ---- HPC_sparsemv.cpp:68-89, TMA path: the main SpMV --------------------------------------------------------
// One stage = SPS consecutive slices = one contiguous block of the vals array and one of the cols array,
// fetched by a single elected thread with two cp.async.bulk (global -> shared, mbarrier complete_tx) and an L2
// evict_first hint (the matrix is read once per SpMV; L2 is for the gathered vector).  NSTAGES stages form a ring:
// while the CTA's SPS*128 threads (one row each) compute from stage i, stages i+1 .. i+NSTAGES-1 are in flight.
// Row sums are accumulated in slot order with un-contracted mul/add, i.e. bit-identical to the register path.
// Rows outside [row_begin,row_end) of a stage are computed but neither stored nor added to the dot product, so
// row ranges need not be stage-aligned.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar,
                                             unsigned long long policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// Multi-GPU: rows [0,interior_begin) and [interior_end,n) reference halo columns (>= n) that the neighbours deliver
// over NVLink WHILE this kernel runs.  The CTAs walk the stages in a rotated order -- interior first, halo-touching
// stages last -- and a CTA waits for the neighbours' stamps (PeerLink mailbox) only when it reaches such a stage, so
// the exchange is hidden behind the interior rows inside ONE launch.  link == nullptr: single GPU, natural order.
struct SpmvHalo {
  PeerLink *link;
  int n;               // local_nrow: columns >= n are halo entries (read with ld.global.cg, never through L1)
  int interior_begin;  // first row that references no halo column
  int interior_end;
  int exch_idx;        // which exchange of this solve delivers the planes (exchange_stamp)
};

__device__ __forceinline__ void tma_prefetch_l2(const void *src_gmem, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}

template <int SLOTS, int SPS, int NSTAGES>
struct SpmvTmaCfg {
  static constexpr int kRows = SPS * kSliceRows;            // rows per stage = threads per CTA
  static constexpr int kValBytes = SLOTS * kRows * 8;
  static constexpr int kColBytes = SLOTS * kRows * 4;
  static constexpr int kSmemBytes = NSTAGES * (kValBytes + kColBytes) + NSTAGES * 8 + 64;
};

template <int SLOTS, int SPS, int NSTAGES, bool DOT>
__global__ void __launch_bounds__(SPS *kSliceRows, 1)
spmv_sell_tma_kernel(const double *__restrict__ vals, const int *__restrict__ cols, const double *__restrict__ x,
                     double *__restrict__ y, int row_begin, int row_end, int stage_begin, int stage_end, double *partials,
                     int partial_offset, int total_partials, unsigned *counter, FinishParams fp, SpmvHalo halo, int l2_ahead) {
  using Cfg = SpmvTmaCfg<SLOTS, SPS, NSTAGES>;
  constexpr int kRows = Cfg::kRows;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *sv = reinterpret_cast<double *>(smem_raw);
  int *sc = reinterpret_cast<int *>(smem_raw + NSTAGES * Cfg::kValBytes);
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem_raw + NSTAGES * (Cfg::kValBytes + Cfg::kColBytes));
  __shared__ double red[kRows / 32];
  if (fp.check_active && fp.st->active == 0) return;

  const int tid = threadIdx.x;
  // logical stage g = blockIdx.x + i * gridDim.x of this launch's T stages; physical stage = stage_begin + (g + rot) % T
  const int T = stage_end - stage_begin;
  const int my_count = (int)blockIdx.x < T ? (T - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  int rot = 0;
  if (halo.link) {
    // first stage that lies completely inside the interior; the order becomes interior.., upper halo rows, lower halo rows
    const int s0 = (halo.interior_begin + kRows - 1) / kRows - stage_begin;
    rot = (s0 > 0 && s0 < T) ? s0 : 0;
  }
  auto phys = [&](int i) {
    int g = (int)blockIdx.x + i * (int)gridDim.x + rot;
    if (g >= T) g -= T;
    return stage_begin + g;
  };
  unsigned long long policy = 0;
  if (tid == 0) {
    for (int s = 0; s < NSTAGES; ++s) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  }
  __syncthreads();
  auto issue = [&](int i) {  // elected thread only
    const int st = i % NSTAGES;
    const long long stage = phys(i);
    mbar_expect_tx(bars + st, Cfg::kValBytes + Cfg::kColBytes);
    tma_bulk_g2s(sv + (size_t)st * SLOTS * kRows, vals + stage * SLOTS * kRows, Cfg::kValBytes, bars + st, policy);
    tma_bulk_g2s(sc + (size_t)st * SLOTS * kRows, cols + stage * SLOTS * kRows, Cfg::kColBytes, bars + st, policy);
    // optional: pull the stage `l2_ahead` turns later into L2 now, so DRAM sees requests beyond the shared-memory ring
    if (l2_ahead > 0 && i + l2_ahead < my_count) {
      const long long ahead = phys(i + l2_ahead);
      tma_prefetch_l2(vals + ahead * SLOTS * kRows, Cfg::kValBytes);
      tma_prefetch_l2(cols + ahead * SLOTS * kRows, Cfg::kColBytes);
    }
  };
  if (tid == 0)
    for (int i = 0; i < NSTAGES && i < my_count; ++i) issue(i);

  // thread -> (slice of the stage, row of the slice): shared-memory offset of slot j is soff + j * kSliceRows
  const int soff = (tid / kSliceRows) * SLOTS * kSliceRows + (tid % kSliceRows);
  double dot = 0.0;
  bool halo_ready = (halo.link == nullptr);
  for (int i = 0; i < my_count; ++i) {
    const int st = i % NSTAGES;
    const int stage = phys(i);
    const int row = stage * kRows + tid;
    const bool touches_halo = halo.link && (stage * kRows < halo.interior_begin || (stage + 1) * kRows > halo.interior_end);
    if (touches_halo && !halo_ready) {
      // the neighbours' planes for THIS exchange (stamp = solve epoch | exchange index) must have landed
      if (tid < halo.link->nnb) {
        const Mailbox *own = halo.link->box[halo.link->rank];
        if (*reinterpret_cast<volatile int *>(&halo.link->error) != 0 || !peer_wait_ge(&own->halo_seq[tid], exchange_stamp(halo.link->epoch, halo.exch_idx)))
          halo.link->error = 2;
      }
      __syncthreads();
      halo_ready = true;
    }
    mbar_wait(bars + st, (unsigned)(i / NSTAGES) & 1u);
    const double *v = sv + (size_t)st * SLOTS * kRows + soff;
    const int *c = sc + (size_t)st * SLOTS * kRows + soff;
    int ci[SLOTS];
    double xv[SLOTS];
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) ci[j] = c[j * kSliceRows];
    // all SLOTS gathers are issued back to back and unconditionally (a padding slot reads x[0] and is discarded
    // below), so no predicate is live across the loads and the compiler keeps every gather of the row in flight
    if (!touches_halo) {
#pragma unroll
      for (int j = 0; j < SLOTS; ++j) xv[j] = __ldg(x + max(ci[j], 0));
    } else {
      // halo entries were written by another GPU during this kernel and must be read at L2 (ld.global.cg), never through
      // L1.  The whole stage gathers with ld.cg: choosing the load per entry (ld.cg above n, ldg below) issues BOTH loads
      // predicated, which made the halo-touching stages -- 3 % of the stages of a 128-plane slab -- twice as expensive
#pragma unroll
      for (int j = 0; j < SLOTS; ++j) xv[j] = __ldcg(x + max(ci[j], 0));
    }
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) {
      const double t = __dadd_rn(sum, __dmul_rn(v[j * kSliceRows], xv[j]));
      sum = ci[j] >= 0 ? t : sum;  // padding slots do not exist in the reference's row (HPC_sparsemv.cpp:83-86)
    }
    if (row >= row_begin && row < row_end) {
      y[row] = sum;
      if (DOT) dot = __dadd_rn(dot, __dmul_rn(__ldg(x + row), sum));
    }
    __syncthreads();  // every thread has read stage st: it may be overwritten
    if (tid == 0 && i + NSTAGES < my_count) issue(i + NSTAGES);
  }
  if (DOT) {
    // block reduction in a fixed order (warp shuffle tree, then warp 0 over the warp sums)
    const int lane = tid & 31, warp = tid >> 5;
    double w = warp_sum(dot);
    if (lane == 0) red[warp] = w;
    __syncthreads();
    double total = 0.0;
    if (warp == 0) {
      total = lane < kRows / 32 ? red[lane] : 0.0;
      total = warp_sum(total);
    }
    publish_and_finish_n<kRows>(total, partials, partial_offset + blockIdx.x, total_partials, counter, fp, red);
  }
}

// ---- pattern-coded SpMV (opt-in format) -------------------------------------------------------
// One 16-bit pattern id per row instead of 12 bytes per stored entry: HPC_sparsemv reads 2 + 16 bytes per row.
//   * warp-uniform fast path (every lane's row has pattern 0, the most frequent one): the pattern's values and deltas
//     are constant-bank operands (Pattern0 is a __grid_constant__ parameter, the loops are fully unrolled), so a stored
//     entry costs an add, an address computation, the gather of x, a multiply and an add -- no matrix load at all;
//   * otherwise each lane reads its pattern's entries from the pattern table (L1/L2-resident, broadcast when lanes
//     share a pattern).
// The arithmetic is the same un-contracted mul/add in stored order on the same values: bit-identical to the SELL paths.
// Rows are walked in tiles of kThreads rows, persistent CTAs; with a PeerLink the tiles are rotated (interior first) and a
// CTA waits for the neighbours' halo stamps only when it reaches a tile that references halo columns.
template <int SLOTS, bool DOT>
__global__ void __launch_bounds__(kThreads)
spmv_pattern_kernel(const unsigned short *__restrict__ pat_id, const double *__restrict__ pat_val,
                    const int *__restrict__ pat_delta, const int *__restrict__ pat_len,
                    const __grid_constant__ Pattern0 p0, const double *__restrict__ x, double *__restrict__ y, int n,
                    int row_begin, int row_end, int tile_begin, int tile_end, double *partials, int partial_offset,
                    int total_partials, unsigned *counter, FinishParams fp, SpmvHalo halo) {
  __shared__ double smem[kThreads / 32];
  if (fp.check_active && fp.st->active == 0) return;
  const int tid = threadIdx.x;
  const int T = tile_end - tile_begin;
  int rot = 0;
  if (halo.link) {
    const int s0 = (halo.interior_begin + kThreads - 1) / kThreads - tile_begin;
    rot = (s0 > 0 && s0 < T) ? s0 : 0;
  }
  double dot = 0.0;
  bool halo_ready = (halo.link == nullptr);
  const bool fast_ok = (p0.len == SLOTS);  // the unrolled fast path assumes a full row
  int it = 0;
  for (int g0 = blockIdx.x; g0 < T; g0 += gridDim.x, ++it) {
    int g = g0 + rot;
    if (g >= T) g -= T;
    const int tile = tile_begin + g;
    // The warps of a CTA run through the tiles without synchronising.  Rows at the ends of a grid line take the slower
    // table path, and they sit at fixed positions of a tile, so the warp -> 32-row-group assignment is rotated from tile
    // to tile; otherwise the same two warps would do all the slow rows and the other six would wait for them at the end.
    const int row = tile * kThreads + ((tid + 32 * it) & (kThreads - 1));
    const bool touches_halo = halo.link && (tile * kThreads < halo.interior_begin || (tile + 1) * kThreads > halo.interior_end);
    if (touches_halo && !halo_ready) {
      if (tid < halo.link->nnb) {
        const Mailbox *own = halo.link->box[halo.link->rank];
        if (*reinterpret_cast<volatile int *>(&halo.link->error) != 0 || !peer_wait_ge(&own->halo_seq[tid], exchange_stamp(halo.link->epoch, halo.exch_idx)))
          halo.link->error = 2;
      }
      __syncthreads();
      halo_ready = true;
    }
    const bool in_range = row >= row_begin && row < row_end;  // row_end <= n
    const int pid = in_range ? (int)pat_id[row] : 0xFFFF;
    double sum = 0.0;
    if (fast_ok && __all_sync(0xffffffffu, pid == 0) && !touches_halo) {
      // every lane's row is a full-length pattern-0 row: values and deltas are constant-bank operands, no predicates
      // (a divergent per-lane version of this test measured 4 % slower at 512^3 and 19 % slower at 256^3)
      double xv[SLOTS];
#pragma unroll
      for (int j = 0; j < SLOTS; ++j) xv[j] = __ldg(x + (row + p0.delta[j]));
#pragma unroll
      for (int j = 0; j < SLOTS; ++j) sum = __dadd_rn(sum, __dmul_rn(p0.value[j], xv[j]));
    } else if (pid != 0xFFFF) {
      const int len = __ldg(pat_len + pid);
      const double *pv = pat_val + (size_t)pid * SLOTS;
      const int *pd = pat_delta + (size_t)pid * SLOTS;
      constexpr int CH = (SLOTS % 9 == 0) ? 9 : SLOTS;
#pragma unroll
      for (int j0 = 0; j0 < SLOTS; j0 += CH) {
        double vj[CH], xv[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          const bool live = j0 + j < len;
          const int c = live ? row + __ldg(pd + j0 + j) : row;
          vj[j] = live ? __ldg(pv + j0 + j) : 0.0;
          xv[j] = (touches_halo && c >= halo.n) ? __ldcg(x + c) : __ldg(x + c);
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          const double t = __dadd_rn(sum, __dmul_rn(vj[j], xv[j]));
          sum = j0 + j < len ? t : sum;
        }
      }
    }
    if (in_range) {
      y[row] = sum;
      if (DOT) dot = __dadd_rn(dot, __dmul_rn(__ldg(x + row), sum));
    }
  }
  if (DOT) {
    const double total = block_sum(dot, smem);
    publish_and_finish(total, partials, partial_offset + blockIdx.x, total_partials, counter, fp, smem);
  }
}

// ---- pattern encoder (one-off, hpccg_dev_matrix_compress) ---------------------------------------------------------
__device__ __forceinline__ unsigned long long pattern_hash(const double *__restrict__ vals, const int *__restrict__ cols,
                                                           long long row, int slots, int *len_out) {
  unsigned long long h = 0x9E3779B97F4A7C15ull;
  int len = 0;
  for (int j = 0; j < slots; ++j) {
    const long long o = sell_offset(row, j, slots);
    const int c = cols[o];
    if (c < 0) continue;
    ++len;
    h = (h ^ (unsigned long long)__double_as_longlong(vals[o])) * 0xff51afd7ed558ccdull;
    h ^= h >> 32;
    h = (h ^ (unsigned long long)(unsigned int)(c - (int)row) ^ ((unsigned long long)j << 40)) * 0xc4ceb9fe1a85ec53ull;
    h ^= h >> 29;
  }
  *len_out = len;
  h ^= (unsigned long long)len << 56;
  return h | 1ull;  // 0 is the empty key
}

// Pass A: every row claims (or finds) the slot of its pattern hash in an open-addressing table.
__global__ void __launch_bounds__(kThreads)
pattern_insert_kernel(const double *__restrict__ vals, const int *__restrict__ cols, int slots, int n,
                      unsigned long long *keys, unsigned mask, int *overflow) {
  for (long long row = (long long)blockIdx.x * kThreads + threadIdx.x; row < n; row += (long long)gridDim.x * kThreads) {
    int len;
    const unsigned long long h = pattern_hash(vals, cols, row, slots, &len);
    unsigned slot = (unsigned)(h >> 17) & mask;
    int probes = 0;
    for (;;) {
      const unsigned long long k = *reinterpret_cast<volatile unsigned long long *>(keys + slot);
      if (k == h) break;
      if (k == 0) {
        const unsigned long long old = atomicCAS(keys + slot, 0ull, h);
        if (old == 0 || old == h) break;
      }
      slot = (slot + 1) & mask;
      if (++probes > 256) {
        *overflow = 1;
        break;
      }
    }
  }
}

// Pass B: occupied slots get consecutive ids.
__global__ void __launch_bounds__(kThreads)
pattern_number_kernel(const unsigned long long *__restrict__ keys, unsigned table_size, int *ids, int *count) {
  for (unsigned s = blockIdx.x * kThreads + threadIdx.x; s < table_size; s += gridDim.x * kThreads)
    ids[s] = keys[s] ? atomicAdd(count, 1) : -1;
}

// Pass C: rows look their id up, leave a representative row per id and count the rows per id (warp-aggregated).
__global__ void __launch_bounds__(kThreads)
pattern_assign_kernel(const double *__restrict__ vals, const int *__restrict__ cols, int slots, int n, long long npad,
                      const unsigned long long *__restrict__ keys, const int *__restrict__ ids, unsigned mask,
                      unsigned short *__restrict__ pat_id, int *rep_row, unsigned long long *freq) {
  for (long long base = (long long)blockIdx.x * kThreads; base < npad; base += (long long)gridDim.x * kThreads) {
    const long long row = base + threadIdx.x;
    int id = -1;
    if (row < n) {
      int len;
      const unsigned long long h = pattern_hash(vals, cols, row, slots, &len);
      unsigned slot = (unsigned)(h >> 17) & mask;
      while (keys[slot] != h) slot = (slot + 1) & mask;
      id = ids[slot];
      rep_row[id] = (int)row;  // any row of the pattern will do (verified afterwards)
    }
    if (row < npad) pat_id[row] = id >= 0 ? (unsigned short)id : (unsigned short)0xFFFF;
    const unsigned peers = __match_any_sync(__activemask(), id);
    if (id >= 0 && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(freq + id, (unsigned long long)__popc(peers));
  }
}

// Pass D: the pattern table from the representative rows.
__global__ void __launch_bounds__(kThreads)
pattern_fill_kernel(const double *__restrict__ vals, const int *__restrict__ cols, int slots, int npat,
                    const int *__restrict__ rep_row, double *pat_val, int *pat_delta, int *pat_len) {
  for (int id = blockIdx.x * kThreads + threadIdx.x; id < npat; id += gridDim.x * kThreads) {
    const long long row = rep_row[id];
    int len = 0;
    for (int j = 0; j < slots; ++j) {
      const long long o = sell_offset(row, j, slots);
      const int c = cols[o];
      if (c < 0) continue;
      pat_val[(size_t)id * slots + len] = vals[o];
      pat_delta[(size_t)id * slots + len] = c - (int)row;
      ++len;
    }
    pat_len[id] = len;
    for (int j = len; j < slots; ++j) {
      pat_val[(size_t)id * slots + j] = 0.0;
      pat_delta[(size_t)id * slots + j] = 0;
    }
  }
}

// Pass E: exact check of every row against its pattern (a 64-bit hash collision must not corrupt the matrix), with the
// final relabelling (the most frequent pattern becomes id 0) applied on the fly.
__global__ void __launch_bounds__(kThreads)
pattern_verify_kernel(const double *__restrict__ vals, const int *__restrict__ cols, int slots, int n,
                      unsigned short *__restrict__ pat_id, int swap_a, int swap_b, const double *__restrict__ pat_val,
                      const int *__restrict__ pat_delta, const int *__restrict__ pat_len, int *mismatch) {
  for (long long row = (long long)blockIdx.x * kThreads + threadIdx.x; row < n; row += (long long)gridDim.x * kThreads) {
    int id = pat_id[row];
    if (id == swap_a) id = swap_b;
    else if (id == swap_b) id = swap_a;
    pat_id[row] = (unsigned short)id;
    int len = 0;
    bool ok = true;
    for (int j = 0; j < slots; ++j) {
      const long long o = sell_offset(row, j, slots);
      const int c = cols[o];
      if (c < 0) continue;
      ok = ok && len < slots && __double_as_longlong(pat_val[(size_t)id * slots + len]) == __double_as_longlong(vals[o]) &&
           pat_delta[(size_t)id * slots + len] == c - (int)row;
      ++len;
    }
    if (!ok || len != pat_len[id]) *mismatch = 1;
  }
}

// ---- ddot.cpp:60-74 ----------------------------------------------------------------------------------
template <bool SAME>
__global__ void __launch_bounds__(kThreads)
dot_kernel(int n, const double *__restrict__ x, const double *__restrict__ y, double *partials, int total_partials,
           unsigned *counter, FinishParams fp) {
  __shared__ double smem[kThreads / 32];
  if (fp.check_active && fp.st->active == 0) return;
  double acc = 0.0;
  const int npair = n >> 1;
  const bool aligned = ((reinterpret_cast<size_t>(x) | reinterpret_cast<size_t>(y)) & 15) == 0;
  if (aligned) {
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < npair; i += gridDim.x * kThreads) {
      const double2 a = ld_stream_f64x2(x + 2 * (long long)i);
      const double2 b = SAME ? a : ld_stream_f64x2(y + 2 * (long long)i);
      acc = __dadd_rn(acc, __dmul_rn(a.x, b.x));
      acc = __dadd_rn(acc, __dmul_rn(a.y, b.y));
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) acc = __dadd_rn(acc, __dmul_rn(x[n - 1], y[n - 1]));
  } else {
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads)
      acc = __dadd_rn(acc, __dmul_rn(x[i], y[i]));
  }
  const double total = block_sum(acc, smem);
  publish_and_finish(total, partials, blockIdx.x, total_partials, counter, fp, smem);
}

// ---- waxpby.cpp:69-93 -------------------------------------------------------------------------------
// BRANCH 0: alpha==1 -> x + beta*y ; 1: beta==1 -> alpha*x + y ; 2: general.  Element-wise, so the
// aliasing the reference uses (w==x, w==y, x==y) is safe; no __restrict__ on purpose.
template <int BRANCH>
__device__ __forceinline__ double waxpby_elem(double alpha, double x, double beta, double y) {
  if (BRANCH == 0) return __dadd_rn(x, __dmul_rn(beta, y));
  if (BRANCH == 1) return __dadd_rn(__dmul_rn(alpha, x), y);
  return __dadd_rn(__dmul_rn(alpha, x), __dmul_rn(beta, y));
}

template <int BRANCH>
__global__ void __launch_bounds__(kThreads)
waxpby_kernel(int n, double alpha, const double *x, double beta, const double *beta_dev, const double *y, double *w,
              const CgState *st_check) {
  if (st_check && st_check->active == 0) return;
  if (beta_dev) beta = *beta_dev;
  const bool aligned =
      ((reinterpret_cast<size_t>(x) | reinterpret_cast<size_t>(y) | reinterpret_cast<size_t>(w)) & 15) == 0;
  if (aligned) {
    const int npair = n >> 1;
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < npair; i += gridDim.x * kThreads) {
      const double2 a = *reinterpret_cast<const double2 *>(x + 2 * (long long)i);
      const double2 b = *reinterpret_cast<const double2 *>(y + 2 * (long long)i);
      double2 o;
      o.x = waxpby_elem<BRANCH>(alpha, a.x, beta, b.x);
      o.y = waxpby_elem<BRANCH>(alpha, a.y, beta, b.y);
      *reinterpret_cast<double2 *>(w + 2 * (long long)i) = o;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) w[n - 1] = waxpby_elem<BRANCH>(alpha, x[n - 1], beta, y[n - 1]);
  } else {
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads)
      w[i] = waxpby_elem<BRANCH>(alpha, x[i], beta, y[i]);
  }
}

// ---- HPCCG.cpp:383-384 fused with the r.r of :367 ------------------------------------------------------
// x = x + alpha*p ; r = r + (-alpha)*Ap ; partial r.r.  alpha is read from device memory.
__global__ void __launch_bounds__(kThreads)
update_xr_dot_kernel(int n, const double *alpha_dev, const double *__restrict__ p, const double *__restrict__ Ap,
                     double *__restrict__ x, double *__restrict__ r, double *partials, int total_partials,
                     unsigned *counter, FinishParams fp) {
  __shared__ double smem[kThreads / 32];
  if (fp.check_active && fp.st->active == 0) return;
  const double alpha = *alpha_dev;
  const double nalpha = -alpha;
  double acc = 0.0;
  const int npair = n >> 1;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < npair; i += gridDim.x * kThreads) {
    const long long e = 2 * (long long)i;
    const double2 pv = ld_stream_f64x2(p + e);
    const double2 av = ld_stream_f64x2(Ap + e);
    double2 xv = *reinterpret_cast<const double2 *>(x + e);
    double2 rv = *reinterpret_cast<const double2 *>(r + e);
    xv.x = __dadd_rn(xv.x, __dmul_rn(alpha, pv.x));
    xv.y = __dadd_rn(xv.y, __dmul_rn(alpha, pv.y));
    rv.x = __dadd_rn(rv.x, __dmul_rn(nalpha, av.x));
    rv.y = __dadd_rn(rv.y, __dmul_rn(nalpha, av.y));
    *reinterpret_cast<double2 *>(x + e) = xv;
    *reinterpret_cast<double2 *>(r + e) = rv;
    acc = __dadd_rn(acc, __dmul_rn(rv.x, rv.x));
    acc = __dadd_rn(acc, __dmul_rn(rv.y, rv.y));
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const int i = n - 1;
    const double xn = __dadd_rn(x[i], __dmul_rn(alpha, p[i]));
    const double rn = __dadd_rn(r[i], __dmul_rn(nalpha, Ap[i]));
    x[i] = xn;
    r[i] = rn;
    acc = __dadd_rn(acc, __dmul_rn(rn, rn));
  }
  const double total = block_sum(acc, smem);
  publish_and_finish(total, partials, blockIdx.x, total_partials, counter, fp, smem);
}

// ---- the loop's two vector kernels with the x update deferred by one kernel -----------------------------------------
// HPCCG.cpp:383 (x += alpha p) only needs p_k, which the NEXT p-update reads anyway, so it moves there: the kernel after the
// SpMV updates r and reduces r.r (24 B/row), the p-update applies the previous iteration's x update and forms the new p
// (40 B/row) -- 64 B/row instead of 48 + 24, every element still computed by the same two un-contracted operations.
// The update of the last executed iteration is applied by x_fixup_kernel when the loop has ended.
// 256-bit accesses (LDG.256 / STG.256, new with sm_100) for the two vector kernels of the loop: 4 doubles per thread and
// stream, half the memory instructions of the 128-bit form.  Pointers must be 32-byte aligned (the launcher checks).
struct double4v {
  double a, b, c, d;
};
__device__ __forceinline__ double4v ld_stream_f64x4(const double *p) {
  double4v v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.a), "=d"(v.b), "=d"(v.c), "=d"(v.d) : "l"(p));
  return v;
}
__device__ __forceinline__ double4v ld_f64x4(const double *p) {
  double4v v;
  asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.a), "=d"(v.b), "=d"(v.c), "=d"(v.d) : "l"(p) : "memory");
  return v;
}
// read-only path, allocating in L1 (the gathered vector is meant to be re-read from L1)
__device__ __forceinline__ double4v ld_f64x4_nc(const double *p) {
  double4v v;
  asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.a), "=d"(v.b), "=d"(v.c), "=d"(v.d) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_f64x4(double *p, const double4v &v) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(v.a), "d"(v.b), "d"(v.c), "d"(v.d) : "memory");
}

template <int VEC>  // 2: 128-bit accesses, 4: 256-bit
__global__ void __launch_bounds__(kThreads)
update_r_dot_kernel(int n, const double *alpha_dev, const double *__restrict__ Ap, double *__restrict__ r, double *partials,
                    int total_partials, unsigned *counter, FinishParams fp) {
  __shared__ double smem[kThreads / 32];
  if (fp.check_active && fp.st->active == 0) return;
  const double nalpha = -(*alpha_dev);
  double acc = 0.0;
  const int nvec = n / VEC;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < nvec; i += gridDim.x * kThreads) {
    const long long e = (long long)VEC * i;
    if (VEC == 4) {
      const double4v av = ld_stream_f64x4(Ap + e);
      double4v rv = ld_f64x4(r + e);
      rv.a = __dadd_rn(rv.a, __dmul_rn(nalpha, av.a));
      rv.b = __dadd_rn(rv.b, __dmul_rn(nalpha, av.b));
      rv.c = __dadd_rn(rv.c, __dmul_rn(nalpha, av.c));
      rv.d = __dadd_rn(rv.d, __dmul_rn(nalpha, av.d));
      st_f64x4(r + e, rv);
      acc = __dadd_rn(acc, __dmul_rn(rv.a, rv.a));
      acc = __dadd_rn(acc, __dmul_rn(rv.b, rv.b));
      acc = __dadd_rn(acc, __dmul_rn(rv.c, rv.c));
      acc = __dadd_rn(acc, __dmul_rn(rv.d, rv.d));
    } else {
      const double2 av = ld_stream_f64x2(Ap + e);
      double2 rv = *reinterpret_cast<const double2 *>(r + e);
      rv.x = __dadd_rn(rv.x, __dmul_rn(nalpha, av.x));
      rv.y = __dadd_rn(rv.y, __dmul_rn(nalpha, av.y));
      *reinterpret_cast<double2 *>(r + e) = rv;
      acc = __dadd_rn(acc, __dmul_rn(rv.x, rv.x));
      acc = __dadd_rn(acc, __dmul_rn(rv.y, rv.y));
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int i = nvec * VEC; i < n; ++i) {  // tail (n not a multiple of VEC)
      const double rn = __dadd_rn(r[i], __dmul_rn(nalpha, Ap[i]));
      r[i] = rn;
      acc = __dadd_rn(acc, __dmul_rn(rn, rn));
    }
  const double total = block_sum(acc, smem);
  publish_and_finish(total, partials, blockIdx.x, total_partials, counter, fp, smem);
}

// ---- the kernel that produces p, with exchange_externals.cpp:103-112 folded in -----------------------------------------
// MODE 1: x += alpha_{k-1} p_{k-1} (deferred HPCCG.cpp:383) ; p_k = r + beta p_{k-1} (HPCCG.cpp:369), alpha / beta from the
//         device state.
// MODE 0: p = src + 0.0 * src, the waxpby copies of HPCCG.cpp:347 (p = x) and :362 (p = r) with their arithmetic kept.
// PUT: every row of p that a neighbour needs is also stored straight into that neighbour's halo tail over NVLink (inverse
// send map, HaloPut).  The CTAs that own a tile of a segment's row range take a per-segment ticket when they are done; the
// last one publishes the exchange stamp in the neighbour's mailbox (release, system scope) -- what the neighbour's SpMV
// waits for at its halo-touching stages.  No separate put kernel, nothing of the exchange on the critical path.
template <bool PUT>
__device__ __forceinline__ bool put_elem(const HaloPut &put, long long e, double v) {
  bool did = false;
  if (PUT) {
#pragma unroll
    for (int s = 0; s < kMaxPeerNb; ++s)
      if (s < put.nseg && e >= put.lo[s] && e < put.hi[s]) {
        const int pos = __ldg(put.inv + put.inv_off[s] + (int)(e - put.lo[s]));
        if (pos >= 0) {
          put.dst[s][pos] = v;
          did = true;
        }
      }
  }
  return did;
}

template <int VEC, int MODE, bool PUT>
__global__ void __launch_bounds__(kThreads)
p_update_x_kernel(int n, const CgState *st, int check_active, const double *__restrict__ r, double *__restrict__ p,
                  double *__restrict__ x, const HaloPut put) {
  if (check_active && st->active == 0) return;
  double alpha = 0.0, beta = 0.0;
  if (MODE == 1) {
    alpha = st->alpha;
    beta = st->beta;
  }
  bool did_put = false;
  const int nvec = (n + VEC - 1) / VEC;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < nvec; i += gridDim.x * kThreads) {
    const long long e = (long long)VEC * i;
    bool near_seg = false;
    if (PUT) {
#pragma unroll
      for (int s = 0; s < kMaxPeerNb; ++s) near_seg = near_seg || (s < put.nseg && e < put.hi[s] && e + VEC > put.lo[s]);
    }
    if (e + VEC > n) {  // tail (n not a multiple of VEC): scalar, same arithmetic
      for (long long q = e; q < n; ++q) {
        double pn;
        if (MODE == 1) {
          const double po = p[q];
          x[q] = __dadd_rn(x[q], __dmul_rn(alpha, po));
          pn = __dadd_rn(r[q], __dmul_rn(beta, po));
        } else {
          pn = __dadd_rn(r[q], __dmul_rn(0.0, r[q]));
        }
        p[q] = pn;
        if (near_seg) did_put = put_elem<PUT>(put, q, pn) || did_put;
      }
    } else if (VEC == 4) {
      const double4v rv = ld_stream_f64x4(r + e);
      double4v pv;
      if (MODE == 1) {
        pv = ld_f64x4(p + e);
        double4v xv = ld_f64x4(x + e);
        xv.a = __dadd_rn(xv.a, __dmul_rn(alpha, pv.a));
        xv.b = __dadd_rn(xv.b, __dmul_rn(alpha, pv.b));
        xv.c = __dadd_rn(xv.c, __dmul_rn(alpha, pv.c));
        xv.d = __dadd_rn(xv.d, __dmul_rn(alpha, pv.d));
        pv.a = __dadd_rn(rv.a, __dmul_rn(beta, pv.a));
        pv.b = __dadd_rn(rv.b, __dmul_rn(beta, pv.b));
        pv.c = __dadd_rn(rv.c, __dmul_rn(beta, pv.c));
        pv.d = __dadd_rn(rv.d, __dmul_rn(beta, pv.d));
        st_f64x4(x + e, xv);
      } else {
        pv.a = __dadd_rn(rv.a, __dmul_rn(0.0, rv.a));
        pv.b = __dadd_rn(rv.b, __dmul_rn(0.0, rv.b));
        pv.c = __dadd_rn(rv.c, __dmul_rn(0.0, rv.c));
        pv.d = __dadd_rn(rv.d, __dmul_rn(0.0, rv.d));
      }
      st_f64x4(p + e, pv);
      if (near_seg) {
        did_put = put_elem<PUT>(put, e, pv.a) || did_put;
        did_put = put_elem<PUT>(put, e + 1, pv.b) || did_put;
        did_put = put_elem<PUT>(put, e + 2, pv.c) || did_put;
        did_put = put_elem<PUT>(put, e + 3, pv.d) || did_put;
      }
    } else {
      const double2 rv = ld_stream_f64x2(r + e);
      double2 pv;
      if (MODE == 1) {
        pv = *reinterpret_cast<const double2 *>(p + e);
        double2 xv = *reinterpret_cast<const double2 *>(x + e);
        xv.x = __dadd_rn(xv.x, __dmul_rn(alpha, pv.x));
        xv.y = __dadd_rn(xv.y, __dmul_rn(alpha, pv.y));
        pv.x = __dadd_rn(rv.x, __dmul_rn(beta, pv.x));
        pv.y = __dadd_rn(rv.y, __dmul_rn(beta, pv.y));
        *reinterpret_cast<double2 *>(x + e) = xv;
      } else {
        pv.x = __dadd_rn(rv.x, __dmul_rn(0.0, rv.x));
        pv.y = __dadd_rn(rv.y, __dmul_rn(0.0, rv.y));
      }
      *reinterpret_cast<double2 *>(p + e) = pv;
      if (near_seg) {
        did_put = put_elem<PUT>(put, e, pv.x) || did_put;
        did_put = put_elem<PUT>(put, e + 1, pv.y) || did_put;
      }
    }
  }
  if (PUT) {
    if (did_put) __threadfence_system();  // my remote stores are visible system-wide before this CTA's ticket
    __syncthreads();
    if (threadIdx.x == 0) {
      const int tile = kThreads * VEC, G = (int)gridDim.x, b = (int)blockIdx.x;
      bool fenced = false;
      for (int s = 0; s < put.nseg; ++s) {
        // tiles [t0, t1] cover the segment's row range; tile t belongs to CTA t % G
        int ctas = 1;
        bool mine = (b == 0);  // an empty segment is still signalled (the neighbour waits for every stamp)
        if (put.hi[s] > put.lo[s]) {
          const int t0 = put.lo[s] / tile, cnt = (put.hi[s] - 1) / tile - t0 + 1;
          ctas = cnt < G ? cnt : G;
          mine = cnt >= G || ((b - t0 % G + G) % G) < cnt;
        }
        if (!mine) continue;
        if (!fenced) {
          __threadfence_system();
          fenced = true;
        }
        const unsigned ticket = atomicInc(&put.link->put_ticket[s], (unsigned)(ctas - 1));  // wraps to 0 for the next exchange
        if (ticket == (unsigned)(ctas - 1)) {
          __threadfence_system();
          st_release_sys(put.link->nb_flag[s], exchange_stamp(put.link->epoch, put.exch_idx));
        }
      }
    }
  }
}

// ---- the loop's two vector kernels with bulk-async (TMA) streaming -------------------------------------------------------
// The SpMV gained 11 % when its matrix stream moved from LDG to cp.async.bulk.  The same for the read+write vector kernels:
// persistent CTAs, the input tiles of a step arrive in a shared-memory ring by cp.async.bulk + mbarrier, the output tiles
// leave by cp.async.bulk shared -> global (one elected thread issues both), so HBM sees whole-tile bursts in both directions
// and the SM's load/store units only touch shared memory.  Same arithmetic per element as the LDG/STG kernels; A/B at 512^3:
// r-update 6342 -> 6775 GB/s with 16 KB tiles x 3 stages (5 stages: 6509).
__device__ __forceinline__ void tma_bulk_s2g(void *dst_gmem, const void *src_smem, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_plain_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// r = r + (-alpha) Ap ; partial r.r            (HPCCG.cpp:384, :367)
struct VecOpRUpdate {
  static constexpr int kIn = 2, kOut = 1;
  static constexpr bool kReduce = true;
  double nalpha;
  __device__ __forceinline__ void load_scalars(const CgState *st) { nalpha = -st->alpha; }
  // in[0] = Ap, in[1] = r ; out[0] = r
  __device__ __forceinline__ double apply(const double (&in)[kIn], double (&out)[kOut]) const {
    out[0] = __dadd_rn(in[1], __dmul_rn(nalpha, in[0]));
    return __dmul_rn(out[0], out[0]);
  }
};
// x += alpha p (deferred HPCCG.cpp:383) ; p = r + beta p (HPCCG.cpp:369)
struct VecOpPUpdate {
  static constexpr int kIn = 3, kOut = 2;
  static constexpr bool kReduce = false;
  double alpha, beta;
  __device__ __forceinline__ void load_scalars(const CgState *st) {
    alpha = st->alpha;
    beta = st->beta;
  }
  // in[0] = r, in[1] = p, in[2] = x ; out[0] = p, out[1] = x
  __device__ __forceinline__ double apply(const double (&in)[kIn], double (&out)[kOut]) const {
    out[1] = __dadd_rn(in[2], __dmul_rn(alpha, in[1]));
    out[0] = __dadd_rn(in[0], __dmul_rn(beta, in[1]));
    return 0.0;
  }
};

struct VecPtrs {
  const double *in[3];
  double *out[2];
};

template <class OP, int TILE, int NSTAGES>
struct VecTmaCfg {
  static constexpr int kTileBytes = TILE * 8;
  static constexpr int kOutBufs = 3;
  static constexpr int kSmemBytes = (OP::kIn * NSTAGES + OP::kOut * kOutBufs) * kTileBytes + NSTAGES * 8 + 64;
};

template <class OP, int TILE, int NSTAGES, bool PUT>
__global__ void __launch_bounds__(kThreads)
vec_stream_tma_kernel(int n, const CgState *st, VecPtrs ptrs, double *partials, int total_partials, unsigned *counter,
                      FinishParams fp, const HaloPut put) {
  using Cfg = VecTmaCfg<OP, TILE, NSTAGES>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *s_in = reinterpret_cast<double *>(smem_raw);                    // [kIn][NSTAGES][TILE]
  double *s_out = s_in + OP::kIn * NSTAGES * TILE;                         // [kOut][kOutBufs][TILE]
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(s_out + OP::kOut * Cfg::kOutBufs * TILE);
  __shared__ double red[kThreads / 32];
  if (st->active == 0) return;
  OP op;
  op.load_scalars(st);
  const int tid = threadIdx.x;
  const int tiles = n / TILE;  // whole tiles; the remainder is done with plain accesses by the CTA that would own tile `tiles`
  const int my_count = (int)blockIdx.x < tiles ? (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  if (tid == 0) {
    for (int s = 0; s < NSTAGES; ++s) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int i) {  // elected thread only
    const int stg = i % NSTAGES;
    const long long e = ((long long)blockIdx.x + (long long)i * gridDim.x) * TILE;
    mbar_expect_tx(bars + stg, OP::kIn * Cfg::kTileBytes);
#pragma unroll
    for (int q = 0; q < OP::kIn; ++q) tma_plain_g2s(s_in + (q * NSTAGES + stg) * TILE, ptrs.in[q] + e, Cfg::kTileBytes, bars + stg);
  };
  if (tid == 0)
    for (int i = 0; i < NSTAGES && i < my_count; ++i) issue(i);
  double acc = 0.0;
  bool did_put = false;
  for (int i = 0; i < my_count; ++i) {
    const int stg = i % NSTAGES, ob = i % Cfg::kOutBufs;
    const long long e0 = ((long long)blockIdx.x + (long long)i * gridDim.x) * TILE;
    mbar_wait(bars + stg, (unsigned)(i / NSTAGES) & 1u);
#pragma unroll
    for (int j = 0; j < TILE / (2 * kThreads); ++j) {
      const int e = j * 2 * kThreads + tid * 2;
      double2 v[OP::kIn];
#pragma unroll
      for (int q = 0; q < OP::kIn; ++q) v[q] = *reinterpret_cast<const double2 *>(s_in + (q * NSTAGES + stg) * TILE + e);
      double ia[OP::kIn], ib[OP::kIn], oa[OP::kOut], obv[OP::kOut];
#pragma unroll
      for (int q = 0; q < OP::kIn; ++q) {
        ia[q] = v[q].x;
        ib[q] = v[q].y;
      }
      const double ra = op.apply(ia, oa), rb = op.apply(ib, obv);
      if (OP::kReduce) {
        acc = __dadd_rn(acc, ra);
        acc = __dadd_rn(acc, rb);
      }
#pragma unroll
      for (int q = 0; q < OP::kOut; ++q) *reinterpret_cast<double2 *>(s_out + (q * Cfg::kOutBufs + ob) * TILE + e) = make_double2(oa[q], obv[q]);
      if (PUT) {  // out[0] is p
        bool near_seg = false;
#pragma unroll
        for (int sg = 0; sg < kMaxPeerNb; ++sg) near_seg = near_seg || (sg < put.nseg && e0 + e < put.hi[sg] && e0 + e + 2 > put.lo[sg]);
        if (near_seg) {
          did_put = put_elem<PUT>(put, e0 + e, oa[0]) || did_put;
          did_put = put_elem<PUT>(put, e0 + e + 1, obv[0]) || did_put;
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // my writes to the out tiles are visible to the bulk stores
    __syncthreads();                                              // stage fully read, out tiles fully written
    if (tid == 0) {
#pragma unroll
      for (int q = 0; q < OP::kOut; ++q) tma_bulk_s2g(ptrs.out[q] + e0, s_out + (q * Cfg::kOutBufs + ob) * TILE, Cfg::kTileBytes);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      // at most this step's stores in flight: the step before has finished READING its out tiles, which are the tiles
      // rewritten two steps from now (three out buffers, one barrier per step)
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      if (i + NSTAGES < my_count) issue(i + NSTAGES);
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all stores complete before the kernel ends
  if ((int)blockIdx.x == tiles % (int)gridDim.x) {  // remainder tile, plain accesses, same arithmetic
    for (int e = tiles * TILE + tid; e < n; e += kThreads) {
      double in[OP::kIn], out[OP::kOut];
#pragma unroll
      for (int q = 0; q < OP::kIn; ++q) in[q] = ptrs.in[q][e];
      const double rr = op.apply(in, out);
      if (OP::kReduce) acc = __dadd_rn(acc, rr);
#pragma unroll
      for (int q = 0; q < OP::kOut; ++q) ptrs.out[q][e] = out[q];
      if (PUT) did_put = put_elem<PUT>(put, e, out[0]) || did_put;
    }
  }
  if (PUT) {
    if (did_put) __threadfence_system();
    __syncthreads();
    if (tid == 0) {
      const int G = (int)gridDim.x, b = (int)blockIdx.x;
      bool fenced = false;
      for (int sg = 0; sg < put.nseg; ++sg) {
        // tiles [t0, t1] (the remainder counts as tile `tiles`) cover the segment's row range; tile t belongs to CTA t % G
        int ctas = 1;
        bool mine = (b == 0);
        if (put.hi[sg] > put.lo[sg]) {
          const int t0 = put.lo[sg] / TILE, cnt = (put.hi[sg] - 1) / TILE - t0 + 1;
          ctas = cnt < G ? cnt : G;
          mine = cnt >= G || ((b - t0 % G + G) % G) < cnt;
        }
        if (!mine) continue;
        if (!fenced) {
          __threadfence_system();
          fenced = true;
        }
        const unsigned ticket = atomicInc(&put.link->put_ticket[sg], (unsigned)(ctas - 1));
        if (ticket == (unsigned)(ctas - 1)) {
          __threadfence_system();
          st_release_sys(put.link->nb_flag[sg], exchange_stamp(put.link->epoch, put.exch_idx));
        }
      }
    }
  }
  if (OP::kReduce) {
    const int lane = tid & 31, warp = tid >> 5;
    double w = warp_sum(acc);
    if (lane == 0) red[warp] = w;
    __syncthreads();
    double total = 0.0;
    if (warp == 0) {
      total = lane < kThreads / 32 ? red[lane] : 0.0;
      total = warp_sum(total);
    }
    publish_and_finish(total, partials, blockIdx.x, total_partials, counter, fp, red);
  }
}

// After the loop: the x update of the last executed iteration (niters >= 1), HPCCG.cpp:383.
__global__ void __launch_bounds__(kThreads)
x_fixup_kernel(int n, const CgState *st, const double *__restrict__ p, double *__restrict__ x) {
  if (st->niters < 1) return;
  const double alpha = st->alpha;
  const int npair = n >> 1;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < npair; i += gridDim.x * kThreads) {
    const long long e = 2 * (long long)i;
    const double2 pv = ld_stream_f64x2(p + e);
    double2 xv = *reinterpret_cast<const double2 *>(x + e);
    xv.x = __dadd_rn(xv.x, __dmul_rn(alpha, pv.x));
    xv.y = __dadd_rn(xv.y, __dmul_rn(alpha, pv.y));
    *reinterpret_cast<double2 *>(x + e) = xv;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) x[n - 1] = __dadd_rn(x[n - 1], __dmul_rn(alpha, p[n - 1]));
}

// ---- HPCCG.cpp:352-353 fused: r = b - Ap (waxpby alpha==1 branch, beta=-1) and r.r ------------------------
__global__ void __launch_bounds__(kThreads)
residual_dot_kernel(int n, const double *__restrict__ b, const double *__restrict__ Ap, double *__restrict__ r,
                    double *partials, int total_partials, unsigned *counter, FinishParams fp) {
  __shared__ double smem[kThreads / 32];
  double acc = 0.0;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    const double rv = __dadd_rn(b[i], __dmul_rn(-1.0, Ap[i]));
    r[i] = rv;
    acc = __dadd_rn(acc, __dmul_rn(rv, rv));
  }
  const double total = block_sum(acc, smem);
  publish_and_finish(total, partials, blockIdx.x, total_partials, counter, fp, smem);
}

// ---- exchange_externals.cpp:103 -----------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
halo_pack_kernel(int count, const int *__restrict__ elements_to_send, const double *__restrict__ x,
                 double *__restrict__ send_buffer, const CgState *st_check) {
  if (st_check && st_check->active == 0) return;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < count; i += gridDim.x * kThreads)
    send_buffer[i] = x[elements_to_send[i]];
}

// ---- exchange_externals.cpp:84-126 over peer memory, stand-alone form ----------------------------------------------
// (used when a send list has no compact inverse map; otherwise the put rides in p_update_x_kernel)
// The gather of exchange_externals.cpp:103 and the MPI_Send of :110-112 in one kernel: element i of the send list is
// stored straight into the neighbour's p vector (its halo tail) through the IPC-mapped pointer, i.e. across NVLink.
// The block that takes the last ticket publishes the new exchange number in each neighbour's mailbox (release, system
// scope), which is what the neighbour's SpMV waits for before it touches rows that reference halo columns.
__global__ void __launch_bounds__(kThreads)
halo_put_kernel(int count, const int *__restrict__ elements_to_send, const double *__restrict__ x, PeerLink *pl,
                const CgState *st_check, int exch_idx) {
  if (st_check && st_check->active == 0) return;
  __shared__ int s_last;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < count; i += gridDim.x * kThreads) {
    int seg = 0;
    while (seg + 1 < pl->nnb && i >= pl->seg_start[seg + 1]) ++seg;
    pl->nb_dst[seg][i - pl->seg_start[seg]] = x[elements_to_send[i]];
  }
  __threadfence_system();  // this block's remote stores are visible system-wide before its ticket
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned ticket = atomicInc(&pl->ticket, gridDim.x - 1);
    s_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if (threadIdx.x == 0) {
    const unsigned long long seq = exchange_stamp(pl->epoch, exch_idx);
    for (int i = 0; i < pl->nnb; ++i) st_release_sys(pl->nb_flag[i], seq);
  }
}

// ---- compute_residual.cpp:59-81 (local part): max |v1-v2| ---------------------------------------------
__global__ void __launch_bounds__(kThreads)
max_abs_diff_kernel(int n, const double *__restrict__ v1, const double *__restrict__ v2, double *partials,
                    int total_partials, unsigned *counter, double *out) {
  __shared__ double smem[kThreads / 32];
  __shared__ int s_last;
  double m = 0.0;
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    const double d = fabs(v1[i] - v2[i]);
    if (d > m) m = d;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) smem[warp] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kThreads / 32; ++w) m = fmax(m, smem[w]);
    partials[blockIdx.x] = m;
    __threadfence();
    const unsigned ticket = atomicInc(counter, (unsigned)(total_partials - 1));
    s_last = (ticket == (unsigned)(total_partials - 1));
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    double g = 0.0;
    for (int i = 0; i < total_partials; ++i) g = fmax(g, __ldcg(partials + i));
    *out = g;
  }
}

// ---- multi-rank scalar step: sum the gathered per-rank contributions in RANK ORDER ----------------------
// (the order of oracle/mpi_shim's MPI_Allreduce; ddot.cpp:77-82), then the same finish action.
__global__ void cg_scalar_kernel(const double *gathered, int nranks, FinishParams fp) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (fp.check_active && fp.st->active == 0) return;
  double g = gathered[0];
  for (int r = 1; r < nranks; ++r) g = __dadd_rn(g, gathered[r]);
  cg_finish(fp, g);
}

__global__ void cg_state_init_kernel(CgState *st, PeerLink *link) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (link) link->epoch = link->epoch + 1;  // a new solve: its exchange stamps are above every earlier one
    st->rtrans = st->oldrtrans = st->alpha = st->neg_alpha = st->beta = st->pAp = st->normr = st->zero = st->local_sum = 0.0;
    st->niters = 0;
    st->active = 1;
    st->counter = 0;
    st->pad = 0;
  }
}

// ---- generate_matrix.cpp:251-289 directly into the device ELL ----------------------------------------------
// One thread per local row; slot j receives the j-th entry the reference's sz,sy,sx loop would store.
__global__ void __launch_bounds__(kThreads)
generate_ell_kernel(int nx, int ny, int nz, long long start_row, long long total_nrow, int stencil7, int slots,
                    long long npad, const int *__restrict__ lower_map, const int *__restrict__ upper_map,
                    double *__restrict__ vals, int *__restrict__ cols) {
  const long long n = (long long)nx * ny * nz;
  const long long plane = (long long)nx * ny;
  for (long long row = (long long)blockIdx.x * kThreads + threadIdx.x; row < npad; row += (long long)gridDim.x * kThreads) {
    int j = 0;
    if (row < n) {
      const int iz = (int)(row / plane);
      const int rem = (int)(row - (long long)iz * plane);
      const int iy = rem / nx, ix = rem - iy * nx;
      const long long currow = start_row + row;
      for (int sz = -1; sz <= 1; ++sz)
        for (int sy = -1; sy <= 1; ++sy)
          for (int sx = -1; sx <= 1; ++sx) {
            const long long curcol = currow + sz * plane + sy * nx + sx;
            if (ix + sx >= 0 && ix + sx < nx && iy + sy >= 0 && iy + sy < ny && curcol >= 0 && curcol < total_nrow) {
              if (!stencil7 || sz * sz + sy * sy + sx * sx <= 1) {
                int lc;
                const int zz = iz + sz;
                const int q = (iy + sy) * nx + (ix + sx);
                if (zz < 0) lc = lower_map[q];
                else if (zz >= nz) lc = upper_map[q];
                else lc = (int)(curcol - start_row);
                const long long o = sell_offset(row, j, slots);
                vals[o] = (curcol == currow) ? 27.0 : -1.0;
                cols[o] = lc;
                ++j;
              }
            }
          }
    }
    for (; j < slots; ++j) {
      const long long o = sell_offset(row, j, slots);
      vals[o] = 0.0;
      cols[o] = -1;
    }
  }
}

// b = 27 - (nnz_row - 1), x = 0, xexact = 1 (generate_matrix.cpp:284-286) for device-only generation.
__global__ void __launch_bounds__(kThreads)
generate_vectors_kernel(int n, int slots, long long npad, const int *__restrict__ cols, double *x, double *b,
                        double *xexact) {
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    int nnzrow = 0;
    for (int j = 0; j < slots; ++j) nnzrow += (cols[sell_offset(i, j, slots)] >= 0) ? 1 : 0;
    if (x) x[i] = 0.0;
    if (b) b[i] = 27.0 - ((double)(nnzrow - 1));
    if (xexact) xexact[i] = 1.0;
  }
}

}  // namespace hpccg
