// persistent_cg.cuh -- HPCCG.cpp:312-402 as ONE kernel of one thread-block cluster, for launch-bound problem sizes.
//
// At 20x30x10 (BASELINE configs[0]) a CG iteration is ~50 KB of vectors and 1.9 MB of matrix: the three kernels of the
// normal loop finish in a few microseconds each, and 447 launches cost more than the arithmetic even when they are replayed
// from a CUDA graph (1.9 ms per solve).  A first single-kernel version synchronised its CTAs through L2 (mailbox words polled
// with gpu-scope loads): 1.1-1.4 ms -- two L2 round trips per reduction are ~1.5 us, and there are two reductions per
// iteration.  This version keeps EVERYTHING inside one cluster of up to 16 CTAs (one GPC):
//   * every CTA owns a contiguous block of rows (one row per thread) and keeps that block of the matrix in SHARED MEMORY for
//     the whole solve (slot-major, so consecutive threads read consecutive words), together with its rows of r and x;
//   * it also keeps a private copy of p for the column WINDOW its rows reference (own rows + the stencil's reach) and updates
//     the whole window itself each iteration: p = r + beta p needs only r of the neighbouring blocks, which it reads straight
//     from the owners' shared memory (distributed shared memory, ~215 cycles) -- p never leaves the SMs;
//   * the two reductions of an iteration (r.r, p.Ap) are: every CTA stores its partial into every CTA's mailbox through
//     DSMEM, one hardware cluster barrier (~380 cycles), every CTA sums the partials in CTA order -- same alpha / beta bits
//     everywhere, same exit decision, no global memory, no atomics.
// Element-wise arithmetic and its order are those of the normal loop (un-contracted mul/add in the reference's stored-entry
// order); the reduction TREE differs (rows are blocked, not grid-strided), i.e. results agree to reduction-order rounding.
#pragma once

#include <cooperative_groups.h>

#include "kernels.cuh"

namespace hpccg {

constexpr int kClusterMax = 16;        // CTAs of the cluster (non-portable size; one GPC)
constexpr int kClusterThreadsMax = 1024;

// per-CTA column window [lo, hi) of the rows [b*rows, (b+1)*rows): one block per CTA-to-be
__global__ void __launch_bounds__(kThreads)
persist_window_kernel(const int *__restrict__ cols, int slots, int n, int rows, int *win_lo, int *win_hi) {
  __shared__ int s_lo[kThreads], s_hi[kThreads];
  const int r0 = blockIdx.x * rows, r1 = min(n, r0 + rows);
  int lo = r0, hi = r1;  // own rows always belong to the window (p of own rows is needed for p.Ap and the x update)
  for (int row = r0 + threadIdx.x; row < r1; row += kThreads)
    for (int j = 0; j < slots; ++j) {
      const int c = cols[sell_offset(row, j, slots)];
      if (c >= 0) {
        lo = min(lo, c);
        hi = max(hi, c + 1);
      }
    }
  s_lo[threadIdx.x] = lo;
  s_hi[threadIdx.x] = hi;
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      s_lo[threadIdx.x] = min(s_lo[threadIdx.x], s_lo[threadIdx.x + o]);
      s_hi[threadIdx.x] = max(s_hi[threadIdx.x], s_hi[threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    win_lo[blockIdx.x] = s_lo[0];
    win_hi[blockIdx.x] = s_hi[0];
  }
}

// Sum over the cluster of one double per thread: fixed order inside the CTA (warp butterflies, warp totals in warp order),
// partials exchanged through DSMEM, one cluster barrier, summed in CTA order.  Result in every thread of every CTA.
__device__ __forceinline__ double cluster_allreduce(double v, double *red /* 32 */, double (*mail)[kClusterMax], int q, int G,
                                                    unsigned rank) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    // the CTA total by one more butterfly over the warp totals (every lane gets it), lane c delivers it to CTA c
    double t = warp_sum(lane < nwarps ? red[lane] : 0.0);
    if (lane < G) {
      double *dst = cluster.map_shared_rank(&mail[q & 1][rank], lane);
      *dst = t;
    }
  }
  cluster.sync();  // release / acquire at cluster scope: every mailbox entry of this reduction has landed
  double mv[kClusterMax];
#pragma unroll
  for (int c = 0; c < kClusterMax; ++c) mv[c] = c < G ? mail[q & 1][c] : 0.0;  // all loads first, then the ordered sum
  double t = mv[0];
#pragma unroll
  for (int c = 1; c < kClusterMax; ++c) t = c < G ? __dadd_rn(t, mv[c]) : t;
  return t;
}

template <int SLOTS>  // 27, 7 or 0 (run-time slot count)
__global__ void __launch_bounds__(kClusterThreadsMax, 1)
cg_cluster_kernel(const double *__restrict__ vals, const int *__restrict__ cols, int slots_rt, int n, int rows, int max_window,
                  const int *__restrict__ win_lo, const int *__restrict__ win_hi, const double *__restrict__ b,
                  double *__restrict__ x, int max_iter, double tol, CgState *st, double *hist) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double red[32];
  __shared__ double mail[2][kClusterMax];
  const int slots = SLOTS > 0 ? SLOTS : slots_rt;
  const unsigned rank = cluster.block_rank();
  const int G = (int)cluster.num_blocks(), tid = threadIdx.x, T = blockDim.x;
  const int r0 = (int)rank * rows, r1 = min(n, r0 + rows), R = max(r1 - r0, 0);
  const int wlo = win_lo[rank], whi = win_hi[rank], W = whi - wlo;
  // shared memory, the same layout in every CTA (the owners' r is read through DSMEM at the same offset):
  //   r [rows] | x [rows] | vals [slots][rows] | cols [slots][rows] | p window [max_window] | r sources [max_window]
  double *s_r = reinterpret_cast<double *>(smem_raw);
  double *s_x = s_r + rows;
  double *s_vals = s_x + rows;
  int *s_cols = reinterpret_cast<int *>(s_vals + (size_t)slots * rows);
  double *s_p = reinterpret_cast<double *>(s_cols + (((size_t)slots * rows + 1) & ~(size_t)1));
  const double **s_src = reinterpret_cast<const double **>(s_p + max_window);
  // where r of each window column lives: my own shared memory or its owner's, through the cluster's shared window
  for (int i = tid; i < W; i += T) {
    const int gi = wlo + i, owner = gi / rows;
    const double *base = (owner == (int)rank) ? s_r : cluster.map_shared_rank(s_r, owner);
    s_src[i] = base + (gi - owner * rows);
  }
  const bool has = tid < R;
  if (has) {
    for (int j = 0; j < slots; ++j) {
      const long long o = sell_offset(r0 + tid, j, slots);
      s_vals[j * rows + tid] = vals[o];
      const int c = cols[o];
      s_cols[j * rows + tid] = c >= 0 ? c - wlo : -1;
    }
    s_x[tid] = x[r0 + tid];
  }
  // ---- set-up (HPCCG.cpp:347-356): p = x ; Ap = A p ; r = b - Ap ; rtrans = r.r ----
  for (int i = tid; i < W; i += T) {
    const double xv = x[wlo + i];
    s_p[i] = __dadd_rn(xv, __dmul_rn(0.0, xv));
  }
  __syncthreads();
  auto spmv_row = [&]() {
    double sum = 0.0;
    if (SLOTS > 0) {
      int c[SLOTS > 0 ? SLOTS : 1];
      double v[SLOTS > 0 ? SLOTS : 1];
#pragma unroll
      for (int j = 0; j < SLOTS; ++j) {
        c[j] = s_cols[j * rows + tid];
        v[j] = s_vals[j * rows + tid];
      }
#pragma unroll
      for (int j = 0; j < SLOTS; ++j) {
        const double t = __dadd_rn(sum, __dmul_rn(v[j], s_p[max(c[j], 0)]));
        sum = c[j] >= 0 ? t : sum;
      }
    } else {
      for (int j = 0; j < slots; ++j) {
        const int c = s_cols[j * rows + tid];
        if (c >= 0) sum = __dadd_rn(sum, __dmul_rn(s_vals[j * rows + tid], s_p[c]));
      }
    }
    return sum;
  };
  double acc = 0.0;
  if (has) {
    const double ap0 = spmv_row();
    const double rv = __dadd_rn(b[r0 + tid], __dmul_rn(-1.0, ap0));
    s_r[tid] = rv;
    acc = __dmul_rn(rv, rv);
  }
  int q = 0;  // reduction number
  double rtrans = cluster_allreduce(acc, red, mail, q++, G, rank);
  double normr = sqrt(rtrans), oldrtrans = 0.0, alpha = 0.0, beta = 0.0;
  int niters = 0;
  const bool lead = rank == 0 && tid == 0;
  if (lead && hist) hist[0] = normr;
  bool active = (1 < max_iter) && (normr > tol);
  if (lead && hist && active) hist[1] = normr;
  // ---- iterations (HPCCG.cpp:358-386) ----
  for (int k = 1; active; ++k) {
    // p = r (k == 1, :362) or p = r + beta p (:369) for the whole window; r of other blocks' rows is read from their owners'
    // shared memory (the cluster barrier of the r.r reduction ordered their update before this read; their next update
    // comes after the p.Ap barrier below, which this CTA reaches only after this loop)
    for (int i0 = 0; i0 < W; i0 += 8 * T) {
      double rv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {  // all of a thread's (remote) loads are in flight before the first is used
        const int i = i0 + u * T + tid;
        rv[u] = i < W ? *s_src[i] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * T + tid;
        if (i < W) s_p[i] = (k == 1) ? __dadd_rn(rv[u], __dmul_rn(0.0, rv[u])) : __dadd_rn(rv[u], __dmul_rn(beta, s_p[i]));
      }
    }
    __syncthreads();
    // Ap = A p ; p.Ap (:379-381).  One row per thread: Ap of my row stays in a register.
    double ap = 0.0, pv = 0.0;
    acc = 0.0;
    if (has) {
      ap = spmv_row();
      pv = s_p[r0 + tid - wlo];
      acc = __dmul_rn(pv, ap);
    }
    const double pAp = cluster_allreduce(acc, red, mail, q++, G, rank);
    alpha = rtrans / pAp;
    niters = k;
    // x += alpha p ; r -= alpha Ap (:383-384) ; the next r.r (:367)
    acc = 0.0;
    if (has) {
      s_x[tid] = __dadd_rn(s_x[tid], __dmul_rn(alpha, pv));
      const double rv = __dadd_rn(s_r[tid], __dmul_rn(-alpha, ap));
      s_r[tid] = rv;
      acc = __dmul_rn(rv, rv);
    }
    const double rr = cluster_allreduce(acc, red, mail, q++, G, rank);
    // loop condition of iteration k+1 (:358): k+1 < max_iter && normr_k > tolerance
    if ((k + 1 < max_iter) && (normr > tol)) {
      oldrtrans = rtrans;
      rtrans = rr;
      beta = rr / oldrtrans;
      normr = sqrt(rr);
      if (lead && hist) hist[k + 1] = normr;
    } else {
      active = false;
    }
  }
  if (has) x[r0 + tid] = s_x[tid];
  if (lead) {
    st->rtrans = rtrans;
    st->oldrtrans = oldrtrans;
    st->alpha = alpha;
    st->neg_alpha = -alpha;
    st->beta = beta;
    st->normr = normr;
    st->niters = niters;
    st->active = 0;
  }
  cluster.sync();  // nobody's shared memory goes away while a neighbour may still read it
}

}  // namespace hpccg
