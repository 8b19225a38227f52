// pattern_march.cuh -- HPC_sparsemv.cpp:68-89 on the pattern-coded mirror when the dominant pattern is a 27- or 7-point
// stencil over rows numbered ix + nx*iy + nx*ny*iz (what generate_matrix.cpp:251-289 produces; detected from the pattern's
// own deltas, hpccg_dev_matrix_compress).
//
// Why: the first pattern kernel (spmv_pattern_kernel, kernels.cuh) issues one scalar gather per stored entry -- 27 LDG per
// row -- and ncu showed it bound by L1 wavefronts (LSU 60 %, DRAM 11 %), not by the 18 bytes per row it moves.  Here
//   * a thread owns FOUR x-consecutive rows: each of the 9 (sy, sz) stencil lines is ONE 256-bit load of the four centre
//     values; the two x-neighbours come from the adjacent lanes by warp shuffle (lanes 0 / 31 load theirs) -- 9 vector
//     loads per 4 rows instead of 108 scalar gathers: 72 B per row through L1 instead of 216+;
//   * a CTA (8 warps = 8 y-lines x 128 x) MARCHES along z through its column, so two thirds of the lines a step reads were
//     read by the same SM one and two steps earlier (L1 hits by construction) and L2 -> SM traffic is ~10 B per row;
//   * boundary rows are not a slow path: a pattern that is a SUB-pattern of the dominant one (same (value, delta) pairs, whole
//     stencil lines and / or the x-1 / x+1 entries missing -- what the faces, edges and corners of a block look like) carries
//     a descriptor and runs through the same code with +0.0 operands in place of the missing entries;
//   * where every off-diagonal value of the dominant pattern is -1.0 (this matrix: 27 / -1, generate_matrix.cpp:268-274) the
//     product -1.0 * x is formed as the exact negation -x, which halves the FP64 instruction count (the un-contracted
//     mul + add per entry otherwise makes the FP64 pipe the next limiter).
// Rows whose pattern is NOT a sub-pattern (halo rows with remapped columns, perturbed rows) take the per-row table path:
// inline when they lie in the interior, in a second phase -- after the neighbours' halo stamps have arrived -- when they
// lie in the halo-touching row ranges.  Entry order, values and operations per row are unchanged: results are
// bit-identical to every other SpMV path here and to the reference.
#pragma once

#include "device_matrix.hpp"
#include "kernels.cuh"

namespace hpccg {

constexpr unsigned kDescGeneric = 0xFFFFFFFFu;  // pattern is not a (regular) sub-pattern of pattern 0
constexpr int kMarchLines = 8;                  // y-lines per CTA step = warps per CTA
constexpr int kMarchRows = 4;                   // x-consecutive rows per thread
constexpr int kMarchWidth = 32 * kMarchRows;    // x-extent of a column

template <int SLOTS>
struct MarchRuns {
  static constexpr int kRuns = SLOTS == 27 ? 9 : 5;
  // first entry index and length (1 = centre only, 3 = x-1, x, x+1) of run k
  __host__ __device__ static constexpr int len(int k) { return SLOTS == 27 ? 3 : (k == 2 ? 3 : 1); }
  __host__ __device__ static constexpr int first(int k) { return SLOTS == 27 ? 3 * k : (k < 2 ? k : (k == 2 ? 2 : k + 2)); }
};

// one entry: s += v * x (un-contracted), or s += -x when v == -1.0 exactly
template <bool NEG1, bool IS_DIAG>
__device__ __forceinline__ double march_term(double s, double v, double xv) {
  if (NEG1 && !IS_DIAG) return __dadd_rn(s, -xv);
  return __dadd_rn(s, __dmul_rn(v, xv));
}

// branch-free conditional loads through the read-only path (allocating in L1): zeros when `pred` is false
__device__ __forceinline__ double4v ld_f64x4_nc_if(const double *p, bool pred) {
  double4v v;
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.b32 q, %5, 0;\n"
      "mov.f64 %0, 0d0000000000000000;\n"
      "mov.f64 %1, 0d0000000000000000;\n"
      "mov.f64 %2, 0d0000000000000000;\n"
      "mov.f64 %3, 0d0000000000000000;\n"
      "@q ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];\n"
      "}"
      : "=d"(v.a), "=d"(v.b), "=d"(v.c), "=d"(v.d)
      : "l"(p), "r"((int)pred));
  return v;
}
__device__ __forceinline__ double ld_f64_nc_if(const double *p, bool pred) {
  double v;
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.b32 q, %2, 0;\n"
      "mov.f64 %0, 0d0000000000000000;\n"
      "@q ld.global.nc.f64 %0, [%1];\n"
      "}"
      : "=d"(v)
      : "l"(p), "r"((int)pred));
  return v;
}
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// per-row table path (same as spmv_pattern_kernel's): entries of pattern `pid` in stored order
template <int SLOTS>
__device__ __forceinline__ double march_generic_row(int pid, long long row, const double *__restrict__ pat_val,
                                                    const int *__restrict__ pat_delta, const int *__restrict__ pat_len,
                                                    const double *__restrict__ x, int n, bool halo_cg) {
  const int len = __ldg(pat_len + pid);
  const double *pv = pat_val + (size_t)pid * SLOTS;
  const int *pd = pat_delta + (size_t)pid * SLOTS;
  double sum = 0.0;
  for (int j = 0; j < len; ++j) {
    const long long c = row + __ldg(pd + j);
    const double xv = (halo_cg && c >= n) ? __ldcg(x + c) : __ldg(x + c);
    sum = __dadd_rn(sum, __dmul_rn(__ldg(pv + j), xv));
  }
  return sum;
}

// Four x-consecutive rows of one thread.  UNIFORM: every row of the warp is a complete pattern-0 row -- no predicate, no
// select.  Otherwise `runs` says which stencil lines exist for my rows, lft_on / rgt_on whether my first row has its x-1
// entries and my last row its x+1 entries; absent entries are fed a +0.0 operand, which leaves the sum unchanged bit for bit
// (s + v * 0.0 == s; s is never -0.0 because it starts from +0.0).
template <int SLOTS, bool NEG1, bool UNIFORM, bool DOT>
__device__ __forceinline__ void march_rows(const double *__restrict__ x, const Pattern0 &p0, const MarchGeom &g, int row0, bool valid,
                                           unsigned runs, bool lft_on, bool rgt_on, int lane, double (&s)[kMarchRows],
                                           double (&xc)[kMarchRows]) {
  using Runs = MarchRuns<SLOTS>;
  constexpr int NR = Runs::kRuns;
  constexpr int kDiagRun = SLOTS == 27 ? 4 : 2;  // the run whose centre entry is the diagonal (checked by the host)
  constexpr int kDiag = SLOTS == 27 ? 13 : 3;
  // every load of the step first -- 9 x 256-bit centre loads per thread, plus the x-neighbours beyond the warp's ends (one
  // instruction per line: lane 0 fetches its left value, lane 31 its right one) -- then the arithmetic in stored order
  double4v c[NR];
  double ev[NR];
  const bool edge_l = lane == 0, edge_r = lane == 31;
#pragma unroll
  for (int k = 0; k < NR; ++k) {
    const bool on = UNIFORM || ((runs >> k) & 1u);
    const double *a = x + (row0 + g.base[k]);
    if (UNIFORM) c[k] = ld_f64x4_nc(a);
    else c[k] = ld_f64x4_nc_if(a, valid && (on || (DOT && k == kDiagRun)));
    if (Runs::len(k) == 3)
      ev[k] = ld_f64_nc_if(a + (edge_l ? -1 : kMarchRows), (UNIFORM || (valid && on)) && ((edge_l && lft_on) || (edge_r && rgt_on)));
  }
#pragma unroll
  for (int i = 0; i < kMarchRows; ++i) s[i] = 0.0;
#pragma unroll
  for (int k = 0; k < NR; ++k) {
    const int e0 = Runs::first(k);
    double4v ck = c[k];
    if (DOT && k == kDiagRun) {
      xc[0] = ck.a;
      xc[1] = ck.b;
      xc[2] = ck.c;
      xc[3] = ck.d;
      if (!UNIFORM && !((runs >> k) & 1u)) ck = double4v{0.0, 0.0, 0.0, 0.0};  // loaded for the dot only
    }
    if (Runs::len(k) == 3) {
      // x-neighbours of my four rows: the adjacent lanes' outer centre values
      double lft = __shfl_up_sync(0xffffffffu, ck.d, 1);
      double rgt = __shfl_down_sync(0xffffffffu, ck.a, 1);
      if (UNIFORM) {
        lft = edge_l ? ev[k] : lft;
        rgt = edge_r ? ev[k] : rgt;
      } else {
        lft = lft_on ? (edge_l ? ev[k] : lft) : 0.0;
        rgt = rgt_on ? (edge_r ? ev[k] : rgt) : 0.0;
      }
      const double v0 = p0.value[e0], v1 = p0.value[e0 + 1], v2 = p0.value[e0 + 2];
      const double xs[6] = {lft, ck.a, ck.b, ck.c, ck.d, rgt};
#pragma unroll
      for (int i = 0; i < kMarchRows; ++i) {
        s[i] = march_term<NEG1, false>(s[i], v0, xs[i]);
        if (e0 + 1 == kDiag) s[i] = march_term<NEG1, true>(s[i], v1, xs[i + 1]);
        else s[i] = march_term<NEG1, false>(s[i], v1, xs[i + 1]);
        s[i] = march_term<NEG1, false>(s[i], v2, xs[i + 2]);
      }
    } else {
      const double v0 = p0.value[e0];
      const double xs[4] = {ck.a, ck.b, ck.c, ck.d};
#pragma unroll
      for (int i = 0; i < kMarchRows; ++i) {
        if (e0 == kDiag) s[i] = march_term<NEG1, true>(s[i], v0, xs[i]);
        else s[i] = march_term<NEG1, false>(s[i], v0, xs[i]);
      }
    }
  }
}

// pattern descriptor (pat_desc[id], built by hpccg_dev_matrix_compress): bits 0..8 = stencil lines present, bit 9 = the x-1
// entries are missing, bit 10 = the x+1 entries are missing; kDescGeneric = not such a sub-pattern of pattern 0
constexpr unsigned kDescLm = 1u << 9, kDescRm = 1u << 10;

template <int SLOTS, bool DOT, bool NEG1>
__global__ void __launch_bounds__(kMarchLines * 32, 2)
spmv_pattern_march_kernel(const unsigned short *__restrict__ pat_id, const unsigned *__restrict__ pat_desc,
                          const double *__restrict__ pat_val, const int *__restrict__ pat_delta,
                          const int *__restrict__ pat_len, const __grid_constant__ Pattern0 p0,
                          const __grid_constant__ MarchGeom g, const double *__restrict__ x, double *__restrict__ y, int n,
                          int ib, int ie, double *partials, int partial_offset, int total_partials, unsigned *counter,
                          FinishParams fp, SpmvHalo halo) {
  using Runs = MarchRuns<SLOTS>;
  constexpr int NR = Runs::kRuns;
  constexpr int kDiagRun = SLOTS == 27 ? 4 : 2;
  __shared__ double smem[kThreads / 32];
  if (fp.check_active && fp.st->active == 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int plane = g.nx * g.ny;
  double dot = 0.0;

  // ---- phase 1: columns x planes, contiguous share of the (column-major, z-minor) step list per CTA ----
  // Rows of the halo-touching ranges [0, ib) and [ie, n) are left to phase 2: their halo columns may still be in flight
  // (a halo row can even BE pattern 0 -- the 7-pt upper plane reads p[row + plane], which is its halo entry).
  const long long steps = (long long)g.cols_x * g.cols_y * g.nz;
  const long long s_begin = steps * blockIdx.x / gridDim.x, s_end = steps * (blockIdx.x + 1) / gridDim.x;
  int col = (int)(s_begin / g.nz), z = (int)(s_begin - (long long)col * g.nz);
  // pattern ids of a step are fetched one step ahead: everything else of a step depends on them
  auto step_row0 = [&](int col_, int z_, bool &valid_) {
    const int cx = col_ % g.cols_x, cy = col_ / g.cols_x;
    const int yy = cy * kMarchLines + warp, x0 = cx * kMarchWidth + lane * kMarchRows;
    valid_ = yy < g.ny && x0 < g.nx;  // nx % 4 == 0: a thread's four rows are in the line or all outside
    return z_ * plane + yy * g.nx + x0;  // n < 2^31
  };
  auto load_ids = [&](int row0_, bool valid_) {
    uint2 raw = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
    if (valid_) raw = __ldg(reinterpret_cast<const uint2 *>(pat_id + row0_));
    return raw;
  };
  bool valid_next = false;
  int row0_next = 0;
  uint2 raw_next = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
  if (s_begin < s_end) {
    row0_next = step_row0(col, z, valid_next);
    raw_next = load_ids(row0_next, valid_next);
  }
  for (long long step = s_begin; step < s_end; ++step) {
    const bool valid = valid_next;
    const int row0 = row0_next;
    const uint2 raw = raw_next;
    if (++z == g.nz) {
      z = 0;
      ++col;
    }
    if (step + 1 < s_end) {
      row0_next = step_row0(col, z, valid_next);
      raw_next = load_ids(row0_next, valid_next);
      // the one plane the next step has not seen yet (its sz = +1 lines) goes to L1 now; the other two thirds of its lines
      // were loaded by this SM during this step and the one before
      if (valid_next && z + 1 < g.nz) {
#pragma unroll
        for (int k = 0; k < NR; ++k)
          if (SLOTS == 27 ? (k >= 6) : (k == 4)) {
            const double *a = x + (row0_next + g.base[k]);
            prefetch_l1(a);
            if (Runs::len(k) == 3 && (lane == 0 || lane == 31)) prefetch_l1(a + (lane == 0 ? -1 : kMarchRows));
          }
      }
    }
    int pid[kMarchRows];
    pid[0] = raw.x & 0xFFFF;
    pid[1] = raw.x >> 16;
    pid[2] = raw.y & 0xFFFF;
    pid[3] = raw.y >> 16;
    // "uniform": every row of the warp is a complete pattern-0 row -- except that the first row of lane 0 may be an x = 0 row
    // (its x-1 entries missing) and the last row of lane 31 an x = nx-1 row: the two ends of a grid line.  Those run through
    // the unpredicated code as well, with the missing neighbour fed as +0.0 (half of the warps of a 512-wide line hold one).
    bool lane_ok = (raw.x | raw.y) == 0u, u_lft = true, u_rgt = true;
    if (!lane_ok && valid) {
      if (lane == 0 && (raw.x >> 16) == 0u && raw.y == 0u) {
        lane_ok = __ldg(pat_desc + pid[0]) == ((SLOTS == 27 ? 0x1FFu : 0x1Fu) | kDescLm);
        u_lft = false;
      } else if (lane == 31 && raw.x == 0u && (raw.y & 0xFFFFu) == 0u) {
        lane_ok = __ldg(pat_desc + pid[3]) == ((SLOTS == 27 ? 0x1FFu : 0x1Fu) | kDescRm);
        u_rgt = false;
      }
    }
    const bool uniform = __all_sync(0xffffffffu, lane_ok);
    double s[kMarchRows], xc[kMarchRows] = {0.0, 0.0, 0.0, 0.0};
    bool simple = true;
    if (uniform) {
      march_rows<SLOTS, NEG1, true, DOT>(x, p0, g, row0, true, 0u, u_lft, u_rgt, lane, s, xc);
    } else {
      unsigned d[kMarchRows];
#pragma unroll
      for (int i = 0; i < kMarchRows; ++i) d[i] = valid ? __ldg(pat_desc + pid[i]) : kDescGeneric;
      unsigned runs = d[0] & 0x1FFu;
      simple = d[1] == runs && d[2] == runs && (d[0] & ~kDescLm) == runs && (d[3] & ~kDescRm) == runs;
      const bool lft_on = !(d[0] & kDescLm), rgt_on = !(d[3] & kDescRm);
      // the shuffles assume my x-neighbours' lanes hold the same lines: every valid lane of the warp must agree
      const unsigned runs0 = __shfl_sync(0xffffffffu, runs, 0);
      simple = __all_sync(0xffffffffu, !valid || (simple && runs == runs0));
      if (!simple) runs = 0u;
      march_rows<SLOTS, NEG1, false, DOT>(x, p0, g, row0, valid, runs, lft_on, rgt_on, lane, s, xc);
    }
    if (!valid) continue;
    // interior rows only; rows that are not (regular) sub-patterns of pattern 0 take the per-row table path right here
    bool store[kMarchRows];
#pragma unroll
    for (int i = 0; i < kMarchRows; ++i) store[i] = row0 + i >= ib && row0 + i < ie;
    if (!simple) {
#pragma unroll
      for (int i = 0; i < kMarchRows; ++i)
        if (store[i]) s[i] = march_generic_row<SLOTS>(pid[i], row0 + i, pat_val, pat_delta, pat_len, x, n, false);
    }
    if (store[0] && store[1] && store[2] && store[3]) {
      st_f64x4(y + row0, double4v{s[0], s[1], s[2], s[3]});
    } else {
#pragma unroll
      for (int i = 0; i < kMarchRows; ++i)
        if (store[i]) y[row0 + i] = s[i];
    }
    if (DOT) {
#pragma unroll
      for (int i = 0; i < kMarchRows; ++i)
        if (store[i]) dot = __dadd_rn(dot, __dmul_rn(xc[i], s[i]));
    }
  }

  // ---- phase 2: rows of the halo-touching ranges that took no part in phase 1 ----
  const long long halo_rows = (long long)ib + (n - ie);
  if (halo_rows > 0) {
    if (halo.link) {
      if (tid < halo.link->nnb) {
        const Mailbox *own = halo.link->box[halo.link->rank];
        if (*reinterpret_cast<volatile int *>(&halo.link->error) != 0 ||
            !peer_wait_ge(&own->halo_seq[tid], exchange_stamp(halo.link->epoch, halo.exch_idx)))
          halo.link->error = 2;
      }
      __syncthreads();
    }
    for (long long q = (long long)blockIdx.x * kThreads + tid; q < halo_rows; q += (long long)gridDim.x * kThreads) {
      const long long row = q < ib ? q : ie + (q - ib);
      const int id = pat_id[row];
      if (id == 0xFFFF) continue;
      const double sum = march_generic_row<SLOTS>(id, row, pat_val, pat_delta, pat_len, x, n, halo.link != nullptr);
      y[row] = sum;
      if (DOT) dot = __dadd_rn(dot, __dmul_rn(__ldg(x + row), sum));
    }
  }
  if (DOT) {
    const double total = block_sum(dot, smem);
    publish_and_finish(total, partials, partial_offset + blockIdx.x, total_partials, counter, fp, smem);
  }
}

}  // namespace hpccg
