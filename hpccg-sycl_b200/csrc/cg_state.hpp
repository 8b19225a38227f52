// cg_state.hpp -- constants and the device-resident scalar block shared by host and device code.
#pragma once

namespace hpccg {

constexpr int kThreads = 256;          // threads per block for every kernel here
constexpr int kMaxPartials = 8192;     // capacity of the per-matrix block-partial array
constexpr int kSliceRows = 128;        // C of the SELL-C layout: rows per slice
constexpr int kRowPad = 512;           // local_nrow is padded to a multiple of this (4 slices)

// SELL-C (C = kSliceRows, sigma = 1: rows keep the reference's order) addressing of the matrix arrays:
// slice s = row / C holds its `slots` x C entries contiguously, slot-major inside the slice, so that
//   * a warp reading one slot of 32/64 consecutive rows touches one contiguous 256/512-byte segment, and
//   * a whole slice (or k consecutive slices) is ONE contiguous block -- what a TMA bulk copy moves.
#if defined(__CUDACC__)
__host__ __device__
#endif
inline long long sell_offset(long long row, int slot, int slots) {
  return ((row / kSliceRows) * slots + slot) * kSliceRows + (row % kSliceRows);
}

// Device-resident scalars of one CG solve (HPCCG.cpp:331-333,366-382 keep these on the host).
struct CgState {
  double rtrans;      // r.r used by the current iteration
  double oldrtrans;
  double alpha;
  double neg_alpha;   // -alpha, the beta argument of waxpby in HPCCG.cpp:384
  double beta;
  double pAp;
  double normr;       // sqrt(rtrans) of the last started iteration (HPCCG.cpp:371)
  double zero;        // constant 0.0 (beta of the k==1 copy, HPCCG.cpp:362)
  double local_sum;   // multi-rank: this rank's contribution before the gather
  int niters;         // HPCCG.cpp:385
  int active;         // loop condition of HPCCG.cpp:358 for the iteration being enqueued
  unsigned counter;   // last-block ticket
  unsigned pad;
};

}  // namespace hpccg
