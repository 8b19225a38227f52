// cg_state.hpp -- constants and the device-resident scalar block shared by host and device code.
#pragma once

namespace hpccg {

constexpr int kThreads = 256;          // threads per block for every kernel here
constexpr int kMaxPartials = 8192;     // capacity of the per-matrix block-partial array
constexpr int kSliceRows = 128;        // C of the SELL-C layout: rows per slice
constexpr int kRowPad = 512;           // local_nrow is padded to a multiple of this (4 slices)
constexpr int kRaggedRows = 32;        // C of the ragged SELL-C-sigma mirror (format 2): one warp per slice

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

// ---- pattern-coded matrix format (opt-in) ----------------------------------------------------------
// A row's PATTERN is its sequence of stored (value, column - row) pairs.  Stencil-like matrices have very few distinct
// patterns (27-pt: one per boundary type, plus the halo variants), so the mirror keeps ONE 16-bit pattern id per row and
// a small pattern table instead of 12 bytes per stored entry.  Lossless: same values, same columns, same order.
constexpr int kMaxPatterns = 65535;   // ids are uint16; 0xFFFF marks the rows of the padding tail
constexpr int kPatternSlots = 32;     // capacity of the constant-bank copy of pattern 0
// Pattern 0 (the most frequent one) as the SpMV kernel receives it: a __grid_constant__ kernel parameter.  In the
// warp-uniform fast path its values and deltas are constant-bank OPERANDS of the multiply / address instructions,
// i.e. the matrix costs no load instruction at all there.
struct Pattern0 {
  double value[kPatternSlots];
  int delta[kPatternSlots];
  int len;
  int pad;
};

// ---- peer-memory communication (multi-GPU, one process per GPU; NVLink / NVSwitch P2P) ------------------------
// Replaces MPI_Allreduce (ddot.cpp:79-80) and MPI_Irecv/Send/Wait (exchange_externals.cpp:87-126) INSIDE the kernels:
// every rank owns a Mailbox in its HBM that its peers write through IPC-mapped pointers.
constexpr int kMaxRanks = 16;     // ranks of one NVSwitch domain
constexpr int kMaxPeerNb = 4;     // halo neighbours served over peer memory (z-slabs have <= 2)
constexpr int kMailSlots = 4;     // ring of reduction slots (a rank can be at most one reduction ahead)

struct Mailbox {
  double value[kMailSlots][kMaxRanks];            // value[slot][r]: rank r's contribution to reduction `seq`
  unsigned long long seq[kMailSlots][kMaxRanks];  // stamp written after the value (release, system scope)
  unsigned long long halo_seq[kMaxPeerNb];        // halo_seq[i]: exchange number of the last plane my neighbour i delivered
};

struct PeerLink {
  int rank, size, nnb, error;
  Mailbox *box[kMaxRanks];                  // box[rank] is this rank's own mailbox (local pointer), the rest are peers'
  double *nb_dst[kMaxPeerNb];               // where my send segment i lands inside neighbour i's p vector
  unsigned long long *nb_flag[kMaxPeerNb];  // neighbour i's halo_seq entry for me
  int seg_start[kMaxPeerNb + 1];            // segment boundaries of elements_to_send
  unsigned long long reduce_seq;            // device-local counter; identical on every rank by construction
  unsigned long long epoch;                 // solves started on this matrix (cg_state_init_kernel); same on every rank
  unsigned int ticket;                      // stand-alone halo_put_kernel: last-block ticket
  unsigned int put_ticket[kMaxPeerNb];      // fused put (p_update_x_kernel): per-segment ticket of the contributing CTAs
  unsigned int pad;
};

// Stamp of exchange number `idx` (1 = the set-up exchange, k + 1 = iteration k) of the current solve.  It is a function of
// device state (epoch) and of a launch-position constant (idx), so a captured CUDA graph replays with fresh stamps.
#if defined(__CUDACC__)
__host__ __device__
#endif
inline unsigned long long exchange_stamp(unsigned long long epoch, int idx) {
  return (epoch << 32) | (unsigned long long)(unsigned int)idx;
}

// exchange_externals.cpp:103-112 folded into the kernel that PRODUCES p: per send segment (= neighbour) an inverse map
// row -> position in the neighbour's halo tail over the row range [lo, hi) the segment draws from, so the thread that
// writes p[row] also stores it across NVLink.  nseg == 0: no put.
struct HaloPut {
  PeerLink *link;
  const int *inv;                 // concatenated inverse maps, -1 = row not sent to this neighbour
  int nseg;
  int exch_idx;
  int lo[kMaxPeerNb], hi[kMaxPeerNb], inv_off[kMaxPeerNb];
  double *dst[kMaxPeerNb];        // neighbour's p + its halo offset for my segment (IPC-mapped)
};

}  // namespace hpccg
