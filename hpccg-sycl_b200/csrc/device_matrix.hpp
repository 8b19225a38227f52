// device_matrix.hpp -- the opaque device mirror behind hpccg_dev_matrix (include/hpccg_b200.h).
#pragma once

#include <cuda_runtime.h>

#include <vector>

#include "cg_state.hpp"

namespace hpccg {
// geometry of a stencil-structured pattern 0 (pattern_march.cuh)
struct MarchGeom {
  int nx, ny, nz;      // rows are numbered ix + nx*iy + nx*ny*iz
  int cols_x, cols_y;  // columns of 128 x 8 rows
  int base[9];         // centre delta of each stencil line (27-pt: 9 lines of 3; 7-pt: 5 lines of 1,1,3,1,1)
  int diag;            // entry index of the diagonal in pattern 0
  int neg1;            // every other value of pattern 0 is -1.0
  int ok;
};
}  // namespace hpccg

// HBM layout of one rank's matrix and solver workspace:
//   vals  : double [npad/128][slots][128]  SELL-C (C = 128, sigma = 1), slot j = j-th stored entry of the row
//   cols  : int32  [npad/128][slots][128]  local column id, -1 = padding (masked, never multiplied)
//   npad  : local_nrow rounded up to 512 rows
//   r, Ap : double [npad] ; p : double [ncol_pad]  (solver temporaries, HPCCG.cpp:327-329)
//   partials : double [kMaxPartials] block partials of the deterministic reductions
//   state : CgState (device scalars) ; hist : double [hist_cap] residual history
//   halo  : elements_to_send int32 [total_to_be_sent], send_buffer double [total_to_be_sent]
struct hpccg_dev_matrix {
  int device = 0;
  int n = 0;        // local_nrow
  int ncol = 0;     // local_ncol = n + externals
  int slots = 0;
  long long npad = 0;
  double *vals = nullptr;
  int *cols = nullptr;

  // format 1 (hpccg_dev_matrix_compress): vals/cols are released and replaced by
  //   pat_id    : uint16 [npad]           pattern id of every row (0xFFFF in the padding tail)
  //   pat_val   : double [npat][slots]    the stored values of each pattern, in the reference's order
  //   pat_delta : int32  [npat][slots]    column id minus row id of each stored entry
  //   pat_len   : int32  [npat]           stored entries of the pattern (row length)
  int format = 0;
  unsigned short *pat_id = nullptr;
  double *pat_val = nullptr;
  int *pat_delta = nullptr;
  int *pat_len = nullptr;
  int npat = 0;
  //   pat_desc  : uint32 [npat]          sub-pattern descriptor relative to pattern 0: stencil lines present, x-1 / x+1 entries
  //                                      missing (pattern_march.cuh); 0xFFFFFFFF: not such a sub-pattern (per-row table path)
  unsigned *pat_desc = nullptr;
  hpccg::MarchGeom march{};             // ok: pattern 0 is a 27- / 7-point stencil over x-fastest rows (z-marching SpMV)
  hpccg::Pattern0 pattern0;             // host copy of pattern 0, passed to the SpMV kernel as a __grid_constant__ parameter

  // format 2 (SELL-C-sigma proper, for matrices whose rows differ in length -- read_HPC_row.cpp:217-373 feeds arbitrary rows):
  //   slice s (kRaggedRows = 32 row POSITIONS, one warp) stores slice_slots[s] = its longest row's length slots, slot-major, at
  //   element offset slice_off[s] of vals / cols; total_elems = slice_off[nslices].  With sigma > 1 the rows of every window
  //   of sigma rows are sorted by decreasing length before they are dealt to slices (less padding); perm[position] =
  //   original row, y is stored through it.  sigma = 1 (no perm) wherever row RANGES must stay contiguous (halo ranges).
  int *slice_slots = nullptr;
  long long *slice_off = nullptr;
  int *perm = nullptr;
  long long total_elems = 0;
  int sigma = 1;

  // rows [0,interior_begin) and [interior_end,n) may reference halo columns (>= n); rows in between do not
  int interior_begin = 0, interior_end = 0;

  // halo plan (make_local_matrix.cpp:445-599)
  int num_neighbors = 0;
  std::vector<int> neighbors, recv_length, send_length;
  int total_to_be_sent = 0;
  int *d_elements_to_send = nullptr;
  double *d_send_buffer = nullptr;
  // inverse send maps for the put fused into the p-producing kernel (built by hpccg_dev_matrix_set_halo when every
  // segment's rows lie in a compact range and no row is sent twice to one neighbour); link / dst are filled by
  // peer_link_create
  int *d_put_inv = nullptr;
  hpccg::HaloPut put_plan{};
  int put_fusable = 0;

  // workspace
  double *partials = nullptr;
  hpccg::CgState *state = nullptr;
  double *gathered = nullptr;   // nranks doubles (multi-rank scalar gather)
  int gathered_cap = 0;
  double *r = nullptr, *p = nullptr, *Ap = nullptr;
  double *hist = nullptr;
  int hist_cap = 0;
  double *scratch_x = nullptr, *scratch_y = nullptr;  // staging for host-pointer calls of the C++ API
  long long scratch_cap = 0;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_p_ready = nullptr, ev_halo_done = nullptr;
  // HPCCG() with host vectors: b arrives on comm_stream while the set-up SpMV runs; the final x leaves in chunks
  static constexpr int kIoChunks = 8;
  cudaEvent_t ev_io[kIoChunks + 2] = {};

  // CUDA-graph replay of a whole solve (HPCCG_SOLVE_GRAPH): the launch sequence is a function of this key only
  cudaGraphExec_t graph_exec = nullptr;
  cudaStream_t graph_stream = nullptr;
  const double *graph_b = nullptr;
  double *graph_x = nullptr;
  int graph_max_iter = 0, graph_flags = 0;
  double graph_tol = 0.0;

  // single-kernel solve for launch-bound sizes (persistent_cg.cuh): plan cached per matrix
  int persist_state = 0;  // 0: not planned yet, 1: usable, -1: does not qualify
  int persist_G = 0, persist_rows = 0, persist_threads = 0, persist_smem = 0, persist_window = 0;
  int *persist_win = nullptr;            // [2][G] column windows

  // peer-memory link (multi-process runs on one NVSwitch domain): mailbox in this rank's HBM, IPC-mapped views of
  // the peers' mailboxes and of the neighbours' p vectors; peer_link == nullptr: NCCL send/recv + gathers are used
  hpccg::Mailbox *mailbox = nullptr;
  hpccg::PeerLink *peer_link = nullptr;   // device copy handed to the kernels
  std::vector<void *> ipc_opened;         // base pointers returned by cudaIpcOpenMemHandle
  int peer_tried = 0;
  int generation = 0;                     // bumped whenever buffers a captured graph refers to are replaced
};

namespace hpccg {

struct Device {
  int sm_count = 148;
  int id = 0;
};
const Device &device_info();

// grid for a streaming/reduction kernel over `work_items` thread-items
int stream_grid(long long work_items, int blocks_per_sm = 8);

int ensure_solver_workspace(hpccg_dev_matrix *m, int max_iter, int nranks);

// Host-vector plumbing of HPCCG() (host_api.cu) threaded through the device loop so that copies overlap compute:
//   b_ready : event after which the device copy of b is complete (the loop waits for it right before the first kernel
//             that reads b, i.e. AFTER p = x and the set-up SpMV, HPCCG.cpp:347-352);
//   x_host  : when set, the final x is produced chunk by chunk (x_fixup_kernel) and each chunk is copied to the host on
//             copy_stream while the next one is computed; the call returns when the last chunk has landed.
struct SolveIO {
  cudaEvent_t b_ready = nullptr;
  double *x_host = nullptr;
  cudaStream_t copy_stream = nullptr;
};
// hpccg_dev_cg_solve without the graph-replay branch, with optional host-vector plumbing (single rank or one NCCL rank)
int cg_solve_io(hpccg_dev_matrix *m, const double *b, double *x, int max_iter, double tolerance, int *niters, double *normr,
                double *hist_host, double *times, double *loop_ms, int flags, cudaStream_t stream, const SolveIO *io);
int ensure_scratch(hpccg_dev_matrix *m, long long doubles);

}  // namespace hpccg
