// context.hpp -- per-thread rank context and the process-wide NCCL communicator.
// Replaces MPI_COMM_WORLD + MPI_Comm_rank/size of the reference (generate_matrix.cpp:207-208,
// make_local_matrix.cpp:75-76, HPCCG.cpp:337, exchange_externals.cpp:68-69).
#pragma once

#include <cuda_runtime.h>

#include "../../include/hpccg_b200.h"

namespace hpccg {

struct RankContext {
  int rank = 0;
  int size = 1;
  hpccg_allgather_fn allgather = nullptr;  // host collective for set-up
  void *allgather_user = nullptr;
  // generate_matrix options the reference fixes at compile time
  int stencil = 27;     // generate_matrix.cpp:219
  int host_arrays = 1;  // 0: device-only generation
  int print_residuals = 1;  // HPCCG() prints the reference's residual lines on rank 0
  int matrix_format = -1;   // -1: environment (HPCCG_B200_FORMAT=pattern -> 1) else 0; 0: SELL int32; 1: pattern-coded
};

RankContext &ctx();  // thread-local

// rank-major allgather of `nbytes` per rank through the installed transport (identity when size==1)
int ctx_allgather(const void *send, long long nbytes, void *recv);

// ---- NCCL (loaded with dlopen so that single-GPU use has no NCCL dependency) ---------------------
bool nccl_ready();
int nccl_rank();
int nccl_size();
// in-place allgather of one double per rank: buf[rank] holds this rank's value on entry
int nccl_allgather_double(double *buf, cudaStream_t stream);
int nccl_halo_exchange(const double *send_buffer, const int *send_length, double *recv_base, const int *recv_length,
                       const int *neighbors, int num_neighbors, cudaStream_t stream);

// allgather of `nbytes` per rank through the NCCL communicator (host buffers in, host buffers out)
int nccl_allgather_host(const void *send, long long nbytes, void *recv);

}  // namespace hpccg

struct hpccg_dev_matrix;
namespace hpccg {
// Builds m->peer_link for the calling rank (collective over the NCCL communicator): exchanges IPC handles of every
// rank's mailbox and p vector, maps the peers' memory and resolves where each send segment lands in its neighbour's p.
// `eligible`: this rank's matrix can be served by the kernels that wait on peer memory; the link exists only when EVERY
// rank is eligible.  Must be called by all ranks of the communicator or by none (it contains collectives).
// Leaves m->peer_link == nullptr (NCCL path) when peer memory cannot be used; returns non-zero only on hard errors.
int peer_link_create(hpccg_dev_matrix *m, bool eligible);
void peer_link_destroy(hpccg_dev_matrix *m);
}  // namespace hpccg
