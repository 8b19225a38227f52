// HPC_Sparse_Matrix.hpp -- the matrix container of the B200-native HPCCG hot path.
//
// Field names and meanings follow the reference struct (HPC_Sparse_Matrix.hpp:54-85 of
// Dart120/HPCCG-SYCL) so that code written against it compiles unchanged.  Differences:
//   * the halo-plan fields the reference only has under -DUSING_MPI are always present: multi-GPU
//     is a run-time property here (rank context, hpccg_b200.h), not a compile-time one;
//   * `device` is the opaque column-major ELLPACK mirror in HBM that the kernels stream
//     (hpccg_dev_matrix, created on first use or by hpccg_api_matrix_device);
//   * the generator descriptor remembers (nx, ny, nz, stencil, rank, size) so that the mirror can be
//     generated directly on the device when the host row arrays were not materialised;
//   * max_external / max_num_messages (reference :49-51) are gone: halo arrays are sized exactly.
#ifndef HPCCG_B200_HPC_SPARSE_MATRIX_HPP
#define HPCCG_B200_HPC_SPARSE_MATRIX_HPP

struct HPC_Sparse_Matrix_STRUCT {
  char *title;
  int start_row;
  int stop_row;
  int total_nrow;
  long long total_nnz;
  int local_nrow;
  int local_ncol;  // local_nrow until make_local_matrix adds the externals
  int local_nnz;   // the reference's claimed 27*local_nrow; saturates at INT_MAX beyond 430^3
  int *nnz_in_row;
  double **ptr_to_vals_in_row;
  int **ptr_to_inds_in_row;
  double **ptr_to_diags;

  // halo plan, filled by make_local_matrix
  int num_external;
  int num_send_neighbors;
  int *external_index;
  int *external_local_index;
  int total_to_be_sent;
  int *elements_to_send;
  int *neighbors;
  int *recv_length;
  int *send_length;
  double *send_buffer;

  double *list_of_vals;
  int *list_of_inds;

  // ---- B200 additions ----
  void *device;     // hpccg_dev_matrix*, owned by this struct
  int gen_nx, gen_ny, gen_nz, gen_stencil;  // generator descriptor (gen_stencil == 0: not generated)
  int rank, size;   // rank context captured by generate_matrix
  int localized;    // make_local_matrix has run
  int host_rows;    // 1: the row arrays above are materialised, 0: device-only generation
};
typedef struct HPC_Sparse_Matrix_STRUCT HPC_Sparse_Matrix;

// Frees the host arrays, the device mirror and A itself; sets A to 0 (reference :88).
void destroyMatrix(HPC_Sparse_Matrix *&A);

#endif
