// hpccg_api.hpp -- the reference-named C++ functions of the B200-native HPCCG hot path.
// Each keeps the signature of the reference header cited beside it, so a caller written against
// Dart120/HPCCG-SYCL switches by changing its include path and linking libhpccg_b200.so.
// Vector arguments may be host pointers (reference usage) or device pointers.
#ifndef HPCCG_B200_API_HPP
#define HPCCG_B200_API_HPP

#include "HPC_Sparse_Matrix.hpp"

// generate_matrix.hpp:58 -- rows in iz,iy,ix order, entries in sz,sy,sx order, global column ids,
// b = A*1, x = 0, xexact = 1.  Rank, size, stencil (27/7) and host-row materialisation come from the
// thread's rank context (hpccg_ctx_set / hpccg_api_set_options).
void generate_matrix(int nx, int ny, int nz, HPC_Sparse_Matrix **A, double **x, double **b, double **xexact);

// read_HPC_row.hpp:56-57 -- the reference's (deprecated, README.md:114-118) matrix-file input: header `total_nrow
// total_nnz`, the row lengths, then per row `nnz (value column)*`, then per row `x b xexact`.  Rows are dealt to the ranks
// in contiguous chunks exactly as read_HPC_row.cpp:257-267; column ids stay global until make_local_matrix.
void read_HPC_row(char *data_file, HPC_Sparse_Matrix **A, double **x, double **b, double **xexact);

// make_local_matrix.hpp:48 -- global -> local column ids, externals numbered per owner in
// first-encounter order, send lists negotiated through the context's set-up collective.
void make_local_matrix(HPC_Sparse_Matrix *A);

// exchange_externals.hpp:49 -- fills x[local_nrow .. local_ncol) from the neighbouring ranks.
void exchange_externals(HPC_Sparse_Matrix *A, const double *x);

// HPC_sparsemv.hpp:55-56 -- y = A x.
int HPC_sparsemv(HPC_Sparse_Matrix *A, const double *const x, double *const y);

// ddot.hpp:55-56 -- *result = sum x_i y_i over all ranks; the gather time is added to time_allreduce.
int ddot(const int n, const double *const x, const double *const y, double *const result, double &time_allreduce);

// waxpby.hpp:51-53 -- w = alpha x + beta y.
int waxpby(const int n, const double alpha, const double *const x, const double beta, const double *const y,
           double *const w);

// HPCCG.hpp:61-63 -- un-preconditioned CG; fills niters, normr and times[0..5].
int HPCCG(HPC_Sparse_Matrix *A, double *const b, double *const x, const int max_iter, const double tolerance,
          int &niters, double &normr, double *times);

// compute_residual.hpp:50-51 -- max_i |v1_i - v2_i| over all ranks.
int compute_residual(const int n, const double *const v1, const double *const v2, double *const residual);

// mytimer.hpp:44 -- wall-clock seconds.
double mytimer(void);

// Releases the three vectors generate_matrix handed out (the reference leaks them, main.cpp:322).
void free_vectors(double *x, double *b, double *xexact);

#endif
