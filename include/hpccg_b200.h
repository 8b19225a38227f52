/* include/hpccg_b200.h -- the C-ABI of libhpccg_b200.so.
 *
 * This is the drop-in boundary of the B200-native HPCCG hot path: plain C types
 * (int, long long, double, pointers), every function returns int (0 = ok,
 * otherwise a CUDA / NCCL / argument error code; hpccg_last_error() has the
 * text), no exceptions and no C++ or torch types cross it.  Above it sit the
 * reference-named C++ functions (hpccg-sycl_b200/csrc/include/*.hpp) and, for
 * tests and bench.py, a ctypes binding (hpccg-sycl_b200/_capi.py).
 *
 * Each entry point cites the reference interface (file:line under the
 * Dart120/HPCCG-SYCL tree) that it replaces.  Vectors handed to hpccg_dev_*
 * functions are DEVICE pointers; the hpccg_api_* functions mirror the reference's
 * C++ calls and take HOST (or device) pointers exactly like the reference does.
 */
#ifndef HPCCG_B200_H
#define HPCCG_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * Library, device and error state
 * ------------------------------------------------------------------------------------------------ */
const char *hpccg_last_error(void);        /* thread-local text of the last non-zero return */
int hpccg_version(void);                   /* 100 * major + minor */
int hpccg_device_count(int *count);
int hpccg_set_device(int device);          /* one process (or thread) per GPU */
int hpccg_device_synchronize(void);

/* Raw device memory for callers without an allocator of their own (tests, the C++ driver). */
int hpccg_dev_malloc(void **ptr, long long bytes);
int hpccg_dev_free(void *ptr);
int hpccg_host_malloc_pinned(void **ptr, long long bytes);
int hpccg_host_free_pinned(void *ptr);
int hpccg_memcpy_h2d(void *dst_dev, const void *src_host, long long bytes, void *stream);
int hpccg_memcpy_d2h(void *dst_host, const void *src_dev, long long bytes, void *stream);
int hpccg_stream_synchronize(void *stream);

/* ------------------------------------------------------------------------------------------------
 * Rank context -- replaces MPI_Comm_rank / MPI_Comm_size(MPI_COMM_WORLD)
 * (generate_matrix.cpp:207-208, make_local_matrix.cpp:75-76, HPCCG.cpp:337,
 * exchange_externals.cpp:68-69).  The context is thread-local; one process per GPU
 * sets it once.  Ranks are stacked in z exactly like the reference's MPI decomposition.
 * ------------------------------------------------------------------------------------------------ */
int hpccg_ctx_set(int rank, int size);
int hpccg_ctx_get(int *rank, int *size);

/* Host-side collective used ONLY during set-up (make_local_matrix): every rank contributes
 * `nbytes` and receives size*nbytes, rank-major.  Replaces the MPI_Allreduce / Irecv / Send / Wait
 * negotiation of make_local_matrix.cpp:185,305,389-411,485-534,546-583.  bench.py installs a
 * torch.distributed (gloo) implementation; the in-process world below installs a thread one. */
typedef int (*hpccg_allgather_fn)(void *user, const void *send, long long nbytes, void *recv);
int hpccg_ctx_set_allgather(hpccg_allgather_fn fn, void *user);

/* In-process world: `size` host threads act as ranks (tests, single-GPU emulation of N ranks). */
int hpccg_local_world_create(int size, void **world);
int hpccg_local_world_bind(void *world, int rank);   /* call on the rank's own thread */
int hpccg_local_world_destroy(void *world);

/* NCCL communicator for the solve (halo send/recv + scalar gathers), one rank per process.
 * Replaces MPI_Irecv/MPI_Send/MPI_Wait in exchange_externals.cpp:87-126 and MPI_Allreduce in
 * ddot.cpp:79-80.  id128 is an ncclUniqueId (128 bytes) created on rank 0 and broadcast by the
 * launcher (bench.py uses torch.distributed). */
int hpccg_nccl_available(void);
int hpccg_nccl_unique_id(void *id128);
int hpccg_nccl_init(const void *id128, int rank, int size);
int hpccg_nccl_finalize(void);

/* ------------------------------------------------------------------------------------------------
 * Device matrix: SELL-C (C = 128 rows per slice, sigma = 1) mirror of HPC_Sparse_Matrix (HPC_Sparse_Matrix.hpp:54-85)
 *   vals[slice][slot][128] (fp64), cols[slice][slot][128] (int32, -1 = padding), slot j = the j-th
 *   STORED entry of the row, so the summation order of HPC_sparsemv.cpp:83-86 is preserved.  A slice is one
 *   contiguous block, which is what the TMA bulk copies of the SpMV kernel move.
 * ------------------------------------------------------------------------------------------------ */
typedef struct hpccg_dev_matrix hpccg_dev_matrix;

/* From assembled host rows (the arrays generate_matrix.cpp:233-289 / make_local_matrix.cpp:239-249
 * leave in the struct).  Column ids must already be local (0 <= col < local_ncol). */
int hpccg_dev_matrix_create(int local_nrow, int local_ncol, const int *nnz_in_row,
                            const double *const *ptr_to_vals_in_row, const int *const *ptr_to_inds_in_row,
                            hpccg_dev_matrix **out);

/* Directly on the device, for sizes the host cannot stage (512^3 is 97 GB of host rows and overflows
 * the reference's `int local_nnz`, generate_matrix.cpp:223).  Rows and their order are those of
 * generate_matrix.cpp:251-289; columns that leave [rank*n, (rank+1)*n) are localised through the two
 * lookup tables (plane position -> local column id), which come from the halo plan
 * (make_local_matrix.cpp:218-249); pass NULL where the rank has no such neighbour. */
int hpccg_dev_matrix_generate(int nx, int ny, int nz, int rank, int size, int stencil /* 27 or 7 */,
                              const int *lower_plane_to_local, const int *upper_plane_to_local, int local_ncol,
                              hpccg_dev_matrix **out);

/* Halo plan (make_local_matrix.cpp:445-599 fields of the struct). Host arrays, copied. */
int hpccg_dev_matrix_set_halo(hpccg_dev_matrix *m, int num_neighbors, const int *neighbors, const int *recv_length,
                              const int *send_length, const int *elements_to_send, int total_to_be_sent);

int hpccg_dev_matrix_destroy(hpccg_dev_matrix *m);

/* Introspection (tests compare the device matrix of both construction routes bit for bit; download returns the
 * canonical column-major [slots][padded_rows] view whatever the device format is). */
int hpccg_dev_matrix_info(const hpccg_dev_matrix *m, int *local_nrow, int *local_ncol, int *slots,
                          long long *padded_rows);
int hpccg_dev_matrix_download(const hpccg_dev_matrix *m, double *vals_host, int *cols_host); /* slots*padded_rows each */
int hpccg_dev_matrix_bytes(const hpccg_dev_matrix *m, long long *bytes);

/* Optional lossless re-encoding of the mirror (SURVEY.md 8 f3 "traffic reduction beyond the algorithmic figure"):
 * a row's PATTERN is its sequence of stored (value, column - row) pairs; the mirror keeps one 16-bit pattern id per
 * row plus the table of distinct patterns, so HPC_sparsemv streams 2 + 16 bytes per row instead of 12*slots + 16.
 * Same values, same columns, same summation order: results are bit-identical to the default format.  A matrix with
 * more than 65535 distinct patterns does not compress and is left as it is (format 0); the call still returns 0. */
int hpccg_dev_matrix_compress(hpccg_dev_matrix *m);
/* format: 0 = SELL-C values + int32 column ids (default, the north-star layout), 1 = pattern-coded. */
int hpccg_dev_matrix_format(const hpccg_dev_matrix *m, int *format, int *patterns);

/* Which data plane the multi-rank solves on this mirror use (decided collectively at the first solve):
 * *peer = 1 when halos (exchange_externals.cpp:84-126) and scalar sums (ddot.cpp:77-82) travel through peer memory
 * inside the kernels, 0 when they are NCCL send/recv + gathers between kernels (or the mirror never took part in a
 * multi-rank solve); *fused_put = 1 when the halo put rides in the kernel that produces p (no exchange launch). */
int hpccg_dev_matrix_comm(const hpccg_dev_matrix *m, int *peer, int *fused_put);

/* ------------------------------------------------------------------------------------------------
 * Kernels (device pointers, asynchronous on `stream` = cudaStream_t or NULL)
 * ------------------------------------------------------------------------------------------------ */
/* y = A x                                       -- HPC_sparsemv.cpp:68-89 */
int hpccg_dev_spmv(const hpccg_dev_matrix *m, const double *x, double *y, void *stream);
/* *result_dev = sum x[i]*y[i], deterministic    -- ddot.cpp:60-74 (local part) */
int hpccg_dev_dot(int n, const double *x, const double *y, double *result_dev, void *stream);
/* w = alpha x + beta y, same three branches     -- waxpby.cpp:69-93 (aliasing w==x, w==y allowed) */
int hpccg_dev_waxpby(int n, double alpha, const double *x, double beta, const double *y, double *w, void *stream);
/* y = A x and *result_dev = x[0..n) . y         -- HPCCG.cpp:379-381 fused */
int hpccg_dev_spmv_dot(const hpccg_dev_matrix *m, const double *x, double *y, double *result_dev, void *stream);
/* x += alpha p ; r -= alpha Ap ; *rr_dev = r.r  -- HPCCG.cpp:383-384 + :367 fused; alpha read from device */
int hpccg_dev_update_xr_dot(int n, const double *alpha_dev, const double *p, const double *Ap, double *x, double *r,
                            double *rr_dev, void *stream);
/* p = r + beta p, beta read from device         -- HPCCG.cpp:369 */
int hpccg_dev_p_update(int n, const double *beta_dev, const double *r, double *p, void *stream);
/* send_buffer[i] = x[elements_to_send[i]]       -- exchange_externals.cpp:103 */
int hpccg_dev_halo_pack(const hpccg_dev_matrix *m, const double *x, double *send_buffer_dev, void *stream);
/* max_i |v1[i]-v2[i]| -> *result_dev            -- compute_residual.cpp:59-81 (local part) */
int hpccg_dev_max_abs_diff(int n, const double *v1, const double *v2, double *result_dev, void *stream);

/* ------------------------------------------------------------------------------------------------
 * The device-resident CG loop -- HPCCG.cpp:312-402.
 *   b, x: device, local_nrow doubles (x is updated in place).  The loop runs without host
 *   synchronisation: alpha, beta, rtrans live on the device, the `normr > tolerance` exit of
 *   HPCCG.cpp:358 is honoured by a device flag that turns the remaining launches into no-ops.
 *   hist_host (may be NULL): max_iter doubles, [0] = initial residual, [k] = normr of iteration k
 *   (what HPCCG.cpp:356,372-373 would print with print_freq = 1), NaN where no iteration ran.
 *   times (may be NULL): 16 doubles; [0..6] as HPCCG.cpp:389-399 (see DESIGN.md for the fused attribution),
 *   [7] seconds in the fused SpMV+p.Ap kernel, [8] in the fused x/r-update+r.r kernel, [9] in the p-update
 *   kernel, [10] number of timed iterations (raw CUDA-event sums; need HPCCG_SOLVE_TIMERS);
 *   loop_ms (may be NULL): CUDA-event time of iterations 1..niters only.
 *   flags: HPCCG_SOLVE_* bits.
 * With an NCCL communicator (hpccg_nccl_init) and ctx size > 1 this is one rank of a z-stacked job.
 * ------------------------------------------------------------------------------------------------ */
#define HPCCG_SOLVE_DEFAULT 0
#define HPCCG_SOLVE_UNFUSED 1   /* literal reference kernel sequence (validation / exact per-kernel times) */
#define HPCCG_SOLVE_NO_OVERLAP 2 /* halo exchange not overlapped with the interior SpMV */
#define HPCCG_SOLVE_TIMERS 4    /* record per-kernel CUDA events for times[1..5] */
#define HPCCG_SOLVE_GRAPH 16    /* no TIMERS: a repeated solve (same b, x, max_iter, tolerance, flags) is captured once into a CUDA graph
                                  * and replayed -- for launch-bound sizes; loop_ms then covers the whole solve.  Single rank, or one
                                  * rank of a multi-GPU job on the peer-memory plane (ranks need not agree on replaying) */
#define HPCCG_SOLVE_EAGER_X 32  /* x += alpha p in the kernel right after the SpMV (48 + 24 B/row) instead of deferred into the next
                                  * p-update (24 + 40 B/row, default); same arithmetic, kept for A/B measurements */
#define HPCCG_SOLVE_PERSISTENT 64 /* single rank, launch-bound sizes (default format, up to ~9600 rows at 27 slots): the WHOLE solve is
                                  * one kernel of one 16-CTA thread-block cluster -- matrix blocks resident in shared memory, partial
                                  * sums and neighbouring r values exchanged through distributed shared memory, two cluster barriers
                                  * per iteration; ignored (normal loop) when the matrix does not qualify; loop_ms = the whole solve */
#define HPCCG_SOLVE_NCCL_ONLY 8 /* multi-GPU: NCCL send/recv + gathers between kernels instead of peer memory inside them */
int hpccg_dev_cg_solve(hpccg_dev_matrix *m, const double *b, double *x, int max_iter, double tolerance, int *niters,
                       double *normr, double *hist_host, double *times, double *loop_ms, int flags, void *stream);

/* Same loop for `nranks` z-stacked ranks that all live in THIS process on the current device,
 * advanced in lock step on one stream (halo = device copies, scalar reduction in rank order).
 * This is how N-rank numerics are checked on a single GPU. */
int hpccg_dev_cg_solve_group(int nranks, hpccg_dev_matrix *const *m, const double *const *b, double *const *x,
                             int max_iter, double tolerance, int *niters, double *normr, double *hist_host,
                             double *loop_ms, int flags, void *stream);

/* Number of kernels this library has launched in this process (bench.py's "gpu_launches"). */
long long hpccg_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * C views of the reference-named C++ API (HOST pointers, reference semantics), for FFI callers.
 * `A` is an HPC_Sparse_Matrix* (hpccg-sycl_b200/csrc/include/HPC_Sparse_Matrix.hpp).
 * ------------------------------------------------------------------------------------------------ */
/* Options the reference fixes at compile time: stencil (generate_matrix.cpp:219, 27 or 7) and whether
 * generate_matrix materialises the host row arrays (0 = device-only, needed beyond 430^3). Thread-local. */
int hpccg_api_set_options(int stencil, int host_arrays);
/* HPCCG() prints "Initial Residual" / "Iteration = k   Residual" on rank 0 like HPCCG.cpp:356,372-373; 0 silences it. */
int hpccg_api_set_print(int on);
/* Format of the device mirrors created from now on by this thread: 0 (default) or 1 (hpccg_dev_matrix_compress is applied
 * when the mirror is built).  The environment variable HPCCG_B200_FORMAT=pattern selects 1 when this was never called. */
int hpccg_api_set_matrix_format(int format);
/* generate_matrix.hpp:58 */
int hpccg_api_generate_matrix(int nx, int ny, int nz, void **A, double **x, double **b, double **xexact);
/* read_HPC_row.hpp:56-57 -- matrix-file input (deprecated upstream, README.md:114-118); rows dealt to the ranks as
 * read_HPC_row.cpp:257-267.  Returns non-zero (file unreadable / malformed) where the reference exit()s. */
int hpccg_api_read_HPC_row(const char *data_file, void **A, double **x, double **b, double **xexact);
/* make_local_matrix.hpp:48 */
int hpccg_api_make_local_matrix(void *A);
/* HPCCG.hpp:61-63 */
int hpccg_api_HPCCG(void *A, double *b, double *x, int max_iter, double tolerance, int *niters, double *normr,
                    double *times);
/* HPC_sparsemv.hpp:55-56 */
int hpccg_api_HPC_sparsemv(void *A, const double *x, double *y);
/* ddot.hpp:55-56 */
int hpccg_api_ddot(int n, const double *x, const double *y, double *result, double *time_allreduce);
/* waxpby.hpp:51-53 */
int hpccg_api_waxpby(int n, double alpha, const double *x, double beta, const double *y, double *w);
/* exchange_externals.hpp:49 */
int hpccg_api_exchange_externals(void *A, double *x);
/* compute_residual.hpp:50-51 */
int hpccg_api_compute_residual(int n, const double *v1, const double *v2, double *residual);
/* HPC_Sparse_Matrix.hpp:88; also frees the vectors generate_matrix handed out when they are passed */
int hpccg_api_destroyMatrix(void *A);
int hpccg_api_free_vectors(double *x, double *b, double *xexact);
/* Struct readers for FFI callers: scalar fields by name, arrays copied out (returns count, <0 = unknown). */
long long hpccg_api_matrix_scalar(const void *A, const char *field);
long long hpccg_api_matrix_array(const void *A, const char *field, void *dst, long long capacity);
/* The opaque device mirror owned by A (created on first use). */
int hpccg_api_matrix_device(void *A, hpccg_dev_matrix **out);
/* Residual history of the last hpccg_api_HPCCG call on this thread (max_iter doubles). */
int hpccg_api_last_history(double *hist, int capacity);
/* YAML report as main.cpp:214-305 builds it, through this library's YAML_Doc; returns length. */
int hpccg_api_yaml_report(int nx, int ny, int nz, int niters, double normr, const double *times, double total_nrow,
                          double total_nnz, int ranks, int omp_threads, const double *t4stats, char *out, int capacity);

#ifdef __cplusplus
}
#endif
#endif /* HPCCG_B200_H */
