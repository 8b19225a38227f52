/* oracle/hpccg_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * A plain-C, CPU restatement of the HPCCG reference's hot path and of the set-up
 * that feeds it.  It exists so that parity tests have a checker that travels to
 * the GPU box even where /root/reference does not.  Every function cites the
 * reference file:line it restates.  The restatement is PINNED: tests/test_oracle.py
 * checks it bit-for-bit against the real reference compiled from /root/reference
 * (oracle/_ref/libhpccg_ref_*.so, see build.sh) and against the committed golden
 * fixtures in tests/golden/ that were generated from that reference.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call
 * this.  Compile with -ffp-contract=off: the reference's g++ -O3 x86-64 build
 * emits no FMA, and neither may this.
 *
 * All `size` ranks of a z-stacked world live in one orc_world; the MPI exchanges
 * of the reference become direct reads of the peer's data, and MPI_Allreduce
 * becomes a sum in rank order (the order oracle/mpi_shim uses).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int start_row, stop_row, total_nrow, local_nrow, local_ncol;
  long long total_nnz, local_nnz; /* claimed 27*n, generate_matrix.cpp:223-226 */
  long long nnz_sum;              /* actual stored entries */
  int *nnz_in_row;
  long long *row_start; /* offset of row i in inds/vals (ptr_to_*_in_row[i] - list_of_*) */
  long long *diag;      /* offset of the diagonal entry (ptr_to_diags[i] - list_of_vals) */
  int *inds;
  double *vals;
  double *x, *b, *xexact;
  /* halo plan, make_local_matrix.cpp */
  int num_external, num_send_neighbors, total_to_be_sent;
  int *external_index, *external_local_index, *elements_to_send;
  int *neighbors, *recv_length, *send_length;
  double *send_buffer;
  /* scratch kept between the two phases of make_local */
  int *new_external;           /* global ids in local-number order */
  int *new_external_processor; /* owner of each, same order */
} orc_rank;

typedef struct {
  int nx, ny, nz, size, stencil7;
  orc_rank *r;
} orc_world;

/* ---- generate_matrix.cpp:196-307 ---------------------------------------------------- */
static void orc_generate(orc_world *w, int rank) {
  const int nx = w->nx, ny = w->ny, nz = w->nz, size = w->size;
  orc_rank *A = &w->r[rank];
  const int local_nrow = nx * ny * nz;                 /* :221 */
  const long long local_nnz = 27LL * local_nrow;       /* :223 (int there; overflows past 430^3) */
  const int total_nrow = local_nrow * size;            /* :225 */
  const int start_row = local_nrow * rank;             /* :228 */
  A->start_row = start_row;
  A->stop_row = start_row + local_nrow - 1;            /* :229 */
  A->total_nrow = total_nrow;
  A->total_nnz = 27LL * (long long)total_nrow;         /* :226 */
  A->local_nrow = local_nrow;
  A->local_ncol = local_nrow;                          /* :303 */
  A->local_nnz = local_nnz;
  A->nnz_in_row = (int *)malloc(sizeof(int) * local_nrow);
  A->row_start = (long long *)malloc(sizeof(long long) * local_nrow);
  A->diag = (long long *)malloc(sizeof(long long) * local_nrow);
  A->inds = (int *)malloc(sizeof(int) * local_nnz);
  A->vals = (double *)malloc(sizeof(double) * local_nnz);
  A->x = (double *)malloc(sizeof(double) * local_nrow);
  A->b = (double *)malloc(sizeof(double) * local_nrow);
  A->xexact = (double *)malloc(sizeof(double) * local_nrow);
  long long cur = 0;
  for (int iz = 0; iz < nz; iz++)
    for (int iy = 0; iy < ny; iy++)
      for (int ix = 0; ix < nx; ix++) {                /* :251-253 */
        const int curlocalrow = iz * nx * ny + iy * nx + ix;
        const int currow = start_row + curlocalrow;
        int nnzrow = 0;
        A->row_start[curlocalrow] = cur;
        for (int sz = -1; sz <= 1; sz++)
          for (int sy = -1; sy <= 1; sy++)
            for (int sx = -1; sx <= 1; sx++) {         /* :259-261 */
              const int curcol = currow + sz * nx * ny + sy * nx + sx;
              /* x and y are bounded by the block, z by the global row range (:266) */
              if (ix + sx >= 0 && ix + sx < nx && iy + sy >= 0 && iy + sy < ny && curcol >= 0 && curcol < total_nrow) {
                if (!w->stencil7 || sz * sz + sy * sy + sx * sx <= 1) { /* :267 */
                  if (curcol == currow) {
                    A->diag[curlocalrow] = cur;
                    A->vals[cur] = 27.0;               /* :270 */
                  } else {
                    A->vals[cur] = -1.0;               /* :273 */
                  }
                  A->inds[cur] = curcol;
                  cur++;
                  nnzrow++;
                }
              }
            }
        A->nnz_in_row[curlocalrow] = nnzrow;
        A->x[curlocalrow] = 0.0;                        /* :284 */
        A->b[curlocalrow] = 27.0 - ((double)(nnzrow - 1)); /* :285 */
        A->xexact[curlocalrow] = 1.0;                   /* :286 */
      }
  A->nnz_sum = cur;
}

/* ---- make_local_matrix.cpp:58-610, receive side (:105-250, :358-367, :466-505) ------- */
static void orc_make_local_recv_side(orc_world *w, int rank, int *seen /* total_nrow ints, all -1 */) {
  orc_rank *A = &w->r[rank];
  const int size = w->size, local_nrow = A->local_nrow;
  int cap = 1024, num_external = 0;
  int *external_index = (int *)malloc(sizeof(int) * cap);

  /* scan: local columns shifted, externals recorded in first-encounter order (:116-153) */
  for (int i = 0; i < local_nrow; i++) {
    int *row = A->inds + A->row_start[i];
    for (int j = 0; j < A->nnz_in_row[i]; j++) {
      const int cur_ind = row[j];
      if (A->start_row <= cur_ind && cur_ind <= A->stop_row) {
        row[j] -= A->start_row;
      } else {
        if (seen[cur_ind] < 0) {
          if (num_external == cap) {
            cap *= 2;
            external_index = (int *)realloc(external_index, sizeof(int) * cap);
          }
          seen[cur_ind] = num_external;
          external_index[num_external++] = cur_ind;
        }
        row[j] = -(row[j] + 1); /* marked external by negation (:138,:150) */
      }
    }
  }
  A->num_external = num_external;
  A->external_index = external_index;

  /* owner of each external: last rank whose start_row <= index (:181-203) */
  int *external_processor = (int *)malloc(sizeof(int) * (num_external + 1));
  for (int i = 0; i < num_external; i++)
    for (int j = size - 1; j >= 0; j--)
      if (w->r[j].start_row <= external_index[i]) {
        external_processor[i] = j;
        break;
      }

  /* number externals so that one owner's are consecutive (:218-230) */
  int *external_local_index = (int *)malloc(sizeof(int) * (num_external + 1));
  int count = local_nrow;
  for (int i = 0; i < num_external; i++) external_local_index[i] = -1;
  for (int i = 0; i < num_external; i++)
    if (external_local_index[i] == -1) {
      external_local_index[i] = count++;
      for (int j = i + 1; j < num_external; j++)
        if (external_processor[j] == external_processor[i]) external_local_index[j] = count++;
    }
  A->external_local_index = external_local_index;

  /* rewrite the marked columns (:239-249) */
  for (int i = 0; i < local_nrow; i++) {
    int *row = A->inds + A->row_start[i];
    for (int j = 0; j < A->nnz_in_row[i]; j++)
      if (row[j] < 0) {
        const int cur_ind = -row[j] - 1;
        row[j] = external_local_index[seen[cur_ind]];
      }
  }

  /* owners and global ids in local-number order (:251-255, :466-469) */
  A->new_external_processor = (int *)malloc(sizeof(int) * (num_external + 1));
  A->new_external = (int *)malloc(sizeof(int) * (num_external + 1));
  for (int i = 0; i < num_external; i++) {
    A->new_external_processor[external_local_index[i] - local_nrow] = external_processor[i];
    A->new_external[external_local_index[i] - local_nrow] = external_index[i];
  }

  /* recv_list = owners in the order their groups appear (:358-367); lengths (:489-505) */
  A->neighbors = (int *)malloc(sizeof(int) * (size + 1));
  A->recv_length = (int *)calloc(size + 1, sizeof(int));
  A->send_length = (int *)calloc(size + 1, sizeof(int));
  int nn = 0;
  for (int i = 0; i < num_external; i++) {
    if (i == 0 || A->new_external_processor[i - 1] != A->new_external_processor[i]) A->neighbors[nn++] = A->new_external_processor[i];
    A->recv_length[nn - 1]++;
  }
  A->num_send_neighbors = nn; /* provisional: send-only neighbours are appended by the send side */

  for (int i = 0; i < num_external; i++) seen[external_index[i]] = -1; /* leave the map clean for the next rank */
  free(external_processor);
  A->local_ncol = local_nrow + num_external; /* :595 */
}

/* ---- make_local_matrix.cpp send side (:283-316, :376-440, :507-587) -------------------- */
static void orc_make_local_send_side(orc_world *w, int rank) {
  orc_rank *A = &w->r[rank];
  const int size = w->size;
  /* Ranks that list me as an owner but are not in my receive list are appended
   * (:418-433).  The reference appends them in message-arrival order; for the
   * symmetric stencils here the set is empty, and this restatement uses rank order. */
  for (int q = 0; q < size; q++) {
    if (q == rank) continue;
    const orc_rank *B = &w->r[q];
    int wants = 0;
    for (int i = 0; i < B->num_external; i++)
      if (B->new_external_processor[i] == rank) { wants = 1; break; }
    if (!wants) continue;
    int found = 0;
    for (int i = 0; i < A->num_send_neighbors; i++)
      if (A->neighbors[i] == q) found = 1;
    if (!found) {
      A->neighbors[A->num_send_neighbors] = q;
      A->recv_length[A->num_send_neighbors] = 0;
      A->num_send_neighbors++;
    }
  }
  /* send_length[i] = how many of my rows neighbour i asked for (:507-519);
   * elements_to_send = their global ids in the asker's order, made local (:545-587) */
  int total = 0;
  for (int i = 0; i < A->num_send_neighbors; i++) {
    const orc_rank *B = &w->r[A->neighbors[i]];
    int len = 0;
    for (int k = 0; k < B->num_external; k++)
      if (B->new_external_processor[k] == rank) len++;
    A->send_length[i] = len;
    total += len;
  }
  A->total_to_be_sent = total; /* equals the allreduce-decoded value of :305-316 */
  A->elements_to_send = (int *)malloc(sizeof(int) * (total + 1));
  int pos = 0;
  for (int i = 0; i < A->num_send_neighbors; i++) {
    const orc_rank *B = &w->r[A->neighbors[i]];
    for (int k = 0; k < B->num_external; k++)
      if (B->new_external_processor[k] == rank) A->elements_to_send[pos++] = B->new_external[k] - A->start_row;
  }
  A->send_buffer = (double *)malloc(sizeof(double) * (total + 1)); /* :598-599 */
}

orc_world *orc_create(int nx, int ny, int nz, int size, int stencil7) {
  if (nx <= 0 || ny <= 0 || nz <= 0 || size <= 0) return NULL;
  orc_world *w = (orc_world *)calloc(1, sizeof(orc_world));
  w->nx = nx; w->ny = ny; w->nz = nz; w->size = size; w->stencil7 = stencil7;
  w->r = (orc_rank *)calloc(size, sizeof(orc_rank));
  for (int r = 0; r < size; r++) orc_generate(w, r);
  if (size > 1) { /* main.cpp:174-182: make_local_matrix only in the MPI build */
    int *seen = (int *)malloc(sizeof(int) * (size_t)w->r[0].total_nrow);
    for (int i = 0; i < w->r[0].total_nrow; i++) seen[i] = -1;
    for (int r = 0; r < size; r++) orc_make_local_recv_side(w, r, seen);
    for (int r = 0; r < size; r++) orc_make_local_send_side(w, r);
    free(seen);
  }
  return w;
}

void orc_destroy(orc_world *w) {
  if (!w) return;
  for (int r = 0; r < w->size; r++) {
    orc_rank *A = &w->r[r];
    free(A->nnz_in_row); free(A->row_start); free(A->diag); free(A->inds); free(A->vals);
    free(A->x); free(A->b); free(A->xexact);
    free(A->external_index); free(A->external_local_index); free(A->elements_to_send);
    free(A->neighbors); free(A->recv_length); free(A->send_length); free(A->send_buffer);
    free(A->new_external); free(A->new_external_processor);
  }
  free(w->r);
  free(w);
}

int orc_variant(void) { return 3; }
int orc_threads(void) { return 1; }

long long orc_scalar(orc_world *w, int rank, const char *s) {
  const orc_rank *A = &w->r[rank];
  if (!strcmp(s, "start_row")) return A->start_row;
  if (!strcmp(s, "stop_row")) return A->stop_row;
  if (!strcmp(s, "total_nrow")) return A->total_nrow;
  if (!strcmp(s, "total_nnz")) return A->total_nnz;
  if (!strcmp(s, "local_nrow")) return A->local_nrow;
  if (!strcmp(s, "local_ncol")) return A->local_ncol;
  if (!strcmp(s, "local_nnz")) return (int)A->local_nnz; /* the reference field is an int */
  if (!strcmp(s, "nnz_sum")) return A->nnz_sum;
  if (!strcmp(s, "num_external")) return A->num_external;
  if (!strcmp(s, "num_send_neighbors")) return A->num_send_neighbors;
  if (!strcmp(s, "total_to_be_sent")) return A->total_to_be_sent;
  return -1;
}

static long long copy_out(const void *src, size_t elt, long long n, void *dst, long long cap) {
  if (dst && cap >= n && n > 0) memcpy(dst, src, elt * (size_t)n);
  return n;
}

long long orc_array(orc_world *w, int rank, const char *s, void *dst, long long cap) {
  const orc_rank *A = &w->r[rank];
  const long long n = A->local_nrow;
  if (!strcmp(s, "nnz_in_row")) return copy_out(A->nnz_in_row, sizeof(int), n, dst, cap);
  if (!strcmp(s, "list_of_inds")) return copy_out(A->inds, sizeof(int), A->nnz_sum, dst, cap);
  if (!strcmp(s, "list_of_vals")) return copy_out(A->vals, sizeof(double), A->nnz_sum, dst, cap);
  if (!strcmp(s, "x")) return copy_out(A->x, sizeof(double), n, dst, cap);
  if (!strcmp(s, "b")) return copy_out(A->b, sizeof(double), n, dst, cap);
  if (!strcmp(s, "xexact")) return copy_out(A->xexact, sizeof(double), n, dst, cap);
  if (!strcmp(s, "ind_offsets") || !strcmp(s, "val_offsets")) return copy_out(A->row_start, sizeof(long long), n, dst, cap);
  if (!strcmp(s, "diag_offsets")) return copy_out(A->diag, sizeof(long long), n, dst, cap);
  if (!strcmp(s, "external_index")) return copy_out(A->external_index, sizeof(int), A->num_external, dst, cap);
  if (!strcmp(s, "external_local_index")) return copy_out(A->external_local_index, sizeof(int), A->num_external, dst, cap);
  if (!strcmp(s, "elements_to_send")) return copy_out(A->elements_to_send, sizeof(int), A->total_to_be_sent, dst, cap);
  if (!strcmp(s, "neighbors")) return copy_out(A->neighbors, sizeof(int), A->num_send_neighbors, dst, cap);
  if (!strcmp(s, "recv_length")) return copy_out(A->recv_length, sizeof(int), A->num_send_neighbors, dst, cap);
  if (!strcmp(s, "send_length")) return copy_out(A->send_length, sizeof(int), A->num_send_neighbors, dst, cap);
  return -1;
}

/* ---- HPC_sparsemv.cpp:68-89 -------------------------------------------------------------- */
static void orc_sparsemv(const orc_rank *A, const double *x, double *y) {
  for (int i = 0; i < A->local_nrow; i++) {
    double sum = 0.0;
    const double *cur_vals = A->vals + A->row_start[i];
    const int *cur_inds = A->inds + A->row_start[i];
    const int cur_nnz = A->nnz_in_row[i];
    for (int j = 0; j < cur_nnz; j++) sum += cur_vals[j] * x[cur_inds[j]]; /* stored order, :83-86 */
    y[i] = sum;
  }
}

/* ---- ddot.cpp:60-88, local part -------------------------------------------------------------- */
static double orc_ddot_local(int n, const double *x, const double *y) {
  double local_result = 0.0;
  if (y == x)
    for (int i = 0; i < n; i++) local_result += x[i] * x[i];
  else
    for (int i = 0; i < n; i++) local_result += x[i] * y[i];
  return local_result;
}

/* ---- waxpby.cpp:69-93 ---------------------------------------------------------------------------- */
int orc_waxpby(int n, double alpha, const double *x, double beta, const double *y, double *w) {
  if (alpha == 1.0)
    for (int i = 0; i < n; i++) w[i] = x[i] + beta * y[i];
  else if (beta == 1.0)
    for (int i = 0; i < n; i++) w[i] = alpha * x[i] + y[i];
  else
    for (int i = 0; i < n; i++) w[i] = alpha * x[i] + beta * y[i];
  return 0;
}

int orc_ddot_raw(int n, const double *x, const double *y, double *result) {
  *result = orc_ddot_local(n, x, y);
  return 0;
}

/* ---- exchange_externals.cpp:51-131, all ranks at once --------------------------------------------- */
static void orc_exchange_all(orc_world *w, double **x) {
  if (w->size == 1) return;
  for (int r = 0; r < w->size; r++) { /* gather, :103 */
    orc_rank *A = &w->r[r];
    for (int i = 0; i < A->total_to_be_sent; i++) A->send_buffer[i] = x[r][A->elements_to_send[i]];
  }
  for (int r = 0; r < w->size; r++) { /* receive into the tail of x in neighbour order, :84-95 */
    orc_rank *A = &w->r[r];
    double *x_external = x[r] + A->local_nrow;
    for (int i = 0; i < A->num_send_neighbors; i++) {
      const orc_rank *B = &w->r[A->neighbors[i]];
      const double *sb = B->send_buffer; /* the slice B sends to r, :109-115 */
      for (int k = 0; k < B->num_send_neighbors; k++) {
        if (B->neighbors[k] == r) break;
        sb += B->send_length[k];
      }
      for (int k = 0; k < A->recv_length[i]; k++) x_external[k] = sb[k];
      x_external += A->recv_length[i];
    }
  }
}

int orc_spmv(orc_world *w, double **x, double **y, int exchange, int reps) {
  for (int it = 0; it < (reps < 1 ? 1 : reps); it++) {
    if (exchange) orc_exchange_all(w, x);
    for (int r = 0; r < w->size; r++) orc_sparsemv(&w->r[r], x[r], y[r]);
  }
  return 0;
}

/* global ddot: local sums added in rank order (mpi_shim's MPI_Allreduce order) */
static double orc_ddot_all(orc_world *w, double **x, double **y) {
  double g = 0.0;
  for (int r = 0; r < w->size; r++) {
    double l = orc_ddot_local(w->r[r].local_nrow, x[r], y[r]);
    g = (r == 0) ? l : g + l;
  }
  return g;
}

int orc_ddot(orc_world *w, double **x, double **y, double *result) {
  double g = orc_ddot_all(w, x, y);
  for (int r = 0; r < w->size; r++) result[r] = g;
  return 0;
}

/* ---- HPCCG.cpp:312-402, every rank in lock step ------------------------------------------------------ */
int orc_solve(orc_world *w, int max_iter, double tol, int hist, double *hist_out, int *niters_out, double *normr_out,
              double *times_out, double **x_out) {
  (void)hist;
  const int S = w->size;
  double **x = (double **)malloc(sizeof(double *) * S), **b = (double **)malloc(sizeof(double *) * S);
  double **r = (double **)malloc(sizeof(double *) * S), **p = (double **)malloc(sizeof(double *) * S);
  double **Ap = (double **)malloc(sizeof(double *) * S);
  for (int q = 0; q < S; q++) {
    orc_rank *A = &w->r[q];
    for (int i = 0; i < A->local_nrow; i++) A->x[i] = 0.0;
    x[q] = A->x; b[q] = A->b;
    r[q] = (double *)malloc(sizeof(double) * A->local_nrow);   /* :327 */
    p[q] = (double *)malloc(sizeof(double) * A->local_ncol);   /* :328 */
    Ap[q] = (double *)malloc(sizeof(double) * A->local_nrow);  /* :329 */
  }
  double normr = 0.0, rtrans = 0.0, oldrtrans = 0.0;
  int niters = 0;
  if (hist_out) for (int k = 0; k < max_iter; k++) hist_out[k] = NAN;

  for (int q = 0; q < S; q++) orc_waxpby(w->r[q].local_nrow, 1.0, x[q], 0.0, x[q], p[q]);  /* :347 */
  orc_exchange_all(w, p);                                                                   /* :349 */
  for (int q = 0; q < S; q++) orc_sparsemv(&w->r[q], p[q], Ap[q]);                          /* :351 */
  for (int q = 0; q < S; q++) orc_waxpby(w->r[q].local_nrow, 1.0, b[q], -1.0, Ap[q], r[q]); /* :352 */
  rtrans = orc_ddot_all(w, r, r);                                                           /* :353 */
  normr = sqrt(rtrans);                                                                     /* :354 */
  if (hist_out && max_iter > 0) hist_out[0] = normr;

  for (int k = 1; k < max_iter && normr > tol; k++) {                                       /* :358 */
    if (k == 1) {
      for (int q = 0; q < S; q++) orc_waxpby(w->r[q].local_nrow, 1.0, r[q], 0.0, r[q], p[q]); /* :362 */
    } else {
      oldrtrans = rtrans;
      rtrans = orc_ddot_all(w, r, r);                                                       /* :367 */
      double beta = rtrans / oldrtrans;                                                     /* :368 */
      for (int q = 0; q < S; q++) orc_waxpby(w->r[q].local_nrow, 1.0, r[q], beta, p[q], p[q]); /* :369 */
    }
    normr = sqrt(rtrans);                                                                   /* :371 */
    if (hist_out) hist_out[k] = normr; /* what :372-373 prints when print_freq is 1 */
    orc_exchange_all(w, p);                                                                 /* :377 */
    for (int q = 0; q < S; q++) orc_sparsemv(&w->r[q], p[q], Ap[q]);                        /* :379 */
    double alpha = orc_ddot_all(w, p, Ap);                                                  /* :381 */
    alpha = rtrans / alpha;                                                                 /* :382 */
    for (int q = 0; q < S; q++) {
      orc_waxpby(w->r[q].local_nrow, 1.0, x[q], alpha, p[q], x[q]);                         /* :383 */
      orc_waxpby(w->r[q].local_nrow, 1.0, r[q], -alpha, Ap[q], r[q]);                       /* :384 */
    }
    niters = k;                                                                             /* :385 */
  }
  if (niters_out) *niters_out = niters;
  if (normr_out) *normr_out = normr;
  if (times_out) for (int i = 0; i < 7; i++) times_out[i] = 0.0;
  for (int q = 0; q < S; q++) {
    if (x_out && x_out[q]) memcpy(x_out[q], x[q], sizeof(double) * w->r[q].local_nrow);
    free(r[q]); free(p[q]); free(Ap[q]);
  }
  free(x); free(b); free(r); free(p); free(Ap);
  return 0;
}

/* ---- compute_residual.cpp:59-81 ------------------------------------------------------------------------ */
int orc_compute_residual(orc_world *w, double **x, double *res_per_rank) {
  double g = 0.0;
  for (int q = 0; q < w->size; q++) {
    const orc_rank *A = &w->r[q];
    double local_residual = 0.0;
    for (int i = 0; i < A->local_nrow; i++) {
      double diff = fabs(x[q][i] - A->xexact[i]);
      if (diff > local_residual) local_residual = diff;
    }
    if (local_residual > g) g = local_residual; /* MPI_MAX, :73 */
  }
  for (int q = 0; q < w->size; q++) res_per_rank[q] = g;
  return 0;
}
