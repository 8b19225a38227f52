// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, not product code.
//
// C-ABI wrapper around the UNMODIFIED reference sources under /root/reference
// (compiled where they lie by oracle/Makefile into oracle/_ref/).  It lets the
// tests and bench.py's cpu_baseline leg call the reference's own
// generate_matrix / make_local_matrix / HPC_sparsemv / ddot / waxpby /
// exchange_externals / HPCCG / compute_residual / YAML_Doc and read back every
// array they produce.  Nothing in the product path may load this.
//
// Build variants (REF_VARIANT): 0 serial, 1 OpenMP (-DUSING_OMP), 2 multi-rank
// (-DUSING_MPI against oracle/mpi_shim, ranks are threads).
//
// Symbols provided by the Makefile's renamed second compilations:
//   generate_matrix_7pt : generate_matrix.cpp with `use_7pt_stencil = true`
//                         (generate_matrix.cpp:219 is a hard-coded local bool)
//   HPCCG_hist          : HPCCG.cpp with print_freq forced to 1 (HPCCG.cpp:342-344)
//                         so that every iteration's residual is printed.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#ifdef USING_OMP
#include <omp.h>
#endif
#ifdef USING_MPI
#include <mpi.h>
#endif

#include "HPCCG.hpp"
#include "HPC_Sparse_Matrix.hpp"
#include "HPC_sparsemv.hpp"
#include "YAML_Doc.hpp"
#include "compute_residual.hpp"
#include "ddot.hpp"
#include "generate_matrix.hpp"
#include "mytimer.hpp"
#include "read_HPC_row.hpp"
#include "waxpby.hpp"
#ifdef USING_MPI
#include "exchange_externals.hpp"
#include "make_local_matrix.hpp"
#endif

void generate_matrix_7pt(int nx, int ny, int nz, HPC_Sparse_Matrix **A, double **x, double **b, double **xexact);
int HPCCG_hist(HPC_Sparse_Matrix *A, double *const b, double *const x, const int max_iter, const double tolerance,
               int &niters, double &normr, double *times);

#ifndef REF_VARIANT
#define REF_VARIANT 0
#endif

namespace {

struct RefRank {
  HPC_Sparse_Matrix *A = nullptr;
  double *x = nullptr, *b = nullptr, *xexact = nullptr;
  long long nnz_sum = 0;
};

struct RefWorld {
  int nx, ny, nz, size, stencil7;
  std::vector<RefRank> ranks;
  std::string file;  // non-empty: matrix read by the reference's read_HPC_row (read_HPC_row.cpp:217-373)
};

void run_ranks(int size, void (*fn)(int, void *), void *arg) {
#ifdef USING_MPI
  hpccg_shim_run(size, fn, arg);
#else
  (void)size;
  fn(0, arg);
#endif
}

void gen_rank(int rank, void *arg) {
  RefWorld *w = static_cast<RefWorld *>(arg);
  RefRank &rr = w->ranks[rank];
  if (!w->file.empty()) {
    read_HPC_row(const_cast<char *>(w->file.c_str()), &rr.A, &rr.x, &rr.b, &rr.xexact);
    // read_HPC_row.cpp never stores the two list pointers in the struct (:356-370); they are the first row's pointers
    rr.A->list_of_vals = rr.A->ptr_to_vals_in_row[0];
    rr.A->list_of_inds = rr.A->ptr_to_inds_in_row[0];
    for (int i = 0; i < rr.A->local_nrow; ++i) rr.A->ptr_to_diags[i] = rr.A->list_of_vals;  // left uninitialised there
  } else if (w->stencil7) generate_matrix_7pt(w->nx, w->ny, w->nz, &rr.A, &rr.x, &rr.b, &rr.xexact);
  else generate_matrix(w->nx, w->ny, w->nz, &rr.A, &rr.x, &rr.b, &rr.xexact);
  long long s = 0;
  for (int i = 0; i < rr.A->local_nrow; ++i) s += rr.A->nnz_in_row[i];
  rr.nnz_sum = s;
#ifdef USING_MPI
  make_local_matrix(rr.A);
#endif
}

template <typename T>
long long copy_out(const T *src, long long n, void *dst, long long cap) {
  if (dst && cap >= n) std::memcpy(dst, src, sizeof(T) * n);
  return n;
}

}  // namespace

extern "C" {

int ref_variant() { return REF_VARIANT; }

int ref_threads() {
#ifdef USING_OMP
  int n = 1;
#pragma omp parallel
  n = omp_get_num_threads();
  return n;
#else
  return 1;
#endif
}

void *ref_create(int nx, int ny, int nz, int size, int stencil7) {
#ifndef USING_MPI
  if (size != 1) return nullptr;
#endif
  RefWorld *w = new RefWorld{nx, ny, nz, size, stencil7, std::vector<RefRank>(size), std::string()};
  run_ranks(size, gen_rank, w);
  return w;
}

void *ref_create_from_file(const char *path, int size) {
#ifndef USING_MPI
  if (size != 1) return nullptr;
#endif
  RefWorld *w = new RefWorld{0, 0, 0, size, 0, std::vector<RefRank>(size), std::string(path)};
  run_ranks(size, gen_rank, w);
  return w;
}

void ref_destroy(void *h) {
  RefWorld *w = static_cast<RefWorld *>(h);
  if (!w) return;
  for (RefRank &rr : w->ranks) {
#ifndef USING_MPI
    destroyMatrix(rr.A);
#else
    destroyMatrix(rr.A);
#endif
    delete[] rr.x;
    delete[] rr.b;
    delete[] rr.xexact;
  }
  delete w;
}

long long ref_scalar(void *h, int rank, const char *name) {
  RefWorld *w = static_cast<RefWorld *>(h);
  const RefRank &rr = w->ranks[rank];
  const HPC_Sparse_Matrix *A = rr.A;
  std::string s(name);
  if (s == "start_row") return A->start_row;
  if (s == "stop_row") return A->stop_row;
  if (s == "total_nrow") return A->total_nrow;
  if (s == "total_nnz") return A->total_nnz;
  if (s == "local_nrow") return A->local_nrow;
  if (s == "local_ncol") return A->local_ncol;
  if (s == "local_nnz") return A->local_nnz;
  if (s == "nnz_sum") return rr.nnz_sum;
#ifdef USING_MPI
  if (s == "num_external") return A->num_external;
  if (s == "num_send_neighbors") return A->num_send_neighbors;
  if (s == "total_to_be_sent") return A->total_to_be_sent;
#else
  if (s == "num_external" || s == "num_send_neighbors" || s == "total_to_be_sent") return 0;
#endif
  return -1;
}

// Copies the named array of `rank` into dst (if cap is large enough) and
// returns its element count.  Offsets are returned as long long.
long long ref_array(void *h, int rank, const char *name, void *dst, long long cap) {
  RefWorld *w = static_cast<RefWorld *>(h);
  const RefRank &rr = w->ranks[rank];
  const HPC_Sparse_Matrix *A = rr.A;
  const long long n = A->local_nrow;
  std::string s(name);
  if (s == "nnz_in_row") return copy_out(A->nnz_in_row, n, dst, cap);
  if (s == "list_of_inds") return copy_out(A->list_of_inds, rr.nnz_sum, dst, cap);
  if (s == "list_of_vals") return copy_out(A->list_of_vals, rr.nnz_sum, dst, cap);
  if (s == "x") return copy_out(rr.x, n, dst, cap);
  if (s == "b") return copy_out(rr.b, n, dst, cap);
  if (s == "xexact") return copy_out(rr.xexact, n, dst, cap);
  if (s == "ind_offsets" || s == "val_offsets" || s == "diag_offsets") {
    if (dst && cap >= n) {
      long long *o = static_cast<long long *>(dst);
      for (long long i = 0; i < n; ++i) {
        if (s == "ind_offsets") o[i] = A->ptr_to_inds_in_row[i] - A->list_of_inds;
        else if (s == "val_offsets") o[i] = A->ptr_to_vals_in_row[i] - A->list_of_vals;
        else o[i] = A->ptr_to_diags[i] - A->list_of_vals;
      }
    }
    return n;
  }
#ifdef USING_MPI
  if (s == "external_index") return copy_out(A->external_index, A->num_external, dst, cap);
  if (s == "external_local_index") return copy_out(A->external_local_index, A->num_external, dst, cap);
  if (s == "elements_to_send") return copy_out(A->elements_to_send, A->total_to_be_sent, dst, cap);
  if (s == "neighbors") return copy_out(A->neighbors, A->num_send_neighbors, dst, cap);
  if (s == "recv_length") return copy_out(A->recv_length, A->num_send_neighbors, dst, cap);
  if (s == "send_length") return copy_out(A->send_length, A->num_send_neighbors, dst, cap);
#else
  if (s == "external_index" || s == "external_local_index" || s == "elements_to_send" || s == "neighbors" ||
      s == "recv_length" || s == "send_length")
    return 0;
#endif
  return -1;
}

// ---- per-kernel entry points -------------------------------------------------

struct SpmvArgs {
  RefWorld *w;
  double **x;
  double **y;
  int exchange;
  int reps;
};

static void spmv_rank(int rank, void *arg) {
  SpmvArgs *a = static_cast<SpmvArgs *>(arg);
  HPC_Sparse_Matrix *A = a->w->ranks[rank].A;
  for (int r = 0; r < a->reps; ++r) {
#ifdef USING_MPI
    if (a->exchange) exchange_externals(A, a->x[rank]);
#endif
    HPC_sparsemv(A, a->x[rank], a->y[rank]);
  }
}

// x[rank] must hold local_ncol doubles (the halo tail is filled when
// exchange != 0 in the multi-rank variant), y[rank] local_nrow doubles.
int ref_spmv(void *h, double **x, double **y, int exchange, int reps) {
  RefWorld *w = static_cast<RefWorld *>(h);
  SpmvArgs a{w, x, y, exchange, reps < 1 ? 1 : reps};
  run_ranks(w->size, spmv_rank, &a);
  return 0;
}

struct DotArgs {
  RefWorld *w;
  double **x;
  double **y;
  double *result;
};

static void dot_rank(int rank, void *arg) {
  DotArgs *a = static_cast<DotArgs *>(arg);
  double t = 0.0;
  ddot(a->w->ranks[rank].A->local_nrow, a->x[rank], a->y[rank], &a->result[rank], t);
}

// result[rank] receives what ddot returned on that rank (the global sum in the
// multi-rank variant, ddot.cpp:77-82).
int ref_ddot(void *h, double **x, double **y, double *result) {
  RefWorld *w = static_cast<RefWorld *>(h);
  DotArgs a{w, x, y, result};
  run_ranks(w->size, dot_rank, &a);
  return 0;
}

// Raw kernels without a world (serial call; aliasing is the caller's business).
int ref_ddot_raw(int n, const double *x, const double *y, double *result) {
#ifdef USING_MPI
  (void)n; (void)x; (void)y; (void)result;
  return -1;  // ddot would call MPI_Allreduce outside a rank thread
#else
  double t = 0.0;
  return ddot(n, x, y, result, t);
#endif
}

int ref_waxpby(int n, double alpha, const double *x, double beta, const double *y, double *w) {
  return waxpby(n, alpha, x, beta, y, w);
}

// ---- full solve ----------------------------------------------------------------

struct SolveArgs {
  RefWorld *w;
  int max_iter;
  double tol;
  int hist;
  std::vector<int> niters;
  std::vector<double> normr;
  std::vector<std::vector<double>> times;
};

static void solve_rank(int rank, void *arg) {
  SolveArgs *a = static_cast<SolveArgs *>(arg);
  RefRank &rr = a->w->ranks[rank];
  for (int i = 0; i < rr.A->local_nrow; ++i) rr.x[i] = 0.0;  // generate_matrix.cpp:284
  int niters = 0;
  double normr = 0.0;
  double *times = a->times[rank].data();
  if (a->hist) HPCCG_hist(rr.A, rr.b, rr.x, a->max_iter, a->tol, niters, normr, times);
  else HPCCG(rr.A, rr.b, rr.x, a->max_iter, a->tol, niters, normr, times);
  a->niters[rank] = niters;
  a->normr[rank] = normr;
}

// hist_out (length max_iter, may be null): hist_out[0] = "Initial Residual",
// hist_out[k] = residual printed at iteration k (NaN where nothing was
// printed).  With hist != 0 every iteration is printed.  times_out: 7 doubles
// of rank 0.  x_out[rank] (may be null): solution copy.
int ref_solve(void *h, int max_iter, double tol, int hist, double *hist_out, int *niters, double *normr,
              double *times_out, double **x_out) {
  RefWorld *w = static_cast<RefWorld *>(h);
  SolveArgs a{w, max_iter, tol, hist, std::vector<int>(w->size, 0), std::vector<double>(w->size, 0.0),
              std::vector<std::vector<double>>(w->size, std::vector<double>(7, 0.0))};
  std::ostringstream captured;
  captured.precision(17);
  std::streambuf *old = std::cout.rdbuf(captured.rdbuf());
  std::streamsize oldprec = std::cout.precision(17);
  run_ranks(w->size, solve_rank, &a);
  std::cout.rdbuf(old);
  std::cout.precision(oldprec);

  if (hist_out) {
    for (int k = 0; k < max_iter; ++k) hist_out[k] = std::nan("");
    std::istringstream in(captured.str());
    std::string line;
    while (std::getline(in, line)) {
      if (line.rfind("Initial Residual = ", 0) == 0) {
        hist_out[0] = std::stod(line.substr(19));
      } else if (line.rfind("Iteration = ", 0) == 0) {
        int k = 0;
        double v = 0;
        char buf[64];
        // "Iteration = 15   Residual = 2.15402e-06"
        if (std::sscanf(line.c_str(), "Iteration = %d   Residual = %63s", &k, buf) == 2) {
          v = std::strtod(buf, nullptr);
          if (k >= 0 && k < max_iter) hist_out[k] = v;
        }
      }
    }
  }
  if (niters) *niters = a.niters[0];
  if (normr) *normr = a.normr[0];
  if (times_out) std::memcpy(times_out, a.times[0].data(), 7 * sizeof(double));
  if (x_out)
    for (int r = 0; r < w->size; ++r)
      if (x_out[r]) std::memcpy(x_out[r], w->ranks[r].x, sizeof(double) * w->ranks[r].A->local_nrow);
  return 0;
}

struct ResArgs {
  RefWorld *w;
  double **x;
  double *res;
};

static void res_rank(int rank, void *arg) {
  ResArgs *a = static_cast<ResArgs *>(arg);
  RefRank &rr = a->w->ranks[rank];
  compute_residual(rr.A->local_nrow, a->x[rank], rr.xexact, &a->res[rank]);
}

// max_i |x_i - xexact_i| through the reference's compute_residual.cpp:59-81.
int ref_compute_residual(void *h, double **x, double *res_per_rank) {
  RefWorld *w = static_cast<RefWorld *>(h);
  ResArgs a{w, x, res_per_rank};
  run_ranks(w->size, res_rank, &a);
  return 0;
}

// ---- YAML golden text ------------------------------------------------------------
// Builds the report tree the way main.cpp:230-298 does, through the reference's
// own YAML_Doc / YAML_Element classes, and returns the text.  `ranks` > 0 adds
// the MPI-only blocks.  generateYAML also writes ./hpccg-1.0_<timestamp>.yaml
// (YAML_Doc.cpp:49-70); callers run this from a scratch directory.
int ref_yaml_report(int nx, int ny, int nz, int niters, double normr, const double *times, double total_nrow,
                    double total_nnz, int ranks, int omp_threads, const double *t4stats, char *out, int cap) {
  double fniters = niters, fnrow = total_nrow, fnnz = total_nnz;
  double fnops_ddot = fniters * 4 * fnrow;
  double fnops_waxpby = fniters * 6 * fnrow;
  double fnops_sparsemv = fniters * 2 * fnnz;
  double fnops = fnops_ddot + fnops_waxpby + fnops_sparsemv;
  YAML_Doc doc("hpccg", "1.0");
  doc.add("Parallelism", "");
  if (ranks > 0) doc.get("Parallelism")->add("Number of MPI ranks", ranks);
  else doc.get("Parallelism")->add("MPI not enabled", "");
  if (omp_threads > 0) doc.get("Parallelism")->add("Number of OpenMP threads", omp_threads);
  else doc.get("Parallelism")->add("OpenMP not enabled", "");
  doc.get("Parallelism")->add("SYCL not enabled", "");
  doc.add("Dimensions", "");
  doc.get("Dimensions")->add("nx", nx);
  doc.get("Dimensions")->add("ny", ny);
  doc.get("Dimensions")->add("nz", nz);
  doc.add("Number of iterations", niters);
  doc.add("Final residual", normr);
  doc.add("#********** Performance Summary (times in sec) ***********", "");
  doc.add("Time Summary", "");
  doc.get("Time Summary")->add("Total   ", times[0]);
  doc.get("Time Summary")->add("DDOT    ", times[1]);
  doc.get("Time Summary")->add("WAXPBY  ", times[2]);
  doc.get("Time Summary")->add("SPARSEMV", times[3]);
  doc.add("FLOPS Summary", "");
  doc.get("FLOPS Summary")->add("Total   ", fnops);
  doc.get("FLOPS Summary")->add("DDOT    ", fnops_ddot);
  doc.get("FLOPS Summary")->add("WAXPBY  ", fnops_waxpby);
  doc.get("FLOPS Summary")->add("SPARSEMV", fnops_sparsemv);
  doc.add("MFLOPS Summary", "");
  doc.get("MFLOPS Summary")->add("Total   ", fnops / times[0] / 1.0E6);
  doc.get("MFLOPS Summary")->add("DDOT    ", fnops_ddot / times[1] / 1.0E6);
  doc.get("MFLOPS Summary")->add("WAXPBY  ", fnops_waxpby / times[2] / 1.0E6);
  doc.get("MFLOPS Summary")->add("SPARSEMV", fnops_sparsemv / (times[3]) / 1.0E6);
  if (ranks > 0) {
    doc.add("DDOT Timing Variations", "");
    doc.get("DDOT Timing Variations")->add("Min DDOT MPI_Allreduce time", t4stats[0]);
    doc.get("DDOT Timing Variations")->add("Max DDOT MPI_Allreduce time", t4stats[1]);
    doc.get("DDOT Timing Variations")->add("Avg DDOT MPI_Allreduce time", t4stats[2]);
    double totalSparseMVTime = times[3] + times[5] + times[6];
    doc.add("SPARSEMV OVERHEADS", "");
    YAML_Element *o = doc.get("SPARSEMV OVERHEADS");
    o->add("SPARSEMV MFLOPS W OVERHEAD", fnops_sparsemv / (totalSparseMVTime) / 1.0E6);
    o->add("SPARSEMV PARALLEL OVERHEAD Time", (times[5] + times[6]));
    o->add("SPARSEMV PARALLEL OVERHEAD Pct", (times[5] + times[6]) / totalSparseMVTime * 100.0);
    o->add("SPARSEMV PARALLEL OVERHEAD Setup Time", (times[6]));
    o->add("SPARSEMV PARALLEL OVERHEAD Setup Pct", (times[6]) / totalSparseMVTime * 100.0);
    o->add("SPARSEMV PARALLEL OVERHEAD Bdry Exch Time", (times[5]));
    o->add("SPARSEMV PARALLEL OVERHEAD Bdry Exch Pct", (times[5]) / totalSparseMVTime * 100.0);
  }
  std::string yaml = doc.generateYAML();
  if ((int)yaml.size() + 1 > cap) return -(int)yaml.size() - 1;
  std::memcpy(out, yaml.c_str(), yaml.size() + 1);
  return (int)yaml.size();
}

}  // extern "C"
