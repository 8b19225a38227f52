// test_HPCCG.cpp -- command-line driver of the B200-native HPCCG hot path, equivalent to the reference's
// main.cpp:99-324: `test_HPCCG nx ny nz` generates the nx*ny*nz problem, runs HPCCG() and prints the YAML report
// (also written to ./hpccg-1.0_<timestamp>.yaml, YAML_Doc.cpp:49-70).  It is written against the reference-named
// C++ API only (the headers in csrc/include have the reference's file names), i.e. it is what the reference's own
// main.cpp looks like after switching its include path and linking libhpccg_b200.so.
//
// Differences from the reference's main, all additive:
//   --iters N      max_iter (reference hard-codes 500, main.cpp:187; upstream and BASELINE.json use 150)
//   --stencil 7    the reference's compile-time `use_7pt_stencil` (generate_matrix.cpp:219)
//   --tolerance T  reference hard-codes 0.0 (main.cpp:188)
//   --device-only  build the matrix directly in HBM (needed beyond 430^3, where the reference's int local_nnz overflows)
//   --check        compute_residual(x, xexact) after the solve (the call the reference has commented out, main.cpp:310-316)
//   extra YAML block "B200" with GFLOP/s, HBM GB/s of the CG loop and its fraction of the roofline.
// Mode 2 (matrix file, read_HPC_row) is deprecated upstream (README.md:114-118) and not supported.
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

#include "HPCCG.hpp"
#include "HPC_Sparse_Matrix.hpp"
#include "YAML_Doc.hpp"
#include "compute_residual.hpp"
#include "generate_matrix.hpp"
#include "hpccg_b200.h"
#include "mytimer.hpp"

using std::cerr;
using std::cout;
using std::endl;

int main(int argc, char *argv[]) {
  HPC_Sparse_Matrix *A;
  double *x, *b, *xexact;
  double times[7] = {0, 0, 0, 0, 0, 0, 0};
  int dims[3] = {0, 0, 0}, ndims = 0;
  int max_iter = 150, stencil = 27, device_only = 0, check = 0;
  double tolerance = 0.0;
  double peak_gbs = 6543.7;  // measured copy bandwidth of this pool's B200s (MEASURED_PEAKS.json); --peak overrides
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--iters" && i + 1 < argc) max_iter = std::atoi(argv[++i]);
    else if (a == "--stencil" && i + 1 < argc) stencil = std::atoi(argv[++i]);
    else if (a == "--tolerance" && i + 1 < argc) tolerance = std::atof(argv[++i]);
    else if (a == "--peak" && i + 1 < argc) peak_gbs = std::atof(argv[++i]);
    else if (a == "--device-only") device_only = 1;
    else if (a == "--check") check = 1;
    else if (ndims < 3 && a[0] != '-') dims[ndims++] = std::atoi(argv[i]);
    else ndims = -1000;
  }
  if (ndims != 3) {
    cerr << "Usage:" << endl
         << "Mode 1: " << argv[0] << " nx ny nz [--iters N] [--stencil 27|7] [--tolerance T] [--device-only] [--check]" << endl
         << "     where nx, ny and nz are the local sub-block dimensions." << endl
         << "Mode 2 (HPC_data_file) of the reference is deprecated upstream and not supported." << endl;
    return 1;
  }
  const int nx = dims[0], ny = dims[1], nz = dims[2];
  const long long n = (long long)nx * ny * nz;
  if (27 * n > 2147483647LL) device_only = 1;
  if (hpccg_api_set_options(stencil, device_only ? 0 : 1)) {
    cerr << hpccg_last_error() << endl;
    return 1;
  }

  generate_matrix(nx, ny, nz, &A, &x, &b, &xexact);

  int niters = 0;
  double normr = 0.0;
  const auto start = std::chrono::high_resolution_clock::now();
  int ierr = HPCCG(A, b, x, max_iter, tolerance, niters, normr, times);
  const auto end = std::chrono::high_resolution_clock::now();
  const std::chrono::duration<double> elapsed = end - start;
  cout << "Elapsed time: " << elapsed.count() << " s\n";
  if (ierr) cerr << "Error in call to CG: " << ierr << ": " << hpccg_last_error() << ".\n" << endl;
  // second solve: the first one paid for the one-off mirror build and workspace allocation inside times[0]
  double times2[7] = {0, 0, 0, 0, 0, 0, 0};
  if (!ierr) {
    for (long long i = 0; i < n; ++i) x[i] = 0.0;
    ierr = HPCCG(A, b, x, max_iter, tolerance, niters, normr, times2);
  }

  const double fniters = niters, fnrow = A->total_nrow, fnnz = (double)A->total_nnz;
  const double fnops_ddot = fniters * 4 * fnrow, fnops_waxpby = fniters * 6 * fnrow, fnops_sparsemv = fniters * 2 * fnnz;
  const double fnops = fnops_ddot + fnops_waxpby + fnops_sparsemv;
  const double *t = ierr ? times : times2;

  YAML_Doc doc("hpccg", "1.0");
  doc.add("Parallelism", "");
  doc.get("Parallelism")->add("MPI not enabled", "");
  doc.get("Parallelism")->add("OpenMP not enabled", "");
  doc.get("Parallelism")->add("SYCL not enabled", "");
  doc.get("Parallelism")->add("Number of B200 GPUs", 1);
  doc.add("Dimensions", "");
  doc.get("Dimensions")->add("nx", nx);
  doc.get("Dimensions")->add("ny", ny);
  doc.get("Dimensions")->add("nz", nz);
  doc.add("Number of iterations", niters);
  doc.add("Final residual", normr);
  doc.add("#********** Performance Summary (times in sec) ***********", "");
  doc.add("Time Summary", "");
  doc.get("Time Summary")->add("Total   ", t[0]);
  doc.get("Time Summary")->add("DDOT    ", t[1]);
  doc.get("Time Summary")->add("WAXPBY  ", t[2]);
  doc.get("Time Summary")->add("SPARSEMV", t[3]);
  doc.add("FLOPS Summary", "");
  doc.get("FLOPS Summary")->add("Total   ", fnops);
  doc.get("FLOPS Summary")->add("DDOT    ", fnops_ddot);
  doc.get("FLOPS Summary")->add("WAXPBY  ", fnops_waxpby);
  doc.get("FLOPS Summary")->add("SPARSEMV", fnops_sparsemv);
  doc.add("MFLOPS Summary", "");
  doc.get("MFLOPS Summary")->add("Total   ", fnops / t[0] / 1.0E6);
  doc.get("MFLOPS Summary")->add("DDOT    ", fnops_ddot / t[1] / 1.0E6);
  doc.get("MFLOPS Summary")->add("WAXPBY  ", fnops_waxpby / t[2] / 1.0E6);
  doc.get("MFLOPS Summary")->add("SPARSEMV", fnops_sparsemv / (t[3]) / 1.0E6);
  // B200 block: the kernel times above are CUDA-event sums, fused kernels split by algorithmic bytes
  const double kernel_s = t[1] + t[2] + t[3];
  const double bytes_per_row = (stencil == 7 ? 7 : 27) * 12.0 + 16.0 + 72.0;
  doc.add("B200", "");
  doc.get("B200")->add("Stencil points", stencil);
  doc.get("B200")->add("First call Total (includes device mirror build)", times[0]);
  doc.get("B200")->add("Host-to-device and back per call (bytes)", 24.0 * (double)n);
  doc.get("B200")->add("CG kernels time", kernel_s);
  doc.get("B200")->add("CG kernels GFLOPS", fnops / kernel_s / 1.0E9);
  doc.get("B200")->add("Algorithmic bytes per row per iteration", bytes_per_row);
  doc.get("B200")->add("CG kernels HBM GB/s", fniters * bytes_per_row * (double)n / kernel_s / 1.0E9);
  doc.get("B200")->add("Fraction of HBM roofline", fniters * bytes_per_row * (double)n / kernel_s / 1.0E9 / peak_gbs);
  doc.get("B200")->add("HBM roofline GB/s", peak_gbs);
  if (check) {
    double residual = 0;
    if (compute_residual(A->local_nrow, x, xexact, &residual)) cerr << "Error in call to compute_residual: " << hpccg_last_error() << endl;
    doc.get("B200")->add("Difference between computed and exact", residual);
  }
  const std::string yaml = doc.generateYAML();
  cout << yaml;

  destroyMatrix(A);
  free_vectors(x, b, xexact);
  return ierr ? 2 : 0;
}
