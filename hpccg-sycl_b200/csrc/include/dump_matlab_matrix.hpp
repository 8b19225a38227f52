// dump_matlab_matrix.hpp -- same include name as the reference (dump_matlab_matrix.hpp:42-46), so that the reference's
// main.cpp compiles against this include directory unchanged.  Debug aid, not on the hot path.
#pragma once
#include "HPC_Sparse_Matrix.hpp"

// Writes the local rows as 1-based "row col value" triplets to mat<rank>.dat for ranks 0..3 (other ranks: no file),
// like dump_matlab_matrix.cpp:58-82.  Needs the host row arrays (not available for device-only matrices: returns 1).
int dump_matlab_matrix(HPC_Sparse_Matrix *A, int rank);
