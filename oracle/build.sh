#!/usr/bin/env bash
# oracle/build.sh -- TEST INFRASTRUCTURE, not product code.
#
# Builds (1) the C restatement of the reference algorithm (hpccg_oracle.c ->
# _ref/libhpccg_oracle.so) and (2), when the reference tree is present, the
# REAL reference compiled from its own sources where they lie under $REF:
#
#   _ref/libhpccg_ref_serial.so   g++ -O3 -DWALL                      (parity oracle)
#   _ref/libhpccg_ref_omp.so      MakefileOMP:83,103,117,133 flags    (CPU baseline)
#   _ref/libhpccg_ref_mpi.so      -DUSING_MPI against mpi_shim/       (multi-rank oracle)
#
# No reference source is copied: each translation unit is preprocessed from
# $REF, piped through sed for the three compile-time switches the reference
# hard-codes, and compiled from stdin.  The switches:
#   * generate_matrix.cpp:219  `bool use_7pt_stencil = false;`  -> second object
#     with the symbol renamed generate_matrix_7pt and the bool set true;
#   * HPCCG.cpp:342-344        print_freq                        -> second object
#     HPCCG_hist printing every iteration (residual history for parity);
#   * HPC_Sparse_Matrix.hpp:49 `max_external = 100000`           -> raised in the
#     multi-rank build only, so xy-planes above 100 000 points do not abort().
# Outputs go only into _ref/ (git-ignored, but shipped to the GPU box).
set -euo pipefail
cd "$(dirname "$0")"
REF="${REF:-/root/reference}"
OUT=_ref
# The image exports CXX=/opt/gcc/bin/g++, a toolchain without libgomp; use the system compiler.
CXX="${HPCCG_ORACLE_CXX:-/usr/bin/g++}"
CC="${HPCCG_ORACLE_CC:-/usr/bin/gcc}"
mkdir -p "$OUT"

# ---- (1) C restatement ---------------------------------------------------------
if [ ! -f "$OUT/libhpccg_oracle.so" ] || [ hpccg_oracle.c -nt "$OUT/libhpccg_oracle.so" ]; then
  "$CC" -O2 -fPIC -shared -ffp-contract=off -o "$OUT/libhpccg_oracle.so" hpccg_oracle.c -lm
  echo "built $OUT/libhpccg_oracle.so"
fi

# ---- (2) the real reference ------------------------------------------------------
if [ ! -f "$REF/HPCCG.cpp" ]; then
  echo "reference tree $REF not present: keeping prebuilt $OUT/libhpccg_ref_*.so"
  exit 0
fi

COMMON_TUS="generate_matrix mytimer HPC_sparsemv HPCCG waxpby ddot compute_residual HPC_Sparse_Matrix YAML_Doc YAML_Element read_HPC_row"
MPI_TUS="make_local_matrix exchange_externals"
SED_NONE='s/^$//'
SED_MAXEXT='s/max_external = 100000/max_external = 2200000/'
SED_7PT='s/use_7pt_stencil = false/use_7pt_stencil = true/'
SED_HIST='s/print_freq>50/print_freq>0/;s/print_freq=50/print_freq=1/'

pipe_compile() {  # dir flags tu sed extra-defs suffix
  local dir="$1" flags="$2" tu="$3" script="$4" defs="$5" suffix="$6"
  # shellcheck disable=SC2086
  "$CXX" -E $flags $defs -I"$REF" "$REF/$tu.cpp" | sed -e "$script" |
    "$CXX" -x c++-cpp-output $flags -c -o "$OUT/$dir/$tu$suffix.o" -
}

build_variant() {  # name flags tus sed
  local name="$1" flags="$2" tus="$3" script="$4"
  local lib="$OUT/libhpccg_ref_$name.so"
  if [ -f "$lib" ] && [ "$lib" -nt ref_driver.cpp ] && [ "$lib" -nt mpi_shim/mpi_shim.cpp ] && [ "$lib" -nt build.sh ]; then
    return
  fi
  rm -rf "$OUT/$name"; mkdir -p "$OUT/$name"
  for tu in $tus; do pipe_compile "$name" "$flags" "$tu" "$script" "" ""; done
  pipe_compile "$name" "$flags" generate_matrix "$script;$SED_7PT" "-Dgenerate_matrix=generate_matrix_7pt" "_7pt"
  pipe_compile "$name" "$flags" HPCCG "$script;$SED_HIST" "-DHPCCG=HPCCG_hist" "_hist"
  # shellcheck disable=SC2086
  "$CXX" $flags -I"$REF" -c -o "$OUT/$name/ref_driver.o" ref_driver.cpp
  if [ "$name" = mpi ]; then
    # shellcheck disable=SC2086
    "$CXX" $flags -c -o "$OUT/$name/mpi_shim.o" mpi_shim/mpi_shim.cpp
  fi
  # shellcheck disable=SC2086
  "$CXX" -shared -Wl,-Bsymbolic $flags -o "$lib" "$OUT/$name"/*.o -lpthread -lm
  rm -rf "$OUT/$name"
  echo "built $lib"
}

BASE="-O3 -DWALL -fPIC -w"
build_variant serial "$BASE -DREF_VARIANT=0" "$COMMON_TUS" "$SED_NONE"
build_variant omp "-O3 -funroll-all-loops -malign-double -fopenmp -DUSING_OMP -DWALL -fPIC -w -DREF_VARIANT=1" "$COMMON_TUS" "$SED_NONE"
build_variant mpi "$BASE -DUSING_MPI -Impi_shim -DREF_VARIANT=2" "$COMMON_TUS $MPI_TUS" "$SED_MAXEXT"

# ---- (3) the reference's OWN main.cpp on the B200 library (drop-in proof, tests/test_gpu_cli.py) ----------------
# /root/reference/main.cpp, unmodified, compiled against hpccg-sycl_b200/csrc/include (the reference's header names) and
# linked with libhpccg_b200.so instead of the reference's kernels.  Nothing of the reference but main() is in this binary.
B200_LIB=../hpccg-sycl_b200/lib
if [ -f "$B200_LIB/libhpccg_b200.so" ]; then
  if [ ! -f "$OUT/test_HPCCG_refmain" ] || [ "$B200_LIB/libhpccg_b200.so" -nt "$OUT/test_HPCCG_refmain" ]; then
    # read from stdin: a quoted #include then looks in the -I directories (this repo's headers under the reference's
    # names), not next to main.cpp where the reference's own headers lie
    "$CXX" -O2 -DWALL -w -I../hpccg-sycl_b200/csrc/include -I../include -x c++ - -o "$OUT/test_HPCCG_refmain" \
      -L"$B200_LIB" -lhpccg_b200 -Wl,-rpath,'$ORIGIN/../../hpccg-sycl_b200/lib' < "$REF/main.cpp"
    echo "built $OUT/test_HPCCG_refmain (reference main.cpp + libhpccg_b200.so)"
  fi
fi
