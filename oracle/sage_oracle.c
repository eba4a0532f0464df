/*
 * sage_oracle.c -- CPU restatement of SAGE-SLAM's factor kernels.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing outside tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may import, link or execute this file.
 * The product path (sage-slam_b200/) never routes through it and fails loudly when its
 * CUDA library is missing.
 *
 * What it is: a plain-C, OpenMP-parallel restatement of the reference's hot path
 * (/root/reference/system/sources/cuda/{photometric,geometric,reprojection}_factor_kernels.cpp),
 * one function per df::*_calculate entry point, each citing the reference lines it follows
 * (see oracle_body.inc).  The reference has NO CPU implementation of this path and ships NO
 * tests or golden vectors, so this restatement is pinned
 * differentially: tests/golden/ holds outputs of the reference's own CUDA kernels
 * (compiled unmodified from /root/reference by oracle/build_ref.py into oracle/_ref/ and run
 * on a B200 by oracle/make_golden.py); tests/test_oracle_golden.py checks this file
 * against them.
 *
 * Two instantiations: *_f32 computes each Jacobian row in float exactly like the kernels;
 * *_f64 computes them in double (used for finite-difference and conditioning checks).
 * Both accumulate J^T J / J^T r / error in double.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define REAL float
#define SUF _f32
#include "oracle_body.inc"
#undef REAL
#undef SUF

#define REAL double
#define SUF _f64
#include "oracle_body.inc"
#undef REAL
#undef SUF

#ifdef _OPENMP
#include <omp.h>
int oracle_num_threads(void) { return omp_get_max_threads(); }
void oracle_set_num_threads(int n) { omp_set_num_threads(n); }
#else
int oracle_num_threads(void) { return 1; }
void oracle_set_num_threads(int n) { (void)n; }
#endif
