// k2_npendulum.cu -- the banded CTA-per-system kernel K2 (k2_band.cuh) for the recursive n-pendulum chain
// (scipyDAE/tests/recursive_pendulum.py:93-127): Jacobian bandwidths kl = 9, ku = 7, analytic or finite-difference Jacobian.
#define K2_KL 9
#define K2_KU 7
#define K2_CHAIN 1
#define K2_ENTRY k2_npendulum
#define K2_SCRATCH_FN k2_scratch_bytes
#include "k2_band.cuh"
