/* gko_oracle.c — TEST INFRASTRUCTURE: CPU parity oracle for libgko_b200.so.
 *
 * A plain-C restatement of the Ginkgo 1.5.0 reference-executor kernels on the
 * sparse-solve hot path.  Pinned against (i) the known-answer
 * tests of the reference's own test-suite, restated in tests/test_oracle_kat.py,
 * and (ii) the unmodified reference compiled from /root/reference into
 * oracle/_ref (oracle/Makefile.ref), compared in tests/test_oracle_vs_ref.py.
 *
 * NOT product code: only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline may load liboracle.so.  The product path never calls it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define V double
#define SFX(name) name##_f64
#define SQRT sqrt
#define FABS fabs
#include "oracle_kernels.inc"
#undef V
#undef SFX
#undef SQRT
#undef FABS

#define V float
#define SFX(name) name##_f32
#define SQRT sqrtf
#define FABS fabsf
#include "oracle_kernels.inc"
#undef V
#undef SFX
#undef SQRT
#undef FABS

int oracle_version(void) { return 100; }
