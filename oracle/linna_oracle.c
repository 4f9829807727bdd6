/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the LINNA emulator-likelihood hot path.
 *
 * A plain-C restatement of the reference algorithm (linna/nn.py, linna/util.py,
 * linna/predictor_gpu.py -- per-function citations in linna_oracle_impl.h), in float32
 * (what the reference computes in) and float64 (error budgeting).  Pinned against golden
 * vectors produced by running the reference itself (tests/golden/make_golden.py); see
 * tests/test_oracle.py.  The reference is pure Python, so there is no oracle/_ref binary:
 * the "real reference" strengthening is the committed goldens.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product (linna_b200/) never does.
 *
 * Build:  make -C oracle      (gcc -O2 -shared -fPIC -> oracle/_build/liblinna_oracle.so)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define REAL float
#define SUF f32
#include "linna_oracle_impl.h"
#undef REAL
#undef SUF

#define REAL double
#define SUF f64
#include "linna_oracle_impl.h"
#undef REAL
#undef SUF

int linna_oracle_abi_version(void) { return 1; }
