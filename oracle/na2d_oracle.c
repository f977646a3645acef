/*
 * TEST INFRASTRUCTURE — CPU oracle for LM-Net's neighbourhood-attention path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (lm-net_b200/) never does.
 *
 * Restates, in plain C + OpenMP, what natten's CPU ops compute for the call at
 * /root/reference/core/modules.py:517 (see na2d_oracle_body.inc for the
 * algorithm statement and citations).  Built twice: *_f32 and *_f64.
 *
 * PARITY UNPINNED against natten (absent offline, un-pinned in the reference);
 * pinned by known-answer properties and against PyTorch FlexAttention with the
 * published NATTEN mask — see tests/test_oracle.py and DESIGN.md §3.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define REAL float
#define AX(name) CAT(name, _f32)
static inline float exp__f32(float x) { return expf(x); }
static inline float log__f32(float x) { return logf(x); }
#include "na2d_oracle_body.inc"
#undef REAL
#undef AX

#define REAL double
#define AX(name) CAT(name, _f64)
static inline double exp__f64(double x) { return exp(x); }
static inline double log__f64(double x) { return log(x); }
#include "na2d_oracle_body.inc"
#undef REAL
#undef AX

int na2d_oracle_abi_version(void) { return 1; }
