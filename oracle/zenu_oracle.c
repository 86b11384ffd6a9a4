/* zenu_oracle.c — CPU ORACLE for the ZeNu CNN-training hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it, and only as the checker / the
 * CPU baseline.  Nothing under zenu_b200/ links, imports or falls back to it.
 *
 * It is a plain-C restatement of the reference's CPU path (Rust, zenu-matrix @ 3ab3e958):
 * im2col + BLAS conv (nn/conv/cpu/{...}.rs), CPU BatchNorm (nn/batch_norm.rs:283-420), GEMM
 * (operation/mul.rs), ReLU, elementwise, pooling, softmax/cross-entropy and the optimizers
 * (zenu-optimizer/src/{sgd,adam,adamw}.rs).  Every function cites the file:line it follows.
 *
 * The reference itself cannot be compiled here (no cargo/rustc, no network for crates), so there is no
 * oracle/_ref.  Parity is PINNED instead against every golden vector the reference's own tests hold
 * for this path (tests/golden/, extracted by tests/golden/make_golden.py; checked by tests/test_oracle_golden.py).
 * Not pinned by any reference test (strides/dilation != 1, N > 1 convs, f64, ResNet-scale shapes):
 * there the restatement, cross-checked against torch-CPU float64, is the authority ("parity unpinned"
 * for those cases — see DESIGN.md).
 *
 * Third-party arithmetic: the reference's GEMM is OpenBLAS (crates cblas 0.4.0 + openblas-src 0.10.8,
 * zenu-matrix/Cargo.toml:12-13; versions otherwise unpinned, Cargo.lock is git-ignored).  zo_set_blas()
 * dlopens an OpenBLAS build (numpy's bundled ILP64 libscipy_openblas64_) for the same library family;
 * without it a plain OpenMP loop nest is used.
 */
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef void (*sgemm_fn)(int, int, int, int64_t, int64_t, int64_t, float, const float *, int64_t, const float *, int64_t,
                         float, float *, int64_t);
typedef void (*dgemm_fn)(int, int, int, int64_t, int64_t, int64_t, double, const double *, int64_t, const double *,
                         int64_t, double, double *, int64_t);
static sgemm_fn g_sgemm = NULL;
static dgemm_fn g_dgemm = NULL;
static void (*g_set_threads)(int) = NULL;

/* Returns 0 when an ILP64 OpenBLAS with scipy_-prefixed cblas symbols was loaded from `path`. */
int zo_set_blas(const char *path) {
  void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h) return 1;
  sgemm_fn s = (sgemm_fn)dlsym(h, "scipy_cblas_sgemm64_");
  dgemm_fn d = (dgemm_fn)dlsym(h, "scipy_cblas_dgemm64_");
  if (!s || !d) return 2;
  g_sgemm = s;
  g_dgemm = d;
  g_set_threads = (void (*)(int))dlsym(h, "scipy_openblas_set_num_threads64_");
  return 0;
}
void zo_unset_blas(void) { g_sgemm = NULL; g_dgemm = NULL; }
int zo_has_blas(void) { return g_sgemm != NULL; }
void zo_set_blas_threads(int n) { if (g_set_threads) g_set_threads(n); }

#define T float
#define FN(name) name##_f32
#define BLAS_GEMM g_sgemm
#define SQRT sqrtf
#define EXP expf
#define LOG logf
#define POW powf
#include "zenu_oracle_impl.inc"
#undef T
#undef FN
#undef BLAS_GEMM
#undef SQRT
#undef EXP
#undef LOG
#undef POW

#define T double
#define FN(name) name##_f64
#define BLAS_GEMM g_dgemm
#define SQRT sqrt
#define EXP exp
#define LOG log
#define POW pow
#include "zenu_oracle_impl.inc"
