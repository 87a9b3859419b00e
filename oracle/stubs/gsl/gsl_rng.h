/* Minimal stand-in for the GNU Scientific Library headers that five Stanford
 * lens/eye cameras of the reference include (src/cameras/pinhole.h:11,
 * idealDiffraction.h:11, perspectiveDiffraction.h:35, realisticEye.cpp:26-27).
 * None of them is on the path-tracing hot path; these stubs only let the
 * reference link without GSL (absent from this image). Calling one aborts.
 * Test infrastructure for building oracle/_ref; not part of the product. */
#ifndef SPT_ORACLE_GSL_STUB_H
#define SPT_ORACLE_GSL_STUB_H
#include <stdio.h>
#include <stdlib.h>
typedef struct { int unused; } gsl_rng_type;
typedef struct { int unused; } gsl_rng;
static const gsl_rng_type *gsl_rng_default = 0;
static inline void gsl_stub_die(const char *what) {
    fprintf(stderr, "GSL stub: %s called; lens cameras are not supported in the oracle build\n", what);
    abort();
}
static inline const gsl_rng_type *gsl_rng_env_setup(void) { return 0; }
static inline gsl_rng *gsl_rng_alloc(const gsl_rng_type *t) { (void)t; return 0; }
static inline void gsl_rng_free(gsl_rng *r) { (void)r; }
static inline void gsl_ran_bivariate_gaussian(const gsl_rng *r, double sx, double sy, double rho,
                                              double *x, double *y) {
    (void)r; (void)sx; (void)sy; (void)rho; (void)x; (void)y;
    gsl_stub_die("gsl_ran_bivariate_gaussian");
}
#define GSL_SUCCESS 0
#define GSL_CONTINUE (-2)
typedef struct { double (*function)(double x, void *params); void *params; } gsl_function;
typedef struct { int unused; } gsl_root_fsolver_type;
typedef struct { int unused; } gsl_root_fsolver;
static const gsl_root_fsolver_type *gsl_root_fsolver_brent = 0;
typedef void gsl_error_handler_t(const char *, const char *, int, int);
static inline gsl_error_handler_t *gsl_set_error_handler_off(void) { return 0; }
static inline gsl_root_fsolver *gsl_root_fsolver_alloc(const gsl_root_fsolver_type *t) {
    (void)t; gsl_stub_die("gsl_root_fsolver_alloc"); return 0; }
static inline int gsl_root_fsolver_set(gsl_root_fsolver *s, gsl_function *f, double lo, double hi) {
    (void)s; (void)f; (void)lo; (void)hi; return -1; }
static inline void gsl_root_fsolver_free(gsl_root_fsolver *s) { (void)s; }
static inline const char *gsl_root_fsolver_name(const gsl_root_fsolver *s) { (void)s; return "stub"; }
static inline int gsl_root_fsolver_iterate(gsl_root_fsolver *s) { (void)s; return -1; }
static inline double gsl_root_fsolver_root(const gsl_root_fsolver *s) { (void)s; return 0; }
static inline double gsl_root_fsolver_x_lower(const gsl_root_fsolver *s) { (void)s; return 0; }
static inline double gsl_root_fsolver_x_upper(const gsl_root_fsolver *s) { (void)s; return 0; }
static inline int gsl_root_test_interval(double lo, double hi, double ea, double er) {
    (void)lo; (void)hi; (void)ea; (void)er; return GSL_SUCCESS; }
static inline const char *gsl_strerror(int e) { (void)e; return "gsl stub"; }
#endif
