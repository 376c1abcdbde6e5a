#define _GNU_SOURCE   /* sincosf */
/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / CPU baseline -- never as the
 * product path (the product is yask_b200/csrc, CUDA only).
 *
 * Parity status: PINNED.  Every function here is checked bit-for-bit against outputs of
 * the unmodified reference (built out-of-tree by oracle/build_ref.sh and driven through
 * its public API by oracle/ref_driver.cpp); the resulting vectors are committed under
 * tests/golden/ together with tests/golden/make_golden.py.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off is REQUIRED: every FMA below is written explicitly (fmaf/fma) so the
 * two variants of the reference's arithmetic can be reproduced exactly:
 *   contract=0 : reference built with -ffp-contract=off (pure IEEE mul/add, DSL order);
 *                bit-exact vs oracle/_ref "iso3dfd-strict" (tests/golden).
 *   contract=1 : canonical FMA form: every "acc + sum*c" and "lhs + acc*v" fused.
 *   contract=2 : what GCC 13.3 -O3 (-ffp-contract=fast) actually emits for the reference's
 *                default build: as contract=1, except that in the FIRST group the *other*
 *                product is the one fused: acc1 = fma(p, c0, sum1*c1) instead of
 *                fma(sum1, c1, p*c0).  Bit-exact vs oracle/_ref "iso3dfd" (tests/golden).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------ */
/* Finite-difference coefficients.                                                        */
/* Follows /root/reference/src/contrib/coefficients/fd_coeff.cpp:53-101 (Fornberg's        */
/* recurrence) with the SAME floating-point operation order (multiply by the reciprocal    */
/* 1.0/c3, scale by c1/c2, -0.0 squashed to 0.0), as called for uniform centred points by   */
/* /root/reference/src/common/fd_coeff2.cpp:49-57 (get_center_fd_coefficients).             */
/* ------------------------------------------------------------------------------------ */

/* delta[m][n][v], m = derivative 0..order, n = number of points used - 1, v = point. */
#define DLT(m, n, v) delta[((size_t)(m) * np + (size_t)(n)) * np + (size_t)(v)]

int yo_fd_coefficients(double* coeff, double eval_point, int order, const double* pts, int np)
{
    if (np < 2 || order < 0) return -1;
    double* delta = (double*)calloc((size_t)(order + 1) * np * np, sizeof(double));
    if (!delta) return -2;
    DLT(0, 0, 0) = 1.0;
    double c1 = 1.0;
    for (int n = 1; n < np; n++) {
        double c2 = 1.0;
        int mmax = n < order ? n : order;
        for (int v = 0; v < n; v++) {
            double c3 = pts[n] - pts[v];
            c2 = c2 * c3;
            for (int m = 0; m <= mmax; m++) {
                double t = (pts[n] - eval_point) * DLT(m, n - 1, v);
                if (m > 0) t -= m * DLT(m - 1, n - 1, v);
                t *= 1.0 / c3;
                if (t == 0.0) t = 0.0; /* squashes -0.0 */
                DLT(m, n, v) = t;
            }
        }
        for (int m = 0; m <= mmax; m++) {
            double t = 0.0;
            if (m > 0) t += m * DLT(m - 1, n - 1, n - 1);
            t -= (pts[n - 1] - eval_point) * DLT(m, n - 1, n - 1);
            t *= c1 / c2;
            if (t == 0.0) t = 0.0;
            DLT(m, n, n) = t;
        }
        c1 = c2;
    }
    for (int i = 0; i < np; i++) coeff[i] = DLT(order, np - 1, i);
    free(delta);
    return 0;
}

int yo_center_fd_coefficients(double* coeff, int order, int radius)
{
    if (radius < 1 || radius > 64) return -1;
    double pts[129];
    for (int i = -radius; i <= radius; i++) pts[i + radius] = (double)i;
    return yo_fd_coefficients(coeff, 0.0, order, pts, 2 * radius + 1);
}

/* The reference's compiler prints every FP constant with 16 significant digits
 * ("setprecision(15) << scientific", /root/reference/src/compiler/lib/Cpp.cpp:39-53)
 * into the generated kernel source; the kernel therefore uses strtod() of THAT string,
 * not the original double.  Integers print as ints. */
double yo_const_roundtrip(double v)
{
    if ((double)(int)v == v) return v;
    char buf[64];
    snprintf(buf, sizeof buf, "%.15e", v);
    return strtod(buf, NULL);
}

/* iso3dfd constants: c[0] = centre coefficient (x3, / 50^2), c[r] for r=1..radius.
 * /root/reference/src/stencils/Iso3dfdStencil.cpp:68-88. */
int yo_iso3dfd_coeffs(double* c, int radius)
{
    double full[129];
    int rc = yo_center_fd_coefficients(full, 2, radius);
    if (rc) return rc;
    double delta_xyz = 50.0;
    double d2 = delta_xyz * delta_xyz;
    for (int i = 0; i <= 2 * radius; i++) {
        if (i == radius) full[i] *= 3.0;
        full[i] /= d2;
    }
    for (int r = 0; r <= radius; r++) c[r] = yo_const_roundtrip(full[radius + r]);
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* iso3dfd, one time-step.                                                                */
/* Equation: /root/reference/src/stencils/Iso3dfdStencil.cpp:63-137 (get_next_p) and        */
/* :140-152 (define); association order = the reference compiler's "-target pseudo"          */
/* output:                                                           */
/*   acc = p*c0;  for r=1..R: acc = acc + ((((((x-r)+(x+r))+(y-r))+(y+r))+(z-r))+(z+r))*c_r  */
/*   p(t+1) = ((2*p) - p(t-1)) + acc*v                                                      */
/* Arrays: p_cur / p_io are (nx+2h, ny+2h, nz+2h) row-major with z unit stride, halo h on     */
/* every side; p_io holds p(t-1) on entry and p(t+1) in its domain points on exit (the       */
/* reference allocates 2 step slots and writes t+1 over t-1); v is      */
/* (nx,ny,nz).  h >= radius.                                                                */
/* ------------------------------------------------------------------------------------ */
#define ISO_STEP_BODY(T, FMA)                                                                   \
    const int64_t sy = nz + 2 * h, sx = (ny + 2 * h) * sy;                                     \
    T c[65];                                                                                   \
    for (int r = 0; r <= radius; r++) c[r] = (T)coef[r];                                       \
    _Pragma("omp parallel for collapse(2) schedule(static)")                                   \
    for (int64_t x = 0; x < nx; x++)                                                           \
        for (int64_t y = 0; y < ny; y++) {                                                     \
            const T* pc = p_cur + (x + h) * sx + (y + h) * sy + h;                             \
            T* pio = p_io + (x + h) * sx + (y + h) * sy + h;                                   \
            const T* vv = v + (x * ny + y) * nz;                                               \
            for (int64_t z = 0; z < nz; z++) {                                                 \
                T acc = pc[z] * c[0];                                                          \
                for (int r = 1; r <= radius; r++) {                                            \
                    T s = pc[z - r * sx] + pc[z + r * sx];                                     \
                    s = s + pc[z - r * sy];                                                    \
                    s = s + pc[z + r * sy];                                                    \
                    s = s + pc[z - r];                                                         \
                    s = s + pc[z + r];                                                         \
                    if (contract == 2 && r == 1) acc = FMA(pc[z], c[0], s * c[1]);                 \
                    else acc = contract ? FMA(s, c[r], acc) : acc + s * c[r];                  \
                }                                                                              \
                T lhs = ((T)2 * pc[z]) - pio[z];                                               \
                pio[z] = contract ? FMA(acc, vv[z], lhs) : lhs + acc * vv[z];                  \
            }                                                                                  \
        }

void yo_iso3dfd_step_f32(const float* p_cur, float* p_io, const float* v, int64_t nx, int64_t ny,
                         int64_t nz, int64_t h, int radius, const double* coef, int contract)
{
    ISO_STEP_BODY(float, fmaf)
}

void yo_iso3dfd_step_f64(const double* p_cur, double* p_io, const double* v, int64_t nx, int64_t ny,
                         int64_t nz, int64_t h, int radius, const double* coef, int contract)
{
    ISO_STEP_BODY(double, fma)
}

/* Run `steps` steps.  p0 = API step 0 = p(t), p1 = API step 1 (consumed as p(t-1) on the
 * first step because slot = imod(t-1, 2) = 1).  Returns 0 if the
 * final p(t+1) is in p0, 1 if it is in p1. */
int yo_iso3dfd_run_f32(float* p0, float* p1, const float* v, int64_t nx, int64_t ny, int64_t nz,
                       int64_t h, int radius, int steps, int contract)
{
    double coef[65];
    if (yo_iso3dfd_coeffs(coef, radius)) return -1;
    float* cur = p0;
    float* io = p1;
    for (int s = 0; s < steps; s++) {
        yo_iso3dfd_step_f32(cur, io, v, nx, ny, nz, h, radius, coef, contract);
        float* t = cur; cur = io; io = t;
    }
    return cur == p0 ? 0 : 1;
}

int yo_iso3dfd_run_f64(double* p0, double* p1, const double* v, int64_t nx, int64_t ny, int64_t nz,
                       int64_t h, int radius, int steps, int contract)
{
    double coef[65];
    if (yo_iso3dfd_coeffs(coef, radius)) return -1;
    double* cur = p0;
    double* io = p1;
    for (int s = 0; s < steps; s++) {
        yo_iso3dfd_step_f64(cur, io, v, nx, ny, nz, h, radius, coef, contract);
        double* t = cur; cur = io; io = t;
    }
    return cur == p0 ? 0 : 1;
}

int yo_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------ */
/* Emitter-generated solutions (oracle/gen/<name>.gen.h, written by                         */
/* yask_b200/emitter/yask_cuda_emit.py from the reference compiler's own output).          */
/* Each part is a list of single-assignment statements in the reference's evaluation       */
/* order; this file supplies the loop nest and the statement vocabulary.  All arithmetic   */
/* is in the element type with no contraction (== reference built with -ffp-contract=off). */
/* ------------------------------------------------------------------------------------ */
#define YO_GEN_MAX_ACC 96
typedef struct {
    int64_t nx, ny, nz;                 /* rank-domain box [0,n) */
    void* ptr[YO_GEN_MAX_ACC];          /* element (0,0,0) of each access' step slot */
    int64_t sx[YO_GEN_MAX_ACC], sy[YO_GEN_MAX_ACC], sz[YO_GEN_MAX_ACC];
    int64_t off[3], gfirst[3], glast[3];   /* for sub-domain conditions (global indices) */
    int64_t bx, by, bz;                 /* first index of the box (0, or -write_halo for scratch parts) */
    int64_t t;                          /* step index (step conditions) */
} yo_gen_args;
typedef struct { const char* name; void (*fn)(const yo_gen_args*); int nacc; } yo_gen_part;

#define YO_GEN_LOOP_BEGIN                                                      \
    _Pragma("omp parallel for collapse(2) schedule(static)")                   \
    for (int64_t x = A->bx; x < A->nx; x++)                                    \
        for (int64_t y = A->by; y < A->ny; y++)                                \
            for (int64_t z = A->bz; z < A->nz; z++) {
#define YO_GEN_LOOP_END }
#define RD(a, m, dx, dy, dz) (((const T*)A->ptr[a])[(x + (dx)) * A->sx[a] + (y + (dy)) * A->sy[a] + (z + (dz)) * A->sz[a]])
#define WR(a, m, v) ((T*)A->ptr[a])[x * A->sx[a] + y * A->sy[a] + z * A->sz[a]] = (v)
#define C(v) ((T)(v))
#define G(i) ((i) == 0 ? x + A->off[0] : ((i) == 1 ? y + A->off[1] : z + A->off[2]))
#define GT A->t
#define GF(i) A->gfirst[i]
#define GL(i) A->glast[i]
#define ADD(a, b) ((a) + (b))
#define SUB(a, b) ((a) - (b))
#define MUL(a, b) ((a) * (b))
#define DIV(a, b) ((a) / (b))
/* The places where the reference's DEFAULT build (GCC -O3, -ffp-contract=fast) fuses a product into the addition that
 * consumes it, as worked out by the emitter (contract_like_gcc): evaluated as two rounded operations by default (== the
 * reference built with -ffp-contract=off) and as one fma after yo_gen_set_contract(1) (== its default build). */
static int yo_gen_contract = 0;
void yo_gen_set_contract(int on) { yo_gen_contract = on; }
#define YO_FMA(a, b, c) (sizeof(T) == 4 ? (T)fmaf((float)(a), (float)(b), (float)(c)) : (T)fma((double)(a), (double)(b), (double)(c)))
#define MAD(a, b, c) (yo_gen_contract ? YO_FMA(a, b, c) : (T)((T)((a) * (b)) + (c)))
#define MSB(a, b, c) (yo_gen_contract ? YO_FMA(a, b, -(c)) : (T)((T)((a) * (b)) - (c)))
#define NMAD(a, b, c) (yo_gen_contract ? YO_FMA(-(a), b, c) : (T)((c) - (T)((a) * (b))))
#define NMSB(a, b, c) (yo_gen_contract ? YO_FMA(-(a), b, -(c)) : (T)((T)(-(T)((a) * (b))) - (c)))
/* DSL math functions: libm of the element type, as the reference's non-SVML build calls them per element
 * (/root/reference/src/kernel/lib/realv.hpp:648-738). */
#define YF1(fd, ff, a) (sizeof(T) == 4 ? (T)ff((float)(a)) : (T)fd((double)(a)))
#define YF_sqrt(a) YF1(sqrt, sqrtf, a)
#define YF_cbrt(a) YF1(cbrt, cbrtf, a)
#define YF_fabs(a) YF1(fabs, fabsf, a)
#define YF_erf(a) YF1(erf, erff, a)
#define YF_exp(a) YF1(exp, expf, a)
#define YF_log(a) YF1(log, logf, a)
#define YF_sin(a) YF1(sin, sinf, a)
#define YF_cos(a) YF1(cos, cosf, a)
#define YF_atan(a) YF1(atan, atanf, a)
#define YF_pow(a, b) (sizeof(T) == 4 ? (T)powf((float)(a), (float)(b)) : (T)pow((double)(a), (double)(b)))
#define YF_min(a, b) ((b) < (a) ? (b) : (a))
#define YF_max(a, b) ((a) < (b) ? (b) : (a))
#define YF_sincos(a, s, c)                                                                        \
    do { if (sizeof(T) == 4) { float s_, c_; sincosf((float)(a), &s_, &c_); (s) = (T)s_; (c) = (T)c_; } \
         else { double s_, c_; sincos((double)(a), &s_, &c_); (s) = (T)s_; (c) = (T)c_; } } while (0)

#include "gen/gen_all.inc"

static const struct { const char* name; const yo_gen_part* parts; int nparts; } yo_gen_table[] = {YO_GEN_TABLE};

/* Run part `part` (0-based, in stage order) of generated solution `stencil` over the box in A. */
int yo_gen_run_part(const char* stencil, int part, const yo_gen_args* A)
{
    for (size_t i = 0; i < sizeof(yo_gen_table) / sizeof(yo_gen_table[0]); i++) {
        if (strcmp(stencil, yo_gen_table[i].name)) continue;
        if (part < 0 || part >= yo_gen_table[i].nparts) return -2;
        yo_gen_table[i].parts[part].fn(A);
        return 0;
    }
    return -1;
}
