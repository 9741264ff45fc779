/*
 * oracle/halma_oracle.c -- CPU restatement of pyHALMA's direct-sum potential kernel.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pyhalma_b200/ may import, link or
 * call this file; it is the checker for the CUDA path (tests/, smoke(), and
 * bench.py's cpu_baseline / --impl reference legs).
 *
 * PARITY STATUS: the reference ships no test, golden vector or fixture for this
 * path and no Fortran compiler exists in the build image,
 * so the KERNEL ARITHMETIC IS "PARITY UNPINNED" by the reference itself.  What
 * pins it here: the known-answer tests in tests/test_oracle_kat.py, and the
 * golden fixtures under tests/golden/ that were produced by running the
 * reference's own Python drivers (halo_gas.RPS, most_bound_particle,
 * halo_properties.escape_velocity_unbinding_fortran) with this file standing in
 * for the f2py module.
 *
 * Every function cites the reference lines it follows, relative to
 * /root/reference/.
 *
 * Floating-point semantics follow fortran_modules/compile-f2py:4
 *   -O3 -fopenmp -mieee-fp -ftree-vectorize -march=native   (no fast-math,
 *   GCC default -ffp-contract=fast, so a*a+b*b+c*c is contracted into FMAs on
 *   any FMA-capable host).  -march=native is pinned to x86-64-v3 (AVX2+FMA) in
 *   oracle/Makefile so the built object travels between hosts.
 *   oracle_potential_f32seq_fma() spells the contraction out with fmaf() and a
 *   test asserts both agree bit for bit, which pins what the compiler did.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------------------------------------------------------------------------
 * particle_subroutines.f90:517-556  serial_brute_force_binding_energy
 *   :541      binding_energy(:) = 0.
 *   :543-544  do ip=1,ntest / do ip2=1,ntotal        (j ascending, f32 accumulator)
 *   :545-547  nested /= on x, then y, then z: pair counts only if all three differ
 *   :548-550  r = sqrt(dx**2 + dy**2 + dz**2),  dx = total_x(ip2) - test_x(ip)
 *   :551      binding_energy(ip) = binding_energy(ip) + total_mass(ip2) / r
 * `real` is 4 bytes: every operation below is float.
 * ------------------------------------------------------------------------- */
static inline float pair_term_f32(float m, float xs, float ys, float zs,
                                  float xt, float yt, float zt)
{
    const float dx = xs - xt;
    const float dy = ys - yt;
    const float dz = zs - zt;
    const float r = sqrtf(dx * dx + dy * dy + dz * dz);
    return m / r;
}

void oracle_potential_f32seq_serial(int64_t ntotal, const float *total_mass,
                                    const float *total_x, const float *total_y,
                                    const float *total_z, int64_t ntest,
                                    const float *test_x, const float *test_y,
                                    const float *test_z, float *binding_energy)
{
    for (int64_t ip = 0; ip < ntest; ++ip) {
        const float xt = test_x[ip], yt = test_y[ip], zt = test_z[ip];
        float be = 0.f;
        for (int64_t ip2 = 0; ip2 < ntotal; ++ip2) {
            if (total_x[ip2] != xt) {
                if (total_y[ip2] != yt) {
                    if (total_z[ip2] != zt) {
                        be = be + pair_term_f32(total_mass[ip2], total_x[ip2],
                                                total_y[ip2], total_z[ip2], xt, yt, zt);
                    }
                }
            }
        }
        binding_energy[ip] = be;
    }
}

/* ---------------------------------------------------------------------------
 * particle_subroutines.f90:466-514  brute_force_binding_energy (OpenMP)
 *   :490      call OMP_SET_NUM_THREADS(ncores)
 *   :494-496  PARALLEL DO over ip with REDUCTION(+:binding_energy)
 * Each ip is owned by one thread whose private copy starts at 0, and the final
 * combine adds exact zeros from the other threads, so the result is the same
 * in-order float sum as the serial routine.
 * ------------------------------------------------------------------------- */
void oracle_potential_f32seq(int ncores, int64_t ntotal, const float *total_mass,
                             const float *total_x, const float *total_y,
                             const float *total_z, int64_t ntest,
                             const float *test_x, const float *test_y,
                             const float *test_z, float *binding_energy)
{
#ifdef _OPENMP
    if (ncores > 0) omp_set_num_threads(ncores);
#else
    (void)ncores;
#endif
#pragma omp parallel for schedule(static)
    for (int64_t ip = 0; ip < ntest; ++ip) {
        const float xt = test_x[ip], yt = test_y[ip], zt = test_z[ip];
        float be = 0.f;
        for (int64_t ip2 = 0; ip2 < ntotal; ++ip2) {
            if (total_x[ip2] != xt) {
                if (total_y[ip2] != yt) {
                    if (total_z[ip2] != zt) {
                        be = be + pair_term_f32(total_mass[ip2], total_x[ip2],
                                                total_y[ip2], total_z[ip2], xt, yt, zt);
                    }
                }
            }
        }
        binding_energy[ip] = be;
    }
}

/* Same loop with the FMA contraction written out: GCC's widening_mul pass turns
 * (dx*dx + dy*dy) + dz*dz into fma(dz,dz, fma(dx,dx, dy*dy)).  The CUDA exact
 * mode uses this exact expression; tests assert it equals the compiled form. */
void oracle_potential_f32seq_fma(int ncores, int64_t ntotal, const float *total_mass,
                                 const float *total_x, const float *total_y,
                                 const float *total_z, int64_t ntest,
                                 const float *test_x, const float *test_y,
                                 const float *test_z, float *binding_energy)
{
#ifdef _OPENMP
    if (ncores > 0) omp_set_num_threads(ncores);
#else
    (void)ncores;
#endif
#pragma omp parallel for schedule(static)
    for (int64_t ip = 0; ip < ntest; ++ip) {
        const float xt = test_x[ip], yt = test_y[ip], zt = test_z[ip];
        float be = 0.f;
        for (int64_t ip2 = 0; ip2 < ntotal; ++ip2) {
            if (total_x[ip2] != xt && total_y[ip2] != yt && total_z[ip2] != zt) {
                const float dx = total_x[ip2] - xt;
                const float dy = total_y[ip2] - yt;
                const float dz = total_z[ip2] - zt;
                const float r2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
                be = be + total_mass[ip2] / sqrtf(r2);
            }
        }
        binding_energy[ip] = be;
    }
}

/* ---------------------------------------------------------------------------
 * f64acc: the same per-term float arithmetic (predicate, r, m/r exactly as
 * :545-551) but the running sum is kept in double.  This is the accuracy
 * anchor for the GPU fast mode (SURVEY.md §7 hard part 1): the reference's own
 * in-order float sum drifts from it by 1e-6..3e-3 relative as N grows.
 * ------------------------------------------------------------------------- */
void oracle_potential_f64acc(int ncores, int64_t ntotal, const float *total_mass,
                             const float *total_x, const float *total_y,
                             const float *total_z, int64_t ntest,
                             const float *test_x, const float *test_y,
                             const float *test_z, double *binding_energy)
{
#ifdef _OPENMP
    if (ncores > 0) omp_set_num_threads(ncores);
#else
    (void)ncores;
#endif
#pragma omp parallel for schedule(static)
    for (int64_t ip = 0; ip < ntest; ++ip) {
        const float xt = test_x[ip], yt = test_y[ip], zt = test_z[ip];
        double be = 0.0;
        for (int64_t ip2 = 0; ip2 < ntotal; ++ip2) {
            if (total_x[ip2] != xt) {
                if (total_y[ip2] != yt) {
                    if (total_z[ip2] != zt) {
                        be += (double)pair_term_f32(total_mass[ip2], total_x[ip2],
                                                    total_y[ip2], total_z[ip2], xt, yt, zt);
                    }
                }
            }
        }
        binding_energy[ip] = be;
    }
}

/* Number of (target, source) pairs the predicate at :545-547 drops; used by the
 * known-answer tests for the lattice cases. */
int64_t oracle_count_excluded(int64_t ntotal, const float *total_x, const float *total_y,
                              const float *total_z, int64_t ntest, const float *test_x,
                              const float *test_y, const float *test_z)
{
    int64_t n = 0;
#pragma omp parallel for schedule(static) reduction(+ : n)
    for (int64_t ip = 0; ip < ntest; ++ip)
        for (int64_t ip2 = 0; ip2 < ntotal; ++ip2)
            if (!(total_x[ip2] != test_x[ip] && total_y[ip2] != test_y[ip] &&
                  total_z[ip2] != test_z[ip]))
                ++n;
    return n;
}

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Wall-clock helper so bench.py times the CPU leg the way BASELINE.md §3 says
 * (omp_get_wtime around the call). */
double oracle_wtime(void)
{
#ifdef _OPENMP
    return omp_get_wtime();
#else
    return 0.0;
#endif
}

/* ===========================================================================
 * §8f-4: the other two routines of the f2py module `particle`.
 * =========================================================================== */

/* particle_subroutines.f90:12-126  DIAGONALISE: eigenvalues of a symmetric 3x3 matrix
 * by cyclic Jacobi rotations, at most 100 sweeps, in double precision.  The input
 * and the output are REAL*4 (:19); only the upper triangle is used. */
void oracle_diagonalise(const float input[3][3], float eigen_out[3])
{
    double a[3][3], d[3], b[3], zacc[3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) a[i][j] = (double)input[i][j];                 /* :29-33 */
    for (int i = 0; i < 3; ++i) {
        b[i] = a[i][i];
        d[i] = b[i];
        zacc[i] = 0.0;
    }                                                                               /* :35-39 */
    double sum_elements = 0.0;
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 3; ++j) sum_elements += fabs(a[i][j]);                  /* :42-47 */
    for (int sweep = 1; sweep <= 100; ++sweep) {                                    /* :50 */
        double off = 0.0;
        for (int i = 0; i < 2; ++i)
            for (int j = i + 1; j < 3; ++j) off += fabs(a[i][j]);
        if (off < (double)1.e-4f * sum_elements) break;        /* :57, `1.e-4` is a REAL*4 literal */
        const double limit = sweep < 4 ? 0.2 * off * off : 0.0;                     /* :59-63 */
        for (int i = 0; i < 2; ++i) {
            for (int j = i + 1; j < 3; ++j) {
                double g = 100.0 * fabs(a[i][j]);                                   /* :66 */
                if (sweep > 4 && fabs(d[i]) + g == fabs(d[i]) && fabs(d[j]) + g == fabs(d[j])) {
                    a[i][j] = 0.0;                                                  /* :68-72 */
                } else if (fabs(a[i][j]) > limit) {
                    double h = d[j] - d[i], t;
                    if (fabs(h) + g == fabs(h)) {
                        t = a[i][j] / h;                                            /* :75-76 */
                    } else {
                        const double theta = 0.5 * h / a[i][j];
                        t = 1.0 / (fabs(theta) + sqrt(1.0 + theta * theta));
                        if (theta < 0.0) t = -t;                                    /* :78-80 */
                    }
                    const double c = 1.0 / sqrt(1.0 + t * t), s = t * c, tau = s / (1.0 + c);
                    h = t * a[i][j];
                    zacc[i] -= h;
                    zacc[j] += h;
                    d[i] -= h;
                    d[j] += h;
                    a[i][j] = 0.0;                                                  /* :82-90 */
                    for (int k = 0; k < i; ++k) {                                   /* :91-96 */
                        const double p = a[k][i], q = a[k][j];
                        a[k][i] = p - s * (q + p * tau);
                        a[k][j] = q + s * (p - q * tau);
                    }
                    for (int k = i + 1; k < j; ++k) {                               /* :97-102 */
                        const double p = a[i][k], q = a[k][j];
                        a[i][k] = p - s * (q + p * tau);
                        a[k][j] = q + s * (p - q * tau);
                    }
                    for (int k = j + 1; k < 3; ++k) {                               /* :103-108 */
                        const double p = a[i][k], q = a[j][k];
                        a[i][k] = p - s * (q + p * tau);
                        a[j][k] = q + s * (p - q * tau);
                    }
                }
            }
        }
        for (int i = 0; i < 3; ++i) {                                               /* :113-117 */
            b[i] += zacc[i];
            d[i] = b[i];
            zacc[i] = 0.0;
        }
    }
    for (int i = 0; i < 3; ++i) eigen_out[i] = (float)d[i];                         /* :121-123 */
}

/* particle_subroutines.f90:129-157  SORT_EIGEN: selection sort that moves the LARGEST
 * value to the front (the header comment says increasing, the `.GE.` test says otherwise). */
void oracle_sort_eigen(float *e, int n)
{
    for (int i = 0; i < n - 1; ++i) {
        int k = i;
        float v = e[i];
        for (int j = i + 1; j < n; ++j)
            if (e[j] >= v) {
                k = j;
                v = e[j];
            }
        if (k != i) {
            e[k] = e[i];
            e[i] = v;
        }
    }
}

/* particle_subroutines.f90:160-214  halo_shape: mass-weighted second-moment tensor of
 * the (already centred) positions, normalised by the total mass, Jacobi eigenvalues,
 * sorted largest first, square roots = semi-axes a >= b >= c.
 * float32 accumulation in particle order (what one OpenMP thread does; with more threads
 * the reference's REDUCTION changes the order and, as written, shares `rvec` between
 * threads -- SURVEY.md §5 -- so only the one-thread semantics are well defined).
 * wide = 1 accumulates in double: the accuracy anchor for the GPU path. */
void oracle_halo_shape(int64_t npart, const float *x, const float *y, const float *z, const float *mass, int wide,
                       float eigenvalues[3])
{
    float t32[3][3] = {{0}}, m32 = 0.f;
    double t64[3][3] = {{0}}, m64 = 0.0;
    for (int64_t ip = 0; ip < npart; ++ip) {
        const float r[3] = {x[ip], y[ip], z[ip]};
        for (int j = 0; j < 3; ++j)
            for (int i = 0; i < 3; ++i) {
                if (wide)
                    t64[i][j] += (double)mass[ip] * (double)r[i] * (double)r[j];
                else
                    t32[i][j] = t32[i][j] + mass[ip] * r[i] * r[j];                 /* :194 */
            }
        if (wide)
            m64 += (double)mass[ip];
        else
            m32 = m32 + mass[ip];                                                   /* :197 */
    }
    float tn[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) tn[i][j] = wide ? (float)(t64[i][j] / m64) : t32[i][j] / m32;   /* :203 */
    oracle_diagonalise(tn, eigenvalues);                                            /* :206 */
    oracle_sort_eigen(eigenvalues, 3);                                              /* :207 */
    for (int i = 0; i < 3; ++i) eigenvalues[i] = sqrtf(eigenvalues[i]);             /* :209-211 */
}

/* particle_subroutines.f90:217-461  sigma_projections.  part_list is 0-based here (the
 * Python wrapper adds 1 for Fortran, halo_properties.py:787).  Maps are n_cell x n_cell,
 * indexed [first + n_cell * second] like the Fortran (first, second) subscripts.
 * wide = 1 keeps the maps and sums in double (accuracy anchor).
 * out5 = SIG_1D_x_05, SIG_1D_y_05, SIG_1D_z_05, V_sigma, lambda. */
static int nearest_cell(const float *grid, int n_cell, float d)
{
    int best = 0;                       /* minloc(abs(grid - d), dim = 1): first minimum */
    float bv = fabsf(grid[0] - d);
    for (int k = 1; k < n_cell; ++k) {
        const float v = fabsf(grid[k] - d);
        if (v < bv) {
            bv = v;
            best = k;
        }
    }
    return best;
}

#include <stdlib.h>
#define SIGMA_BODY(REAL, SQRT, FABS)                                                                          \
    const size_t nn = (size_t)n_cell * n_cell;                                                                \
    REAL *vcm[3], *sd[3], *sig[3];                                                                            \
    int *cnt[3];                                                                                              \
    for (int a = 0; a < 3; ++a) {                                                                             \
        vcm[a] = (REAL *)calloc(nn, sizeof(REAL));                                                            \
        sd[a] = (REAL *)calloc(nn, sizeof(REAL));                                                             \
        sig[a] = (REAL *)calloc(nn, sizeof(REAL));                                                            \
        cnt[a] = (int *)calloc(nn, sizeof(int));                                                              \
    }                                                                                                         \
    int *cell = (int *)malloc(sizeof(int) * 3 * (size_t)(npart > 0 ? npart : 1));                             \
    for (int64_t ip = 0; ip < npart; ++ip) {                                       /* :264-281 */            \
        const int64_t q = part_list[ip];                                                                      \
        const int ix = nearest_cell(grid, n_cell, st_x[q] - cx), iy = nearest_cell(grid, n_cell, st_y[q] - cy), \
                  iz = nearest_cell(grid, n_cell, st_z[q] - cz);                                              \
        cell[3 * ip] = ix; cell[3 * ip + 1] = iy; cell[3 * ip + 2] = iz;                                      \
        const size_t kx = iy + (size_t)n_cell * iz, ky = ix + (size_t)n_cell * iz, kz = ix + (size_t)n_cell * iy; \
        vcm[0][kx] += (REAL)st_vx[q] * (REAL)st_mass[q]; vcm[1][ky] += (REAL)st_vy[q] * (REAL)st_mass[q];         \
        vcm[2][kz] += (REAL)st_vz[q] * (REAL)st_mass[q];                                                          \
        sd[0][kx] += st_mass[q]; sd[1][ky] += st_mass[q]; sd[2][kz] += st_mass[q];                            \
        cnt[0][kx]++; cnt[1][ky]++; cnt[2][kz]++;                                                             \
    }                                                                                                         \
    for (int a = 0; a < 3; ++a)                                                    /* :286-288 */            \
        for (size_t k = 0; k < nn; ++k)                                                                       \
            if (sd[a][k] != 0) vcm[a][k] = vcm[a][k] / sd[a][k];                                              \
    for (int64_t ip = 0; ip < npart; ++ip) {                                       /* :296-306 */            \
        const int64_t q = part_list[ip];                                                                      \
        const int ix = cell[3 * ip], iy = cell[3 * ip + 1], iz = cell[3 * ip + 2];                            \
        const size_t kx = iy + (size_t)n_cell * iz, ky = ix + (size_t)n_cell * iz, kz = ix + (size_t)n_cell * iy; \
        const REAL ax_ = (REAL)st_vx[q] - vcm[0][kx], ay_ = (REAL)st_vy[q] - vcm[1][ky],                      \
                   az_ = (REAL)st_vz[q] - vcm[2][kz];                                                         \
        sig[0][kx] += ax_ * ax_; sig[1][ky] += ay_ * ay_; sig[2][kz] += az_ * az_;                            \
    }                                                                                                         \
    for (int a = 0; a < 3; ++a)                                                    /* :310-312 */            \
        for (size_t k = 0; k < nn; ++k)                                                                       \
            if (cnt[a][k] != 0) sig[a][k] = SQRT(sig[a][k] / cnt[a][k]);                                      \
    REAL s05[3] = {0, 0, 0};                                                                                  \
    int c05[3] = {0, 0, 0};                                                                                   \
    const float r05[3] = {R05x, R05y, R05z};                                                                  \
    for (int64_t ip = 0; ip < npart; ++ip) {                                       /* :326-352 */            \
        const int64_t q = part_list[ip];                                                                      \
        const float dx = st_x[q] - cx, dy = st_y[q] - cy, dz = st_z[q] - cz;                                  \
        const int ix = cell[3 * ip], iy = cell[3 * ip + 1], iz = cell[3 * ip + 2];                            \
        const float dist[3] = {sqrtf(dy * dy + dz * dz), sqrtf(dx * dx + dz * dz), sqrtf(dx * dx + dy * dy)}; \
        const size_t k3[3] = {iy + (size_t)n_cell * iz, ix + (size_t)n_cell * iz, ix + (size_t)n_cell * iy};  \
        for (int a = 0; a < 3; ++a)                                                                           \
            if (dist[a] < r05[a]) { s05[a] += sig[a][k3[a]]; c05[a]++; }                                      \
    }                                                                                                         \
    for (int a = 0; a < 3; ++a) out5[a] = (float)(c05[a] > 0 ? s05[a] / c05[a] : s05[a]);   /* :356-366 */   \
    REAL vs[3] = {0, 0, 0}, lam[3] = {0, 0, 0};                                                               \
    for (int a = 0; a < 3; ++a) {            /* a = 2: XY plane (:380-401), 1: XZ (:404-424), 0: YZ (:427-448) */ \
        REAL sumV = 0, sumS = 0, up = 0, down = 0;                                                            \
        for (int j = 0; j < n_cell; ++j)                                                                      \
            for (int i = 0; i < n_cell; ++i) {                                                                \
                const float rbin = sqrtf(grid[i] * grid[i] + grid[j] * grid[j]);                              \
                if (rbin < r05[a] + 2 * ll) {                                                                 \
                    const size_t k = i + (size_t)n_cell * j;                                                  \
                    sumV += vcm[a][k] * vcm[a][k] * sd[a][k];                                                 \
                    sumS += sig[a][k] * sig[a][k] * sd[a][k];                                                 \
                    up += sd[a][k] * rbin * FABS(vcm[a][k]);                                                  \
                    down += sd[a][k] * rbin * SQRT(vcm[a][k] * vcm[a][k] + sig[a][k] * sig[a][k]);            \
                }                                                                                             \
            }                                                                                                 \
        if (sumS > 0) vs[a] = SQRT(sumV / sumS);                                                              \
        if (down > 0) lam[a] = up / down;                                                                     \
    }                                                                                                         \
    out5[3] = (float)((vs[0] + vs[1] + vs[2]) / 3);                                /* :453-454 */            \
    out5[4] = (float)((lam[0] + lam[1] + lam[2]) / 3);                                                        \
    for (int a = 0; a < 3; ++a) { free(vcm[a]); free(sd[a]); free(sig[a]); free(cnt[a]); }                    \
    free(cell);

void oracle_sigma_projections(int64_t npart, const float *grid, int n_cell, const int32_t *part_list,
                              const float *st_x, const float *st_y, const float *st_z, const float *st_vx,
                              const float *st_vy, const float *st_vz, const float *st_mass, float cx, float cy,
                              float cz, float R05x, float R05y, float R05z, float ll, int wide, float out5[5])
{
    if (wide) {
        SIGMA_BODY(double, sqrt, fabs)
    } else {
        SIGMA_BODY(float, sqrtf, fabsf)
    }
}
