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
