/*
 * halma_unbind.h -- C-ABI of libhalma_unbind.so, the B200 (sm_100a) replacement for
 * pyHALMA's direct-sum potential / unbinding hot path.
 *
 * Plain C: pointers and sizes only, no torch or C++ types.  Every entry point returns
 * HALMA_OK (0) or a negative HALMA_ERR_* code (CUDA failures are reported as
 * HALMA_ERR_CUDA with the text in halma_last_error()).  There is no CPU fallback: a
 * call on a machine without a usable CUDA device fails with HALMA_ERR_NO_DEVICE.
 *
 * Reference interfaces replaced (paths relative to the pyHALMA checkout):
 *   halma_potential_f32        fortran_modules/particle_subroutines.f90:466-514
 *                              (brute_force_binding_energy) and :517-556 (serial twin),
 *                              reached through python_scripts/halo_gas.py:182,208.
 *   halma_plan_* / halma_unbind_*   the per-halo drivers around that kernel:
 *                              python_scripts/halo_properties.py:333-361 (stellar
 *                              energy step and mask), python_scripts/halo_gas.py:299-492
 *                              (gas classes, energy step, mask, mass sums) and
 *                              halo_properties.py:16-60 (mass, centre of mass, bulk
 *                              velocity), iterated to a fixed point.
 *
 * Units are the reference's: comoving Mpc, Msun, km/s.
 */
#ifndef HALMA_UNBIND_H
#define HALMA_UNBIND_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HALMA_ABI_VERSION 2

/* ---- status codes ------------------------------------------------------------------ */
#define HALMA_OK                0
#define HALMA_ERR_INVALID      -1   /* bad argument (null pointer, negative size, bad enum) */
#define HALMA_ERR_NO_DEVICE    -2   /* no CUDA device / device index out of range          */
#define HALMA_ERR_CUDA         -3   /* CUDA runtime error, see halma_last_error()          */
#define HALMA_ERR_ALIGN        -4   /* device pointer not 16-byte aligned                  */
#define HALMA_ERR_TOO_LARGE    -5   /* more than 2^31-16 particles in one call             */
#define HALMA_ERR_STATE        -6   /* call order violated (e.g. run before upload)        */
#define HALMA_ERR_NCCL         -7   /* NCCL could not be loaded or a collective failed     */

/* ---- arithmetic modes --------------------------------------------------------------- *
 * EXACT: bit-faithful to the reference kernel: per pair IEEE sqrt and divide, the
 *        coordinate-inequality predicate on the float32 inputs, r^2 contracted as
 *        fma(dz,dz, fma(dx,dx, dy*dy)) (what GCC emits for compile-f2py:4), one float32
 *        accumulator per target summed in ascending source order; source classes are
 *        added in float32 in class order (halo_gas.py:301-450).
 * FAST:  same predicate and per-pair formula with rsqrt.approx, float32 partial sums
 *        over at most 32 sources flushed into a float64 accumulator; the result is
 *        rounded once to float32.  Agrees with a float64-accumulating evaluation of the
 *        reference terms to better than 1e-6 relative.                                   */
#define HALMA_MODE_FAST   0
#define HALMA_MODE_EXACT  1

/* loop drivers (halma_unbind_config.use_graph) */
#define HALMA_DRIVER_AUTO     0
#define HALMA_DRIVER_GRAPH    1
#define HALMA_DRIVER_ENQUEUE  2
#define HALMA_DRIVER_FUSED    3

const char *halma_last_error(void);
int halma_abi_version(void);
int halma_device_count(int *count);
/* name[0..len) receives the device name; sm_count / clock_khz may be null. */
int halma_device_info(int device, char *name, int len, int *sm_count, int *clock_khz,
                      int64_t *mem_bytes);

/* Pinned host memory for callers that want fast host<->device copies. */
int halma_host_alloc(void **ptr, int64_t bytes);
int halma_host_free(void *ptr);

/* ------------------------------------------------------------------------------------ *
 * brute_force_binding_energy: out_be[i] = sum_j m_j / |r_j - r_i| over the sources whose
 * x, y AND z all differ from the target's (particle_subroutines.f90:497-510).  Positive,
 * Msun/Mpc, no G, no softening.  HOST pointers; float32 like the f2py signature.
 * n_tgt == 0 is a no-op (the wrapper at halo_gas.py:169 never reaches Fortran either).
 * FAST mode, large calls (>= HALMA_POT_PLAN_MIN_PAIRS pairs, default 1e10; 0 = never) whose targets
 * are the sources or a block of them -- gas-gas of RPS, stars-stars of most_bound_particle, the
 * concat(gas, stars, DM) -> stars call of escape_velocity_unbinding_fortran -- are run as a one-pass
 * plan (predicate-free kernel, symmetric self-term): same tolerance, 1.3-1.8x faster.
 * ------------------------------------------------------------------------------------ */
int halma_potential_f32(int device, int mode,
                        const float *src_m, const float *src_x, const float *src_y,
                        const float *src_z, int64_t n_src,
                        const float *tgt_x, const float *tgt_y, const float *tgt_z,
                        int64_t n_tgt, float *out_be);

/* Host-only helper of the above (no GPU needed): the offset of the block of sources whose coordinates
 * equal the targets' bit for bit, in the same order, or -1.  0 with n_tgt == n_src is a self call. */
int64_t halma_find_target_block(const float *src_x, const float *src_y, const float *src_z, int64_t n_src,
                                const float *tgt_x, const float *tgt_y, const float *tgt_z, int64_t n_tgt);

/* Same with DEVICE pointers (16-byte aligned) on `stream` (a cudaStream_t, 0 = default);
 * asynchronous.  workspace: device scratch of halma_potential_workspace_bytes(). */
int64_t halma_potential_workspace_bytes(int64_t n_src, int64_t n_tgt);
int halma_potential_f32_dev(int device, int mode,
                            const float *src_m, const float *src_x, const float *src_y,
                            const float *src_z, int64_t n_src,
                            const float *tgt_x, const float *tgt_y, const float *tgt_z,
                            int64_t n_tgt, float *out_be, void *workspace, void *stream);

/* ------------------------------------------------------------------------------------ *
 * Unbinding plans.  A plan owns the device-resident state of a batch ("catalogue") of
 * haloes: members (targets that are also sources) concatenated with CSR `offsets`, and
 * up to HALMA_MAX_GROUPS groups of fixed external sources, each with its own CSR
 * offsets over the same haloes.
 *
 * Source order / classes per halo:
 *   split_classes = 0 (stellar, halo_properties.py:333-339): one in-order sum over
 *       ext group 0 .. n_pre-1, members, ext group n_pre .. n_groups-1.
 *   split_classes = 1 (gas, halo_gas.py:301-450; most_bound_particle, :508-622): the same
 *       order, but every group and the members are summed on their own and the class sums
 *       are added in float32 in that order (gas: n_pre = 0, members first; most-bound:
 *       n_pre = n_groups, members last).
 * Energy and mask per pass (halo_properties.py:342-359 / halo_gas.py:456-476):
 *       pe = -(float)Phi; pe *= (float)G; pe *= (float)kappa        (float32)
 *       E  = 0.5*((vx-vbx)^2 + (vy-vby)^2 + (vz-vbz)^2) + pe        (float64)
 *       bound = E <= 0
 * Iteration: members <- bound members (stable order); unless vb_fixed,
 * vb <- sum(m v)/sum(m) over the new members (halo_properties.py:45-60); repeat until
 * the member count stops changing, the set is empty, or max_iter passes were made.
 * max_iter = 1 with vb_fixed = 1 is exactly the reference's one-pass function.
 * ------------------------------------------------------------------------------------ */
#define HALMA_MAX_GROUPS 4

typedef struct halma_plan halma_plan;

typedef struct halma_unbind_config {
    int32_t struct_size;     /* sizeof(halma_unbind_config), for ABI checks              */
    int32_t device;
    int32_t mode;            /* HALMA_MODE_*                                             */
    int32_t n_groups;        /* external source groups, 0..HALMA_MAX_GROUPS              */
    int32_t n_pre;           /* groups summed before the members (split_classes = 0)     */
    int32_t split_classes;   /* 0 stellar layout, 1 gas layout                           */
    int32_t vb_fixed;        /* 1: bulk velocity given per halo and held fixed           */
    int32_t max_iter;        /* >= 1                                                     */
    double  G;               /* 4.3e-9 (km/s)^2 Mpc/Msun; cast to float32 like numpy     */
    double  kappa;           /* factor_v**2 (stars) or 2.0 (gas); cast to float32        */
    int32_t rank;            /* split mode: this process's rank ...                      */
    int32_t n_ranks;         /* ... of n_ranks sharing ONE halo by target groups (1 = off) */
    int32_t use_graph;       /* loop driver (the field keeps its round-1 name; HALMA_DRIVER_*):
                                0 AUTO: on one GPU the whole loop runs as ONE persistent cooperative kernel
                                  (fused.cu: every pass inside the launch, grid barriers between the phases, no
                                  host round trip; halma_run_stats.potential_ms comes from in-kernel timestamps);
                                  split mode and the tuning kernel shapes use ENQUEUE.
                                1 GRAPH: one CUDA-graph launch, a WHILE conditional node whose body is one pass
                                  of stand-alone kernels; the scheduling kernel sets the condition on the device.
                                2 ENQUEUE: the host enqueues passes of stand-alone kernels ahead of the device and
                                  learns about convergence from a pinned flag; per-launch CUDA-event timing.
                                3 FUSED: as AUTO, but an error where the persistent kernel cannot serve the plan.
                                All drivers run the same phase code and give bit-identical results.       */
    int32_t symmetric;       /* 1 (FAST mode, plans large enough for the predicate-free kernel's throughput
                                shape): every member x member pair of different 128-member tiles is
                                evaluated once and feeds both particles' sums -- half the rsqrt work of the
                                self term.  The two-sided sums are float64 atomics of addends rounded to a
                                per-halo quantum (2^-37 of M/extent) inside whose window the additions
                                are exact, so runs stay bit-reproducible; a halo whose sums leave the
                                window (potential > ~3e4 M/extent) is recomputed one-sided.  Predicate,
                                tolerances and outputs are unchanged.                                  */
    int32_t cache_external;  /* 1 (FAST mode, predicate-free path, one GPU): the external sources never change
                                during the loop, so the sum over them is evaluated by the first pass only,
                                kept per member in float64 and added in every later pass.  Same terms, same
                                tolerances; halma_run_stats.evaluations counts what was actually evaluated. */
    int32_t incremental;     /* 1 (FAST mode, predicate-free path; also in split mode): when a pass removed at
                                most a third of a halo's members, the next pass of that halo only evaluates
                                survivors x removed members (with the reference's predicate) and subtracts
                                that from the complete float64 potential kept from the pass before: the same
                                per-pair terms, the difference is float32 partial-sum rounding, ~1e-7 of the
                                removed contribution.  If the removed members carried more than 80 % of some
                                member's potential, or the pass fell back to the predicated kernel for any
                                other reason, the halo is recomputed in full instead.                        */
} halma_unbind_config;

typedef struct halma_halo_result {
    int64_t n_bound;         /* members left                                             */
    int32_t n_iter;          /* potential passes made                                    */
    int32_t converged;       /* 1 if the last pass removed nothing (or the set emptied)  */
    double  mass;            /* sum m over the bound set        (halo_properties.py:16)  */
    double  com[3];          /* centre of mass                  (halo_properties.py:26)  */
    double  vb[3];           /* bulk velocity used by the last pass / of the bound set   */
    int64_t pairs;           /* (target, source) pairs evaluated over all passes         */
    int64_t most_bound;      /* local index of the member of the LAST pass with the largest
                                sum m/r, lowest index on ties (halo_gas.py:627-632: argmin of
                                the negated potential); -1 if no pass was made            */
    double  mass_initial;    /* sum m over all members                  (halo_gas.py:483) */
    double  cold_bound_mass; /* T <  cold_T and bound at the end        (halo_gas.py:484) */
    double  unbound_cold_mass; /* T <  cold_T and removed in some pass  (halo_gas.py:489) */
    double  unbound_hot_mass;  /* T >= cold_T and removed in some pass  (halo_gas.py:490) */
} halma_halo_result;

typedef struct halma_run_stats {
    double  total_ms;        /* CUDA-event time of the whole run on the plan's stream    */
    double  potential_ms;    /* time spent evaluating potentials: sum over potential-kernel launches
                                (CUDA events, ENQUEUE driver), or the potential phases of the persistent
                                kernel (its own %globaltimer stamps, FUSED driver); 0 with GRAPH     */
    int32_t potential_launches; /* potential passes run (stand-alone launches, or phases of the one kernel) */
    int32_t launches;        /* all kernel launches issued by the run                    */
    int32_t passes;          /* loop passes executed (max over haloes)                   */
    int32_t driver;          /* HALMA_DRIVER_* that ran                                  */
    int64_t pairs;           /* interactions (target, source) summed over haloes and passes:
                                what the reference's double loop visits                  */
    int64_t evaluations;     /* 1/r evaluations made for them: = pairs, or fewer with
                                symmetric = 1 (one evaluation serves both particles)      */
    double  loop_ms;         /* FUSED: duration of the persistent kernel by its own clock; else 0 */
    double  comm_ms;         /* split mode: CUDA-event time of the NCCL collectives of the run   */
    int64_t comm_bytes;      /* split mode: bytes handed to the collectives, all passes          */
    double  phase_ms[5];     /* FUSED: the persistent kernel's time by phase: prologue (pack, first decisions,
                                first ticket table), potential passes, energy + compaction, commit + ticket
                                tables, epilogue (member lists)                                    */
} halma_run_stats;

/* offsets: int64[n_halo+1], ext_offsets[g]: int64[n_halo+1] for g < n_groups (host). */
int halma_plan_create(const halma_unbind_config *cfg, int64_t n_halo, const int64_t *offsets,
                      const int64_t *const *ext_offsets, halma_plan **out);
void halma_plan_destroy(halma_plan *plan);

/* Host OR device float64 arrays (unified addressing tells them apart) of length offsets[n_halo] (the reference keeps particle data in
 * float64; the float32 cast of positions and masses for the potential is done on the
 * device with round-to-nearest, like np.float32(...) at halo_gas.py:172-178). */
int halma_plan_upload_members(halma_plan *plan, const double *x, const double *y, const double *z,
                              const double *vx, const double *vy, const double *vz,
                              const double *mass);
int halma_plan_upload_group(halma_plan *plan, int group, const double *mass, const double *x,
                            const double *y, const double *z);
/* Optional member temperatures (double[N]) for the cold / hot mass sums of RPS
 * (halo_gas.py:479-492); cold means T < cold_T (5e4 K in the reference). */
int halma_plan_upload_temp(halma_plan *plan, const double *temp, double cold_T);
/* vb: double[3*n_halo]; required when vb_fixed = 1. */
int halma_plan_set_vb(halma_plan *plan, const double *vb);
/* The uploads are asynchronous copies on the plan's stream; this waits for them (callers that pipeline several
 * plans use it to keep one plan's upload from sharing the host link with the next one's). */
int halma_plan_sync(halma_plan *plan);

/* Split mode only: join an NCCL communicator.  unique_id is the 128-byte ncclUniqueId
 * made by rank 0 (halma_nccl_unique_id) and broadcast by the caller. */
int halma_nccl_unique_id(void *unique_id_128);
int halma_plan_join(halma_plan *plan, const void *unique_id_128);
/* A communicator that outlives plans (creating one costs ~0.1 s): create it once per
 * process, hand it to every split-mode plan with halma_plan_use_comm (the plan does not
 * own it), destroy it after the last plan. */
typedef struct halma_comm halma_comm;
int halma_comm_create(int device, int rank, int n_ranks, const void *unique_id_128, halma_comm **out);
void halma_comm_destroy(halma_comm *comm);
int halma_plan_use_comm(halma_plan *plan, halma_comm *comm);

/* Runs the whole loop on the device from the uploaded (pristine) inputs; can be called
 * repeatedly.  Blocks until done; stats may be null. */
int halma_plan_run(halma_plan *plan, halma_run_stats *stats);

/* Tuning aid: phase times (ns) of the first 16 passes of the last FUSED run: out48[3 * pass + k], k = potential,
 * energy + compaction, commit + ticket table. */
int halma_plan_debug_pass_ns(halma_plan *plan, uint32_t *out48);

/* Any output pointer may be null.  mask/be/energy/idx are indexed like the member input
 * arrays: mask uint8[N]; be float32[N] (sum m/r at the last pass the particle took part
 * in); energy float64[N]; idx int32[N]: for halo h, idx[offsets[h] .. offsets[h]+n_bound)
 * holds the ascending local indices of the bound members; halos: [n_halo]. */
int halma_plan_download(halma_plan *plan, uint8_t *mask, float *be, double *energy, int32_t *idx,
                        halma_halo_result *halos);

/* One-shot catalogue call: create + upload + run + download + destroy.  HOST (or device) pointers.
 * offsets: int64[n_halo + 1] CSR over the concatenated member arrays; group_offsets[g]: int64[n_halo + 1] CSR
 * of external group g over its concatenated arrays group_mass[g], group_x[g], ...; vb: double[3 * n_halo]
 * when cfg->vb_fixed, else may be null; temp (optional): member temperatures for the cold / hot mass sums
 * (halo_gas.py:479-492).  Outputs as halma_plan_download; any may be null.  This is the batched form of
 * the per-halo loop at pyHALMA.py:930-1067 for the unbinding step. */
int halma_unbind_catalogue(const halma_unbind_config *cfg, int64_t n_halo, const int64_t *offsets,
                           const double *x, const double *y, const double *z,
                           const double *vx, const double *vy, const double *vz, const double *mass,
                           const int64_t *const *group_offsets, const double *const *group_mass,
                           const double *const *group_x, const double *const *group_y,
                           const double *const *group_z, const double *vb, const double *temp, double cold_T,
                           uint8_t *mask, float *be, double *energy, int32_t *idx,
                           halma_halo_result *halos, halma_run_stats *stats);

/* The same for a single halo: group_n[g] = length of group g's arrays. */
int halma_unbind_halo(const halma_unbind_config *cfg, int64_t n,
                      const double *x, const double *y, const double *z,
                      const double *vx, const double *vy, const double *vz, const double *mass,
                      const int64_t *group_n, const double *const *group_mass,
                      const double *const *group_x, const double *const *group_y,
                      const double *const *group_z, const double *vb,
                      uint8_t *mask, float *be, double *energy, int32_t *idx,
                      halma_halo_result *result, halma_run_stats *stats);

/* ------------------------------------------------------------------------------------ *
 * The other two routines of the f2py module `particle`, so that the
 * module can be swapped without any Fortran toolchain.  HOST pointers, float32 like the
 * f2py signatures; sums are accumulated in float64 on the device (the reference's float32
 * OpenMP reductions are order-dependent, so parity is to ~1e-5, not bitwise).
 *
 * halma_halo_shape_f32: particle_subroutines.f90:160-214 (halo_shape, with DIAGONALISE
 *   :12-126 and SORT_EIGEN :129-157): semi-axes a >= b >= c = square roots of the
 *   eigenvalues of sum(m r_i r_j) / sum(m) of the (already centred) positions.
 * halma_sigma_projections_f32: particle_subroutines.f90:217-461: line-of-sight velocity
 *   dispersion inside R05 for the three projections, V/sigma and lambda_R averaged over
 *   the projections, from n_cell x n_cell maps binned on `grid` (nearest grid point, first
 *   minimum on ties).  part_list is 0-BASED here (the f2py shim subtracts the 1 the
 *   reference's wrapper adds, halo_properties.py:787); n_all = length of the st_* arrays.
 *   out5 = SIG_1D_x_05, SIG_1D_y_05, SIG_1D_z_05, V_sigma, lambda.
 * ------------------------------------------------------------------------------------ */
int halma_halo_shape_f32(int device, const float *x, const float *y, const float *z, const float *mass,
                         int64_t npart, float *eigenvalues3);
int halma_sigma_projections_f32(int device, int64_t npart, const float *grid, int32_t n_cell,
                                const int32_t *part_list, int64_t n_all, const float *st_x, const float *st_y,
                                const float *st_z, const float *st_vx, const float *st_vy, const float *st_vz,
                                const float *st_mass, float cx, float cy, float cz, float R05x, float R05y,
                                float R05z, float ll, float *out5);

/* ------------------------------------------------------------------------------------ *
 * The gather step before the path: python_scripts/halo_gas.py:223-277
 * (st_gas_dm_particles_inside) with :9-52 (patch_to_particles), :56-141
 * (AMRgrid_to_particles), :216-218 (parallel_inside) and the KD-tree ball queries :255,:269.
 * A snapshot (AMR hierarchy + cell fields + DM + star particles) is uploaded ONCE and stays
 * in HBM; every halo then costs one halma_snapshot_gather.
 *
 * halma_snapshot_create: the patches as masclet's grid_data gives them, index 0 = base
 *   grid: level[p], extent nx/ny/nz[p] (cells of level[p]), rx/ry/rz[p] = centre of the
 *   patch's first PARENT cell.  Patches of level 0 are never read (halo_gas.py:107).
 * halma_snapshot_upload_patch: cell fields of one patch, C-ordered (ix slowest, iz fastest =
 *   the reference's loop nest): delta = rho/rho_B - 1, velocity in units of c, temperature
 *   (float32); cr0amr != 0 = not refined, solapst != 0 = not overlapped (uint8).  Patches that
 *   were never uploaded contribute nothing.
 * halma_snapshot_upload_particles: kind 0 = DM, 1 = stars; mass already in Msun; id may be
 *   null (the particle index is reported instead).
 * halma_snapshot_gather: dm_heavy_min = -INFINITY keeps the DM in one group; a finite value splits
 *   it into the heavy species (mass >= dm_heavy_min, `mandatory` at halo_gas.py:341-345) and the
 *   light one, each ascending.  counts4 = n_gas, n_dm (all, or heavy), n_dm_light, n_star.
 *   gas = one particle at the centre of every cell that is strictly
 *   inside the box [c - R, c + R]^3, flagged by cr0amr and solapst, and at distance < R; mass
 *   = (1 + delta) * rho_B * res^3 * mass_scale (mass_scale = rete**3, halo_gas.py:246),
 *   velocity * 3e5; ascending patch, then ix, iy, iz -- bit-identical to the reference's
 *   float64 arrays.  DM / stars = particles with squared distance <= R^2, ASCENDING index (the
 *   reference's KD-tree returns the same set in tree order).
 * halma_snapshot_fetch: copies the last gather's result to host arrays: gas8 = x, y, z, vx,
 *   vy, vz, mass, temp; dm4 / dml4 / st4 = x, y, z, mass; st_id = star ids.  Any pointer may be null.
 * halma_snapshot_result_device: the same result as DEVICE addresses (ptr4[g] + k * counts4[g] is
 *   column k of group g = gas, DM, light DM, stars), valid until the next gather or destroy.
 *   halma_plan_upload_members / _group / _temp accept them directly (they take host or device
 *   pointers), so a halo can go from the resident snapshot through the unbinding loop without
 *   touching the host.
 * halma_snapshot_fetch_star: x, y, z, mass and id of gathered star k (the most bound particle).
 * ------------------------------------------------------------------------------------ */
/* CUDA-event time (ms) of the kernels of the last halma_halo_shape_f32 /
 * halma_sigma_projections_f32 / halma_snapshot_gather call made on this thread. */
double halma_last_kernel_ms(void);

typedef struct halma_snapshot halma_snapshot;
int halma_snapshot_create(int device, double L, int32_t ncoarse, int64_t n_patch, const int32_t *level,
                          const int32_t *nx, const int32_t *ny, const int32_t *nz, const double *rx,
                          const double *ry, const double *rz, halma_snapshot **out);
void halma_snapshot_destroy(halma_snapshot *snap);
int64_t halma_snapshot_cells(const halma_snapshot *snap);
int halma_snapshot_upload_patch(halma_snapshot *snap, int64_t patch, const float *delta, const float *vx,
                                const float *vy, const float *vz, const float *temp, const uint8_t *cr0amr,
                                const uint8_t *solapst);
int halma_snapshot_upload_particles(halma_snapshot *snap, int kind, int64_t n, const double *x, const double *y,
                                    const double *z, const double *mass, const int64_t *id);
int halma_snapshot_gather(halma_snapshot *snap, double cx, double cy, double cz, double R, double rho_B,
                          double mass_scale, double dm_heavy_min, int64_t *counts4);
/* AMRgrid_to_particles alone (halo_gas.py:56-141; call site halo_properties.py:318): gas cells strictly inside
 * the BOX [c - R, c + R]^3, no sphere test, no particles (counts4[1..3] = 0). */
int halma_snapshot_gather_box(halma_snapshot *snap, double cx, double cy, double cz, double R, double rho_B,
                              double mass_scale, int64_t *counts4);
int halma_snapshot_fetch(halma_snapshot *snap, double *const *gas8, double *const *dm4, double *const *dml4,
                         double *const *st4, int64_t *st_id);
int halma_snapshot_result_device(halma_snapshot *snap, double **ptr4, int64_t **st_id, int64_t *counts4);
int halma_snapshot_fetch_star(halma_snapshot *snap, int64_t k, double *xyzm4, int64_t *id);

/* EXACT mode computes rn(m / rn(sqrt(r^2))) with branch-free sequences while the operands are
 * in a safe exponent window and with the IEEE library routines otherwise.  This self-test
 * compares the two on the device: all 2^24 (mantissa, exponent parity) inputs of the square
 * root plus n_random pseudo-random (m, r^2) pairs over the whole window; *mismatches must
 * come back 0. */
int halma_selftest_exact_arith(int device, int64_t n_random, uint64_t seed, int64_t *mismatches);

/* ------------------------------------------------------------------------------------ *
 * Pipe-rate microbenchmark used for the roofline denominator: measures
 * MUFU.RSQ, FFMA and packed FFMA2 issue rates per SM per clock, and the SM clock during
 * the measurement.  out: double[8] = {rsq_per_clk_sm, ffma_per_clk_sm, ffma2_per_clk_sm,
 * sm_clock_mhz (clock64 vs globaltimer inside the kernel), sm_count, and the absolute
 * rates rsq, ffma, ffma2 in 1e9 thread-instructions per second}.
 * ------------------------------------------------------------------------------------ */
int halma_microbench(int device, double *out8);

#ifdef __cplusplus
}
#endif
#endif /* HALMA_UNBIND_H */
