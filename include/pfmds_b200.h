/*
 * pfmds_b200 — C ABI of the B200-native MD inner loop (libpfmds_b200.so).
 *
 * This is the drop-in boundary for the hot path of AlexanderSidorenkov/PFMDS: everything the
 * reference executes inside `do md_step=0,md_step_limit` of md()
 * (code_source/MOLECULAR_DYNAMICS/md_simulation.f90:114-243).  The reference has no FFI of its own
 * (it is one Fortran process); the seam is the set of module procedures md() calls, and each entry
 * point below names the reference procedures it replaces.  Plain C types only: an ISO_C_BINDING
 * Fortran host, the C++ run_md_simulation host in pfmds_b200/host/ and ctypes all bind the same
 * symbols.
 *
 * Conventions
 *  - all reals are FP64, all indices int32, exactly like the reference built with -fdefault-real-8;
 *  - per-atom arrays are xyz-interleaved `a[3*i+k]` in FILE order (Fortran `a(3,N)`), atom and group
 *    numbers are 1-based as in the settings file;
 *  - every call returns 0 on success or a PFMDS_ERR_* code; pfmds_last_error() gives the message
 *    (the reference's own `stop` text where one exists);
 *  - one host thread drives one context; contexts are independent (own CUDA stream), so several
 *    can share a GPU (ensemble mode) or sit one per GPU;
 *  - pfmds_advance only enqueues work; positions, velocities, forces and neighbour lists stay on
 *    the device until pfmds_energies / pfmds_diagnostics / pfmds_download / pfmds_synchronize.
 */
#ifndef PFMDS_B200_H
#define PFMDS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pfmds_ctx pfmds_ctx;

enum {
    PFMDS_OK = 0,
    PFMDS_ERR_INVALID = 1,             /* bad argument / call order */
    PFMDS_ERR_CUDA = 2,                /* CUDA runtime failure (no device, out of memory, ...) */
    PFMDS_ERR_OUT_OF_CELL = 10,        /* "particle out of cell"            md_general.f90:342-364 */
    PFMDS_ERR_TOO_MANY_NEIGHBOURS = 11,/* "error: too many neighbours"      md_neighbours.f90:80,151 */
    PFMDS_ERR_GR_NEIGHBOURS = 12,      /* "not enough / too many gr nearest neibs" graphenenorm.f90:26,30 */
    PFMDS_ERR_NHC_PARAMS = 13,         /* "error: wrong nhc parameters"     md_integrators.f90:187 */
    PFMDS_ERR_LIST_SIZE = 14,          /* nl%N mismatches                   md_neighbours.f90:64,134; LennardJonesCosine.f90:60 */
    PFMDS_ERR_UNKNOWN_INTERACTION = 15,/* "error: unknown interaction name" md_interactions.f90:119-120 */
    PFMDS_ERR_UNSUPPORTED = 20         /* input the reference accepts but this path refuses loudly */
};

enum { PFMDS_NVE = 0, PFMDS_NVT = 1, PFMDS_NVMS = 2 };  /* integrator_name 'nve' (or any other) / 'nvt' / 'nvms' */

/* Replaces read_box_size + read_particles handing their arrays to md() (md_read_write.f90:22-61;
 * types particles / simulation_cell, md_general.f90:9-17).  Arrays are copied. */
int pfmds_create(pfmds_ctx** ctx, int device, int n_atoms, const double* positions, const double* velocities,
                 const double* masses, const double box_size[3]);

/* create_particle_group (md_general.f90:57-80): `indexes` are the 1-based atom numbers in the
 * reference's order (type column first, file order second). */
int pfmds_set_group(pfmds_ctx* ctx, int group_num, int n, const int* indexes);

/* all_moving / xyz_moving / z_moving / all_atoms group numbers (md_simulation.f90:57-60). */
int pfmds_set_roles(pfmds_ctx* ctx, int all_moving_group, int xyz_moving_group, int z_moving_group, int all_atoms_group);

/* One `change_group_num` entry of the settings file (md_simulation.f90:63-71): change_particle_group_N
 * (md_general.f90:82-94) is applied to group `group_to` at the top of every md step (md_simulation.f90:116-119), i.e.
 * group_to%N follows group_from%N up to step change_ts1, grows by one atom at change_ts1 and then every change_frec steps
 * while step < change_ts2 (deposition); the group exposes the first N atoms of the index list given to pfmds_set_group.
 * Entries are applied in the order they were added.  Steps must be advanced in sequence from md step 0. */
int pfmds_add_group_change(pfmds_ctx* ctx, int group_from, int group_to, int change_ts1, int change_ts2, int change_frec);

/* Current group%N (differs from the size given to pfmds_set_group only for targets of pfmds_add_group_change), after the
 * last step queued by pfmds_advance: the writers (write_particle_group, md_read_write.f90:65-107) need it. */
int pfmds_group_size(pfmds_ctx* ctx, int group_num, int* n);

/* create_nose_hoover_chain + set_nose_hoover_chain (md_integrators.f90:165-198), one per settings line. */
int pfmds_add_nhc(pfmds_ctx* ctx, int group_num, double temperature, int M, double q1);

/* zero_momentum_period, invert_z_vel (md_simulation.f90:55,72). */
int pfmds_set_misc(pfmds_ctx* ctx, int zero_momentum_period, int invert_z_vel);

/* One block of create_interactions (md_interactions.f90:59-136).  `name` in lj, lj1g, ljc, morsec,
 * tb, rjl.  `params` in parameter-file order:
 *   lj, lj1g: eps sig R1 R2            ljc: eps sig delt R1 R2 simplified(0/1)
 *   morsec: d r a delt R1 R2 simplified    tb: d s b r0 delt a0 c0 d0 R1 R2    rjl: A0 xi p q r0 R1 R2
 * nl_n neighbour-list lines (2, 1, 3, 3, 1, 1): group_nums[2*nl_n] = g1 g2 per line, then capacity,
 * r_cut and update_period per line.  Call in file order: forces accumulate in that order and
 * ljc/morsec take their nearest-3 list from the first tb interaction added before them. */
int pfmds_add_interaction(pfmds_ctx* ctx, const char* name, int n_params, const double* params, int nl_n,
                          const int* group_nums, const int* neighb_num_max, const double* r_cut, const int* update_period);

/* Runs md steps first_md_step .. first_md_step+n_steps-1 on the device: per step check_positions,
 * invert_z_velocities, [nvt: integrate_nose_hoover_chain], velocity half-kick, drift + wrap,
 * update_interactions_neighbour_lists (cell-binned rebuild when mod(step,period)==0), zero_momentum,
 * zero_forces + calculate_forces, half-kick, [nvt: NHC], [nvms: quench]   (md_simulation.f90:138-186;
 * step 0 only evaluates lists and forces).  Asynchronous. */
int pfmds_advance(pfmds_ctx* ctx, int integrator, double dt, int first_md_step, int n_steps);

/* Same, and the force evaluation of the LAST step also produces the potential energies (md() knows which steps log:
 * md_simulation.f90:188), so the pfmds_energies that follows needs no second sweep over the neighbour lists. */
int pfmds_advance_with_energy(pfmds_ctx* ctx, int integrator, double dt, int first_md_step, int n_steps);

/* md() with period_log = 1 reads the energies after every step (md_simulation.f90:188-199).  This call runs n_steps steps like
 * pfmds_advance; the steps with mod(step, log_period) == 0 evaluate the potential energies in their force pass and append what
 * pfmds_energies would return to a log that stays on the device, and all rows come back with ONE copy at the end of the call:
 * rows[r*row_len + ...] = e_inter[n_interactions], kinetic energy, temperature, e_nhc[n_nhc]  (row_len >= n_interactions+2+n_nhc),
 * *n_rows = number of logged steps.  Same numbers, bit for bit, as pfmds_advance_with_energy(1 step) + pfmds_energies per
 * logged step.  Synchronises once. */
int pfmds_advance_logged(pfmds_ctx* ctx, int integrator, double dt, int first_md_step, int n_steps, int log_period, double* rows,
                         int row_len, int* n_rows);

/* calculate_potential_energies + calculate_temperature(all_moving) + calculate_nose_hoover_chain_energy
 * (md_simulation.f90:191-198).  e_inter[n_interactions], e_nhc[n_nhc].  Synchronises. */
int pfmds_energies(pfmds_ctx* ctx, double* e_inter, double* kinetic_energy, double* temperature, double* e_nhc);

/* calculate_force_sum, calculate_mass_center, calculate_mass_center_velocity over all_atoms,
 * find_max_velocity, nlists_load (md_simulation.f90:212-223; md_interactions.f90:427-444).
 * nl_load has one entry per neighbour-list line in file order.  Synchronises. */
int pfmds_diagnostics(pfmds_ctx* ctx, double force_sum[3], double mass_center[3], double mass_center_velocity[3],
                      double* max_velocity, int* nl_load);

/* Copies state back in file order (x y z per atom, what write_particle_group prints: md_read_write.f90:65-107); any pointer may be
 * NULL.  The device undoes its cell ordering itself and every requested array arrives with one transfer straight into the
 * caller's buffer, so pinned buffers are filled at full PCIe rate.  Synchronises. */
int pfmds_download(pfmds_ctx* ctx, double* positions, double* velocities, double* forces);

/* The reference-shaped view of one neighbour list (type neighbour_list, md_general.f90:30-35):
 * rows in group-1 order, entries = 1-based group-2 local numbers in ascending order,
 * nlist[row*neighb_num_max + p].  For parity tests and debugging.  Synchronises. */
int pfmds_neighbours(pfmds_ctx* ctx, int interaction, int list, int* nlist, int* nnum, int* lessnnum);

/* gr_norm(3,N) of an ljc / morsec interaction, rows in group-1 order (graphenenorm.f90:38-56). */
int pfmds_normals(pfmds_ctx* ctx, int interaction, double* gr_norm);

/* x(M), v(M) of thermostat k (0-based), for exact restarts and tests. */
int pfmds_get_nhc(pfmds_ctx* ctx, int k, double* x, double* v);
int pfmds_set_nhc(pfmds_ctx* ctx, int k, const double* x, const double* v);

/* Exact restart (the reference can only restart from a snapshot xyz, which loses the thermostat chains, the step
 * counter and 3 digits of the velocities: md_simulation.f90:233-236).  pfmds_save_state fills `blob` (pfmds_state_size doubles)
 * with what positions and velocities do not carry: every Nose-Hoover chain (x, v, q, cached kinetic energy) and group%N of
 * every group; take positions / velocities with pfmds_download.  pfmds_restore_state puts all of it into a freshly configured
 * context (same groups, thermostats, interactions), rebuilds every neighbour list from the given positions and evaluates the
 * forces, i.e. leaves the context as it was after the checkpointed md step; continue with pfmds_advance(first = that step + 1).
 * The continuation is bit-identical to the uninterrupted run when the checkpointed step is a rebuild step of every list
 * (mod(step, update_period) == 0) and all lists share one update_period; otherwise it agrees to rounding. */
int pfmds_state_size(pfmds_ctx* ctx, long long* n_doubles);
int pfmds_save_state(pfmds_ctx* ctx, double* blob);
int pfmds_restore_state(pfmds_ctx* ctx, const double* positions, const double* velocities, const double* blob);

/* Seconds spent per phase, from CUDA events, when PFMDS_TIMERS=1 is set in the environment
 * (otherwise zeros): pos_vel, nlists, nlsearch, nldistance, forces, energy — the buckets of the
 * reference's PERFOMANCE table (md_simulation.f90:250-259). */
int pfmds_timers(pfmds_ctx* ctx, double seconds[6]);

/* Overwrite positions and/or velocities (file order, either may be NULL) of a live context, e.g. to
 * restart from a snapshot; every neighbour list is rebuilt at the next step. */
int pfmds_upload(pfmds_ctx* ctx, const double* positions, const double* velocities);

/* Sum of nnum over the rows of one neighbour list (directed pairs), for work models. */
int pfmds_pair_count(pfmds_ctx* ctx, int interaction, int list, long long* pairs);
/* Listed directed pairs whose current min-image distance is below `r` (e.g. R2 of the potential: the pairs of a row between R2 and
 * the list's r_cut leave the pair routines after the distance test and are charged 30 flop, not the full count, by bench.py). */
int pfmds_pair_count_within(pfmds_ctx* ctx, int interaction, int list, double r, long long* pairs);

/* Per-kernel device times: with profiling on, CUDA events bracket every launch of each kernel class on
 * the context's stream.  pfmds_kernel_times fills ms[k], count[k] for k < n (classes named by
 * pfmds_kernel_name).  pfmds_timer_start/stop time a region on the context's stream. */
int pfmds_set_profiling(pfmds_ctx* ctx, int on);
int pfmds_kernel_times(pfmds_ctx* ctx, int n, double* ms, long long* count);
const char* pfmds_kernel_name(int k);
int pfmds_timer_start(pfmds_ctx* ctx);
int pfmds_timer_stop(pfmds_ctx* ctx, double* ms);

/* Roofline denominators measured on the device: FP64 FMA peak (TFLOP/s) and copy bandwidth (GB/s). */
int pfmds_measure_peaks(int device, double* dfma_tflops, double* copy_gbs);

/* ---- slab spatial decomposition of one large cell over the GPUs of a node (BASELINE.json configs[3]; the
 * reference has no counterpart: its only multi-process mode is the ensemble) ------------------------------
 * Rank r owns the atoms with x in [r Lx/P, (r+1) Lx/P); ghost copies, migration at rebuild steps and the
 * sums over ranks (KE, momentum, energies) are handled inside pfmds_advance / pfmds_energies /
 * pfmds_diagnostics with NCCL on the context's stream.  Interactions: lj, lj1g, rjl.
 * pfmds_slab_unique_id: rank 0 creates the id, the host broadcasts the 128 bytes to the other ranks.
 * pfmds_create_slab: this rank's atoms (any order): 1-based global numbers, state, per-atom group bit mask
 * (bit g-1 = member of group g, g <= 31) and the GLOBAL size of every group; `capacity` = slots for local
 * atoms + ghosts.  Then pfmds_set_roles / add_nhc / set_misc / add_interaction / advance as usual
 * (pfmds_set_group is not used).  pfmds_slab_download returns this rank's atoms. */
int pfmds_slab_unique_id(char id[128]);
int pfmds_create_slab(pfmds_ctx** ctx, int device, int rank, int nranks, const char id[128], long long n_global, int n_local,
                      const int* global_index, const double* positions, const double* velocities, const double* masses,
                      const unsigned int* group_mask, int n_groups, const long long* group_sizes, const double box_size[3], int capacity);
/* Atoms this rank owns right now and the ghost copies it holds (they change at every list rebuild: migration, halo re-selection). */
int pfmds_slab_counts(pfmds_ctx* ctx, int* n_local, int* n_ghost);
int pfmds_slab_download(pfmds_ctx* ctx, int* n_local, int* global_index, double* positions, double* velocities, double* forces);
/* Overwrite the state of this rank's atoms, given in the order of the last pfmds_slab_download. */
int pfmds_slab_upload(pfmds_ctx* ctx, int n_local, const double* positions, const double* velocities);

/* Device self-test of the library's FP64 elementary functions against the CUDA math library: max errors
 * [0] exp (relative), [1] cosine switch / sincos (absolute), [2] rsqrt (relative), [3] hardware rsqrt seed. */
int pfmds_selftest_math(int device, double err[4]);
/* Same for the reduced-instruction forms of the second-generation rjl kernels (mathx.cuh exp_m, rsqrt_q, cos_switch_m,
 * half_switch): [0] exp relative on [-40, 40], [1] switch absolute, [2] rsqrt relative, [3] exp relative on [-600, 600]. */
int pfmds_selftest_math2(int device, double err[4]);

/* Contexts of this process alive on `device` (a context that shares its GPU leaves the parallel branches of a step to the others). */
int pfmds_live_contexts(int device);
/* Number of kernel launches issued so far by this context and device-time of the last advance (ms). */
int pfmds_launch_count(pfmds_ctx* ctx, long long* launches);

int pfmds_synchronize(pfmds_ctx* ctx);
const char* pfmds_last_error(pfmds_ctx* ctx);
int pfmds_destroy(pfmds_ctx* ctx);
const char* pfmds_version(void);

#ifdef __cplusplus
}
#endif
#endif
