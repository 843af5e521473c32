/*
 * atlj.h -- C ABI of libatlj.so: the B200 (sm_100a) Lennard-Jones pair-force hot path of
 * Allen-Tildesley/examples.  Plain pointers and sizes only; no torch / C++ types.
 *
 * Every entry point names the reference interface it replaces (file:line relative to the
 * reference tree).  The Python shim in examples_b200/ binds these with ctypes and presents
 * the reference's module API (force/hessian/PotentialType/introduction/conclusion).
 *
 * Conventions
 *   - positions r are (n,3) C-order fp64 in box = 1 units, as in the reference
 *     (python_examples/md_nve_lj.py:162-163); forces/energies in sigma = epsilon = 1 units.
 *   - "host" pointers are ordinary CPU memory (pageable or pinned); "dev" pointers are CUDA
 *     device memory on the library's current device.
 *   - return value: ATLJ_OK, one of the reference's assertion conditions (the shim raises
 *     AssertionError with the reference's message), or ATLJ_ERR_CUDA / ATLJ_ERR_ARG with the
 *     text available from atlj_last_error().
 *   - the library NEVER computes on the CPU: without a usable CUDA device every compute entry
 *     point returns ATLJ_ERR_CUDA.
 */
#ifndef ATLJ_H
#define ATLJ_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ATLJ_OK 0
#define ATLJ_ERR_DIM -1      /* 'Dimension error in force'        md_lj_module.py:77        */
#define ATLJ_ERR_SMALL -2    /* 'System is too small for cells'   md_lj_ll_module.py:107    */
#define ATLJ_ERR_INDEX -3    /* 'Index error'                     md_lj_ll_module.py:109    */
#define ATLJ_ERR_MISMATCH -4 /* 'Dimension mismatch in hessian'   md_lj_module.py:163       */
#define ATLJ_ERR_CUDA -100   /* CUDA runtime failure or no device */
#define ATLJ_ERR_ARG -101    /* invalid argument                  */

/* models */
#define ATLJ_MODEL_LJ 0  /* cut-and-shifted LJ, totals = {cut,pot,vir,lap}   md_lj_module.py    */
#define ATLJ_MODEL_WCA 1 /* WCA + Lees-Edwards, totals = {pot,vir,lap,pyx}   md_lj_le_module.py */

/* number of doubles in one per-step record written by the device-resident integrators */
#define ATLJ_NREC 12
/* record layout: [0..3] totals as above, [4] sum v^2, [5] sum f^2, [6] hessian (NVE, optional),
 * [7] ovr flag (0/1), [8] in-range pair count, [9] sum vx*vy (SLLOD), [10] strain (SLLOD) or
 * conserved-energy thermostat term (NHC), [11] number of owned atoms (multi-GPU)            */

/* ---------------------------------------------------------------------------------------------
 * library / device
 * ------------------------------------------------------------------------------------------- */
int atlj_version(void);
const char *atlj_last_error(void);
int atlj_device_count(void);      /* number of CUDA devices, 0 if none          */
int atlj_set_device(int device);  /* all later calls of this thread use device  */
int64_t atlj_kernel_launches(void); /* kernels launched by this library so far (bench: gpu_launches) */
/* Kernel selection by system size (no reference counterpart): systems of at most `natoms` atoms on one GPU use the
 * eight-lanes-per-atom pair kernel (the tiled kernels cannot fill a B200 there); 0 switches it off.
 * Default 16384 (measured: at N = 32,000 the tiled kernel is the faster one again), or the environment variable ATLJ_MULTI_MAX.  Returns the previous value.  Results are the same
 * pair set with the same exact in-range / overlap decisions whichever kernel runs; tests use it to cover each one. */
int64_t atlj_set_multi_max(int64_t natoms);

/* ---------------------------------------------------------------------------------------------
 * stateless drop-in path: host arrays in, host arrays out (H2D + kernels + D2H per call)
 * ------------------------------------------------------------------------------------------- */

/* Replaces md_lj_module.force (md_lj_module.py:68-149) when sc <= 0 or sc < 4 (all pairs) and
 * md_lj_ll_module.force (md_lj_ll_module.py:69-224) when sc >= 4 (link cells, sc =
 * floor(box/r_cut) computed by the caller with the reference's expression, :106).
 * r_cut_box_sq, box_sq, pot_cut are the host scalars of md_lj_module.py:80-88, passed in so the
 * in-range predicate uses bit-identical thresholds.  check_index != 0 enforces the cell-index
 * assertion of md_lj_ll_module.py:109 even on the all-pairs route.
 * totals = {cut, pot, vir, lap} after the scaling of md_lj_module.py:143-147; f is (n,3). */
int atlj_force_lj(const double *r_host, int64_t n, double box, double r_cut_box_sq, double box_sq, double pot_cut,
                  int sc, int check_index, double *f_host, double totals[4], int *ovr, int64_t *npair);

/* One atom against a set of partners: smc_lj_module.py:100-191 `force_1(ri, box, r_cut, r)` (the single-atom moves of
 * smc_nvt_lj.py:200-231).  ri[3] and r (nj,3) in box = 1 units; f_partial (nj,3) = force on the atom due to each
 * partner; totals = cut, pot, vir, lap with the reference's factors (:185-190).  With an overlap (:146-149) ovr = 1,
 * totals = 0 and f_partial = 0. */
int atlj_force_1(const double *ri_host, const double *r_host, int64_t nj, double box, double r_cut_box_sq,
                 double box_sq, double pot_cut, double *f_partial_host, double totals[4], int *ovr);
/* Replaces md_lj_le_module.force (md_lj_le_module.py:69-150; sc < 4) and md_lj_llle_module.force
 * (md_lj_llle_module.py:70-245; sc >= 4, shift = floor(strain*sc) of :112).  WCA potential,
 * Lees-Edwards image.  totals = {pot, vir, lap, pyx}. */
int atlj_force_le(const double *r_host, int64_t n, double box, double r_cut_box_sq, double box_sq, double strain,
                  int sc, int shift, int check_index, double *f_host, double totals[4], int *ovr, int64_t *npair);

/* Replaces md_lj_module.hessian (md_lj_module.py:151-213) / md_lj_ll_module.hessian
 * (md_lj_ll_module.py:227-357). */
int atlj_hessian_lj(const double *r_host, const double *f_host, int64_t n, double box, double r_cut_box_sq,
                    double box_sq, int sc, int check_index, double *hes);

/* The integer work on its own, for bit-exact tests: wrap (md_lj_ll_module.py:88; le != 0 adds the
 * pre-wrap of md_lj_llle_module.py:94), cell triplets c (int64, :108), and the per-cell atom lists
 * of :116-120 in CSR form: perm (atoms in C order of cells, ascending index inside a cell) and
 * cell_start (sc^3+1).  link_list_module.f90:103-161 is the Fortran twin. */
int atlj_cells(const double *r_host, int64_t n, int sc, int le, double strain, int64_t *c_host, int32_t *perm_host,
               int32_t *cell_start_host);

/* Neighbour-cell enumeration used by the pair kernel (full stencil = the reference's half stencil
 * md_lj_ll_module.py:84-86,128-129 / md_lj_llle_module.py:88-92,131-141 plus its mirror image).
 * nb_host is (sc^3, 30) int32, unused slots -1; count_host (sc^3). */
int atlj_neighbour_cells(int sc, int le, int shift, int32_t *nb_host, int32_t *count_host);

/* ---------------------------------------------------------------------------------------------
 * device-resident path: state stays in HBM between steps
 * (replaces the loop bodies md_nve_lj.py:178-192, bd_nvt_lj.py:221-233, md_nvt_lj.py:270-288,
 *  md_nvt_lj_le.py:226-240 together with the force call inside them)
 * ------------------------------------------------------------------------------------------- */
typedef struct atlj_sim atlj_sim;

/* capacity = maximum number of atoms this rank will ever hold (owned + halo).
 * stream = a cudaStream_t (0 = legacy default stream); all work is enqueued on it. */
atlj_sim *atlj_sim_create(int64_t capacity, void *stream);
/* flags: ATLJ_SIM_IPC = the position buffers and the cell table are allocations of their own, so that a slab rank can
 * export them to its neighbours with CUDA IPC (the neighbours write this rank's halo layers directly) */
#define ATLJ_SIM_IPC 1
atlj_sim *atlj_sim_create_ex(int64_t capacity, void *stream, int flags);
void atlj_sim_destroy(atlj_sim *s);

/* Model and geometry.  sc = floor(box/r_cut) (md_lj_ll_module.py:106); r_cut_box_sq, box_sq, pot_cut as
 * above.  Slab decomposition: this rank owns global cell layers [x_lo, x_hi) along axis 0; for a
 * single GPU x_lo = 0, x_hi = sc. */
int atlj_sim_configure(atlj_sim *s, int model, double box, double r_cut_box_sq, double box_sq, double pot_cut, int sc,
                       int x_lo, int x_hi);

/* Upload owned atoms: r (box units, wrapped), v, global ids (NULL = 0..n-1). */
int atlj_sim_upload(atlj_sim *s, const double *r_host, const double *v_host, const int32_t *id_host, int64_t n);
/* Download owned atoms in the order they are held (sorted by cell); ids tell which is which.
 * Any pointer may be NULL.  Returns the number of owned atoms through n_out. */
int atlj_sim_download(atlj_sim *s, double *r_host, double *v_host, double *f_host, int32_t *id_host, int64_t *n_out);
int64_t atlj_sim_natoms(atlj_sim *s);

/* One force evaluation on the current positions (cell build + pair kernel); rec[ATLJ_NREC]. */
int atlj_sim_force(atlj_sim *s, double strain, int with_hessian, double *rec_host);

/* nsteps velocity-Verlet steps (md_nve_lj.py:178-192); rec_host (nsteps, ATLJ_NREC) or NULL. */
int atlj_sim_run_nve(atlj_sim *s, double dt, int nsteps, int with_hessian, double *rec_host);
/* BAOAB Langevin steps (bd_nvt_lj.py:91-127,221-233) with a counter-based device Gaussian RNG (Philox4x32-10 keyed
 * by seed, counter = (global atom id, step)).  o_decay = exp(-gamma*dt) and o_noise = sqrt(1-exp(-2*gamma*dt)) *
 * sqrt(temperature) are the two scalars of o_propagator (bd_nvt_lj.py:124-127), evaluated by the caller with the
 * reference's expressions. */
int atlj_sim_run_baoab(atlj_sim *s, double dt, double o_decay, double o_noise, uint64_t seed, int nsteps,
                       double *rec_host);
/* Nose-Hoover chain steps (md_nvt_lj.py:101-159,270-288), chain length m <= 8, g = degrees of freedom
 * (md_nvt_lj.py:244); eta, p_eta host in/out, q host in.  rec[10] = the thermostat part of the conserved energy
 * (md_nvt_lj.py:44). */
int atlj_sim_run_nhc(atlj_sim *s, double dt, double temperature, double g, int m, double *eta, double *p_eta,
                     const double *q, int nsteps, double *rec_host);
/* Constant-NPT steps (md_npt_lj.py:101-214 propagators, :344-366 loop): Nose-Hoover chains on the particles and on
 * the barostat, chain length m <= 8, box = box0*exp(eps) updated every step (md_npt_lj.py:125-126) - the library
 * re-derives the force path's scalars and the cell grid sc = floor(box/r_cut) from the new box each step
 * (md_lj_ll_module.py:106; link_list_module.f90:112-120 regrows its cell array the same way).  eta, p_eta, eta_baro,
 * p_eta_baro, eps, p_eps host in/out; q, q_baro host in.  The caller has run atlj_sim_force on the start state
 * (md_npt_lj.py:335).  rec[9] = box after the step, rec[10] = the extended-system part of the conserved quantity
 * (md_npt_lj.py:44-45).  Single domain, LJ. */
int atlj_sim_run_npt(atlj_sim *s, double dt, double temperature, double pressure, double r_cut, double g, int m,
                     double *eta, double *p_eta, const double *q, double *eta_baro, double *p_eta_baro,
                     const double *q_baro, double w_eps, double box0, double *eps, double *p_eps, int nsteps,
                     double *rec_host);
/* Isokinetic SLLOD steps (md_nvt_lj_le.py:84-129,226-240); *strain host in/out. */
int atlj_sim_run_sllod(atlj_sim *s, double dt, double strain_rate, double *strain, int nsteps, double *rec_host);

/* BAOAB with an injected noise array (nsteps, n, 3) indexed by GLOBAL atom id -- deterministic parity
 * tests against bd_nvt_lj.py:127 with the same numbers. */
int atlj_sim_run_baoab_noise(atlj_sim *s, double dt, double o_decay, double o_noise, const double *noise_host,
                             int nsteps, double *rec_host);

/* ---- slab domain decomposition over the GPUs of one box (one rank per GPU; no reference counterpart, the yardstick
 * is the single-domain result).  A rank is a sim (created with ATLJ_SIM_IPC when its neighbours live in other
 * processes) configured with a proper sub-range [x_lo, x_hi) of x cell layers.  Its per-atom arrays hold
 * [owned | left halo layer | right halo layer]; the halo regions of the position buffers and the halo rows of the cell
 * table are written by the NEIGHBOURS' pack kernels over NVLink (peer memory mapped once with CUDA IPC) and published
 * with a system-scope step stamp that the pair kernel acquires when it reaches a boundary layer; migrating atoms
 * (r, v, id) go through a small message arena the same way.  The per-step records are all-reduced once per run
 * through NCCL.  mig_cap = atoms per migration message, halo_cap = atoms per halo layer; all ranks must use the same
 * capacity, mig_cap and halo_cap. ---------- */
int atlj_sim_mg_enable(atlj_sim *s, int64_t mig_cap, int64_t halo_cap);
void atlj_sim_set_total(atlj_sim *s, int64_t n_total); /* atoms of the whole system (RNG / noise indexing) */
/* What a neighbour in another process needs to write into this rank's memory: atlj_sim_mg_blob_bytes() bytes (CUDA IPC
 * handles of the arena, the two position buffers and the cell table + the layout), exchanged by any means ... */
int64_t atlj_sim_mg_blob_bytes(void);
int atlj_sim_mg_ipc_blob(atlj_sim *s, char *out);
int atlj_sim_mg_ipc_open(atlj_sim *s, const char *left_blob, const char *right_blob);
/* ... or the neighbours' sims themselves when they live in the same process (ranks emulated on one GPU). */
int atlj_sim_mg_set_peers(atlj_sim *s, atlj_sim *left, atlj_sim *right);
/* Diagnostics (environment ATLJ_MG_TRACE=1 when the slab is enabled): %globaltimer marks (ns) between the launches of
 * the last nsteps exchange steps, out[nsteps][8]: before / after the integration kernel, after the arrivals, the scan,
 * the sort + gather, the pair kernel (tools/slab_trace.py). */
int atlj_sim_mg_trace(atlj_sim *s, int nsteps, unsigned long long *out);
/* NCCL communicator over all ranks: rank 0 obtains the id, the caller broadcasts the 128 bytes by any means. */
int atlj_nccl_unique_id(char *out128);
int atlj_sim_mg_connect(atlj_sim *s, int rank, int world, const char *id128);
/* nsteps velocity-Verlet steps (md_nve_lj.py:178-192) of the whole system, no host synchronisation per step.
 * rec_host (nsteps, ATLJ_NREC) holds the GLOBAL sums on every rank.  dt = 0 gives a plain force evaluation. */
int atlj_sim_run_nve_mg(atlj_sim *s, double dt, int nsteps, double *rec_host);
/* The same with the hessian sum of md_nve_lj.py:46 (hessian(box, r_cut, r, f), md_lj_ll_module.py:227-357) in rec[6] of
 * every step: after the force kernel each rank stores the forces of its two boundary layers into the neighbours' force
 * arrays (the halo atoms' forces) and the hessian kernel acquires that second stamp at its boundary layers. */
int atlj_sim_run_nve_mg_hess(atlj_sim *s, double dt, int nsteps, double *rec_host);
/* ... on ranks emulated in one process: phase 3 in two halves (force + boundary forces sent | hessian + trailing kick) */
int atlj_sim_mg_phase3_force(atlj_sim *s);
int atlj_sim_mg_phase4_hess(atlj_sim *s, double dt, double *rec_host);
/* The thermostats on slab ranks.  BAOAB (bd_nvt_lj.py:221-233): the noise is keyed by (seed; global atom id, step), so the
 * trajectory does not depend on the decomposition.  Nose-Hoover chain (md_nvt_lj.py:270-288): the chain needs the global
 * sum v^2 once per step - the 12-double step record is all-reduced (NCCL) inside the step; eta, p_eta as in
 * atlj_sim_run_nhc, identical on every rank.  Records are global sums on every rank. */
int atlj_sim_run_baoab_mg(atlj_sim *s, double dt, double o_decay, double o_noise, uint64_t seed, int nsteps,
                          double *rec_host);
int atlj_sim_run_nhc_mg(atlj_sim *s, double dt, double temperature, double g, int m, double *eta, double *p_eta,
                        const double *q, int nsteps, double *rec_host);
/* The same step cut into its three device phases, for driving several ranks of one process in lockstep (tests):
 * phase1 = kick, drift, leavers written to the neighbours; phase2 = append arrivals, bin + sort, boundary layers
 * written into the neighbours' halo regions; phase3 = pair kernel, kick, LOCAL sums.  Every rank must finish a phase
 * before any rank starts the next one. */
int atlj_sim_mg_phase1(atlj_sim *s, double dt);
int atlj_sim_mg_phase2(atlj_sim *s);
int atlj_sim_mg_phase3(atlj_sim *s, double dt, double *rec_host);
/* ... and phase 1 of the thermostatted steps (phase 2 and 3 are the same).  Nose-Hoover in lockstep emulation:
 * atlj_sim_mg_nhc_open (chain state + the GLOBAL sum v^2 of the start velocities) once, then per step phase1_nhc,
 * phase2, phase3 on every rank, the caller sums the records over the ranks and gives the global record to
 * atlj_sim_mg_nhc_close on every rank (chain update; rec[4], rec[10] filled in); atlj_sim_mg_nhc_finish at the end. */
int atlj_sim_mg_phase1_baoab(atlj_sim *s, double dt, double o_decay, double o_noise, uint64_t seed, int first);
int atlj_sim_mg_nhc_open(atlj_sim *s, int m, const double *eta, const double *p_eta, const double *q, double ksq_global);
int atlj_sim_mg_phase1_nhc(atlj_sim *s, double dt, double temperature, double g, int m, int first);
int atlj_sim_mg_nhc_close(atlj_sim *s, double dt, double temperature, double g, int m, double *rec_global_host);
int atlj_sim_mg_nhc_finish(atlj_sim *s, int m, double *eta, double *p_eta);

/* ---- host-buffer step (the end-to-end path of bench.py) --------------------------------------
 * One velocity-Verlet step (md_nve_lj.py:178-192) whose state lives in HOST memory, as it does in the
 * reference driver: r, v, f, id (n atoms, any order, f = forces of r) are copied to the device, the step
 * runs there, and r, v, f, id come back (in cell order; id says which atom is which) with rec[ATLJ_NREC].
 * Pinned buffers from atlj_host_alloc make the copies asynchronous DMA. */
int atlj_sim_nve_step_host(atlj_sim *s, double dt, double *r_host, double *v_host, double *f_host, int32_t *id_host,
                           int64_t n, double *rec_host);
/* The same on a slab rank (the end-to-end path at N > 1 GPUs): *n owned atoms of THIS rank in host memory (device
 * order of the previous call / atlj_sim_download), one exchange step of the whole system (migration, halo, force,
 * NCCL all-reduce of the record), then the rank's owned atoms - *n updated: atoms migrate - come back.  The host
 * arrays must hold the rank's owned capacity (atlj_sim_mg_own_cap).  rec_host: GLOBAL sums, as atlj_sim_run_nve_mg. */
int atlj_sim_nve_step_host_mg(atlj_sim *s, double dt, double *r_host, double *v_host, double *f_host, int32_t *id_host,
                              int64_t *n, double *rec_host);
int64_t atlj_sim_mg_own_cap(const atlj_sim *s);
void *atlj_host_alloc(int64_t bytes); /* page-locked host memory (cudaHostAlloc); NULL on failure */
void atlj_host_free(void *p);

/* ---- configuration files (config_io_module.py:30-48 read_cnf_atoms, :74-95 write_cnf_atoms) ----------------
 * The body of a cnf file is np.savetxt(fmt='%15.10f'): n*cols fields of 15 characters + one separator (' ' between
 * columns, '\n' after the last), cols = 3 (r) or 6 (r, v).  Both directions run on the GPU, one thread per field,
 * bit-exact: format = glibc's correctly rounded '%.10f' (128-bit integer arithmetic on the exact binary value),
 * parse = the correctly rounded value of the decimal string (what np.loadtxt returns).  The two header lines
 * ("{:15d}\n{:15.8f}", n and box) are the caller's.  ATLJ_ERR_ARG (with a message) for values that do not fit a
 * 15-character field, NaN/Inf, or a body that is not in this fixed layout. */
/* r_host (n,3) [times scale_r, separately rounded like the drivers' r*box], v_host (n,3) or NULL ->
 * out_host[n*cols*16] */
int atlj_cnf_format(const double *r_host, const double *v_host, int64_t n, double scale_r, char *out_host);
/* text_host[n*cols*16] -> r_host (n,3), v_host (n,3) or NULL (then columns >= 3 are checked but not returned) */
int atlj_cnf_parse(const char *text_host, int64_t n, int cols, double *r_host, double *v_host);
/* The device-resident state as a cnf body in the ORIGINAL atom order: positions times box (md_nve_lj.py:196 writes
 * r*box), velocities if with_v; the un-sort by atom id is fused into the formatting kernel. */
int atlj_sim_cnf_format(atlj_sim *s, int with_v, char *out_host);

/* ---- measurement helpers ------------------------------------------------------------------ */
/* CUDA-event stopwatch on the sim's stream: start records an event, stop records a second one, waits for it
 * and returns the elapsed device milliseconds. */
int atlj_sim_timer_start(atlj_sim *s);
int atlj_sim_timer_stop(atlj_sim *s, double *ms);
int atlj_sim_sync(atlj_sim *s); /* cudaStreamSynchronize on the sim's stream */
/* Register-resident DFMA loop on every SM: returns achieved fp64 FLOP/s (2 flop per DFMA) -- the
 * roofline denominator for the pair kernel (no fp64 figure in MEASURED_PEAKS.json). */
int atlj_measure_fp64_peak(double *flops_per_s);
/* Device time (ms) of the most recent pair-kernel / cell-build / integration launches, by CUDA events
 * recorded on the sim's stream when timing is enabled. */
int atlj_sim_enable_timing(atlj_sim *s, int on);
/* accumulated ms, then reset: {pair, cells, integrate, other, migration exchange, halo pack, halo exchange + unpack
 * (second stream), boundary-layer pair kernel (second stream)}; the last four are used by the slab path only */
int atlj_sim_timing(atlj_sim *s, double ms[8]);

#ifdef __cplusplus
}
#endif
#endif /* ATLJ_H */
