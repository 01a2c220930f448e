/*
 * mcs_b200.h -- C ABI of libmcs_b200.so: B200-native annealing sweeps behind the
 * MonteCarloSolvers call surface (dtoconnor/MonteCarloSolvers; paths below are relative to
 * the reference tree).
 *
 * The reference exposes the hot path as module-level Cython `cpdef` functions taking NumPy
 * buffers; it has no C header of its own.  Each entry point here names the
 * reference function (file:line) whose loop nest it replaces.  The Python package
 * `montecarlosolvers_b200` binds these with ctypes and reproduces the reference's positional
 * signatures on top (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns 0 on success or a negative
 *     MCS_E* code, with a human-readable message available from mcs_last_error().
 *   - "host" pointers are ordinary (ideally pinned, see mcs_host_alloc) CPU memory; copies to
 *     and from the device happen inside the call.  `mcs_state` keeps a replica batch resident
 *     in HBM between calls for callers that want to avoid the copies.
 *   - spins are int8 (+1 / -1).  Replica batches are C-ordered [R][N][P] (replica, site,
 *     Trotter slice) for PIQMC and [R][N] for SA / SVMC; R = 1 is the reference's single call.
 *   - the neighbour table is the reference's own format: float64 [N][maxnb][2] =
 *     (neighbour index stored as a float, coupling), rows zero padded, a self entry is a local
 *     field (tools.pyx:28-96; consumed at qmc.pyx:114-125, sa.pyx:84-94, svmc.pyx:98-108).
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *     MCS_ENODEVICE.
 *   - threads: the one-shot host-buffer calls (mcs_*_anneal*, mcs_exact_*) may be issued from several threads
 *     (the reference released the GIL in its loops, qmc.pyx:92, sa.pyx:65); calls on the SAME instance are
 *     serialised by a lock inside the instance (they share its scratch batches, staging buffer and stream),
 *     calls on different instances run concurrently.  Calls on an mcs_state are the caller's to order: one
 *     thread at a time per state and per instance.
 */
#ifndef MCS_B200_H
#define MCS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCS_ABI_VERSION 1

enum {
    MCS_OK = 0,
    MCS_EINVAL = -1,     /* bad argument (shape, range, NULL)                               */
    MCS_ENODEVICE = -2,  /* no usable CUDA device / CUDA runtime error (see mcs_last_error) */
    MCS_EZERODIV = -3,   /* temp * P == 0: the reference raises ZeroDivisionError here
                            (qmc.c:3030-3034; qmc.pyx has no cdivision)                     */
    MCS_EUNSUPPORTED = -4, /* valid in the reference but outside this build's kernels      */
    MCS_ENOMEM = -5
};

typedef struct mcs_instance mcs_instance; /* compiled Ising instance resident on one GPU   */
typedef struct mcs_state mcs_state;       /* replica batch resident on the same GPU        */

/* ---- library ------------------------------------------------------------------------- */
int mcs_abi_version(void);
const char *mcs_last_error(void);   /* thread-local message of the last failing call       */
int mcs_device_count(void);         /* 0 when no CUDA device is visible                    */

/* pinned host memory for the batched entry points (pageable memory works, but slower)      */
void *mcs_host_alloc(size_t bytes);
void mcs_host_free(void *p);

/* ---- instance compiler ---------------------------------------------------------------
 * Consumes the table tools.GenerateNeighbors produces (tools.pyx:28-96): drops the zero
 * padding, splits self entries into fields, graph-colours the interaction graph
 * (2-colouring when bipartite -- the Santoro torus and Chimera are -- else greedy), sorts
 * sites by colour and uploads fp32 ELL tables (production kernels) and the fp64 table in
 * the reference's row order (exact / probe kernels).                                        */
int mcs_instance_create(const double *nbs, int64_t nspins, int64_t maxnb, int device,
                        mcs_instance **out);
/* Time-dependent couplings, sa.NoisyAnneal / svmc.NoisySVMC[TF] (sa.pyx:291-378, svmc.pyx:236-448):
 * nbs is float64 [nsteps][N][maxnb][2]; schedule step f of a later *_sweeps / exact call uses table f
 * (the schedule may not be longer than nsteps).  Energies are evaluated with the LAST table.      */
int mcs_instance_create_steps(const double *nbs, int64_t nsteps, int64_t nspins, int64_t maxnb, int device,
                              mcs_instance **out);
void mcs_instance_destroy(mcs_instance *inst);
/* info[0]=nspins info[1]=maxnb info[2]=ncolors info[3]=max degree (fields excluded)
 * info[4]=1 if any local field  info[5]=device  info[6]=1 if the threshold-table PIQMC
 * kernels apply (degree + field <= 8)
 * info[7]=number of tables (1 unless created by mcs_instance_create_steps); bit 32 set when
 *         the instance is dense (max degree >= 48): sweeps then run as blocked tensor-core GEMM
 *         + in-block sequential updates instead of one colour class per site                */
int mcs_instance_info(const mcs_instance *inst, int64_t info[8]);
int mcs_instance_colors(const mcs_instance *inst, int32_t *color /* [nspins] */);
/* dense instances only: enable = 0 routes sweeps through the general coloured kernels instead of the
 * blocked tensor-core path (used to cross-check the two)                                      */
int mcs_instance_set_dense(mcs_instance *inst, int enable);
/* Which dynamics mcs_piqmc_sweeps / mcs_sa_sweeps and the one-shot *_anneal calls run on this instance:
 *   MCS_DYN_COLORED   (default) colour class by colour class, Trotter parity by parity -- the fastest order;
 *                     same Boltzmann distribution, but an anneal relaxes slightly faster per sweep than the
 *                     reference's (residual energies 6-9 % lower on santoro_80x80);
 *   MCS_DYN_REFERENCE the reference's own order in distribution: per (replica, sweep, slice) a fresh uniformly
 *                     random visiting permutation, strictly sequential visits, slices in order (qmc.pyx:99-143,
 *                     sa.pyx:73-99), then the world-line moves in one more permutation (qmc.pyx:405-438).  Executed
 *                     as dependency waves of the orientation the permutation induces on the interaction graph, one
 *                     CTA per replica with the world lines in shared memory (needs N (P > 32 ? 15 : 11) bytes
 *                     <= 227 KB).  Not available for dense instances and the Ohmic-bath sweeps.               */
enum { MCS_DYN_COLORED = 0, MCS_DYN_REFERENCE = 1 };
int mcs_instance_set_dynamics(mcs_instance *inst, int dynamics);
/* CUDA-event stopwatch on the instance's stream: start ... stop returns milliseconds.      */
int mcs_timer_start(mcs_instance *inst);
int mcs_timer_stop(mcs_instance *inst, double *ms);
int mcs_synchronize(mcs_instance *inst);
/* The one-shot host-buffer calls below keep their device batch + staging buffer cached in the
 * instance between calls of the same shape; mcs_instance_trim releases that memory.          */
int mcs_instance_trim(mcs_instance *inst);
/* number of kernels this instance has launched so far (bench.py's gpu_launches)            */
int64_t mcs_launch_count(const mcs_instance *inst);

/* ---- resident replica batches -------------------------------------------------------- */
enum { MCS_KIND_PIQMC = 1, MCS_KIND_SA = 2, MCS_KIND_SVMC = 3 };
/* P is the number of Trotter slices (PIQMC, 2 <= P <= 64) and must be 1 for SA / SVMC.     */
int mcs_state_create(mcs_instance *inst, int kind, int64_t R, int64_t P, mcs_state **out);
void mcs_state_destroy(mcs_state *st);
/* spins: int8 [R][N][P] (PIQMC) or [R][N] (SA); angles: float64 [R][N] (SVMC)              */
int mcs_state_upload_spins(mcs_state *st, const int8_t *host);
int mcs_state_download_spins(mcs_state *st, int8_t *host);
int mcs_state_upload_angles(mcs_state *st, const double *host);
int mcs_state_download_angles(mcs_state *st, double *host);
/* replica r starts from Philox(seed, replica_offset + r) spins, identical across slices
 * (SURVEY.md 8d cfg3); SVMC states start at theta = pi/2.                                   */
int mcs_state_init_random(mcs_state *st, uint64_t seed, uint64_t replica_offset);
/* fixed-order fp64 classical energies, bit-identical to the oracle's definition of
 * tools.ClassicalIsingEnergy (tools.pyx:99-118): out is [R][P] (PIQMC) or [R] (SA).        */
int mcs_state_energies(mcs_state *st, double *host_out);
/* Best Trotter slice of every anneal, evaluated on the device: per-slice energies as above -> arg-min over the
 * slices (first minimum) -> that slice's spins.  This is the post-processing of the reference's example
 * (santoro80.py:290-296: E = min over slices of ClassicalIsingEnergy(confs[:, k])) without moving the world lines:
 * best_energy float64 [R], best_slice int32 [R], best_conf int8 [R][N]; any of them may be NULL.
 * on_device != 0: the pointers are DEVICE pointers on the instance's GPU (optional tensor handoff: e.g.
 * torch.Tensor.data_ptr()); the work is queued on the instance's stream, mcs_synchronize() waits for it.        */
int mcs_state_best(mcs_state *st, double *best_energy, int32_t *best_slice, int8_t *best_conf, int on_device);
/* SVMC energy B*(sum J cos cos + sum h cos) - A*sum sin per replica, out [R]               */
int mcs_state_svmc_energies(mcs_state *st, double a, double b, double *host_out);

/* ---- production sweeps (coloured, Philox-driven; statistical parity) ------------------
 * mcs_piqmc_sweeps replaces the loop nest of qmc.QuantumAnneal (qmc.pyx:93-143) and, with
 * global_moves != 0, of qmc.QuantumAnnealGlobal (qmc.pyx:358-438): for every schedule step
 * `mcsteps` sweeps; a sweep attempts every (site, slice) once, colour class by colour class
 * and Trotter parity by parity, with J_perp = -(P*T/2) ln tanh(A/(P*T)) (qmc.pyx:95) and the
 * Metropolis rule of qmc.pyx:140-143.  `temp` is a C float as in the reference (qmc.pyx:28).
 * sweep_offset numbers the first sweep for the counter-based RNG so that a schedule may be
 * split over several calls (checkpoint / resume = slicing the schedule).                    */
int mcs_piqmc_sweeps(mcs_state *st, const double *A_sched, const double *B_sched, int64_t schedsize,
                     int mcsteps, float temp, int global_moves, uint64_t seed,
                     uint64_t replica_offset, uint64_t sweep_offset);
/* qmc.DissipativeQuantumAnneal[Global] (qmc.pyx:223-278, 523-609): as above plus the Ohmic-bath
 * term sum_{d=1}^{P-1} 2 teff s_k s_{k+d} lookuptable[d-1] (qmc.pyx:268-273), lookuptable float64
 * [P-1].  The bath couples all slices of a world line, so slices are visited in order inside a
 * word; colour classes and replicas stay parallel (coupling planes in registers up to degree +
 * field = 8, a row-walking variant beyond).                                                  */
int mcs_piqmc_sweeps_dissipative(mcs_state *st, const double *A_sched, const double *B_sched,
                                 int64_t schedsize, int mcsteps, float temp, const double *lookuptable,
                                 int global_moves, uint64_t seed, uint64_t replica_offset,
                                 uint64_t sweep_offset);
/* sa.Anneal (sa.pyx:66-101): Metropolis sweeps at temperature sched[itemp]                 */
int mcs_sa_sweeps(mcs_state *st, const double *sched, int64_t schedsize, int mcsteps, uint64_t seed,
                  uint64_t replica_offset, uint64_t sweep_offset);
/* svmc.SpinVectorMonteCarlo (svmc.pyx:78-117); tf != 0: SpinVectorMonteCarloTF (:181-229).
 * replica_offset must be a multiple of 4 (one Philox call serves four consecutive replicas). */
int mcs_svmc_sweeps(mcs_state *st, const double *A_sched, const double *B_sched, int64_t schedsize,
                    int mcsteps, float temp, int tf, uint64_t seed, uint64_t replica_offset,
                    uint64_t sweep_offset);

/* Swendsen-Wang cluster moves on the (space x Trotter) lattice, GPU union-find labelling.  The reference
 * advertises cluster updates (README.md:4) but ships only experimental Wolff variants (qmc.pyx:620-1621)
 * that raise on Linux and no Swendsen-Wang: there is no trajectory oracle, parity is equilibrium
 * statistics (tests/test_gpu_cluster.py).  PIQMC state: bonds -B J_ij/teff in every slice, J_perp/teff
 * along the Trotter ring, fields as bonds to a fixed ghost spin; (a, b) = (Gamma, B) of qmc.pyx:95-96.
 * SA state: a, b ignored, couplings -J/temp.  Needs a symmetric static table.                     */
int mcs_cluster_moves(mcs_state *st, double a, double b, float temp, int nmoves, uint64_t seed,
                      uint64_t replica_offset, uint64_t sweep_offset);
/* The same move for the Ohmic-bath action of the Dissipative solvers (qmc.pyx:268-273): every pair of slices of a
 * world line at ring distance d carries the additional bond K_d = lookuptable[d-1] (in units of teff).  The bond
 * graph the reference's DissaptiveQuantumAnnealWCL / WC2 / WC3 grow their clusters on (qmc.pyx:906-925, 1400-1437,
 * 1598-1610), sampled correctly.  lookuptable float64 [P-1] must be symmetric (table[d-1] == table[P-d-1], as the
 * documented kernel (pi / (P sin(pi d / P)))^2 is): otherwise it does not define an energy -> MCS_EUNSUPPORTED.  */
int mcs_cluster_moves_dissipative(mcs_state *st, double a, double b, float temp, const double *lookuptable,
                                  int nmoves, uint64_t seed, uint64_t replica_offset, uint64_t sweep_offset);

/* one-shot host-buffer forms (upload, sweep, download [, energies]) -- what the Python
 * drop-ins call.  energies_out may be NULL.  mcs_piqmc_anneal with R >= 1024 cuts the batch
 * into replica windows and overlaps the H2D copy / sweeps / D2H copy of neighbouring windows;
 * page-locked host buffers (mcs_host_alloc) are needed for the overlap, pageable ones work.  */
int mcs_piqmc_anneal(mcs_instance *inst, const double *A_sched, const double *B_sched,
                     int64_t schedsize, int mcsteps, float temp, int8_t *confs /* [R][N][P] */,
                     int64_t R, int64_t P, int global_moves, uint64_t seed, uint64_t replica_offset,
                     double *energies_out /* [R][P] */);
/* The example's per-anneal protocol in one call (santoro80.py:286-296): confs = tile(state, P) -> QuantumAnneal
 * [Global] -> best slice.  spins_in is int8 [R][N] when input_tiled != 0 (one start configuration per anneal, copied
 * to all P slices on the device) or the full [R][N][P] otherwise; outputs as mcs_state_best plus the per-slice
 * energies [R][P]; any output may be NULL.  Host traffic R N bytes in and R (N + 8 P + 12) out, not 2 R N P.      */
int mcs_piqmc_anneal_best(mcs_instance *inst, const double *A_sched, const double *B_sched, int64_t schedsize,
                          int mcsteps, float temp, const int8_t *spins_in, int input_tiled, int64_t R, int64_t P,
                          int global_moves, uint64_t seed, uint64_t replica_offset,
                          double *energies_out /* [R][P] */, double *best_energy /* [R] */,
                          int32_t *best_slice /* [R] */, int8_t *best_conf /* [R][N] */);
int mcs_sa_anneal(mcs_instance *inst, const double *sched, int64_t schedsize, int mcsteps,
                  int8_t *svec /* [R][N] */, int64_t R, uint64_t seed, uint64_t replica_offset,
                  double *energies_out /* [R] */);
int mcs_svmc_anneal(mcs_instance *inst, const double *A_sched, const double *B_sched, int64_t schedsize,
                    int mcsteps, float temp, double *svec /* [R][N] */, int64_t R, int tf, uint64_t seed,
                    uint64_t replica_offset);

/* ---- exact (sequential-order) kernels: bit-exact replay of the reference ---------------
 * Each replica runs the reference's own visiting order in fp64: Fisher-Yates shuffle from a glibc
 * rand() stream (qmc.pyx:102-108), sequential Metropolis visits in table order without FMA
 * contraction.  mcs_exact_qmc / mcs_exact_sa give a replica one WARP (state in shared memory): the
 * rand() stream is produced 31 values per step, and runs of shuffle iterations / visits that cannot
 * influence one another (no shared position / no two adjacent sites) are executed at once -- the
 * trajectory is the sequential one bit for bit (7.8e9 attempts/s at 80x80, P = 64, 4096 replicas).
 * Instances that do not fit (N > 65535, rows longer than 64 entries, P > 64) and the SVMC / Wolff
 * replays use one thread per replica.  libc_seeds[r] plays the role of srand(seed) before the
 * r-th reference call; if rand_stream != NULL it is used instead (int32 [R][stream_len] of
 * recorded rand() outputs) and consumed[r] returns how many values replica r used.
 * lookuptable != NULL selects the Dissipative variants (qmc.pyx:149-278, 444-609).          */
int mcs_exact_qmc(mcs_instance *inst, const double *A_sched, const double *B_sched, int64_t schedsize,
                  int mcsteps, float temp, const double *lookuptable, int8_t *confs /* [R][N][P] */,
                  int64_t R, int64_t P, int global_moves, const uint32_t *libc_seeds /* [R] */,
                  const int32_t *rand_stream, int64_t stream_len, int64_t *consumed /* [R] or NULL */);
/* The reference's Wolff-cluster experiments (qmc.pyx:612-1621, "Function under test"), replayed as written:
 *   MCS_WOLFF_WCL       qmc.QuantumAnnealWCL             (qmc.pyx:620-786)    one single-cluster move per step,
 *   MCS_WOLFF_DISS_WCL  qmc.DissaptiveQuantumAnnealWCL   (qmc.pyx:792-1000)   same with Ohmic-bath bonds,
 *   MCS_WOLFF_WC        qmc.QuantumAnnealWC              (qmc.pyx:1006-1225)  bond test on the full energy change,
 *   MCS_WOLFF_DISS_WC2  qmc.DissipativeQuantumAnnealWC2  (qmc.pyx:1231-1446)  local sweep + N bath clusters per step,
 *   MCS_WOLFF_DISS_WC3  qmc.DissipativeQuantumAnnealWC3  (qmc.pyx:1452-1621)  N P bath clusters per step.
 * Single-cluster growth from an explicit stack is sequential, so these exist only as replay: one thread per
 * replica, glibc rand() stream of libc_seeds[r].  As shipped the reference functions raise on Linux (np.intc
 * scratch typed np.int_t); with that dtype corrected they run, and this call reproduces them bit for bit.
 * lookuptable float64 [P-1] for the DISS variants (else NULL).  overrun[r] (may be NULL) is set when replica r made
 * the reference write past its unchecked `cluster` buffer -- undefined behaviour there, well defined here.       */
enum { MCS_WOLFF_WCL = 0, MCS_WOLFF_DISS_WCL = 1, MCS_WOLFF_WC = 2, MCS_WOLFF_DISS_WC2 = 3, MCS_WOLFF_DISS_WC3 = 4 };
int mcs_exact_qmc_wolff(mcs_instance *inst, int variant, const double *A_sched, const double *B_sched,
                        int64_t schedsize, int mcsteps, float temp, const double *lookuptable,
                        int8_t *confs /* [R][N][P] */, int64_t R, int64_t P, const uint32_t *libc_seeds /* [R] */,
                        int64_t *consumed /* [R] or NULL */, int32_t *overrun /* [R] or NULL */);
/* sa.Anneal / Anneal_parallel (sa.pyx:19-101, 201-284); randuni != NULL: sa.AnnealMA
 * (sa.pyx:108-193), float64 [schedsize][mcsteps][N] shared by all replicas                  */
int mcs_exact_sa(mcs_instance *inst, const double *sched, int64_t schedsize, int mcsteps,
                 int8_t *svec /* [R][N] */, int64_t R, const uint32_t *libc_seeds, const double *randuni,
                 int64_t *consumed);
/* svmc.SpinVectorMonteCarlo[TF] (svmc.pyx:21-229) and ...Compact (:455-554): randuni is the
 * float64 [schedsize][mcsteps][N][2] array the reference draws from np.random up front
 * (svmc.pyx:70), shared by all replicas.  serial_stream != 0 reproduces the Compact form's single
 * rand() stream running through the reads one after another (libc_seeds[0] only); randuni ==
 * NULL with tf != 0 is SpinVectorMonteCarloTFCompact (:561-674, all uniforms from rand()).  */
int mcs_exact_svmc(mcs_instance *inst, const double *A_sched, const double *B_sched, int64_t schedsize,
                   int mcsteps, float temp, double *svec /* [R][N] */, int64_t R, int tf,
                   const uint32_t *libc_seeds, const double *randuni, int serial_stream);

/* ---- probes (parity tier a): fp64, reference association order, no FMA ---------------- */
/* ediff of every (site, slice) visit for frozen configurations (qmc.pyx:112-138);
 * out float64 [R][N][P]                                                                     */
int mcs_probe_qmc_delta_e(mcs_instance *inst, double a, double b, float temp, const int8_t *confs,
                          int64_t R, int64_t P, double *out);
/* world-line flip ediff (qmc.pyx:416-431); out float64 [R][N]                              */
int mcs_probe_qmc_delta_e_global(mcs_instance *inst, double b, const int8_t *confs, int64_t R,
                                 int64_t P, double *out);
/* sa.pyx:84-94; out float64 [R][N]                                                         */
int mcs_probe_sa_delta_e(mcs_instance *inst, const int8_t *svec, int64_t R, double *out);

#ifdef __cplusplus
}
#endif
#endif /* MCS_B200_H */
