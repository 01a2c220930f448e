/*
 * mcs_oracle.c -- CPU restatement of the MonteCarloSolvers annealing-sweep hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (montecarlosolvers_b200/, the C-ABI
 * library) may include, link or call this file.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker / the CPU arm.
 *
 * Parity status: PINNED.  The reference has no tests or golden vectors of its own
 * (SURVEY.md section 4); this restatement is pinned against the reference itself, compiled in the
 * build container by oracle/build_ref.py into oracle/_ref/ (tests/test_oracle_vs_reference.py),
 * and against fixtures generated from that compiled reference and committed under
 * tests/golden/ (tests/golden/make_golden.py).
 *
 * Every function cites the reference lines (relative to /root/reference/) it follows.
 * Arithmetic is IEEE fp64 in the reference's exact association order; compile with
 * -O2 -ffp-contract=off (no FMA), link libm (the reference calls libc exp/log/tanh/sin/cos).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * glibc rand() clone (TYPE_3 additive feedback generator, degree 31, separation 3).
 * The reference draws its shuffles and acceptance uniforms from the process-global libc
 * rand() (qmc.pyx:18-19, sa.pyx:11-12, svmc.pyx:13-14); it never calls srand(), so an
 * unseeded process behaves as srand(1).  The clone makes the oracle independent of the libc
 * on the box and lets many independent streams run side by side.
 * -------------------------------------------------------------------------------------- */
typedef struct {
    int32_t r[31];
    int f; /* front index */
    int b; /* rear index  */
} mcs_rand_t;

void mcs_srand(mcs_rand_t *st, uint32_t seed)
{
    int i;
    int32_t word;
    if (seed == 0) seed = 1;
    st->r[0] = (int32_t)seed;
    word = (int32_t)seed;
    for (i = 1; i < 31; ++i) {
        /* 16807 * word % 2147483647 without overflow (Schrage) */
        long hi = word / 127773;
        long lo = word % 127773;
        long w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        word = (int32_t)w;
        st->r[i] = word;
    }
    st->f = 3;
    st->b = 0;
    for (i = 0; i < 310; ++i) {
        st->r[st->f] = (int32_t)((uint32_t)st->r[st->f] + (uint32_t)st->r[st->b]);
        if (++st->f >= 31) st->f = 0;
        if (++st->b >= 31) st->b = 0;
    }
}

static inline int32_t mcs_rand(mcs_rand_t *st)
{
    uint32_t v = (uint32_t)st->r[st->f] + (uint32_t)st->r[st->b];
    st->r[st->f] = (int32_t)v;
    if (++st->f >= 31) st->f = 0;
    if (++st->b >= 31) st->b = 0;
    return (int32_t)(v >> 1);
}

#define MCS_RAND_MAX 2147483647 /* glibc RAND_MAX */

/* fill out[n] with the next n outputs (exported for tests and for feeding the GPU replay) */
void mcs_rand_fill(mcs_rand_t *st, int32_t *out, int64_t n)
{
    int64_t i;
    for (i = 0; i < n; ++i) out[i] = mcs_rand(st);
}

size_t mcs_rand_sizeof(void) { return sizeof(mcs_rand_t); }

/* Fisher-Yates exactly as qmc.pyx:102-108 / sa.pyx:73-79 / svmc.pyx:84-90:
 * identity, then for i = n..1: j = rand() % i; swap(p[i-1], p[j]).                         */
static inline void shuffle(mcs_rand_t *rng, int64_t *p, int n)
{
    int i;
    for (i = 0; i < n; ++i) p[i] = i;
    for (i = n; i > 0; --i) {
        int j = mcs_rand(rng) % i;
        int64_t t = p[i - 1];
        p[i - 1] = p[j];
        p[j] = t;
    }
}

#define NB_IDX(nbs, maxnb, i, si) ((int)(nbs)[((int64_t)(i) * (maxnb) + (si)) * 2])
#define NB_J(nbs, maxnb, i, si) ((nbs)[((int64_t)(i) * (maxnb) + (si)) * 2 + 1])

/* ------------------------------------------------------------------------------------------
 * In-plane part of the PIQMC energy difference, qmc.pyx:112-125: accumulate over the padded
 * neighbour row in table order; self entry is the linear (field) term.
 * confs is [nspins, slices] with element strides (cs0, cs1) like the reference memoryview.
 * -------------------------------------------------------------------------------------- */
static inline double qmc_inplane(const int64_t *confs, int64_t cs0, int64_t cs1, const double *nbs,
                                 int maxnb, int ispin, int islice, double b_coeff, double acc)
{
    int si;
    double s = (double)confs[ispin * cs0 + islice * cs1];
    for (si = 0; si < maxnb; ++si) {
        int spinidx = NB_IDX(nbs, maxnb, ispin, si);
        double jval = NB_J(nbs, maxnb, ispin, si);
        if (spinidx == ispin)
            acc += b_coeff * s * jval;
        else
            acc += b_coeff * s * (jval * (double)confs[spinidx * cs0 + islice * cs1]);
    }
    return acc;
}

/* full local-move energy difference of (ispin, islice): qmc.pyx:112-138 */
static inline double qmc_ediff(const int64_t *confs, int64_t cs0, int64_t cs1, const double *nbs,
                               int maxnb, int slices, int ispin, int islice, double b_coeff,
                               double jperp)
{
    int tleft, tright;
    double s = (double)confs[ispin * cs0 + islice * cs1];
    double e = qmc_inplane(confs, cs0, cs1, nbs, maxnb, ispin, islice, b_coeff, 0.0);
    if (islice == 0) {
        tleft = slices - 1;
        tright = 1;
    } else if (islice == slices - 1) {
        tleft = slices - 2;
        tright = 0;
    } else {
        tleft = islice - 1;
        tright = islice + 1;
    }
    e += 2.0 * s * (jperp * (double)confs[ispin * cs0 + tleft * cs1]);
    e += 2.0 * s * (jperp * (double)confs[ispin * cs0 + tright * cs1]);
    return e;
}

/* J_perp and b_coeff of one schedule step, qmc.pyx:95-96 */
void mcs_oracle_qmc_coeffs(double a, double b, double teff, double *jperp, double *b_coeff)
{
    *jperp = -0.5 * teff * log(tanh(a / teff));
    *b_coeff = -2.0 * b;
}

/* teff = (double)(float)temp * slices, qmc.pyx:85 (temp is a C float in the signature, :28) */
double mcs_oracle_teff(float temp, int slices) { return (double)temp * (double)slices; }

/* ------------------------------------------------------------------------------------------
 * qmc.QuantumAnneal (qmc.pyx:25-143) and qmc.QuantumAnnealGlobal (qmc.pyx:284-438).
 * Returns 0, or -1 for the ZeroDivisionError the Cython code raises when teff == 0
 * (no cdivision in qmc.pyx; qmc.c:3030-3034).
 * -------------------------------------------------------------------------------------- */
int mcs_oracle_qmc_anneal(const double *A, const double *B, int schedsize, int mcsteps, float temp,
                          int64_t *confs, int64_t cs0, int64_t cs1, int nspins, int slices,
                          const double *nbs, int maxnb, int global_moves, mcs_rand_t *rng)
{
    double teff = (double)temp * (double)slices;
    int64_t *ispins;
    int ifield, step, islice, sidx, k;
    if (teff == 0.0 && schedsize > 0) return -1;
    ispins = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nspins > 0 ? nspins : 1));
    for (ifield = 0; ifield < schedsize; ++ifield) {
        double jperp = -0.5 * teff * log(tanh(A[ifield] / teff));
        double b_coeff = -2.0 * B[ifield];
        for (step = 0; step < mcsteps; ++step) {
            for (islice = 0; islice < slices; ++islice) {
                shuffle(rng, ispins, nspins);
                for (sidx = 0; sidx < nspins; ++sidx) {
                    int ispin = (int)ispins[sidx];
                    double e = qmc_ediff(confs, cs0, cs1, nbs, maxnb, slices, ispin, islice,
                                         b_coeff, jperp);
                    if (e <= 0.0)
                        confs[ispin * cs0 + islice * cs1] *= -1;
                    else if (exp(-1.0 * e / teff) > mcs_rand(rng) / (double)MCS_RAND_MAX)
                        confs[ispin * cs0 + islice * cs1] *= -1;
                }
            }
            if (global_moves) { /* qmc.pyx:405-438 */
                shuffle(rng, ispins, nspins);
                for (sidx = 0; sidx < nspins; ++sidx) {
                    int ispin = (int)ispins[sidx];
                    double e = 0.0;
                    for (k = 0; k < slices; ++k)
                        e = qmc_inplane(confs, cs0, cs1, nbs, maxnb, ispin, k, b_coeff, e);
                    if (e <= 0.0) {
                        for (k = 0; k < slices; ++k) confs[ispin * cs0 + k * cs1] *= -1;
                    } else if (exp(-1.0 * e / teff) > mcs_rand(rng) / (double)MCS_RAND_MAX) {
                        for (k = 0; k < slices; ++k) confs[ispin * cs0 + k * cs1] *= -1;
                    }
                }
            }
        }
    }
    free(ispins);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * qmc.DissipativeQuantumAnneal (qmc.pyx:149-278) / DissipativeQuantumAnnealGlobal (:444-609):
 * local move gets the Ohmic-bath term  sum_{k'=1..P-1} 2*teff*(s_i^k * s_i^{(k+k') mod P}) * lookuptable[k'-1]
 * (qmc.pyx:268-273, integer spin product first).
 * -------------------------------------------------------------------------------------- */
int mcs_oracle_qmc_dissipative(const double *A, const double *B, int schedsize, int mcsteps,
                               float temp, const double *lookuptable, int64_t *confs, int64_t cs0,
                               int64_t cs1, int nspins, int slices, const double *nbs, int maxnb,
                               int global_moves, mcs_rand_t *rng);

/* ------------------------------------------------------------------------------------------
 * Energy differences of every (spin, slice) for a frozen configuration: what qmc.pyx:112-138
 * would compute for each visit, without flipping anything.  out is [nspins, slices] C-order.
 * Tier-(a) parity target (bit-exact).
 * -------------------------------------------------------------------------------------- */
void mcs_oracle_qmc_delta_e(double a, double b, float temp, const int64_t *confs, int64_t cs0,
                            int64_t cs1, int nspins, int slices, const double *nbs, int maxnb,
                            double *out)
{
    double teff = (double)temp * (double)slices;
    double jperp = -0.5 * teff * log(tanh(a / teff));
    double b_coeff = -2.0 * b;
    int i, k;
    for (i = 0; i < nspins; ++i)
        for (k = 0; k < slices; ++k)
            out[(int64_t)i * slices + k] =
                qmc_ediff(confs, cs0, cs1, nbs, maxnb, slices, i, k, b_coeff, jperp);
}

/* world-line (global move) energy differences, qmc.pyx:416-431; out is [nspins] */
void mcs_oracle_qmc_delta_e_global(double b, const int64_t *confs, int64_t cs0, int64_t cs1,
                                   int nspins, int slices, const double *nbs, int maxnb,
                                   double *out)
{
    double b_coeff = -2.0 * b;
    int i, k;
    for (i = 0; i < nspins; ++i) {
        double e = 0.0;
        for (k = 0; k < slices; ++k) e = qmc_inplane(confs, cs0, cs1, nbs, maxnb, i, k, b_coeff, e);
        out[i] = e;
    }
}

/* SA energy difference of every spin, sa.pyx:84-94; out is [nspins] */
void mcs_oracle_sa_delta_e(const int64_t *svec, int64_t ss, int nspins, const double *nbs,
                           int maxnb, double *out)
{
    int i, si;
    for (i = 0; i < nspins; ++i) {
        double e = 0.0;
        double s = (double)svec[i * ss];
        for (si = 0; si < maxnb; ++si) {
            int spinidx = NB_IDX(nbs, maxnb, i, si);
            double jval = NB_J(nbs, maxnb, i, si);
            if (spinidx == i)
                e += -2.0 * s * jval;
            else
                e += -2.0 * s * (jval * (double)svec[spinidx * ss]);
        }
        out[i] = e;
    }
}

/* ------------------------------------------------------------------------------------------
 * Total classical Ising energy in the reference's convention (tools.pyx:99-118):
 *   E = sum_{stored (i,j), i != j} J_ij s_i s_j + sum_i h_i s_i ,
 * each bond stored once in J but mirrored into both rows of the neighbour table
 * (tools.pyx:84-92), the field as a self entry.  ClassicalIsingEnergy itself sums with dense
 * BLAS in an implementation-defined order, so the bit-exact definition is this
 * fixed-order one: rows in site order, entries in table order,
 *   E = sum_i s_i * ( 0.5 * sum_{si: j != i} J s_j  +  sum_{si: j == i} h ).
 * tests pin it to ClassicalIsingEnergy within a few ulp * nnz.
 * -------------------------------------------------------------------------------------- */
double mcs_oracle_ising_energy(const int64_t *svec, int64_t ss, int nspins, const double *nbs,
                               int maxnb)
{
    double e = 0.0;
    int i, si;
    for (i = 0; i < nspins; ++i) {
        double pair = 0.0, field = 0.0;
        for (si = 0; si < maxnb; ++si) {
            int spinidx = NB_IDX(nbs, maxnb, i, si);
            double jval = NB_J(nbs, maxnb, i, si);
            if (spinidx == i)
                field += jval;
            else
                pair += jval * (double)svec[spinidx * ss];
        }
        e += (double)svec[i * ss] * (0.5 * pair + field);
    }
    return e;
}

/* ------------------------------------------------------------------------------------------
 * sa.Anneal (sa.pyx:19-101); Anneal_parallel (:201-284) is the same loop nest when built
 * without OpenMP (setup.py:10-11), so it shares this function.
 * randuni == NULL: acceptance uniform = rand()/RAND_MAX drawn only when ediff > 0 (:96-99).
 * randuni != NULL: sa.AnnealMA (:108-193): uniform = randuni[itemp, step, visit position]
 *                  (:190), shuffle still from rand().
 * nbs_step_stride != 0: sa.NoisyAnneal (:291-378): nbs is [sched, nspins, maxnb, 2] and the
 *                  table of step itemp is used (:363-365); NoisyAnneal always uses randuni.
 * -------------------------------------------------------------------------------------- */
void mcs_oracle_sa_anneal(const double *sched, int schedsize, int mcsteps, int64_t *svec,
                          int64_t ss, int nspins, const double *nbs, int maxnb,
                          int64_t nbs_step_stride, const double *randuni, mcs_rand_t *rng)
{
    int64_t *ispins = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nspins > 0 ? nspins : 1));
    int itemp, step, ispin, si;
    double ediff = 0.0;
    for (itemp = 0; itemp < schedsize; ++itemp) {
        double temp = sched[itemp];
        const double *tab = nbs + (int64_t)itemp * nbs_step_stride;
        for (step = 0; step < mcsteps; ++step) {
            shuffle(rng, ispins, nspins);
            for (ispin = 0; ispin < nspins; ++ispin) {
                int sidx = (int)ispins[ispin];
                double s = (double)svec[sidx * ss];
                for (si = 0; si < maxnb; ++si) {
                    int spinidx = NB_IDX(tab, maxnb, sidx, si);
                    double jval = NB_J(tab, maxnb, sidx, si);
                    if (spinidx == sidx)
                        ediff += -2.0 * s * jval;
                    else
                        ediff += -2.0 * s * (jval * (double)svec[spinidx * ss]);
                }
                if (ediff <= 0.0) {
                    svec[sidx * ss] *= -1;
                } else {
                    double u = randuni
                                   ? randuni[((int64_t)itemp * mcsteps + step) * nspins + ispin]
                                   : mcs_rand(rng) / (double)MCS_RAND_MAX;
                    if (exp(-1.0 * ediff / temp) > u) svec[sidx * ss] *= -1;
                }
                ediff = 0.0;
            }
        }
    }
    free(ispins);
}

/* ------------------------------------------------------------------------------------------
 * svmc.SpinVectorMonteCarlo (svmc.pyx:21-117), SpinVectorMonteCarloTF (:123-229),
 * NoisySVMC (:236-334) / NoisySVMCTF (:340-448) via nbs_step_stride, and the batched
 * SpinVectorMonteCarloCompact (:455-554) via numreads > 1 (serial loop over reads, ONE
 * randuni[sched, mcsteps, nspins, 2] shared by all reads, :506,532,551).
 * svec is [numreads, nspins] with element strides (rs, ss).
 * -------------------------------------------------------------------------------------- */
void mcs_oracle_svmc(const double *A, const double *B, int schedsize, int mcsteps, float temp,
                     double *svec, int64_t rs, int64_t ss, int numreads, int nspins,
                     const double *nbs, int maxnb, int64_t nbs_step_stride, const double *randuni,
                     int tf, mcs_rand_t *rng)
{
    int64_t *ispins = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nspins > 0 ? nspins : 1));
    const double pi = 3.141592653589793; /* np.pi */
    int iread, ifield, step, ispin, si;
    double ediff = 0.0;
    for (iread = 0; iread < numreads; ++iread) {
        double *sv = svec + (int64_t)iread * rs;
        for (ifield = 0; ifield < schedsize; ++ifield) {
            double a_coeff = A[ifield];
            double b_coeff = B[ifield];
            const double *tab = nbs + (int64_t)ifield * nbs_step_stride;
            for (step = 0; step < mcsteps; ++step) {
                shuffle(rng, ispins, nspins);
                for (ispin = 0; ispin < nspins; ++ispin) {
                    int sidx = (int)ispins[ispin];
                    const double *ru = randuni + (((int64_t)ifield * mcsteps + step) * nspins + ispin) * 2;
                    double theta_prop, zmagdiff;
                    if (!tf) {
                        theta_prop = pi * ru[0]; /* svmc.pyx:95 */
                    } else {                     /* svmc.pyx:198-207 */
                        double ab_ratio = a_coeff / b_coeff;
                        if (ab_ratio > 1)
                            theta_prop = (2.0 * pi * ru[0]) - pi;
                        else
                            theta_prop = ab_ratio * ((2.0 * pi * ru[0]) - pi);
                        theta_prop = theta_prop + sv[sidx * ss];
                        if (theta_prop < 0)
                            theta_prop = 0.0;
                        else if (theta_prop > pi)
                            theta_prop = pi;
                    }
                    zmagdiff = cos(theta_prop) - cos(sv[sidx * ss]);
                    for (si = 0; si < maxnb; ++si) {
                        int spinidx = NB_IDX(tab, maxnb, sidx, si);
                        double jval = NB_J(tab, maxnb, sidx, si);
                        if (spinidx == sidx)
                            ediff += b_coeff * jval * zmagdiff;
                        else
                            ediff += b_coeff * jval * zmagdiff * cos(sv[spinidx * ss]);
                    }
                    ediff += a_coeff * (sin(sv[sidx * ss]) - sin(theta_prop));
                    if (ediff <= 0.0)
                        sv[sidx * ss] = theta_prop;
                    else if (exp(-1.0 * ediff / temp) > ru[1])
                        sv[sidx * ss] = theta_prop;
                    ediff = 0.0;
                }
            }
        }
    }
    free(ispins);
}

/* ------------------------------------------------------------------------------------------
 * svmc.SpinVectorMonteCarloTFCompact (svmc.pyx:561-674): batched TF variant whose proposal
 * and acceptance uniforms come from rand()/RAND_MAX (:645-647, :671) instead of np.random.
 * Serial over reads (its OpenMP pragma is dead at build, setup.py:17-18).
 * -------------------------------------------------------------------------------------- */
void mcs_oracle_svmc_tf_compact(const double *A, const double *B, int schedsize, int mcsteps,
                                float temp, double *svec, int64_t rs, int64_t ss, int numreads,
                                int nspins, const double *nbs, int maxnb, mcs_rand_t *rng)
{
    int64_t *ispins = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nspins > 0 ? nspins : 1));
    const double pi = 3.141592653589793;
    const double rand_max = (double)MCS_RAND_MAX;
    int iread, ifield, step, ispin, si;
    for (iread = 0; iread < numreads; ++iread) {
        double *sv = svec + (int64_t)iread * rs;
        double ediff = 0.0;
        for (ifield = 0; ifield < schedsize; ++ifield) {
            double a_coeff = A[ifield];
            double b_coeff = B[ifield];
            for (step = 0; step < mcsteps; ++step) {
                shuffle(rng, ispins, nspins);
                for (ispin = 0; ispin < nspins; ++ispin) {
                    int sidx = (int)ispins[ispin];
                    double ab_ratio = a_coeff / b_coeff;
                    double theta_prop, zmagdiff;
                    if (ab_ratio > 1)
                        theta_prop = (2.0 * pi * mcs_rand(rng) / rand_max) - pi;
                    else
                        theta_prop = ab_ratio * ((2.0 * pi * mcs_rand(rng) / rand_max) - pi);
                    theta_prop += sv[sidx * ss];
                    if (theta_prop < 0)
                        theta_prop = 0.0;
                    else if (theta_prop > pi)
                        theta_prop = pi;
                    zmagdiff = cos(theta_prop) - cos(sv[sidx * ss]);
                    for (si = 0; si < maxnb; ++si) {
                        int spinidx = NB_IDX(nbs, maxnb, sidx, si);
                        double jval = NB_J(nbs, maxnb, sidx, si);
                        if (spinidx == sidx)
                            ediff += b_coeff * jval * zmagdiff;
                        else
                            ediff += b_coeff * jval * zmagdiff * cos(sv[spinidx * ss]);
                    }
                    ediff += a_coeff * (sin(sv[sidx * ss]) - sin(theta_prop));
                    if (ediff <= 0.0)
                        sv[sidx * ss] = theta_prop;
                    else if (exp(-1.0 * ediff / temp) > mcs_rand(rng) / rand_max)
                        sv[sidx * ss] = theta_prop;
                    ediff = 0.0;
                }
            }
        }
    }
    free(ispins);
}

/* ------------------------------------------------------------------------------------------
 * SVMC energy (what the sweeps sample), from svmc.pyx:96-110 by integration:
 *   H(theta) = B * ( sum_{bonds} J_ij cos(t_i) cos(t_j) + sum_i h_i cos(t_i) ) - A * sum_i sin(t_i).
 * Fixed order like mcs_oracle_ising_energy.  Used for statistical parity of the production kernel.
 * -------------------------------------------------------------------------------------- */
double mcs_oracle_svmc_energy(double a, double b, const double *sv, int64_t ss, int nspins,
                              const double *nbs, int maxnb)
{
    double ez = 0.0, ex = 0.0;
    int i, si;
    for (i = 0; i < nspins; ++i) {
        double pair = 0.0, field = 0.0;
        for (si = 0; si < maxnb; ++si) {
            int spinidx = NB_IDX(nbs, maxnb, i, si);
            double jval = NB_J(nbs, maxnb, i, si);
            if (spinidx == i)
                field += jval;
            else
                pair += jval * cos(sv[spinidx * ss]);
        }
        ez += cos(sv[i * ss]) * (0.5 * pair + field);
        ex += sin(sv[i * ss]);
    }
    return b * ez - a * ex;
}

/* ------------------------------------------------------------------------------------------
 * Dissipative PIQMC.  Bath term exactly as qmc.pyx:264-273 (local) writes it.
 * -------------------------------------------------------------------------------------- */
int mcs_oracle_qmc_dissipative(const double *A, const double *B, int schedsize, int mcsteps,
                               float temp, const double *lookuptable, int64_t *confs, int64_t cs0,
                               int64_t cs1, int nspins, int slices, const double *nbs, int maxnb,
                               int global_moves, mcs_rand_t *rng)
{
    double teff = (double)temp * (double)slices;
    int64_t *ispins;
    int ifield, step, islice, sidx, k, kp;
    if (teff == 0.0 && schedsize > 0) return -1;
    ispins = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nspins > 0 ? nspins : 1));
    for (ifield = 0; ifield < schedsize; ++ifield) {
        double jperp = -0.5 * teff * log(tanh(A[ifield] / teff));
        double b_coeff = -2.0 * B[ifield];
        for (step = 0; step < mcsteps; ++step) {
            for (islice = 0; islice < slices; ++islice) {
                shuffle(rng, ispins, nspins);
                for (sidx = 0; sidx < nspins; ++sidx) {
                    int ispin = (int)ispins[sidx];
                    double e = qmc_ediff(confs, cs0, cs1, nbs, maxnb, slices, ispin, islice,
                                         b_coeff, jperp);
                    for (kp = 1; kp < slices; ++kp) { /* qmc.pyx:268-273 */
                        int other = (islice + kp) % slices;
                        e += 2.0 * teff *
                             (double)(confs[ispin * cs0 + islice * cs1] * confs[ispin * cs0 + other * cs1]) *
                             lookuptable[kp - 1];
                    }
                    if (e <= 0.0)
                        confs[ispin * cs0 + islice * cs1] *= -1;
                    else if (exp(-1.0 * e / teff) > mcs_rand(rng) / (double)MCS_RAND_MAX)
                        confs[ispin * cs0 + islice * cs1] *= -1;
                }
            }
            if (global_moves) {
                shuffle(rng, ispins, nspins);
                for (sidx = 0; sidx < nspins; ++sidx) {
                    int ispin = (int)ispins[sidx];
                    double e = 0.0;
                    for (k = 0; k < slices; ++k)
                        e = qmc_inplane(confs, cs0, cs1, nbs, maxnb, ispin, k, b_coeff, e);
                    if (e <= 0.0) {
                        for (k = 0; k < slices; ++k) confs[ispin * cs0 + k * cs1] *= -1;
                    } else if (exp(-1.0 * e / teff) > mcs_rand(rng) / (double)MCS_RAND_MAX) {
                        for (k = 0; k < slices; ++k) confs[ispin * cs0 + k * cs1] *= -1;
                    }
                }
            }
        }
    }
    free(ispins);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Coloured-order variants (NOT in the reference): the same fp64 visit arithmetic and the same
 * acceptance rule as mcs_oracle_qmc_anneal / mcs_oracle_sa_anneal, but sites are visited colour
 * class by colour class and, inside a site, even Trotter slices before odd ones -- the order the
 * B200 production kernels use (montecarlosolvers_b200/csrc/mcs_piqmc.cu, mcs_sa.cu).  They exist to
 * separate "different visiting order" from "different arithmetic" in parity tier (c): the GPU
 * kernels must agree with THESE at the standard-error level, while the gap between these and the
 * reference order is a property of the dynamics.
 * order[]: sites sorted by colour; color_start[c] .. color_start[c+1]: the sites of colour c.
 * -------------------------------------------------------------------------------------- */
static inline void qmc_visit(int64_t *confs, int64_t cs0, int64_t cs1, const double *nbs, int maxnb, int slices,
                             int ispin, int islice, double b_coeff, double jperp, double teff, mcs_rand_t *rng)
{
    double e = qmc_ediff(confs, cs0, cs1, nbs, maxnb, slices, ispin, islice, b_coeff, jperp);
    if (e <= 0.0)
        confs[ispin * cs0 + islice * cs1] *= -1;
    else if (exp(-1.0 * e / teff) > mcs_rand(rng) / (double)MCS_RAND_MAX)
        confs[ispin * cs0 + islice * cs1] *= -1;
}

int mcs_oracle_qmc_anneal_colored(const double *A, const double *B, int schedsize, int mcsteps, float temp,
                                  int64_t *confs, int64_t cs0, int64_t cs1, int nspins, int slices,
                                  const double *nbs, int maxnb, int global_moves, const int32_t *order,
                                  const int32_t *color_start, int ncolors, mcs_rand_t *rng)
{
    double teff = (double)temp * (double)slices;
    int ifield, step, c, q, k, parity;
    int last_alone = (slices & 1) ? slices - 1 : -1; /* odd ring: slice P-1 is visited after the two parities */
    (void)nspins;
    if (teff == 0.0 && schedsize > 0) return -1;
    for (ifield = 0; ifield < schedsize; ++ifield) {
        double jperp = -0.5 * teff * log(tanh(A[ifield] / teff));
        double b_coeff = -2.0 * B[ifield];
        for (step = 0; step < mcsteps; ++step) {
            for (c = 0; c < ncolors; ++c) {
                for (q = color_start[c]; q < color_start[c + 1]; ++q) {
                    int ispin = order[q];
                    for (parity = 0; parity < 2; ++parity)
                        for (k = parity; k < slices; k += 2)
                            if (k != last_alone)
                                qmc_visit(confs, cs0, cs1, nbs, maxnb, slices, ispin, k, b_coeff, jperp, teff, rng);
                    if (last_alone >= 0)
                        qmc_visit(confs, cs0, cs1, nbs, maxnb, slices, ispin, last_alone, b_coeff, jperp, teff, rng);
                    if (global_moves) {
                        double e = 0.0;
                        for (k = 0; k < slices; ++k)
                            e = qmc_inplane(confs, cs0, cs1, nbs, maxnb, ispin, k, b_coeff, e);
                        if (e <= 0.0 || exp(-1.0 * e / teff) > mcs_rand(rng) / (double)MCS_RAND_MAX)
                            for (k = 0; k < slices; ++k) confs[ispin * cs0 + k * cs1] *= -1;
                    }
                }
            }
        }
    }
    return 0;
}

void mcs_oracle_sa_anneal_colored(const double *sched, int schedsize, int mcsteps, int64_t *svec, int64_t ss,
                                  int nspins, const double *nbs, int maxnb, const int32_t *order,
                                  const int32_t *color_start, int ncolors, mcs_rand_t *rng)
{
    int itemp, step, c, q, si;
    (void)nspins;
    for (itemp = 0; itemp < schedsize; ++itemp) {
        double temp = sched[itemp];
        for (step = 0; step < mcsteps; ++step) {
            for (c = 0; c < ncolors; ++c) {
                for (q = color_start[c]; q < color_start[c + 1]; ++q) {
                    int sidx = order[q];
                    double s = (double)svec[sidx * ss];
                    double ediff = 0.0;
                    for (si = 0; si < maxnb; ++si) {
                        int spinidx = NB_IDX(nbs, maxnb, sidx, si);
                        double jval = NB_J(nbs, maxnb, sidx, si);
                        if (spinidx == sidx)
                            ediff += -2.0 * s * jval;
                        else
                            ediff += -2.0 * s * (jval * (double)svec[spinidx * ss]);
                    }
                    if (ediff <= 0.0)
                        svec[sidx * ss] *= -1;
                    else if (exp(-1.0 * ediff / temp) > mcs_rand(rng) / (double)MCS_RAND_MAX)
                        svec[sidx * ss] *= -1;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Wolff-cluster experiments of the reference (qmc.pyx:612-1621): mcs_oracle_qmc_wolff
 * -------------------------------------------------------------------------------------- */
#include "mcs_oracle_wolff.c"
