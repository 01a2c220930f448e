/*
 * mcs_oracle_wolff.c -- CPU restatement of the reference's Wolff-cluster experiments
 * (qmc.pyx:612-1621, the block the reference itself titles "Function under test").
 *
 * TEST INFRASTRUCTURE ONLY (same rules as mcs_oracle.c, which #includes this file).
 *
 * Parity status: PINNED.  As shipped these functions raise on Linux (their scratch buffers are allocated
 * with dtype=np.intc but typed np.int_t); with oracle/build_ref.py's mechanical dtype patch they run, and
 * tests/test_oracle_vs_reference.py compares this restatement with the compiled reference bit for bit
 * (configurations and number of rand() draws).
 *
 * The functions are restated AS WRITTEN, including what look like slips of an experiment in progress,
 * because "identical results" is the bar:
 *   - WCL / WC / WC2 do not flip the seed spin (qmc.pyx:704 is commented out), so the seed can join its own
 *     cluster later and the explicit stack can hold one entry more than the reference allocates;
 *   - the padded rows of the neighbour table are walked in full: a padding entry (index 0, J = 0) proposes
 *     site 0 as a neighbour of every short row;
 *   - WC reads `nbs[spinidx, si2, 1]` in its Trotter branches with `spinidx` left over from the previous
 *     spatial loop (qmc.pyx:1130, 1161), WC2 uses `bslice` and `jval` left over from earlier loops
 *     (qmc.pyx:1364, 1414): function-scope variables here too;
 *   - WC never updates r (qmc.pyx:1150 commented out); WC2 UNDOES part of the cluster when the Metropolis
 *     test on the accumulated energy succeeds (qmc.pyx:1443-1446), WC3 undoes it with probability
 *     1 - exp(-E/teff) (qmc.pyx:1617-1621);
 *   - WC3's `for islice in xrange(slices)` body overwrites `islice` (qmc.pyx:1571); Cython iterates on a
 *     temporary, so the next spin starts from the slice the last cluster ended on.
 */

#define CF(i, k) confs[(int64_t)(i) * cs0 + (int64_t)(k) * cs1]
#define URAND(rng) (mcs_rand(rng) / (double)MCS_RAND_MAX)

static inline void trotter_nb(int islice, int slices, int *tl, int *tr)
{ /* qmc.pyx:737-745 */
    if (islice == 0) {
        *tl = slices - 1;
        *tr = 1;
    } else if (islice == slices - 1) {
        *tl = slices - 2;
        *tr = 0;
    } else {
        *tl = islice - 1;
        *tr = islice + 1;
    }
}

/* "add bias energy" loop of qmc.pyx:722-725: every entry of row s that points at s itself */
static inline double wolff_bias(const double *nbs, int maxnb, int s, double b_coeff, int k, double ediff)
{
    int si2;
    for (si2 = 0; si2 < maxnb; ++si2) {
        int spinidx2 = NB_IDX(nbs, maxnb, s, si2);
        if (s == spinidx2) ediff += -2.0 * b_coeff * NB_J(nbs, maxnb, s, si2) * k;
    }
    return ediff;
}

typedef struct {
    int64_t *cl; /* explicit stack: (spin, slice) pairs */
    int stack, stackidx, cluster_count;
    int max_rows; /* largest row index written: the reference allocates nspins*slices rows (WCL, WC) or `slices`
                     rows (WC2, WC3) and does not check */
    double r;
} wolff_t;

static inline void wolff_push(wolff_t *w, int64_t *confs, int64_t cs0, int64_t cs1, int spin, int slice)
{ /* qmc.pyx:731-736 */
    w->cl[2 * w->stackidx] = spin;
    w->cl[2 * w->stackidx + 1] = slice;
    if (w->stackidx > w->max_rows) w->max_rows = w->stackidx;
    CF(spin, slice) *= -1;
    w->stack += 1;
    w->stackidx += 1;
}

/* growth attempt of qmc.pyx:726-736: bond probability p = 1 - exp(ediff/teff), gated by the cumulative r */
static inline void wolff_try(wolff_t *w, int64_t *confs, int64_t cs0, int64_t cs1, int spin, int slice, double ediff,
                             double teff, int update_r, mcs_rand_t *rng)
{
    if (ediff < 0) {
        double p = 1 - exp(ediff / teff);
        if (w->r * p > URAND(rng)) {
            if (update_r) w->r *= p;
            wolff_push(w, confs, cs0, cs1, spin, slice);
        }
    }
}

/* spatial + Trotter growth shared by QuantumAnnealWCL (qmc.pyx:708-781) and DissaptiveQuantumAnnealWCL (:926-995) */
static inline void wcl_grow_space_time(wolff_t *w, int64_t *confs, int64_t cs0, int64_t cs1, const double *nbs,
                                       int maxnb, int slices, int ispin, int islice, int k, double b_coeff,
                                       double jperp, double teff, mcs_rand_t *rng)
{
    int si, tleft, tright;
    double ediff;
    for (si = 0; si < maxnb; ++si) {
        int spinidx = NB_IDX(nbs, maxnb, ispin, si);
        if (CF(spinidx, islice) == k) {
            ediff = 0.0;
            ediff += 2.0 * b_coeff * NB_J(nbs, maxnb, ispin, si);
            ediff = wolff_bias(nbs, maxnb, spinidx, b_coeff, k, ediff);
            wolff_try(w, confs, cs0, cs1, spinidx, islice, ediff, teff, 1, rng);
        }
    }
    trotter_nb(islice, slices, &tleft, &tright);
    if (CF(ispin, tleft) == k) {
        ediff = 0.0;
        ediff += -2.0 * jperp;
        ediff = wolff_bias(nbs, maxnb, ispin, b_coeff, k, ediff);
        wolff_try(w, confs, cs0, cs1, ispin, tleft, ediff, teff, 1, rng);
    }
    if (CF(ispin, tright) == k) {
        ediff = 0.0;
        ediff += -2.0 * jperp;
        ediff = wolff_bias(nbs, maxnb, ispin, b_coeff, k, ediff);
        wolff_try(w, confs, cs0, cs1, ispin, tright, ediff, teff, 1, rng);
    }
}

/* variant: 0 = QuantumAnnealWCL (qmc.pyx:620-786), 1 = DissaptiveQuantumAnnealWCL (:792-1000),
 *          2 = QuantumAnnealWC (:1006-1225), 3 = DissipativeQuantumAnnealWC2 (:1231-1446),
 *          4 = DissipativeQuantumAnnealWC3 (:1452-1621).  lookuptable is read by variants 1, 3, 4.
 * Returns 0, -1 when teff == 0 (the reference divides by it), -2 for slices < 2 (the reference indexes
 * slice 1 unconditionally), and 1 when the run is complete but the REFERENCE would have written past its
 * `cluster` buffer on the way (undefined behaviour there: the seed re-joined a cluster that already held
 * every other node) -- such a case cannot be compared with the compiled reference. */
int mcs_oracle_qmc_wolff(int variant, const double *A, const double *B, int schedsize, int mcsteps, float temp,
                         const double *lookuptable, int64_t *confs, int64_t cs0, int64_t cs1, int nspins, int slices,
                         const double *nbs, int maxnb, mcs_rand_t *rng)
{
    double teff = (double)temp * (double)slices;
    int ifield, step, i, j, si, si2, b, b2, sidx, sidx2, loop_slice;
    /* function-scope variables of the reference (initial values qmc.pyx:1054-1070, 1284-1318, 1506-1540) */
    int ispin = 0, islice = 0, spinidx = 0, spinidx2 = 0, tleft = 0, tright = 0, tleft2 = 0, tright2 = 0;
    int bslice = 0, cslice = 0, k = 0;
    double jval = 0.0, ediff = 0.0, e_total = 0.0;
    int64_t *ispins;
    wolff_t w;
    int capacity = (variant == 3 || variant == 4) ? slices : nspins * slices;
    if (slices < 2) return -2;
    if (teff == 0.0 && schedsize > 0) return -1;
    w.max_rows = 0;
    w.cl = (int64_t *)malloc(sizeof(int64_t) * 2 * ((size_t)nspins * (size_t)slices + 2));
    ispins = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nspins > 0 ? nspins : 1));
    for (i = 0; i < nspins; ++i) ispins[i] = i;
    for (ifield = 0; ifield < schedsize; ++ifield) {
        double jperp = -0.5 * teff * log(tanh(A[ifield] / teff));
        double b_coeff = B[ifield]; /* +B here (qmc.pyx:696), not -2 B */
        for (step = 0; step < mcsteps; ++step) {
            if (variant == 0 || variant == 1 || variant == 2) {
                /* one single-cluster move per step */
                ispin = mcs_rand(rng) % nspins;
                islice = mcs_rand(rng) % slices;
                if (variant == 1) { /* walk to a start point that aligns with the local field, qmc.pyx:879-893 */
                    j = islice * nspins + ispin;
                    for (i = 1; i < nspins * slices; ++i) {
                        ediff = 0.0;
                        for (si = 0; si < maxnb; ++si) {
                            spinidx = NB_IDX(nbs, maxnb, ispin, si);
                            jval = NB_J(nbs, maxnb, ispin, si);
                            if (spinidx == ispin) ediff += -2.0 * jval * b_coeff * CF(ispin, islice);
                        }
                        if (ediff <= 0)
                            break;
                        else if (exp(-1.0 * ediff / teff) > URAND(rng))
                            break;
                        else {
                            ispin = (j + i) % nspins;
                            islice = ((j + i) / nspins) % slices;
                        }
                    }
                }
                w.cl[0] = ispin;
                w.cl[1] = islice;
                k = (int)CF(ispin, islice);
                if (variant == 1) CF(ispin, islice) *= -1; /* qmc.pyx:898; commented out at :704 and absent at :1103 */
                w.stack = 1;
                w.stackidx = 1;
                w.cluster_count = 0;
                w.r = 1.0;
                for (;;) {
                    ispin = (int)w.cl[2 * w.cluster_count];
                    islice = (int)w.cl[2 * w.cluster_count + 1];
                    if (variant == 1) { /* bath neighbours first, qmc.pyx:906-925 */
                        for (b = 1; b < slices; ++b) {
                            bslice = (islice + b) % slices;
                            if (CF(ispin, bslice) == k) {
                                ediff = 0.0;
                                ediff += -2.0 * teff * lookuptable[b - 1];
                                ediff = wolff_bias(nbs, maxnb, ispin, b_coeff, k, ediff);
                                wolff_try(&w, confs, cs0, cs1, ispin, bslice, ediff, teff, 1, rng);
                            }
                        }
                    }
                    if (variant != 2) {
                        wcl_grow_space_time(&w, confs, cs0, cs1, nbs, maxnb, slices, ispin, islice, k, b_coeff, jperp,
                                            teff, rng);
                    } else { /* QuantumAnnealWC: full energy change of the candidate, qmc.pyx:1112-1220 */
                        int side;
                        trotter_nb(islice, slices, &tleft, &tright);
                        for (side = 0; side < 2; ++side) {
                            int ts = side == 0 ? tleft : tright;
                            if (CF(ispin, ts) == k) {
                                ediff = 0.0;
                                for (si2 = 0; si2 < maxnb; ++si2) {
                                    spinidx2 = NB_IDX(nbs, maxnb, ispin, si2);
                                    jval = NB_J(nbs, maxnb, spinidx, si2); /* `spinidx` is stale here (:1130, :1161) */
                                    if (spinidx == spinidx2)
                                        ediff += -2.0 * b_coeff * jval * k;
                                    else
                                        ediff += -2.0 * b_coeff * jval * k * CF(spinidx2, ts);
                                }
                                trotter_nb(ts, slices, &tleft2, &tright2);
                                ediff += 2.0 * jperp * k * CF(ispin, tleft2);
                                ediff += 2.0 * jperp * k * CF(ispin, tright2);
                                wolff_try(&w, confs, cs0, cs1, ispin, ts, ediff, teff, 0, rng);
                            }
                        }
                        for (si = 0; si < maxnb; ++si) {
                            spinidx = NB_IDX(nbs, maxnb, ispin, si);
                            if (CF(spinidx, islice) == k) {
                                ediff = 0.0;
                                for (si2 = 0; si2 < maxnb; ++si2) {
                                    spinidx2 = NB_IDX(nbs, maxnb, spinidx, si2);
                                    jval = NB_J(nbs, maxnb, spinidx, si2);
                                    if (spinidx == spinidx2)
                                        ediff += -2.0 * b_coeff * jval * k;
                                    else
                                        ediff += -2.0 * b_coeff * jval * k * CF(spinidx2, islice);
                                }
                                trotter_nb(islice, slices, &tleft2, &tright2);
                                ediff += 2.0 * jperp * k * CF(spinidx, tleft2);
                                ediff += 2.0 * jperp * k * CF(spinidx, tright2);
                                wolff_try(&w, confs, cs0, cs1, spinidx, islice, ediff, teff, 0, rng);
                            }
                        }
                    }
                    w.cluster_count += 1;
                    w.stack += -1;
                    if (w.stack == 0) break;
                }
            } else if (variant == 3) {
                /* local sweep with the bath term taken from a stale slice (qmc.pyx:1325-1375) ... */
                for (loop_slice = 0; loop_slice < slices; ++loop_slice) {
                    islice = loop_slice;
                    shuffle(rng, ispins, nspins);
                    for (sidx = 0; sidx < nspins; ++sidx) {
                        double s, e = 0.0;
                        ispin = (int)ispins[sidx];
                        s = (double)CF(ispin, islice);
                        for (si = 0; si < maxnb; ++si) {
                            spinidx = NB_IDX(nbs, maxnb, ispin, si);
                            jval = NB_J(nbs, maxnb, ispin, si);
                            if (spinidx == ispin)
                                e += -2.0 * b_coeff * s * jval;
                            else
                                e += -2.0 * b_coeff * s * (jval * (double)CF(spinidx, islice));
                        }
                        trotter_nb(islice, slices, &tleft, &tright);
                        e += 2.0 * s * (jperp * (double)CF(ispin, tleft));
                        e += 2.0 * s * (jperp * (double)CF(ispin, tright));
                        for (b2 = 1; b2 < slices; ++b2) {
                            cslice = (bslice + b2) % slices; /* `bslice`, not islice (:1364) */
                            e += 2.0 * teff * (double)(CF(ispin, islice) * CF(ispin, cslice)) * lookuptable[b2 - 1];
                        }
                        if (e <= 0.0)
                            CF(ispin, islice) *= -1;
                        else if (exp(-1.0 * e / teff) > URAND(rng))
                            CF(ispin, islice) *= -1;
                    }
                }
                /* ... then one bath-only cluster per spin, Metropolis on its accumulated energy (qmc.pyx:1376-1446) */
                shuffle(rng, ispins, nspins);
                for (sidx2 = 0; sidx2 < nspins; ++sidx2) {
                    ispin = (int)ispins[sidx2];
                    islice = mcs_rand(rng) % slices;
                    w.cl[0] = ispin;
                    w.cl[1] = islice;
                    k = (int)CF(ispin, islice);
                    w.stack = 1;
                    w.stackidx = 1;
                    w.cluster_count = 0;
                    w.r = 1.0;
                    e_total = 0.0;
                    for (;;) {
                        ispin = (int)w.cl[2 * w.cluster_count];
                        islice = (int)w.cl[2 * w.cluster_count + 1];
                        for (b = 1; b < slices; ++b) {
                            bslice = (islice + b) % slices;
                            if (CF(ispin, bslice) == k) {
                                double p = 1 - exp(-2.0 * lookuptable[b - 1]);
                                if (w.r * p > URAND(rng)) {
                                    ediff = 0.0;
                                    for (si2 = 0; si2 < maxnb; ++si2) {
                                        spinidx2 = NB_IDX(nbs, maxnb, ispin, si2);
                                        if (ispin == spinidx2)
                                            ediff += -2.0 * b_coeff * NB_J(nbs, maxnb, ispin, si2) * k;
                                        else /* `jval` is whatever the local sweep left behind (:1414) */
                                            ediff += -2.0 * b_coeff * jval * k * CF(spinidx2, bslice);
                                    }
                                    trotter_nb(bslice, slices, &tleft2, &tright2);
                                    ediff += 2.0 * jperp * k * CF(ispin, tleft2);
                                    ediff += 2.0 * jperp * k * CF(ispin, tright2);
                                    for (b2 = 1; b2 < slices; ++b2) {
                                        cslice = (bslice + b2) % slices;
                                        ediff += 2.0 * teff * (double)(k * CF(ispin, cslice)) * lookuptable[b2 - 1];
                                    }
                                    w.r *= p;
                                    e_total += ediff;
                                    wolff_push(&w, confs, cs0, cs1, ispin, bslice);
                                }
                            }
                        }
                        w.cluster_count += 1;
                        w.stack += -1;
                        if (w.stack == 0) break;
                    }
                    if (e_total > 0) {
                        if (exp(-1.0 * e_total / teff) > URAND(rng)) {
                            for (i = 1; i < w.cluster_count; ++i) CF(w.cl[2 * i], w.cl[2 * i + 1]) *= -1;
                        }
                    }
                }
            } else {
                /* DissipativeQuantumAnnealWC3: N P bath clusters per step, qmc.pyx:1546-1621 */
                shuffle(rng, ispins, nspins);
                for (loop_slice = 0; loop_slice < slices; ++loop_slice) {
                    islice = loop_slice;
                    for (sidx2 = 0; sidx2 < nspins; ++sidx2) {
                        e_total = 0.0;
                        ispin = (int)ispins[sidx2];
                        w.cl[0] = ispin;
                        w.cl[1] = islice; /* after the first spin: the slice the previous cluster ended on */
                        k = (int)CF(ispin, islice);
                        CF(ispin, islice) *= -1;
                        w.stack = 1;
                        w.stackidx = 1;
                        w.cluster_count = 0;
                        w.r = 1.0;
                        for (;;) {
                            ispin = (int)w.cl[2 * w.cluster_count];
                            islice = (int)w.cl[2 * w.cluster_count + 1];
                            for (si = 0; si < maxnb; ++si) {
                                spinidx = NB_IDX(nbs, maxnb, ispin, si);
                                jval = NB_J(nbs, maxnb, ispin, si);
                                if (spinidx == ispin)
                                    e_total += -2.0 * b_coeff * (double)k * jval;
                                else
                                    e_total += -2.0 * b_coeff * (double)k * (jval * (double)CF(spinidx, islice));
                            }
                            trotter_nb(islice, slices, &tleft, &tright);
                            e_total += 2.0 * (double)k * (jperp * (double)CF(ispin, tleft));
                            e_total += 2.0 * (double)k * (jperp * (double)CF(ispin, tright));
                            for (b = 1; b < slices; ++b) {
                                bslice = (islice + b) % slices;
                                if (CF(ispin, bslice) == k) {
                                    double p = 1 - exp(-2.0 * lookuptable[b - 1]);
                                    if (w.r * p > URAND(rng)) {
                                        w.r *= p;
                                        wolff_push(&w, confs, cs0, cs1, ispin, bslice);
                                    }
                                }
                            }
                            w.cluster_count += 1;
                            w.stack += -1;
                            if (w.stack == 0) break;
                        }
                        if (e_total > 0.0) {
                            if (1 - exp(-1.0 * e_total / teff) > URAND(rng)) {
                                for (i = 0; i < w.cluster_count; ++i) CF(w.cl[2 * i], w.cl[2 * i + 1]) *= -1;
                            }
                        }
                    }
                }
            }
        }
    }
    free(ispins);
    free(w.cl);
    return w.max_rows >= capacity ? 1 : 0;
}

#undef CF
#undef URAND
