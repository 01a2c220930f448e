// mcs_exact.cu -- sequential-order validation kernels and fp64 probes (sm_100a).
//
// Parity tiers (a) and (b) of the north star: one GPU thread per replica replays the reference's
// OWN update order -- Fisher-Yates shuffle from a glibc rand() stream, strictly sequential
// Metropolis visits, fp64 sums in the neighbour table's row order with the reference's
// association and no FMA contraction (every product/sum goes through __dmul_rn/__dadd_rn) -- so
// that, fed the same rand() values, it reproduces the reference's spin trajectories bit for bit.
// Reference loop nests: qmc.pyx:93-143, 358-438 (and the Ohmic-bath variants 223-278, 523-609),
// sa.pyx:66-101, 153-193, svmc.pyx:78-117, 181-229, 514-554, 624-674.
//
// This file is compiled with -fmad=false as well; it is a validation path, not the fast path.
#include <cmath>
#include <vector>

#include "mcs_common.cuh"

namespace {

// glibc rand(): TYPE_3 additive feedback generator r[i] = r[i-31] + r[i-3], output >> 1.
struct LibcState {
    int32_t r[31];
    int32_t f, b;
};

void host_srand(LibcState &st, uint32_t seed)
{
    if (seed == 0) seed = 1;
    int32_t word = (int32_t)seed;
    st.r[0] = word;
    for (int i = 1; i < 31; ++i) {
        long hi = word / 127773, lo = word % 127773;
        long w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        word = (int32_t)w;
        st.r[i] = word;
    }
    st.f = 3;
    st.b = 0;
    for (int i = 0; i < 310; ++i) {
        st.r[st.f] = (int32_t)((uint32_t)st.r[st.f] + (uint32_t)st.r[st.b]);
        if (++st.f >= 31) st.f = 0;
        if (++st.b >= 31) st.b = 0;
    }
}

inline void host_rand_skip(LibcState &st, uint64_t n)
{
    for (uint64_t i = 0; i < n; ++i) {
        st.r[st.f] = (int32_t)((uint32_t)st.r[st.f] + (uint32_t)st.r[st.b]);
        if (++st.f >= 31) st.f = 0;
        if (++st.b >= 31) st.b = 0;
    }
}

struct Rng {
    LibcState s;
    const int32_t *stream; // recorded rand() outputs, or nullptr
    long long pos, len;
    __device__ __forceinline__ int32_t next()
    {
        if (stream) {
            const int32_t v = pos < len ? stream[pos] : 0;
            ++pos;
            return v;
        }
        const uint32_t v = (uint32_t)s.r[s.f] + (uint32_t)s.r[s.b];
        s.r[s.f] = (int32_t)v;
        if (++s.f >= 31) s.f = 0;
        if (++s.b >= 31) s.b = 0;
        ++pos;
        return (int32_t)(v >> 1);
    }
    __device__ __forceinline__ double uniform() { return __ddiv_rn((double)next(), 2147483647.0); }
};

__device__ __forceinline__ void shuffle(Rng &rng, int32_t *perm, int n)
{
    for (int i = 0; i < n; ++i) perm[i] = i;
    for (int i = n; i > 0; --i) {
        const int j = rng.next() % i;
        const int32_t t = perm[i - 1];
        perm[i - 1] = perm[j];
        perm[j] = t;
    }
}

// in-plane accumulation of one visit, qmc.pyx:112-125
__device__ __forceinline__ double qmc_inplane(const int8_t *conf, int P, const int32_t *tab_idx,
                                              const double *tab_J, int maxnb, int ispin, int islice,
                                              double b_coeff, double acc)
{
    const double bs = __dmul_rn(b_coeff, (double)conf[(long long)ispin * P + islice]);
    for (int si = 0; si < maxnb; ++si) {
        const int spinidx = tab_idx[(long long)ispin * maxnb + si];
        const double jval = tab_J[(long long)ispin * maxnb + si];
        if (spinidx == ispin)
            acc = __dadd_rn(acc, __dmul_rn(bs, jval));
        else
            acc = __dadd_rn(acc, __dmul_rn(bs, __dmul_rn(jval, (double)conf[(long long)spinidx * P + islice])));
    }
    return acc;
}

// full local ediff, qmc.pyx:112-138 (+ bath term :268-273 when lut != nullptr)
__device__ __forceinline__ double qmc_ediff(const int8_t *conf, int P, const int32_t *tab_idx, const double *tab_J,
                                            int maxnb, int ispin, int islice, double b_coeff, double jperp,
                                            double teff, const double *lut)
{
    const double s = (double)conf[(long long)ispin * P + islice];
    double e = qmc_inplane(conf, P, tab_idx, tab_J, maxnb, ispin, islice, b_coeff, 0.0);
    int tleft, tright;
    if (islice == 0) {
        tleft = P - 1;
        tright = 1;
    } else if (islice == P - 1) {
        tleft = P - 2;
        tright = 0;
    } else {
        tleft = islice - 1;
        tright = islice + 1;
    }
    const double s2 = __dmul_rn(2.0, s);
    e = __dadd_rn(e, __dmul_rn(s2, __dmul_rn(jperp, (double)conf[(long long)ispin * P + tleft])));
    e = __dadd_rn(e, __dmul_rn(s2, __dmul_rn(jperp, (double)conf[(long long)ispin * P + tright])));
    if (lut) {
        const double t2 = __dmul_rn(2.0, teff);
        for (int k = 1; k < P; ++k) {
            const int bslice = (islice + k) % P;
            const double ss = (double)((int)conf[(long long)ispin * P + islice] * (int)conf[(long long)ispin * P + bslice]);
            e = __dadd_rn(e, __dmul_rn(__dmul_rn(t2, ss), lut[k - 1]));
        }
    }
    return e;
}

struct ExactQmcArgs {
    int8_t *confs;         // [R][N][P]
    int32_t *perm;         // [R][N] scratch
    const LibcState *st;   // [R]
    const int32_t *stream; // [R][stream_len] or nullptr
    long long stream_len;
    long long *consumed; // [R] or nullptr
    const double *jperp;  // [S]
    const double *bcoef;  // [S]
    const double *lut;    // [P-1] or nullptr
    const int32_t *tab_idx;
    const double *tab_J;
    long long R;
    int N, P, maxnb, S, mcsteps, global_moves;
    double teff;
};

__global__ void exact_qmc_kernel(const ExactQmcArgs a)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.R) return;
    int8_t *conf = a.confs + r * (long long)a.N * a.P;
    int32_t *perm = a.perm + r * (long long)a.N;
    Rng rng;
    rng.s = a.st[r];
    rng.stream = a.stream ? a.stream + r * a.stream_len : nullptr;
    rng.pos = 0;
    rng.len = a.stream_len;
    const int P = a.P;
    for (int f = 0; f < a.S; ++f) {
        const double jperp = a.jperp[f], b_coeff = a.bcoef[f];
        for (int step = 0; step < a.mcsteps; ++step) {
            for (int islice = 0; islice < P; ++islice) {
                shuffle(rng, perm, a.N);
                for (int sidx = 0; sidx < a.N; ++sidx) {
                    const int ispin = perm[sidx];
                    const double e = qmc_ediff(conf, P, a.tab_idx, a.tab_J, a.maxnb, ispin, islice, b_coeff,
                                               jperp, a.teff, a.lut);
                    bool flip = e <= 0.0;
                    if (!flip) flip = exp(__ddiv_rn(__dmul_rn(-1.0, e), a.teff)) > rng.uniform();
                    if (flip) conf[(long long)ispin * P + islice] = -conf[(long long)ispin * P + islice];
                }
            }
            if (a.global_moves) { // qmc.pyx:405-438
                shuffle(rng, perm, a.N);
                for (int sidx = 0; sidx < a.N; ++sidx) {
                    const int ispin = perm[sidx];
                    double e = 0.0;
                    for (int k = 0; k < P; ++k)
                        e = qmc_inplane(conf, P, a.tab_idx, a.tab_J, a.maxnb, ispin, k, b_coeff, e);
                    bool flip = e <= 0.0;
                    if (!flip) flip = exp(__ddiv_rn(__dmul_rn(-1.0, e), a.teff)) > rng.uniform();
                    if (flip)
                        for (int k = 0; k < P; ++k) conf[(long long)ispin * P + k] = -conf[(long long)ispin * P + k];
                }
            }
        }
    }
    if (a.consumed) a.consumed[r] = rng.pos;
}

struct ExactSaArgs {
    int8_t *svec; // [R][N]
    int32_t *perm;
    const LibcState *st;
    long long *consumed;
    const double *sched;   // [S]
    const double *randuni; // [S][mcsteps][N] or nullptr
    const int32_t *tab_idx;
    const double *tab_J;
    long long R;
    long long tab_stride; // elements between the tables of consecutive schedule steps (0: static table)
    int N, maxnb, S, mcsteps;
};

__global__ void exact_sa_kernel(const ExactSaArgs a)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.R) return;
    int8_t *sv = a.svec + r * (long long)a.N;
    int32_t *perm = a.perm + r * (long long)a.N;
    Rng rng;
    rng.s = a.st[r];
    rng.stream = nullptr;
    rng.pos = 0;
    rng.len = 0;
    for (int t = 0; t < a.S; ++t) {
        const double temp = a.sched[t];
        const int32_t *tab_idx = a.tab_idx + t * a.tab_stride; // NoisyAnneal: nbs[itemp] (sa.pyx:363-365)
        const double *tab_J = a.tab_J + t * a.tab_stride;
        for (int step = 0; step < a.mcsteps; ++step) {
            shuffle(rng, perm, a.N);
            for (int ispin = 0; ispin < a.N; ++ispin) {
                const int sidx = perm[ispin];
                const double m2s = __dmul_rn(-2.0, (double)sv[sidx]);
                double e = 0.0;
                for (int si = 0; si < a.maxnb; ++si) { // sa.pyx:84-94
                    const int spinidx = tab_idx[(long long)sidx * a.maxnb + si];
                    const double jval = tab_J[(long long)sidx * a.maxnb + si];
                    if (spinidx == sidx)
                        e = __dadd_rn(e, __dmul_rn(m2s, jval));
                    else
                        e = __dadd_rn(e, __dmul_rn(m2s, __dmul_rn(jval, (double)sv[spinidx])));
                }
                bool flip = e <= 0.0;
                if (!flip) {
                    const double u = a.randuni ? a.randuni[((long long)t * a.mcsteps + step) * a.N + ispin]
                                               : rng.uniform();
                    flip = exp(__ddiv_rn(__dmul_rn(-1.0, e), temp)) > u;
                }
                if (flip) sv[sidx] = -sv[sidx];
            }
        }
    }
    if (a.consumed) a.consumed[r] = rng.pos;
}

struct ExactSvmcArgs {
    double *svec; // [R][N]
    int32_t *perm;
    const LibcState *st;
    const double *A, *B;   // [S]
    const double *randuni; // [S][mcsteps][N][2] or nullptr (TFCompact: rand()-driven)
    const int32_t *tab_idx;
    const double *tab_J;
    long long R;
    long long tab_stride; // elements between the tables of consecutive schedule steps (0: static table)
    int N, maxnb, S, mcsteps, tf;
    int serial; // 1: a single thread walks all reads with ONE stream (svmc.pyx:624-674)
    double temp;
};

__device__ void exact_svmc_read(const ExactSvmcArgs &a, double *sv, int32_t *perm, Rng &rng)
{
    const double pi = 3.141592653589793;
    for (int f = 0; f < a.S; ++f) {
        const double a_coeff = a.A[f], b_coeff = a.B[f];
        const int32_t *tab_idx = a.tab_idx + f * a.tab_stride; // NoisySVMC: nbs[ifield] (svmc.pyx:317-319)
        const double *tab_J = a.tab_J + f * a.tab_stride;
        for (int step = 0; step < a.mcsteps; ++step) {
            shuffle(rng, perm, a.N);
            for (int ispin = 0; ispin < a.N; ++ispin) {
                const int sidx = perm[ispin];
                const double *ru = a.randuni ? a.randuni + (((long long)f * a.mcsteps + step) * a.N + ispin) * 2
                                             : nullptr;
                double theta_prop;
                if (!a.tf) {
                    theta_prop = __dmul_rn(pi, ru[0]); // svmc.pyx:95
                } else {
                    const double ab_ratio = __ddiv_rn(a_coeff, b_coeff);
                    // np.random form (svmc.pyx:198-202): (2.0*pi*u) - pi ; rand() form (:645-647):
                    // (2.0*pi*rand()/rand_max) - pi, evaluated left to right
                    const double span = ru ? __dmul_rn(__dmul_rn(2.0, pi), ru[0])
                                           : __ddiv_rn(__dmul_rn(__dmul_rn(2.0, pi), (double)rng.next()), 2147483647.0);
                    const double d = __dadd_rn(span, -pi);
                    theta_prop = (ab_ratio > 1) ? d : __dmul_rn(ab_ratio, d);
                    theta_prop = __dadd_rn(theta_prop, sv[sidx]);
                    if (theta_prop < 0)
                        theta_prop = 0.0;
                    else if (theta_prop > pi)
                        theta_prop = pi;
                }
                const double zmagdiff = __dadd_rn(cos(theta_prop), -cos(sv[sidx]));
                double e = 0.0;
                for (int si = 0; si < a.maxnb; ++si) { // svmc.pyx:98-108
                    const int spinidx = tab_idx[(long long)sidx * a.maxnb + si];
                    const double jval = tab_J[(long long)sidx * a.maxnb + si];
                    const double bjz = __dmul_rn(__dmul_rn(b_coeff, jval), zmagdiff);
                    if (spinidx == sidx)
                        e = __dadd_rn(e, bjz);
                    else
                        e = __dadd_rn(e, __dmul_rn(bjz, cos(sv[spinidx])));
                }
                e = __dadd_rn(e, __dmul_rn(a_coeff, __dadd_rn(sin(sv[sidx]), -sin(theta_prop))));
                bool acc = e <= 0.0;
                if (!acc) {
                    const double u = ru ? ru[1] : rng.uniform();
                    acc = exp(__ddiv_rn(__dmul_rn(-1.0, e), a.temp)) > u;
                }
                if (acc) sv[sidx] = theta_prop;
            }
        }
    }
}

__global__ void exact_svmc_kernel(const ExactSvmcArgs a)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a.serial) {
        if (r != 0) return;
        Rng rng;
        rng.s = a.st[0];
        rng.stream = nullptr;
        rng.pos = 0;
        rng.len = 0;
        for (long long q = 0; q < a.R; ++q) exact_svmc_read(a, a.svec + q * (long long)a.N, a.perm, rng);
        return;
    }
    if (r >= a.R) return;
    Rng rng;
    rng.s = a.st[r];
    rng.stream = nullptr;
    rng.pos = 0;
    rng.len = 0;
    exact_svmc_read(a, a.svec + r * (long long)a.N, a.perm + r * (long long)a.N, rng);
}

// ---- probes ---------------------------------------------------------------------------------
__global__ void probe_qmc_delta_e_kernel(const int8_t *confs, const int32_t *tab_idx, const double *tab_J,
                                         double *out, long long R, int N, int P, int maxnb, double b_coeff,
                                         double jperp, double teff)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= R * N * P) return;
    const long long r = t / ((long long)N * P);
    const int i = (int)((t / P) % N), k = (int)(t % P);
    out[t] = qmc_ediff(confs + r * (long long)N * P, P, tab_idx, tab_J, maxnb, i, k, b_coeff, jperp, teff, nullptr);
}

__global__ void probe_qmc_delta_e_global_kernel(const int8_t *confs, const int32_t *tab_idx, const double *tab_J,
                                                double *out, long long R, int N, int P, int maxnb, double b_coeff)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= R * N) return;
    const long long r = t / N;
    const int i = (int)(t % N);
    double e = 0.0;
    for (int k = 0; k < P; ++k) e = qmc_inplane(confs + r * (long long)N * P, P, tab_idx, tab_J, maxnb, i, k, b_coeff, e);
    out[t] = e;
}

__global__ void probe_sa_delta_e_kernel(const int8_t *svec, const int32_t *tab_idx, const double *tab_J,
                                        double *out, long long R, int N, int maxnb)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= R * N) return;
    const int8_t *sv = svec + (t / N) * (long long)N;
    const int i = (int)(t % N);
    const double m2s = __dmul_rn(-2.0, (double)sv[i]);
    double e = 0.0;
    for (int si = 0; si < maxnb; ++si) {
        const int spinidx = tab_idx[(long long)i * maxnb + si];
        const double jval = tab_J[(long long)i * maxnb + si];
        if (spinidx == i)
            e = __dadd_rn(e, __dmul_rn(m2s, jval));
        else
            e = __dadd_rn(e, __dmul_rn(m2s, __dmul_rn(jval, (double)sv[spinidx])));
    }
    out[t] = e;
}

// RAII device buffer for the one-shot exact / probe calls
// ------------------------------------------------------------------------------------------------------------
// Wolff-cluster experiments of the reference, qmc.pyx:612-1621 ("Function under test"): QuantumAnnealWCL (:620-786),
// DissaptiveQuantumAnnealWCL (:792-1000), QuantumAnnealWC (:1006-1225), DissipativeQuantumAnnealWC2 (:1231-1446),
// DissipativeQuantumAnnealWC3 (:1452-1621).  Single-cluster growth with an explicit stack is sequential by
// construction, so these are served by the replay path only: one thread per replica, the reference's rand() stream,
// fp64 in the reference's association order.  They are replayed AS WRITTEN (seed spin not flipped in WCL / WC / WC2,
// padded table rows walked in full, `spinidx` / `bslice` / `jval` carried over between loops where the reference
// does so, WC2 / WC3's inverted final tests): the bar is identical results, not a better algorithm -- the working
// cluster move of this library is Swendsen-Wang (mcs_cluster.cu).
// ------------------------------------------------------------------------------------------------------------
struct ExactWolffArgs {
    int8_t *confs;       // [R][N][P]
    int32_t *perm;       // [R][N]        (WC2, WC3)
    int32_t *cl;         // [R][N P + 2]  explicit stack, node = spin * P + slice
    const LibcState *st; // [R]
    long long *consumed; // [R]
    int *overrun;        // [R] 1: the reference would have written past its `cluster` buffer (it does not check)
    const double *jperp; // [S]
    const double *bcoef; // [S]  +B (qmc.pyx:696)
    const double *lut;   // [P-1] or nullptr
    const int32_t *tab_idx;
    const double *tab_J;
    long long R;
    int N, P, maxnb, S, mcsteps, variant;
    double teff;
};

struct WolffCtx {
    const ExactWolffArgs &a;
    int8_t *conf;
    int32_t *cl;
    Rng &rng;
    int stack, stackidx, cluster_count, max_rows;
    double r;

    __device__ __forceinline__ int8_t &cf(int spin, int slice) { return conf[(long long)spin * a.P + slice]; }
    __device__ __forceinline__ int idx(int s, int si) const { return a.tab_idx[(long long)s * a.maxnb + si]; }
    __device__ __forceinline__ double J(int s, int si) const { return a.tab_J[(long long)s * a.maxnb + si]; }
    __device__ __forceinline__ void start(int spin, int slice)
    {
        cl[0] = spin * a.P + slice;
        stack = 1, stackidx = 1, cluster_count = 0, r = 1.0;
    }
    __device__ __forceinline__ void push(int spin, int slice) // qmc.pyx:731-736
    {
        cl[stackidx] = spin * a.P + slice;
        if (stackidx > max_rows) max_rows = stackidx;
        cf(spin, slice) = -cf(spin, slice);
        stack += 1;
        stackidx += 1;
    }
    // growth attempt, qmc.pyx:726-736
    __device__ __forceinline__ void attempt(int spin, int slice, double ediff, bool update_r)
    {
        if (ediff < 0) {
            const double p = __dadd_rn(1.0, -exp(__ddiv_rn(ediff, a.teff)));
            if (__dmul_rn(r, p) > rng.uniform()) {
                if (update_r) r = __dmul_rn(r, p);
                push(spin, slice);
            }
        }
    }
    // "add bias energy", qmc.pyx:722-725
    __device__ __forceinline__ double bias(int s, double b_coeff, int k, double ediff) const
    {
        for (int si2 = 0; si2 < a.maxnb; ++si2)
            if (idx(s, si2) == s)
                ediff = __dadd_rn(ediff, __dmul_rn(__dmul_rn(__dmul_rn(-2.0, b_coeff), J(s, si2)), (double)k));
        return ediff;
    }
};

__device__ __forceinline__ void trotter_nb(int islice, int P, int &tl, int &tr)
{
    if (islice == 0) {
        tl = P - 1;
        tr = 1;
    } else if (islice == P - 1) {
        tl = P - 2;
        tr = 0;
    } else {
        tl = islice - 1;
        tr = islice + 1;
    }
}

__global__ void exact_wolff_kernel(const ExactWolffArgs a)
{
    const long long rep = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (rep >= a.R) return;
    const int N = a.N, P = a.P, maxnb = a.maxnb, variant = a.variant;
    const double teff = a.teff;
    int32_t *perm = a.perm + rep * (long long)N;
    Rng rng;
    rng.s = a.st[rep];
    rng.stream = nullptr;
    rng.pos = 0;
    rng.len = 0;
    WolffCtx w{a, a.confs + rep * (long long)N * P, a.cl + rep * ((long long)N * P + 2), rng, 0, 0, 0, 0, 1.0};
    // function-scope variables of the reference (qmc.pyx:1054-1070, 1284-1318, 1506-1540)
    int ispin = 0, islice = 0, spinidx = 0, tleft = 0, tright = 0, tleft2 = 0, tright2 = 0, bslice = 0, k = 0;
    double jval = 0.0, ediff = 0.0, e_total = 0.0;
    for (int f = 0; f < a.S; ++f) {
        const double jperp = a.jperp[f], b_coeff = a.bcoef[f];
        const double m2b = __dmul_rn(-2.0, b_coeff), p2j = __dmul_rn(2.0, jperp), p2t = __dmul_rn(2.0, teff);
        for (int step = 0; step < a.mcsteps; ++step) {
            if (variant <= 2) { // one single-cluster move per step
                ispin = rng.next() % N;
                islice = rng.next() % P;
                if (variant == 1) { // walk to a start point that aligns with the local field, qmc.pyx:879-893
                    const int j = islice * N + ispin;
                    for (int i = 1; i < N * P; ++i) {
                        ediff = 0.0;
                        for (int si = 0; si < maxnb; ++si) {
                            spinidx = w.idx(ispin, si);
                            jval = w.J(ispin, si);
                            if (spinidx == ispin)
                                ediff = __dadd_rn(ediff, __dmul_rn(__dmul_rn(__dmul_rn(-2.0, jval), b_coeff),
                                                                   (double)w.cf(ispin, islice)));
                        }
                        if (ediff <= 0) break;
                        if (exp(__ddiv_rn(__dmul_rn(-1.0, ediff), teff)) > rng.uniform()) break;
                        ispin = (j + i) % N;
                        islice = ((j + i) / N) % P;
                    }
                }
                w.start(ispin, islice);
                k = w.cf(ispin, islice);
                if (variant == 1) w.cf(ispin, islice) = -w.cf(ispin, islice); // qmc.pyx:898 (commented out at :704)
                for (;;) {
                    ispin = w.cl[w.cluster_count] / P;
                    islice = w.cl[w.cluster_count] % P;
                    if (variant == 1) { // bath neighbours first, qmc.pyx:906-925
                        for (int b = 1; b < P; ++b) {
                            bslice = (islice + b) % P;
                            if (w.cf(ispin, bslice) == k) {
                                ediff = __dadd_rn(0.0, __dmul_rn(__dmul_rn(-2.0, teff), a.lut[b - 1]));
                                w.attempt(ispin, bslice, w.bias(ispin, b_coeff, k, ediff), true);
                            }
                        }
                    }
                    if (variant != 2) { // qmc.pyx:708-781 / 926-995
                        for (int si = 0; si < maxnb; ++si) {
                            spinidx = w.idx(ispin, si);
                            if (w.cf(spinidx, islice) == k) {
                                ediff = __dadd_rn(0.0, __dmul_rn(__dmul_rn(2.0, b_coeff), w.J(ispin, si)));
                                w.attempt(spinidx, islice, w.bias(spinidx, b_coeff, k, ediff), true);
                            }
                        }
                        trotter_nb(islice, P, tleft, tright);
                        if (w.cf(ispin, tleft) == k) {
                            ediff = __dadd_rn(0.0, __dmul_rn(-2.0, jperp));
                            w.attempt(ispin, tleft, w.bias(ispin, b_coeff, k, ediff), true);
                        }
                        if (w.cf(ispin, tright) == k) {
                            ediff = __dadd_rn(0.0, __dmul_rn(-2.0, jperp));
                            w.attempt(ispin, tright, w.bias(ispin, b_coeff, k, ediff), true);
                        }
                    } else { // QuantumAnnealWC: energy change of the candidate, qmc.pyx:1112-1220
                        trotter_nb(islice, P, tleft, tright);
                        for (int side = 0; side < 2; ++side) {
                            const int ts = side == 0 ? tleft : tright;
                            if (w.cf(ispin, ts) == k) {
                                ediff = 0.0;
                                for (int si2 = 0; si2 < maxnb; ++si2) {
                                    const int spinidx2 = w.idx(ispin, si2);
                                    jval = w.J(spinidx, si2); // `spinidx` is left over from the last spatial loop
                                    const double t = __dmul_rn(__dmul_rn(m2b, jval), (double)k);
                                    ediff = __dadd_rn(ediff, spinidx == spinidx2 ? t : __dmul_rn(t, (double)w.cf(spinidx2, ts)));
                                }
                                trotter_nb(ts, P, tleft2, tright2);
                                ediff = __dadd_rn(ediff, __dmul_rn(__dmul_rn(p2j, (double)k), (double)w.cf(ispin, tleft2)));
                                ediff = __dadd_rn(ediff, __dmul_rn(__dmul_rn(p2j, (double)k), (double)w.cf(ispin, tright2)));
                                w.attempt(ispin, ts, ediff, false);
                            }
                        }
                        for (int si = 0; si < maxnb; ++si) {
                            spinidx = w.idx(ispin, si);
                            if (w.cf(spinidx, islice) == k) {
                                ediff = 0.0;
                                for (int si2 = 0; si2 < maxnb; ++si2) {
                                    const int spinidx2 = w.idx(spinidx, si2);
                                    jval = w.J(spinidx, si2);
                                    const double t = __dmul_rn(__dmul_rn(m2b, jval), (double)k);
                                    ediff = __dadd_rn(ediff, spinidx == spinidx2 ? t : __dmul_rn(t, (double)w.cf(spinidx2, islice)));
                                }
                                trotter_nb(islice, P, tleft2, tright2);
                                ediff = __dadd_rn(ediff, __dmul_rn(__dmul_rn(p2j, (double)k), (double)w.cf(spinidx, tleft2)));
                                ediff = __dadd_rn(ediff, __dmul_rn(__dmul_rn(p2j, (double)k), (double)w.cf(spinidx, tright2)));
                                w.attempt(spinidx, islice, ediff, false);
                            }
                        }
                    }
                    w.cluster_count += 1;
                    w.stack -= 1;
                    if (w.stack == 0) break;
                }
            } else if (variant == 3) {
                // local sweep whose bath term starts from a stale slice (qmc.pyx:1325-1375) ...
                for (int loop_slice = 0; loop_slice < P; ++loop_slice) {
                    islice = loop_slice;
                    shuffle(rng, perm, N);
                    for (int sidx = 0; sidx < N; ++sidx) {
                        ispin = perm[sidx];
                        const double s = (double)w.cf(ispin, islice), m2bs = __dmul_rn(m2b, s);
                        double e = 0.0;
                        for (int si = 0; si < maxnb; ++si) {
                            spinidx = w.idx(ispin, si);
                            jval = w.J(ispin, si);
                            if (spinidx == ispin)
                                e = __dadd_rn(e, __dmul_rn(m2bs, jval));
                            else
                                e = __dadd_rn(e, __dmul_rn(m2bs, __dmul_rn(jval, (double)w.cf(spinidx, islice))));
                        }
                        trotter_nb(islice, P, tleft, tright);
                        e = __dadd_rn(e, __dmul_rn(__dmul_rn(2.0, s), __dmul_rn(jperp, (double)w.cf(ispin, tleft))));
                        e = __dadd_rn(e, __dmul_rn(__dmul_rn(2.0, s), __dmul_rn(jperp, (double)w.cf(ispin, tright))));
                        for (int b2 = 1; b2 < P; ++b2) {
                            const int cslice = (bslice + b2) % P; // `bslice`, not islice (qmc.pyx:1364)
                            const double ss = (double)((int)w.cf(ispin, islice) * (int)w.cf(ispin, cslice));
                            e = __dadd_rn(e, __dmul_rn(__dmul_rn(p2t, ss), a.lut[b2 - 1]));
                        }
                        bool flip = e <= 0.0;
                        if (!flip) flip = exp(__ddiv_rn(__dmul_rn(-1.0, e), teff)) > rng.uniform();
                        if (flip) w.cf(ispin, islice) = -w.cf(ispin, islice);
                    }
                }
                // ... then one bath-only cluster per spin, Metropolis on its accumulated energy (qmc.pyx:1376-1446)
                shuffle(rng, perm, N);
                for (int sidx2 = 0; sidx2 < N; ++sidx2) {
                    ispin = perm[sidx2];
                    islice = rng.next() % P;
                    w.start(ispin, islice);
                    k = w.cf(ispin, islice);
                    e_total = 0.0;
                    for (;;) {
                        ispin = w.cl[w.cluster_count] / P;
                        islice = w.cl[w.cluster_count] % P;
                        for (int b = 1; b < P; ++b) {
                            bslice = (islice + b) % P;
                            if (w.cf(ispin, bslice) != k) continue;
                            const double p = __dadd_rn(1.0, -exp(__dmul_rn(-2.0, a.lut[b - 1])));
                            if (!(__dmul_rn(w.r, p) > rng.uniform())) continue;
                            ediff = 0.0;
                            for (int si2 = 0; si2 < maxnb; ++si2) {
                                const int spinidx2 = w.idx(ispin, si2);
                                if (ispin == spinidx2)
                                    ediff = __dadd_rn(ediff, __dmul_rn(__dmul_rn(m2b, w.J(ispin, si2)), (double)k));
                                else // `jval` is whatever the local sweep left behind (qmc.pyx:1414)
                                    ediff = __dadd_rn(ediff, __dmul_rn(__dmul_rn(__dmul_rn(m2b, jval), (double)k),
                                                                       (double)w.cf(spinidx2, bslice)));
                            }
                            trotter_nb(bslice, P, tleft2, tright2);
                            ediff = __dadd_rn(ediff, __dmul_rn(__dmul_rn(p2j, (double)k), (double)w.cf(ispin, tleft2)));
                            ediff = __dadd_rn(ediff, __dmul_rn(__dmul_rn(p2j, (double)k), (double)w.cf(ispin, tright2)));
                            for (int b2 = 1; b2 < P; ++b2) {
                                const int cslice = (bslice + b2) % P;
                                ediff = __dadd_rn(ediff, __dmul_rn(__dmul_rn(p2t, (double)(k * (int)w.cf(ispin, cslice))),
                                                                   a.lut[b2 - 1]));
                            }
                            w.r = __dmul_rn(w.r, p);
                            e_total = __dadd_rn(e_total, ediff);
                            w.push(ispin, bslice);
                        }
                        w.cluster_count += 1;
                        w.stack -= 1;
                        if (w.stack == 0) break;
                    }
                    if (e_total > 0 && exp(__ddiv_rn(__dmul_rn(-1.0, e_total), teff)) > rng.uniform())
                        for (int i = 1; i < w.cluster_count; ++i) w.conf[w.cl[i]] = -w.conf[w.cl[i]]; // node = flat index
                }
            } else { // DissipativeQuantumAnnealWC3: N P bath clusters per step, qmc.pyx:1546-1621
                shuffle(rng, perm, N);
                for (int loop_slice = 0; loop_slice < P; ++loop_slice) {
                    islice = loop_slice; // the body overwrites islice; Cython iterates on a temporary
                    for (int sidx2 = 0; sidx2 < N; ++sidx2) {
                        e_total = 0.0;
                        ispin = perm[sidx2];
                        w.start(ispin, islice);
                        k = w.cf(ispin, islice);
                        w.cf(ispin, islice) = -w.cf(ispin, islice);
                        const double m2bk = __dmul_rn(m2b, (double)k), p2k = __dmul_rn(2.0, (double)k);
                        for (;;) {
                            ispin = w.cl[w.cluster_count] / P;
                            islice = w.cl[w.cluster_count] % P;
                            for (int si = 0; si < maxnb; ++si) {
                                spinidx = w.idx(ispin, si);
                                jval = w.J(ispin, si);
                                if (spinidx == ispin)
                                    e_total = __dadd_rn(e_total, __dmul_rn(m2bk, jval));
                                else
                                    e_total = __dadd_rn(e_total, __dmul_rn(m2bk, __dmul_rn(jval, (double)w.cf(spinidx, islice))));
                            }
                            trotter_nb(islice, P, tleft, tright);
                            e_total = __dadd_rn(e_total, __dmul_rn(p2k, __dmul_rn(jperp, (double)w.cf(ispin, tleft))));
                            e_total = __dadd_rn(e_total, __dmul_rn(p2k, __dmul_rn(jperp, (double)w.cf(ispin, tright))));
                            for (int b = 1; b < P; ++b) {
                                bslice = (islice + b) % P;
                                if (w.cf(ispin, bslice) != k) continue;
                                const double p = __dadd_rn(1.0, -exp(__dmul_rn(-2.0, a.lut[b - 1])));
                                if (__dmul_rn(w.r, p) > rng.uniform()) {
                                    w.r = __dmul_rn(w.r, p);
                                    w.push(ispin, bslice);
                                }
                            }
                            w.cluster_count += 1;
                            w.stack -= 1;
                            if (w.stack == 0) break;
                        }
                        if (e_total > 0.0 &&
                            __dadd_rn(1.0, -exp(__ddiv_rn(__dmul_rn(-1.0, e_total), teff))) > rng.uniform())
                            for (int i = 0; i < w.cluster_count; ++i) w.conf[w.cl[i]] = -w.conf[w.cl[i]];
                    }
                }
            }
        }
    }
    a.consumed[rep] = rng.pos;
    a.overrun[rep] = w.max_rows >= ((variant == 3 || variant == 4) ? P : N * P) ? 1 : 0;
}

struct DevBuf {
    void *p = nullptr;
    ~DevBuf()
    {
        if (p) cudaFree(p);
    }
    int alloc(size_t bytes)
    {
        MCS_CUDA(cudaMalloc(&p, bytes ? bytes : 1));
        return MCS_OK;
    }
    int put(const void *src, size_t bytes, cudaStream_t s)
    {
        MCS_TRY(alloc(bytes));
        if (bytes) MCS_CUDA(cudaMemcpyAsync(p, src, bytes, cudaMemcpyHostToDevice, s));
        return MCS_OK;
    }
    template <typename T>
    T *as()
    {
        return (T *)p;
    }
};

int make_states(const uint32_t *seeds, int64_t R, std::vector<LibcState> &out)
{
    out.resize((size_t)R);
    for (int64_t r = 0; r < R; ++r) host_srand(out[r], seeds[r]);
    return MCS_OK;
}

} // namespace

extern "C" int mcs_exact_qmc(mcs_instance *inst, const double *A, const double *B, int64_t S, int mcsteps, float temp,
                             const double *lookuptable, int8_t *confs, int64_t R, int64_t P, int global_moves,
                             const uint32_t *libc_seeds, const int32_t *rand_stream, int64_t stream_len,
                             int64_t *consumed)
{
    MCS_REQUIRE(inst && confs && R > 0 && (S == 0 || (A && B)), MCS_EINVAL, "mcs_exact_qmc: bad argument");
    MCS_REQUIRE(P >= 2, MCS_EINVAL, "mcs_exact_qmc: P >= 2 required (P=1 reads out of bounds in the reference)");
    MCS_REQUIRE(libc_seeds || rand_stream, MCS_EINVAL, "mcs_exact_qmc: need libc_seeds or rand_stream");
    MCS_REQUIRE(inst->nsteps == 1, MCS_EUNSUPPORTED, "mcs_exact_qmc: time-dependent tables are an SA / SVMC feature");
    const double teff = (double)temp * (double)P;
    MCS_REQUIRE(teff != 0.0 || S == 0, MCS_EZERODIV, "float division");
    MCS_CUDA(cudaSetDevice(inst->device));
    cudaStream_t s = inst->stream;
    std::vector<double> jperp((size_t)S), bcoef((size_t)S);
    for (int64_t f = 0; f < S; ++f) {
        jperp[f] = -0.5 * teff * log(tanh(A[f] / teff)); // qmc.pyx:95, host libm like the reference
        bcoef[f] = -2.0 * B[f];                          // qmc.pyx:96
    }
    std::vector<LibcState> states;
    if (libc_seeds)
        make_states(libc_seeds, R, states);
    else
        states.assign((size_t)R, LibcState());
    DevBuf d_conf, d_perm, d_st, d_stream, d_cons, d_jp, d_bc, d_lut;
    const size_t cbytes = (size_t)R * inst->N * P;
    MCS_TRY(d_conf.put(confs, cbytes, s));
    MCS_TRY(d_perm.alloc((size_t)R * inst->N * sizeof(int32_t)));
    MCS_TRY(d_st.put(states.data(), states.size() * sizeof(LibcState), s));
    if (rand_stream) MCS_TRY(d_stream.put(rand_stream, (size_t)R * stream_len * sizeof(int32_t), s));
    MCS_TRY(d_cons.alloc((size_t)R * sizeof(long long)));
    MCS_TRY(d_jp.put(jperp.data(), jperp.size() * sizeof(double), s));
    MCS_TRY(d_bc.put(bcoef.data(), bcoef.size() * sizeof(double), s));
    if (lookuptable) MCS_TRY(d_lut.put(lookuptable, (size_t)(P - 1) * sizeof(double), s));
    ExactQmcArgs a;
    a.confs = d_conf.as<int8_t>();
    a.perm = d_perm.as<int32_t>();
    a.st = d_st.as<LibcState>();
    a.stream = rand_stream ? d_stream.as<int32_t>() : nullptr;
    a.stream_len = stream_len;
    a.consumed = d_cons.as<long long>();
    a.jperp = d_jp.as<double>();
    a.bcoef = d_bc.as<double>();
    a.lut = lookuptable ? d_lut.as<double>() : nullptr;
    a.tab_idx = inst->d_tab_idx;
    a.tab_J = inst->d_tab_J;
    a.R = R;
    a.N = (int)inst->N;
    a.P = (int)P;
    a.maxnb = (int)inst->maxnb;
    a.S = (int)S;
    a.mcsteps = mcsteps;
    a.global_moves = global_moves ? 1 : 0;
    a.teff = teff;
    exact_qmc_kernel<<<(unsigned)((R + 31) / 32), 32, 0, s>>>(a);
    inst->launches++;
    MCS_CUDA(cudaGetLastError());
    MCS_CUDA(cudaMemcpyAsync(confs, d_conf.p, cbytes, cudaMemcpyDeviceToHost, s));
    if (consumed)
        MCS_CUDA(cudaMemcpyAsync(consumed, d_cons.p, (size_t)R * sizeof(long long), cudaMemcpyDeviceToHost, s));
    MCS_CUDA(cudaStreamSynchronize(s));
    return MCS_OK;
}

extern "C" int mcs_exact_qmc_wolff(mcs_instance *inst, int variant, const double *A, const double *B, int64_t S,
                                   int mcsteps, float temp, const double *lookuptable, int8_t *confs, int64_t R,
                                   int64_t P, const uint32_t *libc_seeds, int64_t *consumed, int32_t *overrun)
{
    MCS_REQUIRE(inst && confs && R > 0 && libc_seeds && (S == 0 || (A && B)), MCS_EINVAL,
                "mcs_exact_qmc_wolff: bad argument");
    MCS_REQUIRE(variant >= MCS_WOLFF_WCL && variant <= MCS_WOLFF_DISS_WC3, MCS_EINVAL,
                "mcs_exact_qmc_wolff: unknown variant %d", variant);
    MCS_REQUIRE(P >= 2, MCS_EINVAL, "mcs_exact_qmc_wolff: P >= 2 required (the reference indexes slice 1)");
    const bool bath = variant == MCS_WOLFF_DISS_WCL || variant == MCS_WOLFF_DISS_WC2 || variant == MCS_WOLFF_DISS_WC3;
    MCS_REQUIRE(!bath || lookuptable, MCS_EINVAL, "mcs_exact_qmc_wolff: this variant needs lookuptable[P-1]");
    MCS_REQUIRE(inst->nsteps == 1, MCS_EUNSUPPORTED, "mcs_exact_qmc_wolff: static tables only");
    MCS_REQUIRE(inst->N * P < (1ll << 30), MCS_EUNSUPPORTED, "mcs_exact_qmc_wolff: N * P too large");
    const double teff = (double)temp * (double)P;
    MCS_REQUIRE(teff != 0.0 || S == 0, MCS_EZERODIV, "float division");
    MCS_CUDA(cudaSetDevice(inst->device));
    cudaStream_t s = inst->stream;
    std::vector<double> jperp((size_t)S), bcoef((size_t)S);
    for (int64_t f = 0; f < S; ++f) {
        jperp[f] = -0.5 * teff * log(tanh(A[f] / teff)); // qmc.pyx:695, host libm like the reference
        bcoef[f] = B[f];                                 // qmc.pyx:696: +B in these functions
    }
    std::vector<LibcState> states;
    make_states(libc_seeds, R, states);
    DevBuf d_conf, d_perm, d_cl, d_st, d_cons, d_over, d_jp, d_bc, d_lut;
    const size_t cbytes = (size_t)R * inst->N * P;
    MCS_TRY(d_conf.put(confs, cbytes, s));
    MCS_TRY(d_perm.alloc((size_t)R * inst->N * sizeof(int32_t)));
    MCS_TRY(d_cl.alloc((size_t)R * ((size_t)inst->N * P + 2) * sizeof(int32_t)));
    MCS_TRY(d_st.put(states.data(), states.size() * sizeof(LibcState), s));
    MCS_TRY(d_cons.alloc((size_t)R * sizeof(long long)));
    MCS_TRY(d_over.alloc((size_t)R * sizeof(int)));
    MCS_TRY(d_jp.put(jperp.data(), jperp.size() * sizeof(double), s));
    MCS_TRY(d_bc.put(bcoef.data(), bcoef.size() * sizeof(double), s));
    if (bath) MCS_TRY(d_lut.put(lookuptable, (size_t)(P - 1) * sizeof(double), s));
    ExactWolffArgs a;
    a.confs = d_conf.as<int8_t>();
    a.perm = d_perm.as<int32_t>();
    a.cl = d_cl.as<int32_t>();
    a.st = d_st.as<LibcState>();
    a.consumed = d_cons.as<long long>();
    a.overrun = d_over.as<int>();
    a.jperp = d_jp.as<double>();
    a.bcoef = d_bc.as<double>();
    a.lut = bath ? d_lut.as<double>() : nullptr;
    a.tab_idx = inst->d_tab_idx;
    a.tab_J = inst->d_tab_J;
    a.R = R;
    a.N = (int)inst->N;
    a.P = (int)P;
    a.maxnb = (int)inst->maxnb;
    a.S = (int)S;
    a.mcsteps = mcsteps;
    a.variant = variant;
    a.teff = teff;
    exact_wolff_kernel<<<(unsigned)((R + 31) / 32), 32, 0, s>>>(a);
    inst->launches++;
    MCS_CUDA(cudaGetLastError());
    MCS_CUDA(cudaMemcpyAsync(confs, d_conf.p, cbytes, cudaMemcpyDeviceToHost, s));
    if (consumed)
        MCS_CUDA(cudaMemcpyAsync(consumed, d_cons.p, (size_t)R * sizeof(long long), cudaMemcpyDeviceToHost, s));
    if (overrun) MCS_CUDA(cudaMemcpyAsync(overrun, d_over.p, (size_t)R * sizeof(int), cudaMemcpyDeviceToHost, s));
    MCS_CUDA(cudaStreamSynchronize(s));
    return MCS_OK;
}

extern "C" int mcs_exact_sa(mcs_instance *inst, const double *sched, int64_t S, int mcsteps, int8_t *svec, int64_t R,
                            const uint32_t *libc_seeds, const double *randuni, int64_t *consumed)
{
    MCS_REQUIRE(inst && svec && R > 0 && libc_seeds && (S == 0 || sched), MCS_EINVAL, "mcs_exact_sa: bad argument");
    MCS_REQUIRE(inst->nsteps == 1 || S <= inst->nsteps, MCS_EINVAL, "mcs_exact_sa: schedule longer than the tables");
    MCS_CUDA(cudaSetDevice(inst->device));
    cudaStream_t s = inst->stream;
    std::vector<LibcState> states;
    make_states(libc_seeds, R, states);
    DevBuf d_sv, d_perm, d_st, d_cons, d_sched, d_ru;
    const size_t bytes = (size_t)R * inst->N;
    MCS_TRY(d_sv.put(svec, bytes, s));
    MCS_TRY(d_perm.alloc((size_t)R * inst->N * sizeof(int32_t)));
    MCS_TRY(d_st.put(states.data(), states.size() * sizeof(LibcState), s));
    MCS_TRY(d_cons.alloc((size_t)R * sizeof(long long)));
    MCS_TRY(d_sched.put(sched, (size_t)S * sizeof(double), s));
    if (randuni) MCS_TRY(d_ru.put(randuni, (size_t)S * mcsteps * inst->N * sizeof(double), s));
    ExactSaArgs a;
    a.svec = d_sv.as<int8_t>();
    a.perm = d_perm.as<int32_t>();
    a.st = d_st.as<LibcState>();
    a.consumed = d_cons.as<long long>();
    a.sched = d_sched.as<double>();
    a.randuni = randuni ? d_ru.as<double>() : nullptr;
    a.tab_idx = inst->d_tab_idx;
    a.tab_J = inst->d_tab_J;
    a.R = R;
    a.tab_stride = inst->nsteps > 1 ? inst->N * inst->maxnb : 0;
    a.N = (int)inst->N;
    a.maxnb = (int)inst->maxnb;
    a.S = (int)S;
    a.mcsteps = mcsteps;
    exact_sa_kernel<<<(unsigned)((R + 31) / 32), 32, 0, s>>>(a);
    inst->launches++;
    MCS_CUDA(cudaGetLastError());
    MCS_CUDA(cudaMemcpyAsync(svec, d_sv.p, bytes, cudaMemcpyDeviceToHost, s));
    if (consumed)
        MCS_CUDA(cudaMemcpyAsync(consumed, d_cons.p, (size_t)R * sizeof(long long), cudaMemcpyDeviceToHost, s));
    MCS_CUDA(cudaStreamSynchronize(s));
    return MCS_OK;
}

extern "C" int mcs_exact_svmc(mcs_instance *inst, const double *A, const double *B, int64_t S, int mcsteps, float temp,
                              double *svec, int64_t R, int tf, const uint32_t *libc_seeds, const double *randuni,
                              int serial_stream)
{
    MCS_REQUIRE(inst && svec && R > 0 && libc_seeds && (S == 0 || (A && B)), MCS_EINVAL,
                "mcs_exact_svmc: bad argument");
    MCS_REQUIRE(inst->nsteps == 1 || S <= inst->nsteps, MCS_EINVAL,
                "mcs_exact_svmc: schedule longer than the tables");
    MCS_REQUIRE(randuni || tf, MCS_EINVAL,
                "mcs_exact_svmc: randuni == NULL is only defined for the TF form (SpinVectorMonteCarloTFCompact)");
    MCS_CUDA(cudaSetDevice(inst->device));
    cudaStream_t s = inst->stream;
    std::vector<LibcState> states((size_t)R);
    int serial_kernel = 0;
    if (!serial_stream) {
        make_states(libc_seeds, R, states);
    } else if (randuni) {
        // Compact form: one stream through the reads; every read consumes exactly S*mcsteps*N shuffle
        // draws (acceptance uniforms come from randuni), so each read's start state is known up front.
        LibcState st;
        host_srand(st, libc_seeds[0]);
        const uint64_t per_read = (uint64_t)S * mcsteps * inst->N;
        for (int64_t r = 0; r < R; ++r) {
            states[r] = st;
            host_rand_skip(st, per_read);
        }
    } else {
        host_srand(states[0], libc_seeds[0]); // data-dependent draw counts: replay serially
        serial_kernel = 1;
    }
    DevBuf d_sv, d_perm, d_st, d_A, d_B, d_ru;
    const size_t bytes = (size_t)R * inst->N * sizeof(double);
    MCS_TRY(d_sv.put(svec, bytes, s));
    MCS_TRY(d_perm.alloc((size_t)R * inst->N * sizeof(int32_t)));
    MCS_TRY(d_st.put(states.data(), states.size() * sizeof(LibcState), s));
    MCS_TRY(d_A.put(A, (size_t)S * sizeof(double), s));
    MCS_TRY(d_B.put(B, (size_t)S * sizeof(double), s));
    if (randuni) MCS_TRY(d_ru.put(randuni, (size_t)S * mcsteps * inst->N * 2 * sizeof(double), s));
    ExactSvmcArgs a;
    a.svec = d_sv.as<double>();
    a.perm = d_perm.as<int32_t>();
    a.st = d_st.as<LibcState>();
    a.A = d_A.as<double>();
    a.B = d_B.as<double>();
    a.randuni = randuni ? d_ru.as<double>() : nullptr;
    a.tab_idx = inst->d_tab_idx;
    a.tab_J = inst->d_tab_J;
    a.R = R;
    a.tab_stride = inst->nsteps > 1 ? inst->N * inst->maxnb : 0;
    a.N = (int)inst->N;
    a.maxnb = (int)inst->maxnb;
    a.S = (int)S;
    a.mcsteps = mcsteps;
    a.tf = tf ? 1 : 0;
    a.serial = serial_kernel;
    a.temp = (double)temp; // C float in the signature (svmc.pyx:24), promoted in -1.0*ediff/temp
    exact_svmc_kernel<<<(unsigned)((R + 31) / 32), 32, 0, s>>>(a);
    inst->launches++;
    MCS_CUDA(cudaGetLastError());
    MCS_CUDA(cudaMemcpyAsync(svec, d_sv.p, bytes, cudaMemcpyDeviceToHost, s));
    MCS_CUDA(cudaStreamSynchronize(s));
    return MCS_OK;
}

extern "C" int mcs_probe_qmc_delta_e(mcs_instance *inst, double a, double b, float temp, const int8_t *confs,
                                     int64_t R, int64_t P, double *out)
{
    MCS_REQUIRE(inst && confs && out && R > 0 && P >= 2, MCS_EINVAL, "mcs_probe_qmc_delta_e: bad argument");
    const double teff = (double)temp * (double)P;
    MCS_REQUIRE(teff != 0.0, MCS_EZERODIV, "float division");
    MCS_CUDA(cudaSetDevice(inst->device));
    cudaStream_t s = inst->stream;
    const double jperp = -0.5 * teff * log(tanh(a / teff));
    const long long n = (long long)R * inst->N * P;
    DevBuf d_conf, d_out;
    MCS_TRY(d_conf.put(confs, (size_t)n, s));
    MCS_TRY(d_out.alloc((size_t)n * sizeof(double)));
    probe_qmc_delta_e_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_conf.as<int8_t>(), inst->d_tab_idx,
                                                                        inst->d_tab_J, d_out.as<double>(), R,
                                                                        (int)inst->N, (int)P, (int)inst->maxnb,
                                                                        -2.0 * b, jperp, teff);
    inst->launches++;
    MCS_CUDA(cudaGetLastError());
    MCS_CUDA(cudaMemcpyAsync(out, d_out.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
    MCS_CUDA(cudaStreamSynchronize(s));
    return MCS_OK;
}

extern "C" int mcs_probe_qmc_delta_e_global(mcs_instance *inst, double b, const int8_t *confs, int64_t R, int64_t P,
                                            double *out)
{
    MCS_REQUIRE(inst && confs && out && R > 0 && P >= 1, MCS_EINVAL, "mcs_probe_qmc_delta_e_global: bad argument");
    MCS_CUDA(cudaSetDevice(inst->device));
    cudaStream_t s = inst->stream;
    const long long n = (long long)R * inst->N;
    DevBuf d_conf, d_out;
    MCS_TRY(d_conf.put(confs, (size_t)n * P, s));
    MCS_TRY(d_out.alloc((size_t)n * sizeof(double)));
    probe_qmc_delta_e_global_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(
        d_conf.as<int8_t>(), inst->d_tab_idx, inst->d_tab_J, d_out.as<double>(), R, (int)inst->N, (int)P,
        (int)inst->maxnb, -2.0 * b);
    inst->launches++;
    MCS_CUDA(cudaGetLastError());
    MCS_CUDA(cudaMemcpyAsync(out, d_out.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
    MCS_CUDA(cudaStreamSynchronize(s));
    return MCS_OK;
}

extern "C" int mcs_probe_sa_delta_e(mcs_instance *inst, const int8_t *svec, int64_t R, double *out)
{
    MCS_REQUIRE(inst && svec && out && R > 0, MCS_EINVAL, "mcs_probe_sa_delta_e: bad argument");
    MCS_CUDA(cudaSetDevice(inst->device));
    cudaStream_t s = inst->stream;
    const long long n = (long long)R * inst->N;
    DevBuf d_sv, d_out;
    MCS_TRY(d_sv.put(svec, (size_t)n, s));
    MCS_TRY(d_out.alloc((size_t)n * sizeof(double)));
    probe_sa_delta_e_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_sv.as<int8_t>(), inst->d_tab_idx,
                                                                       inst->d_tab_J, d_out.as<double>(), R,
                                                                       (int)inst->N, (int)inst->maxnb);
    inst->launches++;
    MCS_CUDA(cudaGetLastError());
    MCS_CUDA(cudaMemcpyAsync(out, d_out.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
    MCS_CUDA(cudaStreamSynchronize(s));
    return MCS_OK;
}
