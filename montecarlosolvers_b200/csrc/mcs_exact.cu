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
#include <cstdio>
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
    long long *prof = nullptr; // MCS_EXACT_PROF=1: [R][4] cycle / window counters of the warp kernel
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

// ------------------------------------------------------------------------------------------------------------
// Warp-per-replica replay (the path mcs_exact_qmc / mcs_exact_sa take whenever a replica fits shared memory).
//
// The reference's trajectory is a strictly sequential program, but three parts of it parallelise over the 32
// lanes of a warp WITHOUT changing a single bit of the result:
//   1. glibc rand() is the additive generator x_n = x_{n-31} + x_{n-3} (mod 2^32), output x_n >> 1.  With
//      y_l = x_{n0-31+l} known for a whole block, x_{n0+l} = y_l + x_{n0+l-3}: three interleaved prefix sums,
//      i.e. 31 outputs per five warp shuffles.
//   2. The Fisher-Yates loop `for i = N..1: j = rand() % i; swap(p[i-1], p[j])` (qmc.pyx:102-108): consecutive
//      iterations commute unless they touch the same position.  A window of up to 32 iterations is cut at the
//      first lane whose {i-1, j} meets an earlier lane's j (match.any + redux.or); the lanes before the cut
//      swap at once.
//   3. The visits: a visit reads only its site's table neighbours (and, for PIQMC, other slices of its own
//      world line, which nobody else touches during that slice's sweep), so a run of consecutive visits without
//      two adjacent sites gives the same ediffs whether executed in order or at once.  The window is cut at the
//      first lane with a neighbour visited earlier in the window (byte marks in shared memory).  The uniform of
//      a visit is only drawn when ediff > 0 (qmc.pyx:140-143): its index in the stream is the window's base plus
//      the number of earlier lanes that drew one (ballot + popc).
// One warp owns one replica: spins bit-packed (PIQMC: one word per site, bit k = slice k) or as bytes (SA) in
// shared memory together with the permutation (u16), the marks and a 256-entry window of the rand() stream.
// Arithmetic is the thread-per-replica kernels': fp64, table order, __dmul_rn / __dadd_rn.
// ------------------------------------------------------------------------------------------------------------
constexpr unsigned kFull = 0xFFFFFFFFu;

struct WarpRng {
    uint32_t *ring; // shared [256]: raw generator words x_n at index n & 255 (x_{-31..-1} = the seeded state)
    long long pos;  // values consumed so far
    long long gen;  // values generated so far (multiple of 31)
    const int32_t *stream; // recorded rand() outputs instead of a seeded state (or nullptr)
    long long len;

    __device__ __forceinline__ void init(const LibcState &s, const int32_t *str, long long n, int lane)
    {
        pos = gen = 0;
        stream = str;
        len = n;
        if (lane < 31) ring[(lane - 31) & 255] = (uint32_t)s.r[(s.f + lane) % 31]; // oldest first: x_{n-31} sits at r[f]
        __syncwarp();
    }
    // make values pos .. pos + n - 1 available (n <= 64); warp-uniform
    __device__ __forceinline__ void ensure(int n, int lane)
    {
        if (stream) return;
        while (gen - pos < n) {
            uint32_t v = ring[(gen - 31 + lane) & 255]; // lane 31 reads x_{gen}: garbage, never stored
            if (lane < 3) v += ring[(gen - 3 + lane) & 255];
#pragma unroll
            for (int d = 3; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(kFull, v, d);
                if (lane >= d) v += t;
            }
            if (lane < 31) ring[(gen + lane) & 255] = v;
            gen += 31;
            __syncwarp();
        }
    }
    __device__ __forceinline__ int32_t at(long long idx) const
    {
        if (stream) return idx < len ? stream[idx] : 0;
        return (int32_t)(ring[idx & 255] >> 1);
    }
};

struct WarpCtx {
    WarpRng rng;
    uint16_t *perm; // shared [N]
    uint8_t *mark;  // shared [N], 0 outside a window
    int32_t *ring_idx; // shared [64][8]: table rows of the visits ahead (index column)
    double *ring_J;    // shared [64][8]: ... coupling column
    const int32_t *tab_idx;
    const double *tab_J;
    int N, maxnb, lane;
    uint32_t lt; // lanes below this one
    long long pf[4] = {0, 0, 0, 0}; // MCS_EXACT_PROF: cycles in shuffles, cycles in visits, shuffle windows, visit windows
    bool prof = false;
};

// r % i for 0 <= r < 2^31, 1 <= i < 2^16 without the 140-cycle integer division: fp32 quotient estimate (within 5
// of the true one for i >= 256), exact remainder by correction
__device__ __forceinline__ int fast_mod(int32_t r, int i)
{
    if (i < 256) return r % i;
    const uint32_t q = __float2uint_rz(__uint2float_rz((uint32_t)r) * __frcp_rn((float)i));
    int rem = (int)((uint32_t)r - q * (uint32_t)i);
    while (rem < 0) rem += i;
    while (rem >= i) rem -= i;
    return rem;
}

// qmc.pyx:102-108 / sa.pyx:73-79
__device__ __forceinline__ void warp_shuffle(WarpCtx &c)
{
    const int N = c.N, lane = c.lane;
    const long long t0 = c.prof ? clock64() : 0;
    for (int i = lane; i < N; i += 32) c.perm[i] = (uint16_t)i;
    __syncwarp();
    int cur = N;
    while (cur > 0) {
        c.pf[2] += 1;
        c.rng.ensure(32, lane);
        const int i = cur - lane;
        const bool valid = i >= 1;
        const int j = valid ? fast_mod(c.rng.at(c.rng.pos + lane), i) : 0;
        const int A = i - 1;
        // (a) an earlier lane has the same j.  Cheap test first (the marks are all zero between visits): every
        // lane writes its number at mark[j]; if every lane reads its own number back no two lanes share a j.
        // match.any (379 cycles on B200) only runs for windows that do have a duplicate.
        if (valid) c.mark[j] = (uint8_t)(lane + 1);
        __syncwarp();
        const bool lost = valid && c.mark[j] != (uint8_t)(lane + 1);
        const uint32_t any_lost = __ballot_sync(kFull, lost);
        if (valid) c.mark[j] = 0;
        uint32_t same = 0u;
        if (any_lost) same = __match_any_sync(kFull, valid ? j : 0x10000 + lane) & c.lt;
        // (b) an earlier lane's j is this lane's own position i - 1
        const int t = cur - 1 - j; // the lane whose own position is j
        const uint32_t hit = __reduce_or_sync(kFull, (valid && t > lane && t < 32) ? (1u << t) : 0u);
        const bool conflict = !valid || same != 0u || ((hit >> lane) & 1u);
        const uint32_t cb = __ballot_sync(kFull, conflict);
        const int ncommit = cb ? __ffs(cb) - 1 : 32; // >= 1: lane 0 never conflicts
        uint16_t pa = 0, pj = 0;
        if (lane < ncommit) {
            pa = c.perm[A];
            pj = c.perm[j];
        }
        __syncwarp();
        if (lane < ncommit) {
            c.perm[A] = pj;
            c.perm[j] = pa;
        }
        __syncwarp();
        c.rng.pos += ncommit;
        cur -= ncommit;
    }
    if (c.prof) c.pf[0] += clock64() - t0;
}

// Table rows of the visits ahead, staged in shared memory by cp.async: slot (position & 63) holds the row of the
// site visited at that position of the sweep (RW entries of the index and of the coupling column; rows longer than
// 8 entries are read from the table directly instead, RW = 0).  A window needs positions [cur, cur + 32); the next
// one starts at most 32 further on, so while a window is being decided the rows up to cur + nrun + 32 are requested
// and have a whole window's arithmetic to arrive: no table latency on the critical path.
constexpr int kRowRing = 64;

template <int RW>
struct RowRing {
    int32_t *idx; // shared [64][RW]
    double *J;    // shared [64][RW]
};

__device__ __forceinline__ void cp_async(void *smem, const void *gmem, int bytes)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem);
    if (bytes == 16)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem) : "memory");
    else if (bytes == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem) : "memory");
}

// request the table row of `site` into slot `pos & 63`
template <int RW>
__device__ __forceinline__ void stage_row(const WarpCtx &c, const RowRing<RW> &ring, int pos, int site)
{
    const int32_t *ri = c.tab_idx + (long long)site * c.maxnb;
    const double *rj = c.tab_J + (long long)site * c.maxnb;
    int32_t *di = ring.idx + (pos & (kRowRing - 1)) * RW;
    double *dj = ring.J + (pos & (kRowRing - 1)) * RW;
    if ((c.maxnb & 3) == 0) { // rows are 16-byte multiples (both arrays come from cudaMalloc)
        for (int si = 0; si < c.maxnb; si += 4) {
            cp_async(di + si, ri + si, 16);
            cp_async(dj + si, rj + si, 16);
            cp_async(dj + si + 2, rj + si + 2, 16);
        }
    } else {
        for (int si = 0; si < c.maxnb; ++si) {
            cp_async(di + si, ri + si, 4);
            cp_async(dj + si, rj + si, 8);
        }
    }
}

// One sweep's visits in permutation order.  V: double ediff(site, idx row, J row); void flip(site); `scale` = the
// Metropolis temperature (teff or sched[t]); ru != nullptr: acceptance uniforms indexed by visit position
// (sa.pyx:190) instead of rand().
template <int RW, typename V>
__device__ __forceinline__ void warp_visits(WarpCtx &c, V &v, double scale, const double *ru)
{
    const int N = c.N, lane = c.lane;
    const long long t0 = c.prof ? clock64() : 0;
    RowRing<RW> ring{c.ring_idx, c.ring_J};
    const float inv_scale = 1.0f / (float)scale;
    int filled = 0; // rows of positions < filled are staged or on their way
    if (RW > 0) {
        for (int p = lane; p < min(N, kRowRing); p += 32) stage_row<RW>(c, ring, p, (int)c.perm[p]);
        asm volatile("cp.async.commit_group;" ::: "memory");
        filled = min(N, kRowRing);
    }
    int cur = 0;
    while (cur < N) {
        c.pf[3] += 1;
        if (!ru) c.rng.ensure(32, lane);
        const int sidx = cur + lane;
        const bool valid = sidx < N;
        const int site = valid ? (int)c.perm[sidx] : 0;
        if (valid) c.mark[site] = (uint8_t)(lane + 1);
        if (RW > 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        const int32_t *ri = RW > 0 ? ring.idx + (sidx & (kRowRing - 1)) * RW : c.tab_idx + (long long)site * c.maxnb;
        const double *rj = RW > 0 ? ring.J + (sidx & (kRowRing - 1)) * RW : c.tab_J + (long long)site * c.maxnb;
        bool conflict = !valid;
        if (valid) { // a table neighbour is visited earlier in this window
            for (int si = 0; si < c.maxnb; ++si) {
                const int j = ri[si];
                const int m = c.mark[j];
                if (j != site && m != 0 && m - 1 < lane) conflict = true;
            }
        }
        const uint32_t cb = __ballot_sync(kFull, conflict);
        const int nrun = cb ? __ffs(cb) - 1 : 32; // >= 1
        if (valid) c.mark[site] = 0;
        if (RW > 0) { // rows the next window can reach: positions < cur + nrun + 32 (at most 32 new ones)
            const int upto = min(N, cur + nrun + 32), pp = filled + lane;
            if (pp < upto) stage_row<RW>(c, ring, pp, (int)c.perm[pp]);
            asm volatile("cp.async.commit_group;" ::: "memory");
            filled = max(filled, upto);
        }
        __syncwarp();
        const bool active = lane < nrun;
        const double e = active ? v.ediff(site, ri, rj) : 0.0;
        const bool draws = active && !(e <= 0.0); // the `elif` of qmc.pyx:142 is evaluated (NaN included)
        const uint32_t db = __ballot_sync(kFull, draws);
        bool flip = active && e <= 0.0;
        if (draws) {
            // exp(-e / scale) > u, decided in fp32 whenever the two sides are further apart than fp32 can blur
            // (relative error of either side < 4e-5 for |e / scale| < 100, beyond which exp underflows): the
            // fp64 division, exp and division of the reference expression run only for near ties (and NaN)
            const int32_t r = ru ? 0 : c.rng.at(c.rng.pos + __popc(db & c.lt));
            const float uf = ru ? (float)ru[sidx] : (float)r * (1.0f / 2147483647.0f);
            const float evf = __expf(-(float)e * inv_scale);
            if (fabsf(evf - uf) > 1e-3f * fmaxf(evf, uf)) {
                flip = evf > uf;
            } else {
                const double u = ru ? ru[sidx] : __ddiv_rn((double)r, 2147483647.0);
                flip = exp(__ddiv_rn(__dmul_rn(-1.0, e), scale)) > u;
            }
        }
        if (!ru) c.rng.pos += __popc(db);
        __syncwarp(); // every ediff of the window is computed before any spin of it changes
        if (flip) v.flip(site);
        __syncwarp();
        cur += nrun;
    }
    if (RW > 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (c.prof) c.pf[1] += clock64() - t0;
}

template <typename V>
__device__ __forceinline__ void warp_visits_any(WarpCtx &c, V &v, double scale, const double *ru)
{
    if (c.maxnb <= 4)
        warp_visits<4>(c, v, scale, ru);
    else if (c.maxnb <= 8)
        warp_visits<8>(c, v, scale, ru);
    else
        warp_visits<0>(c, v, scale, ru);
}

template <typename WT>
struct QmcVisit {
    WT *w; // shared [N]: bit k = slice k, bit set <=> spin -1
    const int32_t *tab_idx;
    const double *tab_J;
    const double *lut;
    int maxnb, P, k;
    bool global;
    double b_coeff, jperp, teff;

    __device__ __forceinline__ double spin(WT word, int k_) const { return ((word >> k_) & (WT)1) ? -1.0 : 1.0; }
    // qmc.pyx:112-125; (ri, rj) = the site's table row (shared-memory copy or the table itself)
    __device__ __forceinline__ double inplane(int site, const int32_t *ri, const double *rj, WT wi, int k_,
                                              double acc) const
    {
        const double bs = __dmul_rn(b_coeff, spin(wi, k_));
        for (int si = 0; si < maxnb; ++si) {
            const int spinidx = ri[si];
            const double jval = rj[si];
            if (spinidx == site)
                acc = __dadd_rn(acc, __dmul_rn(bs, jval));
            else
                acc = __dadd_rn(acc, __dmul_rn(bs, __dmul_rn(jval, spin(w[spinidx], k_))));
        }
        return acc;
    }
    __device__ __forceinline__ double ediff(int site, const int32_t *ri, const double *rj) const
    {
        const WT wi = w[site];
        if (global) { // qmc.pyx:416-431
            double e = 0.0;
            for (int k_ = 0; k_ < P; ++k_) e = inplane(site, ri, rj, wi, k_, e);
            return e;
        }
        double e = inplane(site, ri, rj, wi, k, 0.0);
        const int tleft = k == 0 ? P - 1 : (k == P - 1 ? P - 2 : k - 1); // qmc.pyx:127-135
        const int tright = k == 0 ? 1 : (k == P - 1 ? 0 : k + 1);
        const double s2 = __dmul_rn(2.0, spin(wi, k));
        e = __dadd_rn(e, __dmul_rn(s2, __dmul_rn(jperp, spin(wi, tleft))));
        e = __dadd_rn(e, __dmul_rn(s2, __dmul_rn(jperp, spin(wi, tright))));
        if (lut) { // qmc.pyx:268-273
            const double t2 = __dmul_rn(2.0, teff);
            for (int d = 1; d < P; ++d) {
                const int bslice = (k + d) % P;
                const double ss = (((wi >> k) ^ (wi >> bslice)) & (WT)1) ? -1.0 : 1.0;
                e = __dadd_rn(e, __dmul_rn(__dmul_rn(t2, ss), lut[d - 1]));
            }
        }
        return e;
    }
    __device__ __forceinline__ void flip(int site) const
    {
        const WT pmask = P == (int)(8 * sizeof(WT)) ? (WT)~(WT)0 : (WT)(((WT)1 << P) - (WT)1);
        w[site] ^= global ? pmask : (WT)((WT)1 << k);
    }
};

// shared memory: row ring (J | idx) | w[N] words | perm[N] u16 | rand() window [256] | mark[N] u8
template <typename WT>
__global__ void __launch_bounds__(32) exact_qmc_warp_kernel(const ExactQmcArgs a)
{
    extern __shared__ __align__(16) unsigned char ex_smem[];
    const int N = a.N, P = a.P, lane = threadIdx.x;
    const long long r = blockIdx.x;
    const int rw = a.maxnb <= 4 ? 4 : (a.maxnb <= 8 ? 8 : 0); // staged row length (0: rows read from the table)
    WarpCtx c;
    c.ring_J = reinterpret_cast<double *>(ex_smem);
    c.ring_idx = reinterpret_cast<int32_t *>(c.ring_J + kRowRing * rw);
    WT *w = reinterpret_cast<WT *>(c.ring_idx + kRowRing * rw);
    c.perm = reinterpret_cast<uint16_t *>(w + N);
    c.rng.ring = reinterpret_cast<uint32_t *>(c.perm + ((N + 1) & ~1));
    c.mark = reinterpret_cast<uint8_t *>(c.rng.ring + 256);
    c.tab_idx = a.tab_idx;
    c.tab_J = a.tab_J;
    c.N = N;
    c.maxnb = a.maxnb;
    c.lane = lane;
    c.lt = (1u << lane) - 1u;
    c.prof = a.prof != nullptr;
    int8_t *conf = a.confs + r * (long long)N * P;
    for (int i = lane; i < N; i += 32) {
        WT word = 0;
        for (int k = 0; k < P; ++k)
            if (conf[(long long)i * P + k] < 0) word |= (WT)1 << k;
        w[i] = word;
        c.mark[i] = 0;
    }
    c.rng.init(a.st[r], a.stream ? a.stream + r * a.stream_len : nullptr, a.stream_len, lane);
    QmcVisit<WT> v;
    v.w = w;
    v.tab_idx = a.tab_idx;
    v.tab_J = a.tab_J;
    v.lut = a.lut;
    v.maxnb = a.maxnb;
    v.P = P;
    v.teff = a.teff;
    for (int f = 0; f < a.S; ++f) {
        v.jperp = a.jperp[f];
        v.b_coeff = a.bcoef[f];
        for (int step = 0; step < a.mcsteps; ++step) {
            v.global = false;
            for (int islice = 0; islice < P; ++islice) {
                warp_shuffle(c);
                v.k = islice;
                warp_visits_any(c, v, a.teff, nullptr);
            }
            if (a.global_moves) { // qmc.pyx:405-438
                warp_shuffle(c);
                v.global = true;
                warp_visits_any(c, v, a.teff, nullptr);
            }
        }
    }
    __syncwarp();
    for (int i = lane; i < N; i += 32) {
        const WT word = w[i];
        for (int k = 0; k < P; ++k) conf[(long long)i * P + k] = ((word >> k) & (WT)1) ? -1 : 1;
    }
    if (a.consumed && lane == 0) a.consumed[r] = c.rng.pos;
    if (a.prof && lane == 0)
        for (int q = 0; q < 4; ++q) a.prof[r * 4 + q] = c.pf[q];
}

struct SaVisit {
    int8_t *sv; // shared [N]
    const int32_t *tab_idx;
    const double *tab_J;
    int maxnb;
    __device__ __forceinline__ double ediff(int site, const int32_t *ri, const double *rj) const
    { // sa.pyx:84-94
        const double m2s = __dmul_rn(-2.0, (double)sv[site]);
        double e = 0.0;
        for (int si = 0; si < maxnb; ++si) {
            const int spinidx = ri[si];
            const double jval = rj[si];
            if (spinidx == site)
                e = __dadd_rn(e, __dmul_rn(m2s, jval));
            else
                e = __dadd_rn(e, __dmul_rn(m2s, __dmul_rn(jval, (double)sv[spinidx])));
        }
        return e;
    }
    __device__ __forceinline__ void flip(int site) const { sv[site] = -sv[site]; }
};

// shared memory: row ring (J | idx) | perm[N] u16 | rand() window [256] | sv[N] i8 | mark[N] u8
__global__ void __launch_bounds__(32) exact_sa_warp_kernel(const ExactSaArgs a)
{
    extern __shared__ __align__(16) unsigned char ex_smem[];
    const int N = a.N, lane = threadIdx.x;
    const long long r = blockIdx.x;
    const int rw = a.maxnb <= 4 ? 4 : (a.maxnb <= 8 ? 8 : 0);
    WarpCtx c;
    c.ring_J = reinterpret_cast<double *>(ex_smem);
    c.ring_idx = reinterpret_cast<int32_t *>(c.ring_J + kRowRing * rw);
    c.perm = reinterpret_cast<uint16_t *>(c.ring_idx + kRowRing * rw);
    c.rng.ring = reinterpret_cast<uint32_t *>(c.perm + ((N + 1) & ~1));
    int8_t *sv = reinterpret_cast<int8_t *>(c.rng.ring + 256);
    c.mark = reinterpret_cast<uint8_t *>(sv + N);
    c.N = N;
    c.maxnb = a.maxnb;
    c.lane = lane;
    c.lt = (1u << lane) - 1u;
    int8_t *gsv = a.svec + r * (long long)N;
    for (int i = lane; i < N; i += 32) {
        sv[i] = gsv[i];
        c.mark[i] = 0;
    }
    c.rng.init(a.st[r], nullptr, 0, lane);
    SaVisit v;
    v.sv = sv;
    v.maxnb = a.maxnb;
    for (int t = 0; t < a.S; ++t) {
        c.tab_idx = v.tab_idx = a.tab_idx + t * a.tab_stride; // NoisyAnneal: nbs[itemp] (sa.pyx:363-365)
        c.tab_J = v.tab_J = a.tab_J + t * a.tab_stride;
        for (int step = 0; step < a.mcsteps; ++step) {
            warp_shuffle(c);
            warp_visits_any(c, v, a.sched[t], a.randuni ? a.randuni + ((long long)t * a.mcsteps + step) * N : nullptr);
        }
    }
    __syncwarp();
    for (int i = lane; i < N; i += 32) gsv[i] = sv[i];
    if (a.consumed && lane == 0) a.consumed[r] = c.rng.pos;
}

// Launch helper: returns false when a replica does not fit (caller falls back to the thread-per-replica kernel)
template <typename K, typename A>
int launch_warp_replay(K kernel, const A &a, long long R, size_t smem, cudaStream_t s)
{
    if (smem > 48 * 1024)
        MCS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<(unsigned)R, 32, smem, s>>>(a);
    return MCS_OK;
}

bool warp_replay_wanted(const mcs_instance *inst, size_t smem)
{
    if (getenv("MCS_EXACT_LEGACY")) return false; // tests: the thread-per-replica kernels
    // u16 permutation entries; very long rows make every window one visit long (complete graphs): no gain
    return inst->N <= 65535 && smem <= 227 * 1024 - 1024 && inst->maxnb <= 64;
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
    std::lock_guard<std::recursive_mutex> one_call_at_a_time(inst->call_mutex);
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
    const bool trace = getenv("MCS_TRACE_EXACT") != nullptr;
    if (trace) MCS_CUDA(cudaEventRecord(inst->ev0, s));
    // per replica in shared memory: words | u16 permutation | rand() window | marks
    const size_t wbytes = P > 32 ? 8 : 4;
    const size_t rowring = (size_t)64 * 12 * (inst->maxnb <= 4 ? 4 : (inst->maxnb <= 8 ? 8 : 0));
    const size_t smem =
        rowring + (size_t)inst->N * wbytes + (((size_t)inst->N + 1) & ~(size_t)1) * 2 + 1024 + (size_t)inst->N;
    DevBuf d_prof;
    if (getenv("MCS_EXACT_PROF")) {
        MCS_TRY(d_prof.alloc((size_t)R * 4 * sizeof(long long)));
        MCS_CUDA(cudaMemsetAsync(d_prof.p, 0, (size_t)R * 4 * sizeof(long long), s));
        a.prof = d_prof.as<long long>();
    }
    if (P <= 64 && warp_replay_wanted(inst, smem)) {
        if (P > 32)
            MCS_TRY(launch_warp_replay(exact_qmc_warp_kernel<uint64_t>, a, R, smem, s));
        else
            MCS_TRY(launch_warp_replay(exact_qmc_warp_kernel<uint32_t>, a, R, smem, s));
    } else {
        exact_qmc_kernel<<<(unsigned)((R + 31) / 32), 32, 0, s>>>(a);
    }
    if (a.prof) {
        std::vector<long long> hp((size_t)R * 4);
        MCS_CUDA(cudaMemcpyAsync(hp.data(), d_prof.p, hp.size() * sizeof(long long), cudaMemcpyDeviceToHost, s));
        MCS_CUDA(cudaStreamSynchronize(s));
        double q[4] = {0, 0, 0, 0};
        for (int64_t r = 0; r < R; ++r)
            for (int x = 0; x < 4; ++x) q[x] += (double)hp[(size_t)r * 4 + x] / (double)R;
        fprintf(stderr, "[mcs exact] per replica: shuffles %.0f cycles in %.0f windows (%.0f each), visits %.0f cycles in "
                        "%.0f windows (%.0f each)\n", q[0], q[2], q[0] / std::max(1.0, q[2]), q[1], q[3],
                q[1] / std::max(1.0, q[3]));
    }
    if (trace) {
        float ms = 0.0f;
        MCS_CUDA(cudaEventRecord(inst->ev1, s));
        MCS_CUDA(cudaEventSynchronize(inst->ev1));
        MCS_CUDA(cudaEventElapsedTime(&ms, inst->ev0, inst->ev1));
        fprintf(stderr, "[mcs exact] qmc replay kernel: %.3f ms, %.3e attempts/s (R %lld, N %lld, P %lld, %lld sweeps)\n", ms,
                (double)R * inst->N * P * S * mcsteps / (ms * 1e-3), (long long)R, (long long)inst->N, (long long)P,
                (long long)(S * mcsteps));
    }
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
    std::lock_guard<std::recursive_mutex> one_call_at_a_time(inst->call_mutex);
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
    std::lock_guard<std::recursive_mutex> one_call_at_a_time(inst->call_mutex);
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
    const size_t rowring = (size_t)64 * 12 * (inst->maxnb <= 4 ? 4 : (inst->maxnb <= 8 ? 8 : 0));
    const size_t smem = rowring + (((size_t)inst->N + 1) & ~(size_t)1) * 2 + 1024 + 2 * (size_t)inst->N;
    if (warp_replay_wanted(inst, smem))
        MCS_TRY(launch_warp_replay(exact_sa_warp_kernel, a, R, smem, s));
    else
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
    std::lock_guard<std::recursive_mutex> one_call_at_a_time(inst->call_mutex);
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
