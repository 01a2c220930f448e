// mcs_sa.cu -- classical simulated-annealing sweeps, multi-spin coded over restarts (sm_100a).
//
// Replaces the loop nest of sa.Anneal (reference sa.pyx:66-101).
//
// Data layout in HBM:  V[site][word] : uint32, bit b of word g = restart 32 g + b, bit set <=>
// spin -1.  A warp owns (site, 32 consecutive words) = up to 1024 restarts of one site, so the
// couplings and the 2^deg-entry acceptance-threshold table of the site are warp-uniform and the
// 32 lanes read/write 128 contiguous bytes.  All 32 bits of a word are attempted in one launch
// (same site, different restarts); neighbours are in other colour classes and frozen.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <set>
#include <tuple>
#include <vector>

#include "mcs_common.cuh"

namespace {

constexpr int kWarps = 4;

struct SaPass {
    uint32_t *V;
    const int32_t *ell_idx;
    const float *ell_J;
    const float *h;
    const int32_t *sites;
    int nsites;
    int dpad;
    int nq;
    int field;
    int chunks; // ceil(G / 32) warps per site
    long long G;
    float nl2e_over_t; // -log2(e)/T
    mcs_philox_keys keys;
    mcs_pow2_table pow2;
    uint32_t sweep_lo, sweep_hi;
    uint32_t word_offset; // replica_offset / 32
    uint32_t tie_thr;     // lazily refined uniforms (mcs_common.cuh)
    long long Gs;         // words per row of V (= G unless the launch covers a chunk of the words)
    int wpt;              // MULTI: words per thread ...
    uint32_t wstep;       // ... the thread's k-th word is g0 + k wstep
};

// Index word of group Q (restarts 8 i + 7 - Q of the word, i = 0..3): plane p's bit of restart 8 i + 7 - Q goes
// to bit 8 i + SH + p, i.e. the plane is shifted by SH + p - 7 + Q and merged with one LOP3.  Every bit of
// every plane is used here (no Trotter parity), so planes cannot be interleaved first as in mcs_piqmc.cu.
template <int NPL, int Q>
__device__ __forceinline__ uint32_t sa_gather_index(const uint32_t (&pl)[NPL], const mcs_pow2_table &pow2)
{
    constexpr int SH = NPL <= 6 ? 2 : 0;
    uint32_t acc = 0;
#define MCS_SA_PLANE(p)                                                                                       \
    if (p < NPL) acc |= mcs_plane_shift<SH + p - 7 + Q>(pl[p < NPL ? p : 0], pow2) & (0x01010101u << (SH + p));
    MCS_SA_PLANE(0) MCS_SA_PLANE(1) MCS_SA_PLANE(2) MCS_SA_PLANE(3)
    MCS_SA_PLANE(4) MCS_SA_PLANE(5) MCS_SA_PLANE(6) MCS_SA_PLANE(7)
#undef MCS_SA_PLANE
    return acc;
}

// WARPS warps per CTA, all on the same site; grid = (CTAs per site, sites of the colour class).  The table
// (complemented thresholds, see mcs_common.cuh) sits at a compile-time shared address and, for up to 6 planes,
// the pattern index is kept pre-multiplied by 4 (= the LDS byte offset).  Groups 2q and 2q+1 share one Philox
// call (lazily refined uniforms).  FLD: the instance has (1) / has no (0) field plane.
// MULTI: a thread takes a.wpt words one after the other and shares the site's set-up (as the PIQMC pass kernel does:
// one-warp CTAs, no register cap).
template <int NPL, int WARPS, int FLD, bool MULTI = false>
__global__ void __launch_bounds__(WARPS * 32, 8) sa_lut_pass_kernel(const __grid_constant__ SaPass a)
{
    constexpr int ENT = 1 << NPL, NQ = NPL - FLD;
    constexpr int SH = NPL <= 6 ? 2 : 0;
    __shared__ uint32_t s_lut[ENT];
    __shared__ uint2 s_bounce[4 * WARPS * 32]; // [call][thread]: private slots for the index bytes
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    mcs_pdl_launch_dependents();
    const unsigned si = blockIdx.y + 65535u * blockIdx.z;
    if (si >= (unsigned)a.nsites) return; // only when the colour class has more than 65535 sites (CTA-uniform)
    const int site = __ldg(&a.sites[si]);
    const uint32_t g0 = ((uint32_t)blockIdx.x * WARPS + warp) * 32 + lane;

    float c[NPL];
    int nb[NPL];
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
        if (j < NQ) {
            nb[j] = __ldg(&a.ell_idx[(uint64_t)(uint32_t)site * (uint32_t)a.dpad + j]);
            c[j] = -2.0f * __ldg(&a.ell_J[(uint64_t)(uint32_t)site * (uint32_t)a.dpad + j]); // sa.pyx:91-94
        } else {
            nb[j] = site;
            c[j] = -2.0f * __ldg(&a.h[site]);
        }
    }
    for (int e = threadIdx.x; e < ENT; e += WARPS * 32) {
        float dE = 0.0f;
#pragma unroll
        for (int j = 0; j < NPL; ++j) dE += ((e >> j) & 1) ? -c[j] : c[j];
        s_lut[e] = ~mcs_accept_threshold(dE, a.nl2e_over_t);
    }
    // the state loads are issued BEFORE the barrier that publishes the table: their latency hides behind it
    const uint32_t G32 = (uint32_t)a.G, Gs32 = (uint32_t)a.Gs;
    mcs_pdl_wait(); // everything above depends on the instance and the schedule only
    const int wpt = MULTI ? a.wpt : 1;
    for (int kw = 0; kw < wpt; ++kw) {
    const uint32_t g = g0 + (MULTI ? (uint32_t)kw * a.wstep : 0u);
    const bool live = g < G32;
    uint32_t *Vg = a.V + (live ? g : 0u);
    const uint32_t v = Vg[(uint64_t)(uint32_t)site * Gs32];
    uint32_t pl[NPL];
#pragma unroll
    for (int j = 0; j < NPL; ++j) pl[j] = j < NQ ? v ^ Vg[(uint64_t)(uint32_t)nb[j] * Gs32] : v;
    if (kw == 0) {
        if (WARPS == 1)
            __syncwarp();
        else
            __syncthreads();
    }
    if (!live) continue;
    const uint32_t c0 = a.word_offset + g, c1 = (uint32_t)site, c2 = a.sweep_lo, c3hi = a.sweep_hi << 8;
    uint2 *bounce = s_bounce + threadIdx.x;
    uint32_t rej = 0, flags = 0;
#define MCS_SA_CALL(q)                                                                                        \
    {                                                                                                         \
        uint32_t chA, chB;                                                                                    \
        mcs_decide_call<SH>(chA, chB, flags, sa_gather_index<NPL, 2 * (q)>(pl, a.pow2),                       \
                            sa_gather_index<NPL, 2 * (q) + 1>(pl, a.pow2), s_lut, c0, c1, c2,                 \
                            c3hi | (uint32_t)(2 * (q)), a.keys, a.pow2, a.tie_thr, bounce + (q) * WARPS * 32);\
        rej = chA * a.pow2.up[7 - 2 * (q)] + rej;                                                             \
        rej = chB * a.pow2.up[6 - 2 * (q)] + rej;                                                             \
    }
    MCS_SA_CALL(0) MCS_SA_CALL(1) MCS_SA_CALL(2) MCS_SA_CALL(3)
#undef MCS_SA_CALL
    if (flags) { // rare: Horner order, call 0 ended at bit 3 ... call 3 at bit 0
#define MCS_SA_REFINE(q)                                                                                      \
    if (flags & (8u >> (q))) {                                                                                \
        const uint2 ch = mcs_refine_call<SH>(sa_gather_index<NPL, 2 * (q)>(pl, a.pow2),                       \
                                             sa_gather_index<NPL, 2 * (q) + 1>(pl, a.pow2), s_lut, c0, c1, c2,\
                                             c3hi | (uint32_t)(2 * (q)), a.keys.rk[0], a.keys.rk[1]);         \
        rej = (rej & ~(0x03030303u << (6 - 2 * (q)))) | (ch.x << (7 - 2 * (q))) | (ch.y << (6 - 2 * (q)));    \
    }
        MCS_SA_REFINE(0) MCS_SA_REFINE(1) MCS_SA_REFINE(2) MCS_SA_REFINE(3)
#undef MCS_SA_REFINE
    }
    Vg[(uint64_t)(uint32_t)site * Gs32] = v ^ ~rej;
    } // words of this thread
}

// general-degree pass: energy differences accumulated over the ELL row, one register per restart
__global__ void __launch_bounds__(kWarps * 32) sa_direct_pass_kernel(const __grid_constant__ SaPass a)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * kWarps + warp;
    if (item >= (long long)a.nsites * a.chunks) return;
    const int site = a.sites[item / a.chunks];
    const long long g = (item % a.chunks) * 32 + lane;
    if (g >= a.G) return;
    uint32_t v = a.V[(long long)site * a.G + g];
    const float hc = a.field ? -2.0f * __ldg(&a.h[site]) : 0.0f;
    float e[32];
#pragma unroll
    for (int b = 0; b < 32; ++b) e[b] = ((v >> b) & 1u) ? -hc : hc;
    for (int j = 0; j < a.dpad; ++j) {
        const float cj = -2.0f * __ldg(&a.ell_J[(long long)site * a.dpad + j]);
        if (cj == 0.0f) continue;
        const int nbj = __ldg(&a.ell_idx[(long long)site * a.dpad + j]);
        const uint32_t x = v ^ a.V[(long long)nbj * a.G + g];
#pragma unroll
        for (int b = 0; b < 32; ++b) e[b] += __uint_as_float(__float_as_uint(cj) ^ (((x >> b) & 1u) << 31));
    }
    const uint32_t c0 = a.word_offset + (uint32_t)g, c1 = (uint32_t)site;
    uint32_t flip = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        uint32_t rnd[4];
        mcs_philox4x32_rk(c0, c1, a.sweep_lo, (a.sweep_hi << 8) | (uint32_t)q, a.keys, rnd);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int b = 8 * i + 7 - q;
            if (mcs_accepts(rnd[i], mcs_accept_threshold(e[b], a.nl2e_over_t))) flip |= 1u << b;
        }
    }
    a.V[(long long)site * a.G + g] = v ^ flip;
}

// in: int8 [R][N]; thread per (site, word)
__global__ void sa_pack_kernel(const int8_t *__restrict__ in, uint32_t *__restrict__ V, long long N, long long R,
                               long long G)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * G) return;
    const long long g = t / N, i = t % N;
    uint32_t v = 0;
    for (int b = 0; b < 32; ++b) {
        const long long r = g * 32 + b;
        if (r < R) v |= (uint32_t)(in[r * N + i] < 0) << b;
    }
    V[i * G + g] = v;
}

__global__ void sa_unpack_kernel(const uint32_t *__restrict__ V, int8_t *__restrict__ out, long long N,
                                 long long R, long long G)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * G) return;
    const long long g = t / N, i = t % N;
    const uint32_t v = V[i * G + g];
    for (int b = 0; b < 32; ++b) {
        const long long r = g * 32 + b;
        if (r < R) out[r * N + i] = ((v >> b) & 1u) ? -1 : 1;
    }
}

__global__ void sa_init_kernel(uint32_t *V, long long N, long long R, long long G, uint32_t key0, uint32_t key1,
                               uint32_t replica_offset)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * G) return;
    const long long i = t / G, g = t % G;
    uint32_t v = 0;
    for (int b = 0; b < 32; ++b) {
        const long long r = g * 32 + b;
        if (r >= R) break;
        uint32_t rnd[4];
        // same draw as piqmc_init_kernel: replica r starts from the same spins in both solvers
        mcs_philox4x32(replica_offset + (uint32_t)r, (uint32_t)i, 0u, MCS_TAG_INIT, key0, key1, rnd);
        v |= (rnd[0] & 1u) << b;
    }
    V[i * G + g] = v;
}

// fixed-order fp64 energy per restart (see piqmc_energy_kernel)
__global__ void sa_energy_kernel(const uint32_t *__restrict__ V, const int32_t *__restrict__ tab_idx,
                                 const double *__restrict__ tab_J, double *__restrict__ out, long long N,
                                 int maxnb, long long R, long long G)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const long long g = r >> 5;
    const int b = (int)(r & 31);
    double e = 0.0;
    for (long long i = 0; i < N; ++i) {
        double pair = 0.0, field = 0.0;
        for (int s = 0; s < maxnb; ++s) {
            const int j = tab_idx[i * maxnb + s];
            const double jv = tab_J[i * maxnb + s];
            if (j == i) {
                field = __dadd_rn(field, jv);
            } else {
                const double sj = ((V[(long long)j * G + g] >> b) & 1u) ? -1.0 : 1.0;
                pair = __dadd_rn(pair, __dmul_rn(jv, sj));
            }
        }
        const double si = ((V[i * G + g] >> b) & 1u) ? -1.0 : 1.0;
        e = __dadd_rn(e, __dmul_rn(si, __dadd_rn(__dmul_rn(0.5, pair), field)));
    }
    out[r] = e;
}

// Fixed-order energy by table (mcs_energy_tables, see mcs_piqmc.cu): one fp64 addition per site and restart.
constexpr int kSaLTile = 128;
template <int CH>
__global__ void __launch_bounds__(128) sa_energy_lut_kernel(const uint32_t *__restrict__ V,
                                                            const double *__restrict__ etab,
                                                            const int32_t *__restrict__ etab_j,
                                                            double *__restrict__ out, long long N, long long R,
                                                            long long G)
{
    __shared__ __align__(16) double s_t[kSaLTile][16];
    __shared__ __align__(16) int32_t s_j[kSaLTile][4];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const long long r = (long long)blockIdx.x * blockDim.x + tid;
    const long long g = min(r >> 5, G - 1);
    const int bit = (int)(r & 31);
    const uint32_t *Vg = V + g;
    double e = 0.0;
    for (long long tile0 = 0; tile0 < N; tile0 += kSaLTile) {
        const int nt = (int)min((long long)kSaLTile, N - tile0);
        __syncthreads();
        {
            const double2 *src = reinterpret_cast<const double2 *>(etab + tile0 * 16);
            double2 *dst = reinterpret_cast<double2 *>(&s_t[0][0]);
#pragma unroll 8
            for (int x = tid; x < nt * 8; x += nthr) dst[x] = __ldg(&src[x]);
        }
        for (int x = tid; x < nt * 4; x += nthr) (&s_j[0][0])[x] = __ldg(&etab_j[tile0 * 4 + x]);
        __syncthreads();
        uint32_t b[2][CH][4], w[2][CH];
        // every load unconditional, raw words kept and shifted at use (see piqmc_energy_lut_kernel)
        auto load_chunk = [&](int c0, uint32_t (&bb)[CH][4], uint32_t (&ww)[CH]) {
#pragma unroll
            for (int u = 0; u < CH; ++u) {
                const int i = min(c0 + u, nt - 1);
                const int4 j = *reinterpret_cast<const int4 *>(s_j[i]);
                bb[u][0] = __ldg(&Vg[(long long)j.x * G]);
                bb[u][1] = __ldg(&Vg[(long long)j.y * G]);
                bb[u][2] = __ldg(&Vg[(long long)j.z * G]);
                bb[u][3] = __ldg(&Vg[(long long)j.w * G]);
                ww[u] = __ldg(&Vg[(tile0 + i) * G]);
            }
        };
        auto sum_chunk = [&](int c0, const uint32_t (&bb)[CH][4], const uint32_t (&ww)[CH]) {
#pragma unroll
            for (int u = 0; u < CH; ++u) {
                const int i = c0 + u;
                if (i < nt) {
                    const uint32_t idx = ((bb[u][0] >> bit) & 1u) | (((bb[u][1] >> bit) & 1u) << 1) |
                                         (((bb[u][2] >> bit) & 1u) << 2) | (((bb[u][3] >> bit) & 1u) << 3);
                    const double t = s_t[i][idx];
                    e = __dadd_rn(e, ((ww[u] >> bit) & 1u) ? -t : t);
                }
            }
        };
        load_chunk(0, b[0], w[0]);
        for (int c0 = 0; c0 < nt; c0 += 2 * CH) {
            load_chunk(c0 + CH, b[1], w[1]);
            sum_chunk(c0, b[0], w[0]);
            load_chunk(c0 + 2 * CH, b[0], w[0]);
            sum_chunk(c0 + CH, b[1], w[1]);
        }
    }
    if (r < R) out[r] = e;
}

// ------------------------------------------------------------------------------------------------------------
// Small batches, cluster-resident: the whole schedule in ONE launch without a grid barrier.  Restarts are
// independent, so the batch is cut into groups of Wc words (32 Wc restarts) and each group is owned by one
// thread-block CLUSTER for the whole schedule: CTA r of the cluster keeps a contiguous slice of every colour class
// (all Wc words of those sites) in its shared memory, reads the neighbour words that live in the other CTAs through
// distributed shared memory, and the colour passes are separated by the hardware cluster barrier (~0.2 us) instead of
// a kernel boundary or a grid barrier (~2.3 us).  The lanes of a warp are 32 SITES of one word here, so every site
// has its own threshold table, kept column-wise (mcs_lut_col: conflict-free) and rebuilt per schedule step between
// arriving at the barrier and waiting on it; a pattern and its complement have opposite energy differences, so half
// the entries need no exponential.  Same Philox counters (global word, site, sweep), same thresholds, same decision
// code as sa_lut_pass_kernel: bit-identical to the multi-launch path (tests).
// ------------------------------------------------------------------------------------------------------------
constexpr int kClMaxColors = 16;

struct SaCluster {
    uint32_t *V;
    const int32_t *ell_idx;
    const float *ell_J; // [nsteps][N][dpad]
    const float *h;     // [nsteps][N]
    long long ellJ_stride, h_stride;
    const int32_t *order; // sites sorted by colour
    const int32_t *pos;   // inverse of order
    const float *nl2e;    // [S]: -log2(e) / sched[t]
    long long *prof;      // MCS_CLUSTER_PROF=1: cycles of thread 0 of CTA 0 in {items, arrive + tables, wait}
    int color_start[kClMaxColors + 1];
    int per[kClMaxColors];          // sites of the colour per CTA = ceil(n_c / C)
    int loc_base[kClMaxColors + 1]; // local index of a CTA's first site of the colour
    int ncolors, dpad, S, mcsteps, Wc, nloc_pad, csize, prof_cta;
    long long G;
    uint64_t sweep_offset;
    mcs_philox_keys keys;
    mcs_pow2_table pow2;
    uint32_t word_offset, tie_thr;
};

// x / d and x % d for 0 <= x < 2^22, d >= 1 (rcp = 1.0f / d): no integer division in the pass loop
__device__ __forceinline__ void sa_divmod(int x, int d, float rcp, int &q, int &r)
{
    q = __float2int_rz(__int2float_rn(x) * rcp);
    r = x - q * d;
    if (r < 0) r += d, --q;
    if (r >= d) r -= d, ++q;
}

// TMAX = 512: CTAs of at most 512 threads compile without the 64-register cap (measured: 2.8 against 3.2 us per pass
// at 448 restarts, 3.4 against 3.6 at 640; with 800 - 1000 items per CTA the 1024-thread form is faster).
template <int NPL, int FLD, int TMAX>
__global__ void __launch_bounds__(TMAX, 1) sa_cluster_kernel(const __grid_constant__ SaCluster a)
{
    constexpr int ENT = 1 << NPL, NQ = NPL - FLD;
    static_assert(NPL <= 6, "byte index fields hold pattern * 4");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = blockDim.x, tid = threadIdx.x;
    const uint32_t NP = (uint32_t)a.nloc_pad;
    uint32_t *s_lut = reinterpret_cast<uint32_t *>(smem_raw);        // [ENT][NP]
    uint32_t *s_V = s_lut + (size_t)ENT * NP;                       // [Wc][NP]
    uint32_t *s_nba = s_V + (size_t)a.Wc * NP;                      // [NQ][NP] shared::cluster address of word 0
    float *s_c = reinterpret_cast<float *>(s_nba + (size_t)NQ * NP); // [NPL][NP] -2 J, -2 h (static instances)
    int32_t *s_site = reinterpret_cast<int32_t *>(s_c + (size_t)NPL * NP); // [NP] global site, -1: none
    uint2 *s_bounce = reinterpret_cast<uint2 *>(s_site + NP);       // [4][T]
    const uint32_t rank = mcs_cluster_ctarank();
    const uint32_t cid = blockIdx.x / (uint32_t)a.csize;
    const uint32_t w0 = cid * (uint32_t)a.Wc;
    const uint32_t G32 = (uint32_t)a.G;
    const int nw = (int)min((uint32_t)a.Wc, G32 - w0);
    const bool fixed = a.ellJ_stride == 0 && a.h_stride == 0;
    const int nloc = a.loc_base[a.ncolors];

    // ---- resident data: sites, neighbour addresses, couplings, state ----
    for (int loc = tid; loc < (int)NP; loc += T) {
        int site = -1;
        if (loc < nloc) {
            int c = 0;
            while (c + 1 < a.ncolors && loc >= a.loc_base[c + 1]) ++c;
            const int q = (int)rank * a.per[c] + (loc - a.loc_base[c]);
            if (q < a.color_start[c + 1] - a.color_start[c]) site = __ldg(&a.order[a.color_start[c] + q]);
        }
        s_site[loc] = site;
        if (site < 0) {
            for (int w = 0; w < a.Wc; ++w) s_V[(uint32_t)w * NP + loc] = 0u;
            continue;
        }
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
            if (j < NQ) {
                const int nbs = __ldg(&a.ell_idx[(uint64_t)(uint32_t)site * (uint32_t)a.dpad + j]);
                const int p = __ldg(&a.pos[nbs]);
                int c = 0;
                while (c + 1 < a.ncolors && p >= a.color_start[c + 1]) ++c;
                const int q = p - a.color_start[c];
                const int r = q / a.per[c];
                const int l = a.loc_base[c] + (q - r * a.per[c]);
                s_nba[(uint32_t)j * NP + loc] =
                    mcs_mapa((uint32_t)__cvta_generic_to_shared(s_V + l), (uint32_t)r);
                s_c[(uint32_t)j * NP + loc] = -2.0f * __ldg(&a.ell_J[(uint64_t)(uint32_t)site * (uint32_t)a.dpad + j]);
            } else {
                s_c[(uint32_t)j * NP + loc] = -2.0f * __ldg(&a.h[site]);
            }
        }
        for (int w = 0; w < a.Wc; ++w)
            s_V[(uint32_t)w * NP + loc] = w < nw ? a.V[(uint64_t)(uint32_t)site * G32 + w0 + w] : 0u;
    }
    __syncthreads();

    // Threshold tables of this CTA's colour-`col` sites at schedule step t, spread over threads [first, first + nthr):
    // one work item = (site, pattern e with its top bit clear) fills the entries of e and of its complement, whose
    // energy difference is the exact negation (same terms, same order, opposite signs): one exponential per pair.
    auto build_tables = [&](int col, int t, int first, int nthr) {
        const int nc = a.color_start[col + 1] - a.color_start[col];
        const int pv = max(0, min(a.per[col], nc - (int)rank * a.per[col]));
        const int work = pv * (ENT / 2);
        if (tid < first || work == 0) return;
        const float rcp = 1.0f / (float)pv;
        const float nl2e = __ldg(&a.nl2e[t]);
        const float *ellJ = a.ell_J + (size_t)t * a.ellJ_stride;
        const float *hrow = a.h + (size_t)t * a.h_stride;
        for (int idx = tid - first; idx < work; idx += nthr) {
            int e, q;
            sa_divmod(idx, pv, rcp, e, q);
            const int loc = a.loc_base[col] + q;
            float dE = 0.0f;
            if (fixed) {
#pragma unroll
                for (int j = 0; j < NPL; ++j) {
                    const float cj = s_c[(uint32_t)j * NP + loc];
                    dE += ((e >> j) & 1) ? -cj : cj;
                }
            } else {
                const int site = s_site[loc];
#pragma unroll
                for (int j = 0; j < NPL; ++j) {
                    const float cj = -2.0f * (j < NQ ? __ldg(&ellJ[(uint64_t)(uint32_t)site * (uint32_t)a.dpad + j])
                                                     : __ldg(&hrow[site]));
                    dE += ((e >> j) & 1) ? -cj : cj;
                }
            }
            const uint32_t nthr_ = ~mcs_accept_threshold(fabsf(dE), nl2e); // dE == 0: always; NaN: never, both ways
            s_lut[(uint32_t)e * NP + loc] = !(dE <= 0.0f) ? nthr_ : 0u;
            s_lut[(uint32_t)(ENT - 1 - e) * NP + loc] = !(dE >= 0.0f) ? nthr_ : 0u;
        }
    };
    auto items_of = [&](int col) {
        const int nc = a.color_start[col + 1] - a.color_start[col];
        return max(0, min(a.per[col], nc - (int)rank * a.per[col])) * nw;
    };

    const long long npass = (long long)a.S * a.mcsteps * a.ncolors;
    build_tables(0, 0, 0, T);
    mcs_cluster_arrive(); // every CTA's state is in place before the first remote read
    __syncthreads();
    mcs_cluster_wait();

    uint2 *bounce = s_bounce + tid;
    long long pf[4] = {0, 0, 0, 0}, tq = clock64();
    int col = 0, step_in = 0, t = 0; // pass p = ((t * mcsteps) + step_in) * ncolors + col
    uint64_t sweep = a.sweep_offset;
    for (long long p = 0; p < npass; ++p) {
        const int nc = a.color_start[col + 1] - a.color_start[col];
        const int pv = max(0, min(a.per[col], nc - (int)rank * a.per[col]));
        const int items = pv * nw;
        const uint32_t c2 = (uint32_t)sweep, c3hi = (uint32_t)(sweep >> 32) << 8;
        const float rcp = pv > 0 ? 1.0f / (float)pv : 0.0f;
        // Tables of the NEXT pass (another colour: nobody reads those columns during this pass).  With enough
        // threads left over after the items (at least a quarter of the CTA) they build them now, under the decisions;
        // otherwise every thread takes a share between arriving at the barrier and waiting on it.
        int ncol = col + 1, nstep = step_in, nt = t;
        if (ncol == a.ncolors) {
            ncol = 0;
            if (++nstep == a.mcsteps) nstep = 0, ++nt;
        }
        const bool need_tables = p + 1 < npass && nstep == 0 && (ncol != col || items == 0);
        const bool tables_now = need_tables && a.ncolors > 1 && 4 * (T - items) >= T;
        if (tables_now) build_tables(ncol, nt, items, T - items);
        for (int item = tid; item < items; item += T) {
            int w, q;
            sa_divmod(item, pv, rcp, w, q);
            const uint32_t loc = (uint32_t)(a.loc_base[col] + q);
            uint32_t *own = s_V + (uint32_t)w * NP + loc;
            const uint32_t v = *own;
            uint32_t pl[NPL];
#pragma unroll
            for (int j = 0; j < NPL; ++j)
                pl[j] = j < NQ ? v ^ mcs_ld_cluster_u32(s_nba[(uint32_t)j * NP + loc] + (uint32_t)w * NP * 4u) : v;
            const uint32_t *lcol = s_lut + loc;
            const uint32_t c0 = a.word_offset + w0 + (uint32_t)w, c1 = (uint32_t)s_site[loc];
            uint32_t rej = 0, flags = 0;
#define MCS_SA_CALL(qq)                                                                                       \
    {                                                                                                         \
        uint32_t chA, chB;                                                                                    \
        mcs_decide_call_col(chA, chB, flags, sa_gather_index<NPL, 2 * (qq)>(pl, a.pow2),                      \
                            sa_gather_index<NPL, 2 * (qq) + 1>(pl, a.pow2), lcol, NP, c0, c1, c2,             \
                            c3hi | (uint32_t)(2 * (qq)), a.keys, a.pow2, a.tie_thr, bounce + (qq) * T);       \
        rej = chA * a.pow2.up[7 - 2 * (qq)] + rej;                                                            \
        rej = chB * a.pow2.up[6 - 2 * (qq)] + rej;                                                            \
    }
            MCS_SA_CALL(0) MCS_SA_CALL(1) MCS_SA_CALL(2) MCS_SA_CALL(3)
#undef MCS_SA_CALL
            if (flags) {
#define MCS_SA_REFINE(qq)                                                                                     \
    if (flags & (8u >> (qq))) {                                                                               \
        const uint2 ch = mcs_refine_call_col(sa_gather_index<NPL, 2 * (qq)>(pl, a.pow2),                      \
                                             sa_gather_index<NPL, 2 * (qq) + 1>(pl, a.pow2), lcol, NP, c0, c1,\
                                             c2, c3hi | (uint32_t)(2 * (qq)), a.keys.rk[0], a.keys.rk[1]);    \
        rej = (rej & ~(0x03030303u << (6 - 2 * (qq)))) | (ch.x << (7 - 2 * (qq))) | (ch.y << (6 - 2 * (qq))); \
    }
                MCS_SA_REFINE(0) MCS_SA_REFINE(1) MCS_SA_REFINE(2) MCS_SA_REFINE(3)
#undef MCS_SA_REFINE
            }
            *own = v ^ ~rej;
        }
        // next pass
        if (++col == a.ncolors) {
            col = 0;
            ++sweep;
            if (++step_in == a.mcsteps) step_in = 0, ++t;
        }
        long long tn = clock64();
        pf[0] += tn - tq, tq = tn;
        mcs_cluster_arrive();
        // the tables of the next colour are not in use by anyone (their last readers passed a cluster barrier); a
        // single-colour instance rebuilds its only table here, after the CTA's readers are done
        if (p + 1 < npass && step_in == 0 && !tables_now) {
            if (a.ncolors == 1) __syncthreads();
            build_tables(col, t, 0, T);
        }
        tn = clock64();
        pf[1] += tn - tq, tq = tn;
        __syncthreads();
        tn = clock64();
        pf[2] += tn - tq, tq = tn;
        mcs_cluster_wait();
        tn = clock64();
        pf[3] += tn - tq, tq = tn;
    }
    if (a.prof && blockIdx.x == (unsigned)a.prof_cta && tid == 0)
        for (int q = 0; q < 4; ++q) a.prof[q] = pf[q];
    // no remote read is in flight after the last barrier: write the state back
    for (int loc = tid; loc < nloc; loc += T) {
        const int site = s_site[loc];
        if (site < 0) continue;
        for (int w = 0; w < nw; ++w) a.V[(uint64_t)(uint32_t)site * G32 + w0 + w] = s_V[(uint32_t)w * NP + loc];
    }
}

template <int NPL>
const void *sa_cluster_fn(int field, int tmax)
{
    if (tmax == 512)
        return field ? (const void *)sa_cluster_kernel<NPL, 1, 512> : (const void *)sa_cluster_kernel<NPL, 0, 512>;
    return field ? (const void *)sa_cluster_kernel<NPL, 1, 1024> : (const void *)sa_cluster_kernel<NPL, 0, 1024>;
}

const void *sa_cluster_fn(int npl, int field, int tmax)
{
    switch (npl) {
    case 1: return sa_cluster_fn<1>(field, tmax);
    case 2: return sa_cluster_fn<2>(field, tmax);
    case 3: return sa_cluster_fn<3>(field, tmax);
    case 4: return sa_cluster_fn<4>(field, tmax);
    case 5: return sa_cluster_fn<5>(field, tmax);
    default: return sa_cluster_fn<6>(field, tmax);
    }
}

template <int NPL, int FLD>
void launch_lut_wf(int warps, cudaStream_t s, const SaPass &a)
{
    // a.chunks = warps of work per site; `warps` divides it, so a CTA never straddles two sites
    const unsigned ny = (unsigned)std::min(a.nsites, 65535), nz = (unsigned)((a.nsites + 65534) / 65535);
    if (a.wpt > 1) { // one-warp CTAs, a.wpt words per thread (the last slab may be short: guarded by g < G)
        const dim3 g1((unsigned)(a.wstep / 32), ny, nz);
        mcs_launch_pdl(sa_lut_pass_kernel<NPL, 1, FLD, true>, g1, dim3(32), s, a);
        return;
    }
    const dim3 grid((unsigned)(a.chunks / warps), ny, nz);
    if (warps == 4)
        mcs_launch_pdl(sa_lut_pass_kernel<NPL, 4, FLD>, grid, dim3(128), s, a);
    else if (warps == 2)
        mcs_launch_pdl(sa_lut_pass_kernel<NPL, 2, FLD>, grid, dim3(64), s, a);
    else
        mcs_launch_pdl(sa_lut_pass_kernel<NPL, 1, FLD>, grid, dim3(32), s, a);
}

template <int NPL>
void launch_lut_w(int warps, cudaStream_t s, const SaPass &a)
{
    if (a.field)
        launch_lut_wf<NPL, 1>(warps, s, a);
    else
        launch_lut_wf<NPL, 0>(warps, s, a);
}

void launch_lut(int npl, int warps, cudaStream_t s, const SaPass &a)
{
    switch (npl) {
    case 1: launch_lut_w<1>(warps, s, a); break;
    case 2: launch_lut_w<2>(warps, s, a); break;
    case 3: launch_lut_w<3>(warps, s, a); break;
    case 4: launch_lut_w<4>(warps, s, a); break;
    case 5: launch_lut_w<5>(warps, s, a); break;
    case 6: launch_lut_w<6>(warps, s, a); break;
    case 7: launch_lut_w<7>(warps, s, a); break;
    default: launch_lut_w<8>(warps, s, a); break;
    }
}


// Cluster-resident schedule (sa_cluster_kernel) when the batch is small enough for every cluster to be resident at
// once; returns MCS_OK with *done = false when the shape does not qualify (the caller takes the multi-launch path).
int sa_try_cluster(mcs_state *st, const double *sched, int64_t S, int mcsteps, uint64_t sweep_offset, int npl,
                   const SaPass &a, bool *done)
{
    *done = false;
    mcs_instance *inst = st->inst;
    const char *env = getenv("MCS_CLUSTER");
    if (env && env[0] == '0') return MCS_OK;
    const bool forced = env && env[0] == '1';
    const long long passes = (long long)S * mcsteps * inst->ncolors;
    if (npl > 6 || inst->ncolors > kClMaxColors || passes < 8 || st->G <= 0) return MCS_OK;
    int csize = 16; // measured on B200 (benchmarks/sa_cluster_probe.py): 16 beats 8 and 4 at every batch size
    if (const char *e = getenv("MCS_CLUSTER_SIZE")) csize = atoi(e);
    if (csize != 1 && csize != 2 && csize != 4 && csize != 8 && csize != 16) return MCS_OK;
    SaCluster c;
    int max_per = 0;
    c.loc_base[0] = 0;
    for (int k = 0; k < inst->ncolors; ++k) {
        const int nc = inst->color_start[k + 1] - inst->color_start[k];
        c.color_start[k] = inst->color_start[k];
        c.per[k] = std::max(1, (nc + csize - 1) / csize);
        c.loc_base[k + 1] = c.loc_base[k] + c.per[k];
        max_per = std::max(max_per, c.per[k]);
    }
    c.color_start[inst->ncolors] = inst->color_start[inst->ncolors];
    const int nloc = c.loc_base[inst->ncolors];
    c.nloc_pad = (nloc + 31) / 32 * 32;
    if (max_per > 1024) return MCS_OK;
    const void *fn = sa_cluster_fn(npl, a.field, 1024);
    const int ent = 1 << npl, nq = npl - a.field;
    auto smem_for = [&](int wc, int threads) {
        return (size_t)c.nloc_pad * 4u * (size_t)(ent + wc + nq + npl + 1) + (size_t)threads * 32u;
    };
    // 1024 threads whatever the item count: the threads without an item build the next pass's tables meanwhile
    auto threads_for = [&](int wc) { return wc >= 0 ? 1024 : 0; };
    // function attributes and occupancy answers are asked for ONCE per (device, function, geometry): repeated
    // cudaFuncSetAttribute / cudaOccupancyMaxActiveClusters calls made single small calls erratically slow (spikes
    // of 100 ms in the drop-in sa.Anneal on one configuration)
    static std::mutex plan_mutex;
    static std::set<std::tuple<int, const void *, bool>> attr_done;
    static std::map<std::tuple<int, const void *, int, int, size_t>, int> occupancy_cache;
    static std::map<int, int> smem_cache;
    std::lock_guard<std::mutex> plan_lock(plan_mutex);
    int smem_max = 0;
    if (smem_cache.count(inst->device)) {
        smem_max = smem_cache[inst->device];
    } else {
        MCS_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, inst->device));
        smem_cache[inst->device] = smem_max;
    }
    if (smem_for(1, threads_for(1)) > (size_t)smem_max) return MCS_OK;
    auto prepare_fn = [&](const void *f) -> cudaError_t {
        const auto akey = std::make_tuple(inst->device, f, csize > 8);
        if (attr_done.count(akey)) return cudaSuccess;
        cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
        if (e == cudaSuccess && csize > 8) e = cudaFuncSetAttribute(f, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e == cudaSuccess) attr_done.insert(akey);
        return e;
    };
    MCS_CUDA(prepare_fn(fn));
    // how many clusters are resident together (one CTA per SM: 1024 threads x 64 registers)
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cfg.stream = inst->stream;
    // words per cluster: as few as keep every cluster resident at once (more clusters = more SMs at work), at most
    // what one CTA's threads (one item per thread and pass) and shared memory hold
    int wc_max = std::max(1, 1024 / max_per);
    while (wc_max > 1 && smem_for(wc_max, threads_for(wc_max)) > (size_t)smem_max) --wc_max;
    int wc = 0, resident = 0;
    for (int w = 1; w <= wc_max; ++w) {
        cfg.gridDim = dim3((unsigned)csize);
        cfg.blockDim = dim3((unsigned)threads_for(w));
        cfg.dynamicSmemBytes = smem_for(w, threads_for(w));
        int ncl = 0;
        const auto key = std::make_tuple(inst->device, fn, csize, threads_for(w), (size_t)cfg.dynamicSmemBytes);
        if (occupancy_cache.count(key)) {
            ncl = occupancy_cache[key];
        } else {
            if (cudaOccupancyMaxActiveClusters(&ncl, fn, &cfg) != cudaSuccess) {
                cudaGetLastError();
                return MCS_OK;
            }
            occupancy_cache[key] = ncl;
        }
        resident = ncl;
        if ((st->G + w - 1) / w <= ncl) {
            wc = w;
            break;
        }
    }
    // Measured on B200 (80 x 80 torus, profiles/r02_sa_cluster.log): per SM the cluster kernel decides as fast as the
    // pass kernels, but only the 7 x 16 SMs that hold resident clusters work: 2.6 us per pass up to 224 restarts, 3.2 at
    // 448, 4.1 at 896 against 4.4-4.5 us for one launch per pass at any of these sizes (and ONE launch instead of
    // thousands on the host); at 1000 items per CTA (1024 restarts, cfg2) the two tie, so the pass kernels keep it.
    // Batches whose clusters do not all fit at once take the pass kernels too.
    if (!forced && wc > 0 && (long long)max_per * wc > 800) wc = 0;
    if (const char *e = getenv("MCS_CLUSTER_WORDS")) {
        const int w = atoi(e);
        if (w >= 1 && w <= wc_max) wc = w;
    }
    if (wc == 0) {
        if (!forced) return MCS_OK; // more clusters than fit at once: the batch is large, one launch per pass is faster
        wc = wc_max;
    }
    const long long nclusters = (st->G + wc - 1) / wc;
    int threads = threads_for(wc);
    if ((long long)max_per * wc <= 512 && !getenv("MCS_CLUSTER_1024")) { // the form without the register cap
        threads = 512;
        fn = sa_cluster_fn(npl, a.field, 512);
        MCS_CUDA(prepare_fn(fn));
    }
    std::vector<float> nl((size_t)S);
    for (int64_t t = 0; t < S; ++t) nl[(size_t)t] = (float)(-1.4426950408889634 / sched[t]); // sa.pyx:98
    // the schedule goes to the batch's staging buffer (stream ordered with the conversions that also use it)
    MCS_TRY(mcs_state_reserve_stage(st, (size_t)S * sizeof(float) + 64));
    float *d_nl = reinterpret_cast<float *>(st->d_stage);
    MCS_CUDA(cudaMemcpyAsync(d_nl, nl.data(), (size_t)S * sizeof(float), cudaMemcpyHostToDevice, inst->stream));
    c.V = st->d_V;
    c.ell_idx = inst->d_ell_idx;
    c.ell_J = inst->d_ell_J;
    c.h = inst->d_h;
    c.ellJ_stride = inst->nsteps > 1 ? (long long)inst->N * inst->dpad : 0;
    c.h_stride = inst->nsteps > 1 ? (long long)inst->N : 0;
    c.order = inst->d_order;
    c.pos = inst->d_pos;
    c.nl2e = d_nl;
    c.prof = getenv("MCS_CLUSTER_PROF") ? reinterpret_cast<long long *>(d_nl + ((S + 1) / 2) * 2) : nullptr;
    c.prof_cta = c.prof ? atoi(getenv("MCS_CLUSTER_PROF")) : 0;
    c.ncolors = inst->ncolors;
    c.dpad = inst->dpad;
    c.S = (int)S;
    c.mcsteps = mcsteps;
    c.Wc = wc;
    c.csize = csize;
    c.G = st->G;
    c.sweep_offset = sweep_offset;
    c.keys = a.keys;
    c.pow2 = a.pow2;
    c.word_offset = a.word_offset;
    c.tie_thr = a.tie_thr;
    cfg.gridDim = dim3((unsigned)(nclusters * csize));
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem_for(wc, threads);
    void *args[] = {&c};
    // the host copy of the schedule must outlive the asynchronous upload: pageable memory is staged by the runtime
    // before cudaMemcpyAsync returns
    const cudaError_t le = cudaLaunchKernelExC(&cfg, fn, args);
    inst->launches++;
    if (c.prof && le == cudaSuccess) {
        long long hp[4] = {0, 0, 0, 0};
        MCS_CUDA(cudaMemcpyAsync(hp, c.prof, sizeof(hp), cudaMemcpyDeviceToHost, inst->stream));
        MCS_CUDA(cudaStreamSynchronize(inst->stream));
        fprintf(stderr, "[mcs cluster] per pass: %.0f cycles item, %.0f arrive + tables, %.0f CTA barrier, %.0f cluster wait\n",
                (double)hp[0] / passes, (double)hp[1] / passes, (double)hp[2] / passes, (double)hp[3] / passes);
    }
    MCS_CUDA(le);
    if (getenv("MCS_CLUSTER_VERBOSE"))
        fprintf(stderr, "[mcs cluster] %lld clusters (%d resident) of %d CTAs x %d threads, %d words per cluster, %zu B smem, %lld passes\n",
                nclusters, resident, csize, threads, wc, (size_t)cfg.dynamicSmemBytes, passes);
    *done = true;
    return MCS_OK;
}

} // namespace

int mcs_launch_sa_sweeps(mcs_state *st, const double *sched, int64_t S, int mcsteps, uint64_t seed,
                         uint64_t replica_offset, uint64_t sweep_offset)
{
    mcs_instance *inst = st->inst;
    MCS_REQUIRE((replica_offset & 31) == 0, MCS_EINVAL,
                "mcs_sa_sweeps: replica_offset must be a multiple of 32 (restarts are packed 32 per word)");
    MCS_CUDA(cudaSetDevice(inst->device));
    if (inst->dynamics == MCS_DYN_REFERENCE)
        return mcs_launch_refdyn_sweeps(st, MCS_KIND_SA, sched, nullptr, S, mcsteps, 0.0f, 0, seed, replica_offset,
                                        sweep_offset);
    if (mcs_dense_supported(inst, 1))
        return mcs_launch_dense_sweeps(st, MCS_KIND_SA, sched, nullptr, S, mcsteps, 0.0f, 0, seed, replica_offset,
                                       sweep_offset);
    SaPass a;
    a.V = st->d_V;
    a.ell_idx = inst->d_ell_idx;
    a.ell_J = inst->d_ell_J;
    a.h = inst->d_h;
    a.dpad = inst->dpad;
    a.nq = inst->maxdeg;
    a.field = inst->has_field ? 1 : 0;
    a.G = st->G;
    a.Gs = st->G;
    a.wpt = 1;
    a.wstep = 0;
    a.chunks = (int)((st->G + 31) / 32);
    a.keys = mcs_philox_expand(seed);
    a.pow2 = mcs_pow2_make();
    a.word_offset = (uint32_t)(replica_offset >> 5);
    a.tie_thr = mcs_tie_threshold();
    const int npl = std::max(1, inst->maxdeg + (inst->has_field ? 1 : 0));
    const bool lut = npl <= 8;
    const int warps = (a.chunks % 4 == 0) ? 4 : (a.chunks % 2 == 0) ? 2 : 1;
    uint64_t sweep = sweep_offset;
    MCS_REQUIRE(inst->nsteps == 1 || S <= inst->nsteps, MCS_EINVAL,
                "time-dependent instance has %lld tables but the schedule has %lld steps", (long long)inst->nsteps,
                (long long)S);
    // ---- small batches: the whole schedule resident in thread-block clusters (sa_cluster_kernel) ----
    if (lut) {
        bool done = false;
        MCS_TRY(sa_try_cluster(st, sched, S, mcsteps, sweep_offset, npl, a, &done));
        if (done) return MCS_OK;
    }
    // Larger batches (more than 2048 restarts: three or more warps of words per site; at 2048 two streams of one-warp
    // launches are limited by the host's launch rate: 0.97e12 against 1.22e12 attempts/s): as in the PIQMC launcher, the words are
    // cut into two chunks whose colour passes alternate on two streams, and a one-warp CTA takes all the words of its
    // site and chunk one after the other (up to 64 per thread), sharing the site's set-up.  Same counters (global
    // word index): identical results (tests).  MCS_SA_WPT=1 keeps one word per thread on one stream.
    const int nwarps = a.chunks; // 32-word warps per site
    int nchunk = 1, multi = 0;
    if (lut && nwarps >= 3 && !(getenv("MCS_SA_WPT") && atoi(getenv("MCS_SA_WPT")) <= 1)) {
        multi = 1;
        nchunk = getenv("MCS_ONE_STREAM") ? 1 : 2;
        if (const char *e = getenv("MCS_STREAMS")) nchunk = std::max(1, std::min(std::min(4, nwarps / 2), atoi(e)));
    }
    if (nchunk > 1 && !inst->ev_aux0) MCS_CUDA(cudaEventCreateWithFlags(&inst->ev_aux0, cudaEventDisableTiming));
    for (int q = 0; q + 1 < nchunk; ++q) {
        if (!inst->s_aux[q]) {
            MCS_CUDA(cudaStreamCreateWithFlags(&inst->s_aux[q], cudaStreamNonBlocking));
            MCS_CUDA(cudaEventCreateWithFlags(&inst->ev_aux1[q], cudaEventDisableTiming));
        }
    }
    if (nchunk > 1) {
        MCS_CUDA(cudaEventRecord(inst->ev_aux0, inst->stream));
        for (int q = 0; q + 1 < nchunk; ++q) MCS_CUDA(cudaStreamWaitEvent(inst->s_aux[q], inst->ev_aux0, 0));
    }
    uint32_t *const V0 = a.V;
    const uint32_t woff0 = a.word_offset;
    const long long Gall = a.G;
    for (int64_t t = 0; t < S; ++t) {
        a.ell_J = inst->ell_J_at(t); // sa.NoisyAnneal: nbs[itemp] (sa.pyx:363-365)
        a.h = inst->h_at(t);
        // exp(-ediff/temp), sa.pyx:98; temp == 0 gives -inf here -> threshold "never" for ediff > 0
        a.nl2e_over_t = (float)(-1.4426950408889634 / sched[t]);
        for (int step = 0; step < mcsteps; ++step, ++sweep) {
            a.sweep_lo = (uint32_t)sweep;
            a.sweep_hi = (uint32_t)(sweep >> 32);
            for (int c = 0; c < inst->ncolors; ++c) {
                a.sites = inst->d_order + inst->color_start[c];
                a.nsites = inst->color_start[c + 1] - inst->color_start[c];
                if (a.nsites == 0) continue;
                if (multi) {
                    for (int q = 0; q < nchunk; ++q) {
                        const int w0 = nwarps * q / nchunk, w1 = nwarps * (q + 1) / nchunk; // warps of this chunk
                        a.V = V0 + 32ll * w0;
                        a.word_offset = woff0 + 32u * (uint32_t)w0;
                        a.G = std::min(Gall, 32ll * w1) - 32ll * w0;
                        int want = 64;
                        if (const char *e = getenv("MCS_SA_WPT")) want = atoi(e);
                        a.wpt = 1;
                        while (2 * a.wpt <= want && 2 * a.wpt <= w1 - w0) a.wpt *= 2;
                        a.wstep = 32u * (uint32_t)((w1 - w0 + a.wpt - 1) / a.wpt);
                        if (a.wpt == 1) a.wpt = 2; // (the MULTI kernel with a second, out-of-range word: skipped)
                        launch_lut(npl, 1, q ? inst->s_aux[q - 1] : inst->stream, a);
                        inst->launches++;
                    }
                    continue;
                }
                const long long items = (long long)a.nsites * a.chunks;
                if (lut)
                    launch_lut(npl, warps, inst->stream, a);
                else
                    sa_direct_pass_kernel<<<(unsigned)((items + kWarps - 1) / kWarps), kWarps * 32, 0,
                                            inst->stream>>>(a);
                inst->launches++;
            }
        }
    }
    for (int q = 0; q + 1 < nchunk; ++q) {
        MCS_CUDA(cudaEventRecord(inst->ev_aux1[q], inst->s_aux[q]));
        MCS_CUDA(cudaStreamWaitEvent(inst->stream, inst->ev_aux1[q], 0));
    }
    MCS_CUDA(mcs_take_launch_error());
    MCS_CUDA(cudaGetLastError());
    return MCS_OK;
}

int mcs_sa_pack(mcs_state *st, const int8_t *d_in)
{
    mcs_instance *inst = st->inst;
    const long long n = inst->N * st->G;
    sa_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, inst->stream>>>(d_in, st->d_V, inst->N, st->R, st->G);
    inst->launches++;
    MCS_CUDA(cudaGetLastError());
    return MCS_OK;
}

int mcs_sa_unpack(mcs_state *st, int8_t *d_out)
{
    mcs_instance *inst = st->inst;
    const long long n = inst->N * st->G;
    sa_unpack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, inst->stream>>>(st->d_V, d_out, inst->N, st->R, st->G);
    inst->launches++;
    MCS_CUDA(cudaGetLastError());
    return MCS_OK;
}

int mcs_sa_init(mcs_state *st, uint64_t seed, uint64_t replica_offset)
{
    mcs_instance *inst = st->inst;
    const long long n = inst->N * st->G;
    sa_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, inst->stream>>>(
        st->d_V, inst->N, st->R, st->G, (uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)replica_offset);
    inst->launches++;
    MCS_CUDA(cudaGetLastError());
    return MCS_OK;
}

int mcs_sa_energy(mcs_state *st, double *d_out)
{
    mcs_instance *inst = st->inst;
    const char *chain = getenv("MCS_ENERGY_CHAIN"); // tests: the chain kernel
    if (!chain && mcs_energy_tables(inst)) {
        const int threads = st->R >= 128 * 296 ? 128 : (st->R >= 64 * 296 ? 64 : 32);
        sa_energy_lut_kernel<8><<<(unsigned)((st->R + threads - 1) / threads), threads, 0, inst->stream>>>(
            st->d_V, inst->d_etab, inst->d_etab_j, d_out, inst->N, st->R, st->G);
        inst->launches++;
        MCS_CUDA(cudaGetLastError());
        return MCS_OK;
    }
    sa_energy_kernel<<<(unsigned)((st->R + 63) / 64), 64, 0, inst->stream>>>(
        st->d_V, inst->tab_idx_at(inst->nsteps - 1), inst->tab_J_at(inst->nsteps - 1), d_out, inst->N, (int)inst->maxnb, st->R, st->G);
    inst->launches++;
    MCS_CUDA(cudaGetLastError());
    return MCS_OK;
}
