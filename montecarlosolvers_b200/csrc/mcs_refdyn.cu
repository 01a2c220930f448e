// mcs_refdyn.cu -- "reference dynamics" production sweeps (sm_100a).
//
// The reference visits the sites of a slice strictly one after another in a fresh random permutation
// (Fisher-Yates, qmc.pyx:102-108; sa.pyx:73-79), slice after slice (qmc.pyx:99), then (QuantumAnnealGlobal)
// the world lines in one more permutation (qmc.pyx:405-438).  The coloured kernels of mcs_piqmc.cu / mcs_sa.cu
// sample the same Boltzmann distribution but in checkerboard order, and an anneal is a non-equilibrium
// process: their residual energies come out 6-9 % lower than the reference's.
//
// This file reproduces the reference's dynamics IN DISTRIBUTION at production speed.  The outcome of a
// sequential sweep in visiting order pi depends on pi only through the orientation it induces on the edges of
// the interaction graph ("which endpoint is visited first"): any execution that respects that acyclic
// orientation produces the identical state.  So, per (replica, sweep, slice):
//   1. every site draws a random 16-bit priority from Philox (ties broken by the site index: the induced
//      orientation is that of a uniformly random permutation up to 2^-16 ties per edge);
//   2. a site is visited as soon as all its neighbours with a smaller priority have been visited -- the sites
//      whose predecessor count is zero form the ready queue of the current round, finishing a site decrements
//      the counters of its later neighbours and pushes those that reach zero into the next round's queue.
//      A slice takes about 15 rounds on the 80x80 lattice instead of 6400 sequential visits;
//   3. slices are processed in order, so slice k sees slice k-1 already updated and slice k+1 stale
//      (qmc.pyx:127-138), exactly like the reference.
// One CTA owns one replica for the WHOLE schedule: its world lines live bit-packed in shared memory (51 KB at
// N = 6400, P = 64), the schedule loop runs inside the kernel (one launch per call), and the only global traffic
// is the coupling table (L2-resident, shared by all CTAs).
//
// The spin-vector solvers (svmc.pyx:78-117, 181-229) shuffle and visit the same way (svmc.pyx:83-91): refdyn_svmc_kernel
// runs the rotor attempt on the same wave machinery (theta and cos theta of one read in shared memory).
//
// The acceptance arithmetic is the production one: fp32 dE in ELL order, threshold T = ceil(exp(-dE/teff) 2^32) - 1,
// accept iff a 32-bit uniform u <= T (qmc.pyx:140-143); u = (v << 16) | r with r drawn lazily only when v alone
// does not decide.  `sequential = 1` (test hook) lets thread 0 walk the sites in increasing (priority, index)
// order instead: the two executions must agree bit for bit (tests/test_gpu_refdyn.py).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <type_traits>
#include <vector>

#include "mcs_common.cuh"

namespace {

constexpr uint32_t kTagPrio = 0x20u;   // c3 low byte: priorities + upper half of the acceptance uniforms
constexpr uint32_t kTagRefine = 0x21u; // lower half of the acceptance uniforms (lazy)
constexpr uint32_t kTagProp = 0x22u;   // SVMC proposals

struct RefdynArgs {
    uint64_t *W;  // PIQMC [N][Rpad] (window base), bit k = slice k
    uint32_t *V;  // SA    [N][G], bit b of word g = replica 32 g + b
    float *theta, *cosz; // SVMC [N][Rpad]
    const int32_t *ell_idx; // [N][dpad]
    const float *ell_J;     // [nsteps][N][dpad]
    const float *h;         // [nsteps][N]
    long long ellJ_stride, h_stride; // elements between the tables of consecutive schedule steps (0: static)
    const float *bcoef;  // [S]  -2 B          (qmc.pyx:96; SA: -2)
    const float *jperp2; // [S]  2 J_perp      (qmc.pyx:95)
    const float *nl2e;   // [S]  -log2(e)/teff (SA: -log2(e)/sched[t])
    const float *acoef;  // [S]  SVMC only: A; there bcoef = B and jperp2 = min(1, A/B) (svmc.pyx:198-202)
    long long Rpad, G;
    int N, Npad, dpad, field, P, S, mcsteps, global_moves, tf, sequential;
    uint32_t replica_offset;
    uint64_t sweep_offset;
    mcs_philox_keys keys;
    long long *prof; // MCS_REFDYN_PROF=1: per-CTA cycle counters {priorities, counts, rounds, #rounds, #passes}
    int *err; // device flag: set when a pass could not make progress (cannot happen for an acyclic orientation)
};

__device__ __forceinline__ int rd_popc(uint32_t x) { return __popc(x); }
__device__ __forceinline__ int rd_popc(uint64_t x) { return __popcll(x); }

// Warp-aggregated append of up to DP sites per lane: every lane of a converged warp calls this with the bit mask
// of its ready neighbours (bit j <-> nb[j]).  One shared-memory atomic per warp.
template <int DP>
__device__ __forceinline__ void rd_push_many(uint32_t rmask, const int (&nb)[DP], uint16_t *queue, int *counter,
                                             int base, int lane)
{
    if (__ballot_sync(0xFFFFFFFFu, rmask != 0u) == 0u) return;
    const int n = __popc(rmask);
    int incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += t;
    }
    int pos = 0;
    if (lane == 31) pos = atomicAdd(counter, incl);
    pos = __shfl_sync(0xFFFFFFFFu, pos, 31) + base + incl - n;
#pragma unroll
    for (int j = 0; j < DP; ++j)
        if ((rmask >> j) & 1u) queue[pos++] = (uint16_t)nb[j];
}

__device__ __forceinline__ void rd_push(bool cond, uint32_t site, uint16_t *queue, int *counter, int base, int lane)
{
    const uint32_t mask = __ballot_sync(0xFFFFFFFFu, cond);
    if (mask == 0u) return;
    const int leader = __ffs(mask) - 1;
    int pos = 0;
    if (lane == leader) pos = atomicAdd(counter, __popc(mask));
    pos = __shfl_sync(0xFFFFFFFFu, pos, leader);
    if (cond) queue[base + pos + __popc(mask & ((1u << lane) - 1u))] = (uint16_t)site;
}

// Metropolis rule of qmc.pyx:140-143 / sa.pyx:96-99 on a lazily refined 32-bit uniform.  Strict "never" when
// exp(-dE/teff) underflows to 0 or dE is NaN (the reference compares 0 > rand()/RAND_MAX: false).
template <typename Refine>
__device__ __forceinline__ bool rd_accept(float dE, float nl2e, uint32_t v16, Refine refine)
{
    if (dE <= 0.0f) return true;
    const float t = ceilf(exp2f(dE * nl2e) * 4294967296.0f);
    if (!(t >= 1.0f)) return false;
    if (t >= 4294967296.0f) return true;
    const uint32_t T = (uint32_t)t - 1u, Thi = T >> 16;
    if (v16 < Thi) return true;
    if (v16 > Thi) return false;
    return ((v16 << 16) | refine()) <= T;
}

enum { RD_LOCAL = 0, RD_GLOBAL = 1, RD_SA = 2 };

// Lower 16 bits of site i's acceptance uniform.  Needed once in 2^16 attempts: out of line, so that ptxas cannot
// if-convert the Philox call into the hot path (it did: 15 % of the executed instructions, ncu r2_refdyn_v2).
static __device__ __noinline__ uint32_t rd_refine_draw(uint32_t c0, uint32_t i, uint32_t c2, uint32_t c3, uint32_t k0,
                                                       uint32_t k1)
{
    uint32_t x[4];
    mcs_philox4x32(c0, i >> 2, c2, c3, k0, k1, x);
    return x[i & 3u] >> 16;
}

// Entries [j0, j0 + DP) of site i's ELL row, all loads issued together (the table is L2-resident: one round trip
// per chunk instead of one per entry).  Entries beyond the row read as (i, 0): the padding convention of the table.
template <int DP, bool WITH_J>
__device__ __forceinline__ void rd_load_row(const RefdynArgs &a, const float *ellJ, int i, int j0, int (&nb)[DP],
                                            float (&J)[DP])
{
    const int32_t *idx = a.ell_idx + (size_t)i * a.dpad + j0;
    const float *Jr = ellJ + (size_t)i * a.dpad + j0;
    if (DP == 4 && a.dpad == 4) { // 16-byte rows (the 2-D lattices)
        const int4 v = __ldg(reinterpret_cast<const int4 *>(idx));
        nb[0] = v.x, nb[1] = v.y, nb[2] = v.z, nb[3 % DP] = v.w;
        if (WITH_J) {
            const float4 f = __ldg(reinterpret_cast<const float4 *>(Jr));
            J[0] = f.x, J[1] = f.y, J[2] = f.z, J[3 % DP] = f.w;
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < DP; ++j) {
        const bool in = j0 + j < a.dpad;
        nb[j] = in ? __ldg(idx + j) : i;
        if (WITH_J) J[j] = in ? __ldg(Jr + j) : 0.0f;
    }
}

// In-plane part of the energy difference over one chunk of the row (fp32, ELL order)
template <typename WT, int MODE, int DP>
__device__ __forceinline__ float rd_row_dE(const WT *w, WT wi, const int (&nb)[DP], const float (&J)[DP], float bcoef,
                                           int k, int P, WT pmask)
{
    float dE = 0.0f;
#pragma unroll
    for (int j = 0; j < DP; ++j) {
        const float cj = bcoef * J[j]; // padding: nb == i, J = 0
        const WT x = wi ^ w[nb[j]];
        if (MODE == RD_GLOBAL)
            dE += cj * (float)(P - 2 * rd_popc((WT)(x & pmask)));
        else
            dE += ((x >> k) & (WT)1) ? -cj : cj;
    }
    return dE;
}

// Field and Trotter terms, acceptance, flip
template <typename WT, int MODE>
__device__ __forceinline__ void rd_finish(const RefdynArgs &a, WT *w, const uint32_t *pu, int i, WT wi, float dE, int k,
                                          int P, WT pmask, const float *hrow, float bcoef, float jperp2, float nl2e,
                                          uint32_t c0, uint32_t c2, uint32_t c3base)
{
    if (a.field) {
        const float hc = bcoef * __ldg(hrow + i);
        if (MODE == RD_GLOBAL)
            dE += hc * (float)(P - 2 * rd_popc((WT)(wi & pmask)));
        else
            dE += ((wi >> k) & (WT)1) ? -hc : hc;
    }
    if (MODE == RD_LOCAL) { // qmc.pyx:127-138 (for P == 2 both neighbours are the same slice, both counted)
        const int kl = k == 0 ? P - 1 : k - 1, kr = k == P - 1 ? 0 : k + 1;
        const int anti = (int)(((wi >> kl) ^ (wi >> k)) & (WT)1) + (int)(((wi >> kr) ^ (wi >> k)) & (WT)1);
        dE += jperp2 * (float)(2 - 2 * anti);
    }
    const bool acc = rd_accept(dE, nl2e, pu[i] >> 16, [&]() -> uint32_t {
        return rd_refine_draw(c0, (uint32_t)i, c2, c3base | kTagRefine, a.keys.rk[0], a.keys.rk[1]);
    });
    if (acc) w[i] = wi ^ (MODE == RD_GLOBAL ? pmask : (WT)((WT)1 << k));
}

// ---- the wave machinery, shared by the Ising and the rotor kernels ---------------------------------------------
// Per-CTA bookkeeping of the dependency waves.  Shared memory: pu[Npad] (upper half of the acceptance uniform : 16 |
// priority : 16), cnt[Npad] bytes (open predecessors, packed four to a word), queue[Npad] u16, cntr[3].
struct RdWaves {
    uint32_t *pu;
    uint32_t *cnt32;
    uint16_t *queue;
    int *cntr;   // round q counts its pushes in cntr[q % 3]
    int rnd = 0; // running round number (uniform over the CTA)
    long long pf0 = 0, pf1 = 0, pf2 = 0;
    int pf_rounds = 0, pf_passes = 0;
};

// A `Site` policy supplies the solver-specific part of a visit:
//   draw(q)              priorities + uniforms of sites 4q .. 4q+3 for this pass (into wv.pu and its own arrays)
//   Acc begin(i)         start of a visit (reads the site's own state)
//   row(acc, nb, J)      one chunk of the coupling row (reads the neighbours' state)
//   finish(i, acc)       remaining terms, acceptance, write-back
template <int DP, typename Site>
__device__ __forceinline__ void rd_visit(const RefdynArgs &a, const float *ellJ, Site &site, int i)
{
    typename Site::Acc acc = site.begin(i);
    for (int j0 = 0; j0 < a.dpad; j0 += DP) {
        int nb[DP];
        float J[DP];
        rd_load_row<DP, true>(a, ellJ, i, j0, nb, J);
        site.row(acc, nb, J);
    }
    site.finish(i, acc);
}

// one pass = every site visited once, in the order of this pass's priorities
template <int DP, typename Site>
__device__ __forceinline__ void rd_pass(const RefdynArgs &a, RdWaves &wv, const float *ellJ, Site &site)
{
    const int N = a.N, Npad = a.Npad, T = blockDim.x, tid = threadIdx.x, lane = tid & 31, wbase = tid & ~31;
    const bool one_chunk = a.dpad <= DP;
    const uint32_t *pu = wv.pu;
    uint32_t *cnt32 = wv.cnt32;
    uint8_t *cnt8 = reinterpret_cast<uint8_t *>(cnt32);
    uint16_t *queue = wv.queue;
    long long t0 = a.prof ? clock64() : 0;
    for (int q = tid; q < Npad / 4; q += T) site.draw(q);
    __syncthreads();
    auto key = [&](int s) -> uint32_t { return (pu[s] << 16) | (uint32_t)s; };
    if (a.sequential) { // test hook: thread 0 walks the sites in increasing (priority, index) order
        if (tid == 0) {
            for (int i = 0; i < N; ++i) cnt8[i] = 0;
            for (int n = 0; n < N; ++n) {
                int best = -1;
                uint32_t bk = 0xFFFFFFFFu;
                for (int i = 0; i < N; ++i)
                    if (!cnt8[i] && key(i) <= bk) {
                        bk = key(i);
                        best = i;
                    }
                rd_visit<DP>(a, ellJ, site, best);
                cnt8[best] = 1;
            }
        }
        __syncthreads();
        return;
    }
    if (a.prof) {
        const long long t1 = clock64();
        wv.pf0 += t1 - t0;
        t0 = t1;
    }
    // round 0: open-predecessor counts (independent iterations: the loads of several sites overlap) ...
#pragma unroll 2
    for (int i = tid; i < N; i += T) {
        const uint32_t ki = key(i);
        int c = 0;
        for (int j0 = 0; j0 < a.dpad; j0 += DP) {
            int nb[DP];
            float J[DP];
            rd_load_row<DP, false>(a, ellJ, i, j0, nb, J);
#pragma unroll
            for (int j = 0; j < DP; ++j) c += (nb[j] != i && key(nb[j]) < ki) ? 1 : 0;
        }
        cnt8[i] = (uint8_t)c;
    }
    // ... then the sites without predecessors start the queue (each thread re-reads its own bytes)
    int *cn = &wv.cntr[wv.rnd % 3];
    for (int i0 = wbase; i0 < N; i0 += T) {
        const int i = i0 + lane;
        rd_push(i < N && cnt8[i] == 0, (uint32_t)i, queue, cn, 0, lane);
    }
    __syncthreads();
    int head = 0, tail = *(volatile int *)cn;
    if (tid == 0) wv.cntr[(wv.rnd + 2) % 3] = 0; // last read before the barrier above; next used in round rnd + 2
    ++wv.rnd;
    if (a.prof) {
        const long long t1 = clock64();
        wv.pf1 += t1 - t0;
        t0 = t1;
        wv.pf_passes += 1;
    }
    // notifications of a finished site: decrement the open-predecessor count of its later neighbours and queue
    // those that reach zero
    auto notify = [&](bool active, int i, uint32_t ki, const int (&nb)[DP]) {
        uint32_t rmask = 0u;
#pragma unroll
        for (int j = 0; j < DP; ++j) {
            const int nbj = nb[j];
            if (active && nbj != i && key(nbj) > ki) {
                const int sh = 8 * (nbj & 3);
                const uint32_t old = atomicSub(&cnt32[nbj >> 2], 1u << sh);
                if (((old >> sh) & 0xFFu) == 1u) rmask |= 1u << j;
            }
        }
        rd_push_many<DP>(rmask, nb, queue, cn, tail, lane);
    };
    while (head < N) {
        wv.pf_rounds += 1;
        if (tail == head) { // no progress: impossible for an acyclic orientation
            if (tid == 0 && a.err) atomicExch(a.err, 1);
            break;
        }
        cn = &wv.cntr[wv.rnd % 3];
        for (int e0 = head + wbase; e0 < tail; e0 += T) {
            const int e = e0 + lane;
            const bool active = e < tail;
            const int i = active ? (int)queue[e] : 0;
            const uint32_t ki = key(i);
            typename Site::Acc acc = site.begin(i);
            for (int j0 = 0; j0 < a.dpad; j0 += DP) {
                int nb[DP];
                float J[DP];
                rd_load_row<DP, true>(a, ellJ, i, j0, nb, J);
                if (active) site.row(acc, nb, J);
                if (!one_chunk) continue;
                // rows of one chunk: attempt, then tell the later neighbours, from the same registers
                if (active) site.finish(i, acc);
                notify(active, i, ki, nb);
            }
            if (one_chunk) continue;
            if (active) site.finish(i, acc);
            for (int j0 = 0; j0 < a.dpad; j0 += DP) { // long rows: second walk for the notifications
                int nb[DP];
                float J[DP];
                rd_load_row<DP, false>(a, ellJ, i, j0, nb, J);
                notify(active, i, ki, nb);
            }
        }
        __syncthreads();
        head = tail;
        tail += *(volatile int *)cn;
        if (tid == 0) wv.cntr[(wv.rnd + 2) % 3] = 0;
        ++wv.rnd;
    }
    if (a.prof) wv.pf2 += clock64() - t0;
}

__device__ __forceinline__ void rd_store_prof(const RefdynArgs &a, const RdWaves &wv)
{
    if (a.prof && threadIdx.x == 0) {
        long long *o = a.prof + (size_t)blockIdx.x * 5;
        o[0] = wv.pf0, o[1] = wv.pf1, o[2] = wv.pf2, o[3] = wv.pf_rounds, o[4] = wv.pf_passes;
    }
}

// ---- Ising sites (PIQMC slices, world lines, SA) ----------------------------------------------------------------
template <typename WT, int MODE>
struct IsingSite {
    struct Acc {
        WT wi;
        float dE;
    };
    const RefdynArgs &a;
    WT *w;
    uint32_t *pu;
    int k, P;
    WT pmask;
    const float *hrow;
    float bcoef, jperp2, nl2e;
    uint32_t c0, c2, c3base;

    __device__ __forceinline__ void draw(int q) const
    { // priorities and the upper halves of the uniforms
        uint32_t x[4];
        mcs_philox4x32_rk(c0, (uint32_t)q, c2, c3base | kTagPrio, a.keys, x);
        reinterpret_cast<uint4 *>(pu)[q] = make_uint4(x[0], x[1], x[2], x[3]);
    }
    __device__ __forceinline__ Acc begin(int i) const { return Acc{w[i], 0.0f}; }
    template <int DP>
    __device__ __forceinline__ void row(Acc &acc, const int (&nb)[DP], const float (&J)[DP]) const
    {
        acc.dE += rd_row_dE<WT, MODE, DP>(w, acc.wi, nb, J, bcoef, k, P, pmask);
    }
    __device__ __forceinline__ void finish(int i, const Acc &acc) const
    {
        rd_finish<WT, MODE>(a, w, pu, i, acc.wi, acc.dE, k, P, pmask, hrow, bcoef, jperp2, nl2e, c0, c2, c3base);
    }
};

// Shared memory: w[Npad] | pu[Npad] | cnt[Npad] bytes | queue[Npad] u16
// DP: chunk of the ELL row held in registers (4: rows of at most 4 entries are one chunk; 8 otherwise)
template <typename WT, bool SA, int DP>
__global__ void __launch_bounds__(512) refdyn_ising_kernel(const __grid_constant__ RefdynArgs a)
{
    extern __shared__ __align__(16) unsigned char rd_smem[];
    __shared__ int s_cntr[3];
    const int N = a.N, Npad = a.Npad, T = blockDim.x, tid = threadIdx.x;
    WT *w = reinterpret_cast<WT *>(rd_smem);
    RdWaves wv;
    wv.pu = reinterpret_cast<uint32_t *>(w + Npad);
    wv.cnt32 = wv.pu + Npad;
    wv.queue = reinterpret_cast<uint16_t *>(wv.cnt32 + Npad / 4);
    wv.cntr = s_cntr;
    const long long r = blockIdx.x;
    const int P = SA ? 1 : a.P;
    const WT pmask = (P == (int)(8 * sizeof(WT))) ? (WT)~(WT)0 : (WT)(((WT)1 << P) - (WT)1);
    const uint32_t c0 = a.replica_offset + (uint32_t)r;

    for (int i = tid; i < Npad; i += T) {
        WT v = 0;
        if (i < N) v = SA ? (WT)((a.V[(size_t)i * a.G + (r >> 5)] >> (r & 31)) & 1u) : (WT)a.W[(size_t)i * a.Rpad + r];
        w[i] = v;
    }
    if (tid < 3) s_cntr[tid] = 0;
    __syncthreads();

    for (int f = 0; f < a.S; ++f) {
        const float *ellJ = a.ell_J + (size_t)f * a.ellJ_stride;
        const float *hrow = a.h + (size_t)f * a.h_stride;
        const float bcoef = __ldg(&a.bcoef[f]), jperp2 = __ldg(&a.jperp2[f]), nl2e = __ldg(&a.nl2e[f]);
        for (int step = 0; step < a.mcsteps; ++step) {
            const uint64_t sweep = a.sweep_offset + (uint64_t)f * (uint64_t)a.mcsteps + (uint64_t)step;
            const uint32_t c2 = (uint32_t)sweep, c3hi = (uint32_t)(sweep >> 32) << 16;
            if (SA) {
                IsingSite<WT, RD_SA> site{a, w, wv.pu, 0, P, pmask, hrow, bcoef, jperp2, nl2e, c0, c2, c3hi};
                rd_pass<DP>(a, wv, ellJ, site);
            } else {
                for (int k = 0; k < P; ++k) {
                    IsingSite<WT, RD_LOCAL> site{a,     w,      wv.pu, k,  P,  pmask,
                                                 hrow,  bcoef,  jperp2, nl2e, c0, c2, c3hi | ((uint32_t)k << 8)};
                    rd_pass<DP>(a, wv, ellJ, site);
                }
                if (a.global_moves) {
                    IsingSite<WT, RD_GLOBAL> site{a,     w,      wv.pu, 0,  P,  pmask,
                                                  hrow,  bcoef,  jperp2, nl2e, c0, c2, c3hi | (64u << 8)};
                    rd_pass<DP>(a, wv, ellJ, site);
                }
            }
        }
    }
    __syncthreads();
    rd_store_prof(a, wv);
    for (int i = tid; i < N; i += T) {
        if (SA) {
            const uint32_t bit = 1u << (r & 31);
            uint32_t *p = &a.V[(size_t)i * a.G + (r >> 5)];
            if (w[i] & (WT)1)
                atomicOr(p, bit);
            else
                atomicAnd(p, ~bit);
        } else {
            a.W[(size_t)i * a.Rpad + r] = (uint64_t)w[i];
        }
    }
}

// ---- rotor sites (svmc.pyx:78-117, TF: 181-229) ----------------------------------------------------------------
// Visit of svmc.pyx:92-115 in fp32: z = sum_j J_ij cos(theta_j) + h_i over the ELL row, proposal from a 24-bit
// uniform (theta' = pi u, or the TF displacement clamped to [0, pi]), dE = B (cos theta' - cos theta_i) z +
// A (sin theta_i - sin theta'), Metropolis rule as above.
struct RotorSite {
    using Acc = float;
    const RefdynArgs &a;
    float *th, *cz;  // shared memory: theta and its cosine
    uint32_t *pu;    // as above
    uint32_t *prop;  // 32-bit proposal uniform per site
    const float *hrow;
    float acoef, bcoef, tfscale, nl2e;
    uint32_t c0, c2, c3base;

    __device__ __forceinline__ void draw(int q) const
    {
        uint32_t x[4];
        mcs_philox4x32_rk(c0, (uint32_t)q, c2, c3base | kTagPrio, a.keys, x);
        reinterpret_cast<uint4 *>(pu)[q] = make_uint4(x[0], x[1], x[2], x[3]);
        mcs_philox4x32_rk(c0, (uint32_t)q, c2, c3base | kTagProp, a.keys, x);
        reinterpret_cast<uint4 *>(prop)[q] = make_uint4(x[0], x[1], x[2], x[3]);
    }
    __device__ __forceinline__ Acc begin(int i) const { return a.field ? __ldg(hrow + i) : 0.0f; }
    template <int DP>
    __device__ __forceinline__ void row(Acc &z, const int (&nb)[DP], const float (&J)[DP]) const
    {
#pragma unroll
        for (int j = 0; j < DP; ++j) z = fmaf(J[j], cz[nb[j]], z); // padding: J = 0
    }
    __device__ __forceinline__ void finish(int i, Acc z) const
    {
        const float kPi = 3.14159265358979323846f;
        const float thi = th[i], ci = cz[i];
        const float f0 = (float)(prop[i] >> 8) * (1.0f / 16777216.0f); // [0, 1)
        float thp;
        if (!a.tf) {
            thp = kPi * f0; // svmc.pyx:95
        } else {            // svmc.pyx:198-207
            thp = thi + tfscale * (2.0f * kPi * f0 - kPi);
            thp = fminf(fmaxf(thp, 0.0f), kPi);
        }
        float sp, cp;
        __sincosf(thp, &sp, &cp);
        const float dE = bcoef * (cp - ci) * z + acoef * (__sinf(thi) - sp); // svmc.pyx:96-110
        const bool acc = rd_accept(dE, nl2e, pu[i] >> 16, [&]() -> uint32_t {
            return rd_refine_draw(c0, (uint32_t)i, c2, c3base | kTagRefine, a.keys.rk[0], a.keys.rk[1]);
        });
        if (acc) {
            th[i] = thp;
            cz[i] = cp;
        }
    }
};

// Shared memory: theta[Npad] | cos[Npad] | pu[Npad] | prop[Npad] | cnt[Npad] bytes | queue[Npad] u16
template <int DP>
__global__ void __launch_bounds__(512) refdyn_svmc_kernel(const __grid_constant__ RefdynArgs a)
{
    extern __shared__ __align__(16) unsigned char rd_smem[];
    __shared__ int s_cntr[3];
    const int N = a.N, Npad = a.Npad, T = blockDim.x, tid = threadIdx.x;
    float *th = reinterpret_cast<float *>(rd_smem), *cz = th + Npad;
    RdWaves wv;
    wv.pu = reinterpret_cast<uint32_t *>(cz + Npad);
    uint32_t *prop = wv.pu + Npad;
    wv.cnt32 = prop + Npad;
    wv.queue = reinterpret_cast<uint16_t *>(wv.cnt32 + Npad / 4);
    wv.cntr = s_cntr;
    const long long r = blockIdx.x;
    const uint32_t c0 = a.replica_offset + (uint32_t)r;

    for (int i = tid; i < Npad; i += T) {
        const float t = i < N ? a.theta[(size_t)i * a.Rpad + r] : 0.0f;
        th[i] = t;
        cz[i] = __cosf(t);
    }
    if (tid < 3) s_cntr[tid] = 0;
    __syncthreads();

    for (int f = 0; f < a.S; ++f) {
        const float *ellJ = a.ell_J + (size_t)f * a.ellJ_stride;
        const float *hrow = a.h + (size_t)f * a.h_stride;
        // schedule arrays: bcoef = B, jperp2 = min(1, A / B) (TF scale), acoef = A
        const float bcoef = __ldg(&a.bcoef[f]), tfscale = __ldg(&a.jperp2[f]), nl2e = __ldg(&a.nl2e[f]);
        const float acoef = __ldg(&a.acoef[f]);
        for (int step = 0; step < a.mcsteps; ++step) {
            const uint64_t sweep = a.sweep_offset + (uint64_t)f * (uint64_t)a.mcsteps + (uint64_t)step;
            const uint32_t c2 = (uint32_t)sweep, c3hi = (uint32_t)(sweep >> 32) << 16;
            RotorSite site{a, th, cz, wv.pu, prop, hrow, acoef, bcoef, tfscale, nl2e, c0, c2, c3hi};
            rd_pass<DP>(a, wv, ellJ, site);
        }
    }
    __syncthreads();
    rd_store_prof(a, wv);
    for (int i = tid; i < N; i += T) {
        a.theta[(size_t)i * a.Rpad + r] = th[i];
        a.cosz[(size_t)i * a.Rpad + r] = cz[i];
    }
}

template <typename K>
int rd_launch(K kernel, const RefdynArgs &a, long long replicas, int threads, size_t smem, cudaStream_t s)
{
    // opt-in dynamic shared memory: the attribute is per (kernel, device) and cheap to set, so set it whenever needed
    if (smem > 48 * 1024)
        MCS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<(unsigned)replicas, threads, smem, s>>>(a);
    return MCS_OK;
}

template <typename WT, bool SA>
int launch_ising(const RefdynArgs &a, long long replicas, int threads, size_t smem, cudaStream_t s)
{
    return a.dpad <= 4 ? rd_launch(refdyn_ising_kernel<WT, SA, 4>, a, replicas, threads, smem, s)
                       : rd_launch(refdyn_ising_kernel<WT, SA, 8>, a, replicas, threads, smem, s);
}

} // namespace

// kind = MCS_KIND_PIQMC: (A, B, temp) as mcs_piqmc_sweeps; MCS_KIND_SA: A = the temperature schedule, B = nullptr;
// MCS_KIND_SVMC: (A, B, temp) as mcs_svmc_sweeps, `variant` = tf; otherwise `variant` = global_moves
int mcs_launch_refdyn_sweeps(mcs_state *st, int kind, const double *A, const double *B, int64_t S, int mcsteps,
                             float temp, int variant, uint64_t seed, uint64_t replica_offset, uint64_t sweep_offset)
{
    mcs_instance *inst = st->inst;
    MCS_REQUIRE(!inst->dense, MCS_EUNSUPPORTED,
                "reference dynamics: dense instances have a total visiting order (no site parallelism); use the "
                "blocked tensor-core sweeps or the exact kernel");
    MCS_REQUIRE(inst->N <= 65536, MCS_EUNSUPPORTED, "reference dynamics: at most 65536 sites");
    MCS_REQUIRE(inst->maxdeg <= 255, MCS_EUNSUPPORTED, "reference dynamics: degree <= 255");
    MCS_REQUIRE(inst->nsteps == 1 || S <= inst->nsteps, MCS_EINVAL,
                "time-dependent instance has %lld tables but the schedule has %lld steps", (long long)inst->nsteps,
                (long long)S);
    if (S == 0 || mcsteps == 0) return MCS_OK;
    MCS_CUDA(cudaSetDevice(inst->device));
    const int P = kind == MCS_KIND_PIQMC ? (int)st->P : 1;
    const bool wide = P > 32;
    const int N = (int)inst->N, Npad = (N + 3) & ~3;
    // per site: state | pu | open-predecessor byte | queue entry (rotors: theta, cos, pu, proposal uniform)
    const size_t smem = (size_t)Npad * (kind == MCS_KIND_SVMC ? 16 + 1 + 2 : (wide ? 8 : 4) + 4 + 1 + 2);
    MCS_REQUIRE(smem <= 227 * 1024 - 64, MCS_EUNSUPPORTED,
                "reference dynamics: a replica's state (%zu bytes) does not fit one SM's shared memory", smem);
    std::vector<float> sched((size_t)4 * S);
    float *bc = sched.data(), *jp = bc + S, *nl = jp + S, *ac = nl + S;
    if (kind == MCS_KIND_PIQMC) {
        const double teff = (double)temp * (double)P; // qmc.pyx:85
        MCS_REQUIRE(teff != 0.0, MCS_EZERODIV, "float division");
        for (int64_t f = 0; f < S; ++f) {
            bc[f] = (float)(-2.0 * B[f]);                                    // qmc.pyx:96
            jp[f] = (float)(2.0 * (-0.5 * teff * log(tanh(A[f] / teff))));   // qmc.pyx:95
            nl[f] = (float)(-1.4426950408889634 / teff);
            ac[f] = 0.0f;
        }
    } else if (kind == MCS_KIND_SA) {
        for (int64_t t = 0; t < S; ++t) {
            bc[t] = -2.0f; // sa.pyx:84-94
            jp[t] = 0.0f;
            nl[t] = (float)(-1.4426950408889634 / A[t]); // exp(-ediff/temp), sa.pyx:98
            ac[t] = 0.0f;
        }
    } else {
        for (int64_t f = 0; f < S; ++f) {
            const double ab = A[f] / B[f]; // cdivision: inf / nan allowed (svmc.pyx:198)
            bc[f] = (float)B[f];
            jp[f] = (float)((ab > 1) ? 1.0 : ab);
            nl[f] = (float)(-1.4426950408889634 / (double)temp); // temp is a C float (svmc.pyx:24)
            ac[f] = (float)A[f];
        }
    }
    float *d_sched = nullptr;
    int *d_err = nullptr;
    MCS_CUDA(cudaMallocAsync((void **)&d_sched, sched.size() * sizeof(float) + sizeof(int), inst->stream));
    d_err = reinterpret_cast<int *>(d_sched + sched.size());
    MCS_CUDA(cudaMemcpyAsync(d_sched, sched.data(), sched.size() * sizeof(float), cudaMemcpyHostToDevice, inst->stream));
    MCS_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), inst->stream));
    RefdynArgs a;
    a.W = kind == MCS_KIND_PIQMC ? st->d_W + st->win_lo() : nullptr;
    a.V = kind == MCS_KIND_SA ? st->d_V : nullptr;
    a.theta = kind == MCS_KIND_SVMC ? st->d_theta : nullptr;
    a.cosz = kind == MCS_KIND_SVMC ? st->d_cosz : nullptr;
    a.ell_idx = inst->d_ell_idx;
    a.ell_J = inst->d_ell_J;
    a.h = inst->d_h;
    a.ellJ_stride = inst->nsteps > 1 ? (long long)inst->N * inst->dpad : 0;
    a.h_stride = inst->nsteps > 1 ? (long long)inst->N : 0;
    a.bcoef = d_sched;
    a.jperp2 = d_sched + S;
    a.nl2e = d_sched + 2 * S;
    a.acoef = d_sched + 3 * S;
    a.Rpad = st->Rpad;
    a.G = st->G;
    a.N = N;
    a.Npad = Npad;
    a.dpad = inst->dpad;
    a.field = inst->has_field ? 1 : 0;
    a.P = P;
    a.S = (int)S;
    a.mcsteps = mcsteps;
    a.global_moves = kind == MCS_KIND_PIQMC && variant ? 1 : 0;
    a.tf = kind == MCS_KIND_SVMC && variant ? 1 : 0;
    a.sequential = getenv("MCS_REFDYN_SEQUENTIAL") != nullptr ? 1 : 0;
    a.replica_offset = (uint32_t)(replica_offset + (kind == MCS_KIND_PIQMC ? (uint64_t)st->win_lo() : 0ull));
    a.sweep_offset = sweep_offset;
    a.keys = mcs_philox_expand(seed);
    a.err = d_err;
    a.prof = nullptr;
    const long long replicas = kind == MCS_KIND_PIQMC ? st->win_valid() : st->R;
    int threads = std::min(512, std::max(32, ((N / 6 + 31) / 32) * 32));
    if (const char *e = getenv("MCS_REFDYN_THREADS")) threads = std::max(32, std::min(512, atoi(e) / 32 * 32));
    long long *d_prof = nullptr;
    if (getenv("MCS_REFDYN_PROF")) {
        MCS_CUDA(cudaMalloc((void **)&d_prof, (size_t)replicas * 5 * sizeof(long long)));
        a.prof = d_prof;
    }
    int rc;
    if (kind == MCS_KIND_SVMC)
        rc = a.dpad <= 4 ? rd_launch(refdyn_svmc_kernel<4>, a, replicas, threads, smem, inst->stream)
                         : rd_launch(refdyn_svmc_kernel<8>, a, replicas, threads, smem, inst->stream);
    else if (kind == MCS_KIND_SA)
        rc = launch_ising<uint32_t, true>(a, replicas, threads, smem, inst->stream);
    else if (wide)
        rc = launch_ising<uint64_t, false>(a, replicas, threads, smem, inst->stream);
    else
        rc = launch_ising<uint32_t, false>(a, replicas, threads, smem, inst->stream);
    MCS_TRY(rc);
    inst->launches++;
    MCS_CUDA(cudaGetLastError());
    if (d_prof) {
        std::vector<long long> hp((size_t)replicas * 5);
        MCS_CUDA(cudaMemcpyAsync(hp.data(), d_prof, hp.size() * sizeof(long long), cudaMemcpyDeviceToHost, inst->stream));
        MCS_CUDA(cudaStreamSynchronize(inst->stream));
        cudaFree(d_prof);
        double s[5] = {0, 0, 0, 0, 0};
        for (long long r = 0; r < replicas; ++r)
            for (int q = 0; q < 5; ++q) s[q] += (double)hp[(size_t)r * 5 + q] / (double)replicas;
        fprintf(stderr, "[mcs refdyn] %lld CTAs x %d threads, smem %zu: per pass %.0f cycles priorities, %.0f counts, "
                        "%.0f rounds (%.1f rounds, %.0f cycles each); %.0f passes\n",
                replicas, threads, smem, s[0] / s[4], s[1] / s[4], s[2] / s[4], s[3] / s[4], s[2] / s[3], s[4]);
    }
    if (getenv("MCS_REFDYN_CHECK")) { // tests: surface the (impossible) stalled-pass flag
        int h_err = 0;
        MCS_CUDA(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, inst->stream));
        MCS_CUDA(cudaStreamSynchronize(inst->stream));
        MCS_CUDA(cudaFreeAsync(d_sched, inst->stream));
        MCS_REQUIRE(h_err == 0, MCS_EINVAL, "reference dynamics: a pass stalled (cyclic orientation?)");
        return MCS_OK;
    }
    MCS_CUDA(cudaFreeAsync(d_sched, inst->stream));
    return MCS_OK;
}
