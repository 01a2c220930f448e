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
template <int NPL, int WARPS, int FLD>
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
    const uint32_t g = ((uint32_t)blockIdx.x * WARPS + warp) * 32 + lane;

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
    const uint32_t G32 = (uint32_t)a.G;
    const bool live = g < G32;
    uint32_t *Vg = a.V + (live ? g : 0u);
    mcs_pdl_wait(); // everything above depends on the instance and the schedule only
    const uint32_t v = Vg[(uint64_t)(uint32_t)site * G32];
    uint32_t pl[NPL];
#pragma unroll
    for (int j = 0; j < NPL; ++j) pl[j] = j < NQ ? v ^ Vg[(uint64_t)(uint32_t)nb[j] * G32] : v;
    if (WARPS == 1)
        __syncwarp();
    else
        __syncthreads();
    if (!live) return;
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
    Vg[(uint64_t)(uint32_t)site * G32] = v ^ ~rej;
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
            if (rnd[i] <= mcs_accept_threshold(e[b], a.nl2e_over_t)) flip |= 1u << b;
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

template <int NPL, int FLD>
void launch_lut_wf(int warps, cudaStream_t s, const SaPass &a)
{
    // a.chunks = warps of work per site; `warps` divides it, so a CTA never straddles two sites
    const unsigned ny = (unsigned)std::min(a.nsites, 65535), nz = (unsigned)((a.nsites + 65534) / 65535);
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
    sa_energy_kernel<<<(unsigned)((st->R + 63) / 64), 64, 0, inst->stream>>>(
        st->d_V, inst->tab_idx_at(inst->nsteps - 1), inst->tab_J_at(inst->nsteps - 1), d_out, inst->N, (int)inst->maxnb, st->R, st->G);
    inst->launches++;
    MCS_CUDA(cudaGetLastError());
    return MCS_OK;
}
