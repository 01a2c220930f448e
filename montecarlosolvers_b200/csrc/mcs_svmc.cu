// mcs_svmc.cu -- spin-vector Monte Carlo (O(2) rotors, theta in [0, pi]) sweeps (sm_100a).
//
// Replaces the loop nest of svmc.SpinVectorMonteCarlo (reference svmc.pyx:78-117) and of
// svmc.SpinVectorMonteCarloTF (svmc.pyx:181-229); the batched forms (svmc.pyx:455-674) are the
// replica axis.
//
// Data layout in HBM: theta[site][replica], cosz[site][replica] fp32, replica fastest: a warp owns
// (site, 32 replicas), couplings are warp-uniform, every neighbour read is one coalesced 128-byte
// line of cos(theta_j).  Colour classes are launched one after another; neighbours are frozen.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "mcs_common.cuh"

namespace {

constexpr int kWarps = 4;

struct SvmcPass {
    float *theta;
    float *cosz;
    const int32_t *ell_idx;
    const float *ell_J;
    const float *h;
    const int32_t *sites;
    int nsites;
    int dpad;
    int field;
    int G; // Rpad / 32
    long long Rpad;
    float a_coef, b_coef;
    float nl2e_over_t;
    int tf;
    float tf_scale; // min(1, A/B) as the reference evaluates it (svmc.pyx:198-202)
    mcs_philox_keys keys;
    uint32_t sweep_lo, sweep_hi;
    uint32_t replica_offset;
    int always_refine; // MCS_ALWAYS_REFINE=1: evaluate the refinement call for every thread (timing / test hook)
};

// Rotor attempt (svmc.pyx:92-115) given the z-field of the site, from ONE 32-bit Philox word x:
//   * proposal: the high 20 bits, theta' = pi (x >> 12) 2^-20 (or the TF displacement);
//   * acceptance: exp(-dE/T) > u (svmc.pyx:112-115) with u = (v + r) / 4096, v = the low 12 bits of x and
//     r in [0, 1) 32 more bits of a SECOND Philox call that is evaluated only when it matters (lazily refined
//     uniform, as in the Ising kernels): with t = 4096 exp(-dE/T) the move is certainly accepted if v + 1 <= t,
//     certainly rejected if v >= t, and only for floor(t) == v (probability 2^-12 per attempt) is r needed.
//     dE <= 0 is accepted outright (t >= 4096 > v anyway, except 0 * inf at T = 0); NaN (inf - inf) is rejected.
// Returns true if the attempt is undecided; th / cz are updated for decided accepts, (thp, cp, t - v) are kept
// for the refinement.
__device__ __forceinline__ bool svmc_decide(const SvmcPass &a, float zfield, float &th, float &cz, uint32_t x,
                                            float &thp_out, float &cp_out, float &gap)
{
    const float kPi = 3.14159265358979323846f;
    const float f0 = (float)(x >> 12) * (1.0f / 1048576.0f); // [0, 1)
    float thp;
    if (!a.tf) {
        thp = kPi * f0; // svmc.pyx:95
    } else {            // svmc.pyx:198-207
        thp = th + a.tf_scale * (2.0f * kPi * f0 - kPi);
        thp = fminf(fmaxf(thp, 0.0f), kPi);
    }
    float sp, cp;
    __sincosf(thp, &sp, &cp);
    const float si = __sinf(th);
    const float dE = a.b_coef * (cp - cz) * zfield + a.a_coef * (si - sp); // svmc.pyx:96-110
    const float t = exp2f(fmaf(dE, a.nl2e_over_t, 12.0f));                 // 4096 exp(-dE/T)
    const float v = (float)(x & 0xFFFu);
    const bool yes = (dE <= 0.0f) | (v + 1.0f <= t), no = !yes & (v >= t); // dE <= 0 explicit: T = 0 gives 0 * inf
    if (yes) {
        th = thp;
        cz = cp;
    }
    thp_out = thp;
    cp_out = cp;
    gap = t - v; // undecided: accept iff r < gap
    return !(yes | no);
}

// A lane owns FOUR consecutive replicas of one site (128-bit loads of theta, cos theta and of every
// neighbour's cos theta; the coupling row is read once per four attempts); a warp = 128 replicas of one
// site; the kWarps warps of a CTA take kWarps consecutive sites of the colour class, which on Chimera-like
// graphs share most of their neighbours (their lines stay in L1).  Rpad is a multiple of 128.
// CACHE: neighbours read the cached cos(theta) (batches that stay in L2); otherwise they read theta and take the
// cosine themselves -- the pass is memory bound, and without the cache it moves 32 instead of 40 bytes per attempt
// (measured: +16 % at 16384 reads = 268 MB of state, -5 % at 2048 reads = 34 MB).
template <bool CACHE>
__global__ void __launch_bounds__(kWarps * 32) svmc_pass_kernel(const __grid_constant__ SvmcPass a)
{
    mcs_pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long sblocks = (a.nsites + kWarps - 1) / kWarps;
    const long long grp = blockIdx.x / sblocks; // group of 128 replicas
    const long long srank = (blockIdx.x % sblocks) * kWarps + warp;
    if (srank >= a.nsites) return;
    const int site = a.sites[srank];
    const long long r = grp * 128 + 4 * lane;
    float4 *th_ptr = reinterpret_cast<float4 *>(a.theta + (size_t)site * a.Rpad + r);
    float4 *cz_ptr = reinterpret_cast<float4 *>(a.cosz + (size_t)site * a.Rpad + r);
    mcs_pdl_wait(); // the state is first read here
    float4 th = *th_ptr, cz;
    if (CACHE)
        cz = *cz_ptr;
    else
        cz = make_float4(__cosf(th.x), __cosf(th.y), __cosf(th.z), __cosf(th.w));
    const float h0 = a.field ? __ldg(&a.h[site]) : 0.0f;
    float4 z = make_float4(h0, h0, h0, h0); // sum_j J_ij cos(theta_j) + h_i
    const float *jrow = a.ell_J + (size_t)site * a.dpad;
    const int *irow = a.ell_idx + (size_t)site * a.dpad;
    const float *czr = (CACHE ? a.cosz : a.theta) + r;
#pragma unroll 2
    for (int j = 0; j < a.dpad; ++j) { // padding entries have J = 0 and point at the site itself
        const float jv = __ldg(&jrow[j]);
        float4 c = *reinterpret_cast<const float4 *>(czr + (size_t)__ldg(&irow[j]) * a.Rpad);
        if (!CACHE) c = make_float4(__cosf(c.x), __cosf(c.y), __cosf(c.z), __cosf(c.w));
        z.x = fmaf(jv, c.x, z.x);
        z.y = fmaf(jv, c.y, z.y);
        z.z = fmaf(jv, c.z, z.z);
        z.w = fmaf(jv, c.w, z.w);
    }
    // one Philox call per four replicas; the counter is the GLOBAL quad index, so a shard starting at a multiple
    // of four reproduces the un-sharded run
    uint32_t x[4];
    const uint32_t c0 = (a.replica_offset >> 2) + (uint32_t)(r >> 2);
    const uint32_t c3 = (a.sweep_hi << 8) | MCS_TAG_SVMC;
    mcs_philox4x32_rk(c0, (uint32_t)site, a.sweep_lo, c3, a.keys, x);
    float tp[4], cp[4], gap[4];
    const bool u0 = svmc_decide(a, z.x, th.x, cz.x, x[0], tp[0], cp[0], gap[0]);
    const bool u1 = svmc_decide(a, z.y, th.y, cz.y, x[1], tp[1], cp[1], gap[1]);
    const bool u2 = svmc_decide(a, z.z, th.z, cz.z, x[2], tp[2], cp[2], gap[2]);
    const bool u3 = svmc_decide(a, z.w, th.w, cz.w, x[3], tp[3], cp[3], gap[3]);
    if ((u0 | u1 | u2 | u3) || a.always_refine) { // rare (2^-10 per thread): the low half of the uniforms
        uint32_t f[4];
        mcs_philox4x32_rk(c0, (uint32_t)site, a.sweep_lo, c3 | MCS_TAG_REFINE, a.keys, f);
        const float k32 = 1.0f / 4294967296.0f;
        if (u0 && (float)f[0] * k32 < gap[0]) { th.x = tp[0]; cz.x = cp[0]; }
        if (u1 && (float)f[1] * k32 < gap[1]) { th.y = tp[1]; cz.y = cp[1]; }
        if (u2 && (float)f[2] * k32 < gap[2]) { th.z = tp[2]; cz.z = cp[2]; }
        if (u3 && (float)f[3] * k32 < gap[3]) { th.w = tp[3]; cz.w = cp[3]; }
    }
    *th_ptr = th;
    if (CACHE) *cz_ptr = cz;
}

// host float64 [R][N] -> theta/cos [N][Rpad]
__global__ void svmc_pack_kernel(const double *__restrict__ in, float *__restrict__ theta,
                                 float *__restrict__ cosz, long long N, long long R, long long Rpad)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * R) return;
    const long long r = t / N, i = t % N;
    const double th = in[r * N + i];
    theta[i * Rpad + r] = (float)th;
    cosz[i * Rpad + r] = (float)cos(th);
}

__global__ void svmc_unpack_kernel(const float *__restrict__ theta, double *__restrict__ out, long long N,
                                   long long R, long long Rpad)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * R) return;
    const long long r = t / N, i = t % N;
    out[r * N + i] = (double)theta[i * Rpad + r];
}

__global__ void svmc_init_kernel(float *theta, float *cosz, long long n)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    theta[t] = 1.57079632679489661923f; // pi/2: all rotors along x, the transverse-field ground state
    cosz[t] = 0.0f;
}

// H = B (sum_bonds J cos cos + sum_i h cos) - A sum_i sin, fp64 from the stored angles
__global__ void svmc_energy_kernel(const float *__restrict__ theta, const int32_t *__restrict__ tab_idx,
                                   const double *__restrict__ tab_J, double *__restrict__ out, long long N,
                                   int maxnb, long long R, long long Rpad, double a, double b)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    double ez = 0.0, ex = 0.0;
    for (long long i = 0; i < N; ++i) {
        double pair = 0.0, field = 0.0;
        for (int s = 0; s < maxnb; ++s) {
            const int j = tab_idx[i * maxnb + s];
            const double jv = tab_J[i * maxnb + s];
            if (j == i)
                field += jv;
            else if (jv != 0.0)
                pair += jv * cos((double)theta[(long long)j * Rpad + r]);
        }
        const double th = (double)theta[i * Rpad + r];
        ez += cos(th) * (0.5 * pair + field);
        ex += sin(th);
    }
    out[r] = b * ez - a * ex;
}

} // namespace

int mcs_launch_svmc_sweeps(mcs_state *st, const double *A, const double *B, int64_t S, int mcsteps, float temp,
                           int tf, uint64_t seed, uint64_t replica_offset, uint64_t sweep_offset)
{
    mcs_instance *inst = st->inst;
    if (inst->dynamics == MCS_DYN_REFERENCE) // random-permutation sequential sweeps (svmc.pyx:83-115 in distribution)
        return mcs_launch_refdyn_sweeps(st, MCS_KIND_SVMC, A, B, S, mcsteps, temp, tf, seed, replica_offset,
                                        sweep_offset);
    MCS_REQUIRE((replica_offset & 3) == 0, MCS_EINVAL, "mcs_svmc_sweeps: replica_offset must be a multiple of 4");
    MCS_CUDA(cudaSetDevice(inst->device));
    SvmcPass a;
    a.theta = st->d_theta;
    a.cosz = st->d_cosz;
    a.ell_idx = inst->d_ell_idx;
    a.ell_J = inst->d_ell_J;
    a.h = inst->d_h;
    a.dpad = inst->dpad;
    a.field = inst->has_field ? 1 : 0;
    a.G = (int)(st->Rpad / 32);
    a.Rpad = st->Rpad;
    a.tf = tf ? 1 : 0;
    a.keys = mcs_philox_expand(seed);
    a.replica_offset = (uint32_t)replica_offset;
    a.always_refine = mcs_tie_threshold() == 0xFFFFFFFFu ? 1 : 0;
    // theta + cos(theta) of the batch: keep the cosine cache while both fit the L2 comfortably
    const bool cache = (size_t)inst->N * (size_t)st->Rpad * 8u <= (size_t)96 << 20;
    a.nl2e_over_t = (float)(-1.4426950408889634 / (double)temp); // temp is a C float (svmc.pyx:24)
    uint64_t sweep = sweep_offset;
    MCS_REQUIRE(inst->nsteps == 1 || S <= inst->nsteps, MCS_EINVAL,
                "time-dependent instance has %lld tables but the schedule has %lld steps", (long long)inst->nsteps,
                (long long)S);
    // Mid-size batches (BASELINE cfg4: 2048 reads): a colour pass is 5 us of work behind 4.4 us of kernel-boundary
    // latency.  Replicas are independent, so the batch is cut into two chunks whose passes alternate on two
    // streams: one chunk computes while the other sits in its launch gap (+16 %; with three or four chunks the host
    // cannot issue the launches fast enough: 2.3e11 / 1.7e11 attempts/s against 2.6e11 with two; MCS_STREAMS=n).
    const long long groups = st->Rpad / 128;
    long long max_sites = 0;
    for (int c = 0; c < inst->ncolors; ++c)
        max_sites = std::max(max_sites, (long long)(inst->color_start[c + 1] - inst->color_start[c]));
    int nchunk = 1;
    if (groups * max_sites <= 65536 && !getenv("MCS_ONE_STREAM")) nchunk = (int)std::max(1ll, std::min(2ll, groups / 2));
    if (const char *e = getenv("MCS_STREAMS")) nchunk = (int)std::max(1ll, std::min(std::min(4ll, groups), atoll(e)));
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
    float *const theta0 = a.theta, *const cosz0 = a.cosz;
    const uint32_t roff0 = a.replica_offset;
    for (int64_t f = 0; f < S; ++f) {
        a.ell_J = inst->ell_J_at(f); // svmc.NoisySVMC: nbs[ifield] (svmc.pyx:317-319)
        a.h = inst->h_at(f);
        a.a_coef = (float)A[f];
        a.b_coef = (float)B[f];
        const double ab = A[f] / B[f]; // cdivision: inf / nan allowed (svmc.pyx:198)
        a.tf_scale = (float)((ab > 1) ? 1.0 : ab);
        for (int step = 0; step < mcsteps; ++step, ++sweep) {
            a.sweep_lo = (uint32_t)sweep;
            a.sweep_hi = (uint32_t)(sweep >> 32);
            for (int c = 0; c < inst->ncolors; ++c) {
                a.sites = inst->d_order + inst->color_start[c];
                a.nsites = inst->color_start[c + 1] - inst->color_start[c];
                if (a.nsites == 0) continue;
                for (int q = 0; q < nchunk; ++q) {
                    const long long g0 = groups * q / nchunk, ng = groups * (q + 1) / nchunk - g0;
                    a.theta = theta0 + g0 * 128; // replica is the fastest axis: a chunk is a column offset
                    a.cosz = cosz0 + g0 * 128;
                    a.replica_offset = roff0 + (uint32_t)(g0 * 128);
                    cudaStream_t s = q ? inst->s_aux[q - 1] : inst->stream;
                    const long long ctas = (long long)((a.nsites + kWarps - 1) / kWarps) * ng;
                    if (cache)
                        mcs_launch_pdl(svmc_pass_kernel<true>, dim3((unsigned)ctas), dim3(kWarps * 32), s, a);
                    else
                        mcs_launch_pdl(svmc_pass_kernel<false>, dim3((unsigned)ctas), dim3(kWarps * 32), s, a);
                    inst->launches++;
                }
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

int mcs_svmc_pack(mcs_state *st, const double *d_in)
{
    mcs_instance *inst = st->inst;
    const long long n = inst->N * st->R;
    svmc_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, inst->stream>>>(d_in, st->d_theta, st->d_cosz, inst->N,
                                                                           st->R, st->Rpad);
    inst->launches++;
    MCS_CUDA(cudaGetLastError());
    return MCS_OK;
}

int mcs_svmc_unpack(mcs_state *st, double *d_out)
{
    mcs_instance *inst = st->inst;
    const long long n = inst->N * st->R;
    svmc_unpack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, inst->stream>>>(st->d_theta, d_out, inst->N, st->R,
                                                                             st->Rpad);
    inst->launches++;
    MCS_CUDA(cudaGetLastError());
    return MCS_OK;
}

int mcs_svmc_init(mcs_state *st)
{
    mcs_instance *inst = st->inst;
    const long long n = inst->N * st->Rpad;
    svmc_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, inst->stream>>>(st->d_theta, st->d_cosz, n);
    inst->launches++;
    MCS_CUDA(cudaGetLastError());
    return MCS_OK;
}

int mcs_svmc_energy(mcs_state *st, double a, double b, double *d_out)
{
    mcs_instance *inst = st->inst;
    svmc_energy_kernel<<<(unsigned)((st->R + 63) / 64), 64, 0, inst->stream>>>(
        st->d_theta, inst->tab_idx_at(inst->nsteps - 1), inst->tab_J_at(inst->nsteps - 1), d_out, inst->N, (int)inst->maxnb, st->R, st->Rpad, a, b);
    inst->launches++;
    MCS_CUDA(cudaGetLastError());
    return MCS_OK;
}
