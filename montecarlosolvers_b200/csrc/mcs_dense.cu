// mcs_dense.cu -- sweeps for DENSE coupling matrices (SK-like instances; BASELINE configs[4]) (sm_100a).
//
// Same reference loop nests as the sparse kernels (qmc.pyx:93-143 / 358-438, sa.pyx:66-101), different
// parallel decomposition: on a (near-)complete graph every site conflicts with every other, so a
// colour class is a single site and the sparse kernels degenerate into N tiny launches per sweep with
// O(N) work per attempt.  Here the sites are visited in order (a valid sequential sweep) in BLOCKS of 128:
//
//   A. local fields of the block for all columns (column = one (replica, slice) pair) as a GEMM
//        Hb[128, cols] = J[block rows, :] . S[:, cols]
//      on the tensor cores: S holds +-1 (exact in bf16), J is split J = hi + lo into two bf16 matrices
//      (16 mantissa bits), accumulation in fp32.  This is the only dense contraction of the solver and
//      it carries N^2 C of the sweep's N^2 C + O(128 N C) multiply-adds.
//   B. inside the block, site after site: Metropolis decision for every column from Hb (even slices,
//      then odd slices, then the world-line move), followed by a rank-1 correction
//        Hb[m', cols] += J[m', m] * (s_new - s_old)            for the later rows m' of the block,
//      everything held in shared memory -- the CTA owns its 64 columns (whole replicas) for the block.
//
// Spins live as S[column][site] bf16 (+1 / -1) between blocks; the bit-packed W / V arrays of the state
// are expanded on entry and re-packed on exit of a sweeps call.
#include <cuda.h>
#include <cuda_bf16.h>
#include <mma.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "mcs_common.cuh"

namespace {

constexpr int kBS = 128;      // sites per block
constexpr int kTC = 64;       // columns per CTA (a multiple of every supported P)
constexpr int kThreads = 256; // 8 warps
constexpr int kHld = kTC + 4; // leading dimension of the field tile (multiple of 4 floats for WMMA stores)
constexpr int kJld = kBS + 4; // rows 16-byte aligned (cp.async); the decision warps read J by broadcast

struct DensePass {
    const __nv_bfloat16 *Jhi, *Jlo; // [Npad][Npad] row i = couplings of site i (bf16 split of J)
    const float *Jf;                // [Npad][Npad] fp32
    const float *h;                 // [Npad]
    __nv_bfloat16 *S;               // [Cpad][Npad]
    int Npad, N, C, P, i0;
    int PS;           // columns per replica: P rounded up to a power of two (<= 32) or 64; columns k >= P of a replica
                      // are dead (spin +1, never attempted, no part in the ring or the world-line sum)
    int trotter;      // 1: PIQMC (columns of a replica form a ring of P slices), 0: SA
    float bcoef;      // -2 B (PIQMC, qmc.pyx:96) or -2 (SA, sa.pyx:91-94)
    float jperp2;     // 2 J_perp
    float nl2e_over_t;
    mcs_philox_keys keys;
    uint32_t sweep_lo, sweep_hi, replica_offset;
    int global_moves;
    float gscale, ginv; // world-line sums in fixed point: 2^K and 2^-K, K such that P |dE|_max 2^K < 2^30
    unsigned long long *trace; // MCS_DENSE_TRACE=1: phase timestamps of CTA 0 (ns), else nullptr
    // tcgen05 variant: block kernels of consecutive steps overlap.  ver[c] = number of steps column group c has
    // completed in this call (its spins written back); ver[groups] is an error flag (a wait that timed out)
    unsigned *ver;
    unsigned step;   // index of this launch within the call: sweep * blocks + block
    int nblocks;     // blocks per sweep
};

__device__ __forceinline__ void dense_stamp(const DensePass &a, int slot)
{
    if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        a.trace[slot] = t;
    }
}

__device__ __forceinline__ void bar_decide() { asm volatile("bar.sync 1, %0;" ::"n"(kTC)); }

// ---- phase B: the block's sites one after another -------------------------------------------------
// The sequential chain (128 decisions per block, each followed by a rank-1 correction of the later rows) is
// what bounds the dense sweep, so it is cut into STRIPS of 16 rows:
//   * the two decision warps (lane = column) take the strip's 16 field values of their column into registers,
//     decide the 16 sites one after another and apply each rank-1 correction to the REST OF THE STRIP
//     themselves (<= 15 register FMAs, J broadcast from shared memory): no CTA barrier inside a strip;
//   * after the strip one barrier, then all eight warps apply the strip's rank-16 correction
//     Hb[row, :] += sum_j J[row, r0 + j] * delta_j[:] to the later rows of the block (shared memory tile), one
//     barrier, next strip: 16 barriers per block instead of 256;
//   * every uniform the block will need is drawn up front, in parallel (the Philox counters do not depend on
//     the state) and stored as lu = log2((u+1) 2^-32): the Metropolis test  u/2^32 < exp(-dE/teff)
//     (qmc.pyx:140-143) becomes  dE <= 0  or  dE * (-log2 e / teff) >= lu  -- one multiply and a compare on
//     the critical path instead of an exponential;
//   * for P <= 32 a replica's slices are lanes of one warp: Trotter neighbours and the world-line sum go
//     through shuffles.  P == 64 (a replica spans both decision warps) keeps the older row-by-row scheme
//     (dense_phase_b_rows) with shared memory + a 64-thread barrier between the parity phases.
constexpr int kSB = 16;                                     // rows per strip
constexpr int kScrJd = 0;                                   // float [kBS][kJld]
constexpr int kScrUloc = kScrJd + kBS * kJld * 4;           // float [kBS][kTC]: log2 of the local uniforms
constexpr int kScrSb = kScrUloc + kBS * kTC * 4;            // int8 [kBS][kTC]
constexpr int kScrFrow = kScrSb + kBS * kTC;                // float [kTC]
constexpr int kScrDelta = kScrFrow + kTC * 4;               // float [kTC]
constexpr int kScrGterm = kScrDelta + kTC * 4;              // float [kTC]
constexpr int kScrHblk = kScrGterm + kTC * 4;               // float [kBS]: local fields h_i of the block
constexpr int kScrDstrip = kScrHblk + kBS * 4;              // float [kSB][kTC]: spin changes of the strip
constexpr int kScrBytes = kScrDstrip + kSB * kTC * 4;
constexpr int kUglobBytes = kBS * (kTC / 2) * 4;            // float [kBS][kTC/2]: log2 of the world-line uniforms

// PT: compile-time P for the in-warp cases (1, 2, 4, 8, 16, 32: a replica's slices sit in one warp);
// PT == 0: P == 64, the replica spans both decision warps (shared memory + a 64-thread barrier).
template <int PT>
__device__ __forceinline__ void dense_phase_b_rows(const DensePass &a, float *Hb, unsigned char *scr, float *uglob,
                                                   int col0)
{
    constexpr bool INWARP = PT != 0;
    float *Jd = reinterpret_cast<float *>(scr + kScrJd);
    float *uloc = reinterpret_cast<float *>(scr + kScrUloc);
    float *hblk = reinterpret_cast<float *>(scr + kScrHblk);
    signed char *sb = reinterpret_cast<signed char *>(scr + kScrSb);
    float *frow = reinterpret_cast<float *>(scr + kScrFrow);
    float *delta = reinterpret_cast<float *>(scr + kScrDelta);
    float *gterm = reinterpret_cast<float *>(scr + kScrGterm);
    const int tid = threadIdx.x;
    const int i0 = a.i0;
    const long long ld = a.Npad;
    const int P = INWARP ? PT : a.P;
    const int PS = INWARP ? PT : a.PS; // columns per replica (>= P)
    const int mend = min(kBS, a.N - i0);

    // field strip of this thread -> registers; the tile's shared memory then holds the world-line uniforms
    const int row = tid >> 1, cbase = (tid & 1) * (kTC / 2);
    float hreg[kTC / 2];
#pragma unroll
    for (int q = 0; q < kTC / 2; ++q) hreg[q] = Hb[row * kHld + cbase + q];
    __syncthreads();
    if (tid < kBS) hblk[tid] = __ldg(&a.h[i0 + tid]);

#pragma unroll 4
    for (int e = tid; e < kBS * kBS; e += kThreads) {
        const int r = e / kBS, c = e % kBS;
        Jd[r * kJld + c] = __ldg(&a.Jf[(long long)(i0 + r) * ld + i0 + c]);
    }
    for (int e = tid; e < kBS * kTC; e += kThreads) {
        const int c = e / kBS, m = e % kBS; // consecutive threads -> consecutive sites of one column
        sb[m * kTC + c] = __bfloat162float(a.S[(long long)(col0 + c) * ld + i0 + m]) < 0.0f ? -1 : 1;
    }
#pragma unroll 1
    for (int e = tid; e < kBS * kTC; e += kThreads) { // all uniforms of the block
        const int m = e / kTC, c = e % kTC;
        const int col = col0 + c, k = col % PS;
        const uint32_t rep = a.replica_offset + (uint32_t)(col / PS);
        if ((k & 3) == 0) {
            uint32_t rnd[4];
            mcs_philox4x32_rk(rep, (uint32_t)(i0 + m), a.sweep_lo, (a.sweep_hi << 8) | (uint32_t)(k >> 2), a.keys, rnd);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (k + j < P) uloc[m * kTC + c + j] = __log2f((float)rnd[j] + 1.0f) - 32.0f;
        }
        if (a.global_moves && k == 0) {
            uint32_t rnd[4];
            mcs_philox4x32_rk(rep, (uint32_t)(i0 + m), a.sweep_lo, (a.sweep_hi << 8) | MCS_TAG_GLOBAL, a.keys, rnd);
            uglob[m * (kTC / 2) + c / PS] = __log2f((float)rnd[0] + 1.0f) - 32.0f;
        }
    }
    if (row == 0) {
#pragma unroll
        for (int q = 0; q < kTC / 2; ++q) frow[cbase + q] = hreg[q];
    }
    __syncthreads();

    const int c = tid; // decision threads: tid < kTC
    const int col = col0 + c;
    const int k = col % PS; // slice (dead column if k >= P)
    const int cl = c - k + (k == 0 ? P - 1 : k - 1), cr = k < P ? c - k + (k == P - 1 ? 0 : k + 1) : c;
    const int crep = c / PS;
    const bool valid = col < a.C && k < P;
    const bool odd = (k & 1) != 0;
    const bool lastodd = (P & 1) != 0 && P > 1 && k == P - 1; // odd ring: slice P-1 neighbours slice 0, visited alone
    const float sc = a.nl2e_over_t, bco = a.bcoef, jp2 = a.jperp2;
    const bool trotter = a.trotter != 0, glob = a.global_moves != 0;
#define MCS_ACCEPT(dE, lg) (valid && ((dE) <= 0.0f || (dE) * sc >= (lg)))
    for (int m = 0; m < mend; ++m) {
        if (tid < kTC) {
            const float bf = bco * (frow[c] + hblk[m]); // -2B * local field
            const float lu = uloc[m * kTC + c];
            const float s_init = (float)sb[m * kTC + c];
            float s = s_init;
            if (INWARP) {
                if (trotter) { // even slices, then odd slices (P is even); neighbours by shuffle
                    float nb = __shfl_sync(0xffffffffu, s, cl & 31) + __shfl_sync(0xffffffffu, s, cr & 31);
                    float dE = s * fmaf(jp2, nb, bf);
                    if (!odd && MCS_ACCEPT(dE, lu)) s = -s;
                    nb = __shfl_sync(0xffffffffu, s, cl & 31) + __shfl_sync(0xffffffffu, s, cr & 31);
                    dE = s * fmaf(jp2, nb, bf);
                    if (odd && MCS_ACCEPT(dE, lu)) s = -s;
                } else {
                    const float dE = s * bf;
                    if (MCS_ACCEPT(dE, lu)) s = -s;
                }
                if (glob) { // world-line move: all P slices of the replica (qmc.pyx:405-438)
                    float dE = s * bf;
#pragma unroll
                    for (int off = 1; off < PT; off <<= 1) dE += __shfl_xor_sync(0xffffffffu, dE, off);
                    if (MCS_ACCEPT(dE, uglob[m * (kTC / 2) + crep])) s = -s;
                }
            } else { // P == 64: the replica spans both decision warps -> shared memory + a 64-thread barrier
#pragma unroll 1
                for (int parity = 0; parity < ((P & 1) && P > 1 ? 3 : 2); ++parity) {
                    if (parity == 2 ? lastodd : ((k & 1) == parity && !lastodd)) {
                        const float dE = s * fmaf(jp2, (float)(sb[m * kTC + cl] + sb[m * kTC + cr]), bf);
                        if (MCS_ACCEPT(dE, lu)) {
                            s = -s;
                            sb[m * kTC + c] = (signed char)s;
                        }
                    }
                    bar_decide();
                }
                if (glob) {
                    gterm[c] = valid ? s * bf : 0.0f;
                    bar_decide();
                    float dE = 0.0f;
                    for (int q = 0; q < P; ++q) dE += gterm[c - k + q];
                    if (MCS_ACCEPT(dE, uglob[m * (kTC / 2) + crep])) s = -s;
                }
            }
            sb[m * kTC + c] = (signed char)s;
            delta[c] = s - s_init;
        }
        __syncthreads();
        if (row > m) {
            const float jv = Jd[row * kJld + m];
#pragma unroll
            for (int q4 = 0; q4 < kTC / 8; ++q4) {
                const float4 d = *reinterpret_cast<const float4 *>(delta + cbase + 4 * q4);
                hreg[4 * q4 + 0] = fmaf(jv, d.x, hreg[4 * q4 + 0]);
                hreg[4 * q4 + 1] = fmaf(jv, d.y, hreg[4 * q4 + 1]);
                hreg[4 * q4 + 2] = fmaf(jv, d.z, hreg[4 * q4 + 2]);
                hreg[4 * q4 + 3] = fmaf(jv, d.w, hreg[4 * q4 + 3]);
            }
            if (row == m + 1) {
#pragma unroll
                for (int q4 = 0; q4 < kTC / 8; ++q4)
                    *reinterpret_cast<float4 *>(frow + cbase + 4 * q4) =
                        make_float4(hreg[4 * q4], hreg[4 * q4 + 1], hreg[4 * q4 + 2], hreg[4 * q4 + 3]);
            }
        }
        __syncthreads();
    }
#undef MCS_ACCEPT
    // write the block's spins back
    for (int e = tid; e < kBS * kTC; e += kThreads) {
        const int cc = e / kBS, m = e % kBS;
        a.S[(long long)(col0 + cc) * ld + i0 + m] = __float2bfloat16((float)sb[m * kTC + cc]);
    }
}

// Strip scheme (see above), P in {1, 2, 4, 8, 16, 32}.
template <int PT>
__device__ __forceinline__ void dense_phase_b_strips(const DensePass &a, float *Hb, unsigned char *scr, float *uglob,
                                                     int col0)
{
    float *Jd = reinterpret_cast<float *>(scr + kScrJd);
    float *uloc = reinterpret_cast<float *>(scr + kScrUloc);
    float *hblk = reinterpret_cast<float *>(scr + kScrHblk);
    signed char *sb = reinterpret_cast<signed char *>(scr + kScrSb);
    float *dstrip = reinterpret_cast<float *>(scr + kScrDstrip);
    const int tid = threadIdx.x;
    const int i0 = a.i0;
    const long long ld = a.Npad;
    const int P = a.P; // slices that exist; PT = columns per replica (a power of two >= P)

    // Prologue: everything phase B needs besides the fields, fetched with as many requests in flight as possible
    // (it is pure latency: 64 KB of J, 16 KB of spins, 2304 Philox calls per CTA).
    // J block: 128 rows x 512 B as 16-byte cp.async chunks, all issued before anything is waited for
    for (int e = tid; e < kBS * (kBS / 4); e += kThreads) {
        const int r = e / (kBS / 4), c4 = (e % (kBS / 4)) * 4;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(Jd + r * kJld + c4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(a.Jf + (long long)(i0 + r) * ld + i0 + c4) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (tid < kBS) hblk[tid] = __ldg(&a.h[i0 + tid]);
    // spins: item = (column, 8 consecutive sites) = one 16-byte load of bf16, stored transposed as int8
    {
        uint4 v[kBS * kTC / 8 / kThreads];
#pragma unroll
        for (int q = 0; q < kBS * kTC / 8 / kThreads; ++q) {
            const int e = tid + q * kThreads, cc = e / (kBS / 8), m8 = (e % (kBS / 8)) * 8;
            v[q] = __ldcg(reinterpret_cast<const uint4 *>(a.S + (long long)(col0 + cc) * ld + i0 + m8));
        }
#pragma unroll
        for (int q = 0; q < kBS * kTC / 8 / kThreads; ++q) {
            const int e = tid + q * kThreads, cc = e / (kBS / 8), m8 = (e % (kBS / 8)) * 8;
            const uint32_t w[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
            for (int t = 0; t < 8; ++t) // bf16 sign bit: bit 15 of each half word
                sb[(m8 + t) * kTC + cc] = ((w[t >> 1] >> (16 * (t & 1) + 15)) & 1u) ? -1 : 1;
        }
    }
    // uniforms: one Philox call = four consecutive slices of one (row, replica); calls spread evenly over the threads
    constexpr int G4 = (PT + 3) / 4;              // calls per (row, replica)
    constexpr int NREP = kTC / PT;                // replicas of this CTA
#pragma unroll 2
    for (int e = tid; e < kBS * NREP * G4; e += kThreads) {
        const int m = e / (NREP * G4), rr = (e / G4) % NREP, g = e % G4;
        const uint32_t rep = a.replica_offset + (uint32_t)((col0 + rr * PT) / PT);
        uint32_t rnd[4];
        mcs_philox4x32_rk(rep, (uint32_t)(i0 + m), a.sweep_lo, (a.sweep_hi << 8) | (uint32_t)g, a.keys, rnd);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (4 * g + j < PT) uloc[m * kTC + rr * PT + 4 * g + j] = __log2f((float)rnd[j] + 1.0f) - 32.0f;
    }
    if (a.global_moves) {
        for (int e = tid; e < kBS * NREP; e += kThreads) {
            const int m = e / NREP, rr = e % NREP;
            const uint32_t rep = a.replica_offset + (uint32_t)((col0 + rr * PT) / PT);
            uint32_t rnd[4];
            mcs_philox4x32_rk(rep, (uint32_t)(i0 + m), a.sweep_lo, (a.sweep_hi << 8) | MCS_TAG_GLOBAL, a.keys, rnd);
            uglob[m * (kTC / 2) + rr] = __log2f((float)rnd[0] + 1.0f) - 32.0f;
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    dense_stamp(a, 2);

    const int c = tid & (kTC - 1); // column of a decision thread (tid < kTC)
    const int col = col0 + c;
    const int k = col % PT; // slice (dead column if k >= P)
    const int cl = c - k + (k == 0 ? P - 1 : k - 1), cr = k < P ? c - k + (k == P - 1 ? 0 : k + 1) : c;
    const int crep = c / PT;
    const bool valid = col < a.C && k < P;
    const bool oddP = (P & 1) != 0 && P > 1; // odd ring: slice P-1 neighbours slice 0 and is visited alone, last
    const bool lastodd = oddP && k == P - 1;
    const bool odd = (k & 1) != 0 && !lastodd, even = (k & 1) == 0 && !lastodd;
    const float sc = a.nl2e_over_t, bco = a.bcoef, jp2 = a.jperp2;
    const bool trotter = a.trotter != 0 && P > 1, glob = a.global_moves != 0;
    const unsigned segmask = PT >= 32 ? 0xffffffffu : (((1u << (PT & 31)) - 1u) << ((tid & 31) & ~(PT - 1)));
// branch-free (bitwise, no short circuit): the decision chain is latency bound
#define MCS_ACCEPT(dE, lg) (valid & (((dE) <= 0.0f) | ((dE) * sc >= (lg))))
    // Rows beyond N (last block of an N that is not a multiple of 128) are decided like the others: their couplings
    // and fields are zero padding and their spins are never read back, so the strip code has no per-row branch --
    // one basic block per strip, every operand of the 16 steps prefetched into registers up front.
#pragma unroll 1
    for (int r0 = 0; r0 < kBS; r0 += kSB) {
        if (tid < kTC) {
            float f[kSB], lu[kSB], hb[kSB], s0[kSB], ug[kSB];
#pragma unroll
            for (int j = 0; j < kSB; ++j) {
                f[j] = Hb[(r0 + j) * kHld + c];
                lu[j] = uloc[(r0 + j) * kTC + c];
                hb[j] = hblk[r0 + j];
                s0[j] = (float)sb[(r0 + j) * kTC + c];
                ug[j] = glob ? uglob[(r0 + j) * (kTC / 2) + crep] : 0.0f;
            }
#pragma unroll
            for (int j = 0; j < kSB; ++j) {
                const int m = r0 + j;
                const float bf = bco * (f[j] + hb[j]); // -2B * local field
                float s = s0[j];
                if (trotter) { // even slices, then odd slices (P is even); neighbours by shuffle
                    float nb = __shfl_sync(0xffffffffu, s, cl & 31) + __shfl_sync(0xffffffffu, s, cr & 31);
                    float dE = s * fmaf(jp2, nb, bf);
                    s = (even & MCS_ACCEPT(dE, lu[j])) ? -s : s;
                    nb = __shfl_sync(0xffffffffu, s, cl & 31) + __shfl_sync(0xffffffffu, s, cr & 31);
                    dE = s * fmaf(jp2, nb, bf);
                    s = (odd & MCS_ACCEPT(dE, lu[j])) ? -s : s;
                    if (oddP) { // warp-uniform
                        nb = __shfl_sync(0xffffffffu, s, cl & 31) + __shfl_sync(0xffffffffu, s, cr & 31);
                        dE = s * fmaf(jp2, nb, bf);
                        s = (lastodd & MCS_ACCEPT(dE, lu[j])) ? -s : s;
                    }
                } else {
                    const float dE = s * bf;
                    s = MCS_ACCEPT(dE, lu[j]) ? -s : s;
                }
                if (glob) { // world-line move: all P slices of the replica (qmc.pyx:405-438)
                    // sum over the replica's lanes in fixed point with one redux.sync instead of log2(P)
                    // dependent shuffles (the scale is chosen per launch from the instance's largest possible
                    // field: resolution P |dE|_max 2^-30, finer than the fp32 butterfly it replaces)
                    const int part = valid ? __float2int_rn(s * bf * a.gscale) : 0;
                    const float dE = (float)__reduce_add_sync(segmask, part) * a.ginv;
                    s = MCS_ACCEPT(dE, ug[j]) ? -s : s;
                }
                sb[m * kTC + c] = (signed char)s;
                const float d = s - s0[j];
                // rank-1 correction of the rest of the strip (this column only, registers)
#pragma unroll
                for (int j2 = j + 1; j2 < kSB; ++j2) f[j2] = fmaf(Jd[(r0 + j2) * kJld + m], d, f[j2]);
                dstrip[j * kTC + c] = d;
            }
        }
        __syncthreads();
        if (r0 == 0) dense_stamp(a, 3);
        // rank-16 correction of the later rows of the block: item = (row, 4 columns)
        const int nrows = kBS - (r0 + kSB);
#pragma unroll 1
        for (int it = tid; it < nrows * (kTC / 4); it += kThreads) {
            const int row = r0 + kSB + it / (kTC / 4), c4 = (it % (kTC / 4)) * 4;
            float4 h = *reinterpret_cast<const float4 *>(Hb + row * kHld + c4);
            const float *jr = Jd + row * kJld + r0;
#pragma unroll
            for (int j = 0; j < kSB; ++j) {
                const float jv = jr[j];
                const float4 dd = *reinterpret_cast<const float4 *>(dstrip + j * kTC + c4);
                h.x = fmaf(jv, dd.x, h.x);
                h.y = fmaf(jv, dd.y, h.y);
                h.z = fmaf(jv, dd.z, h.z);
                h.w = fmaf(jv, dd.w, h.w);
            }
            *reinterpret_cast<float4 *>(Hb + row * kHld + c4) = h;
        }
        __syncthreads();
        if (r0 == 0) dense_stamp(a, 4);
    }
    dense_stamp(a, 5);
#undef MCS_ACCEPT
    // write the block's spins back
#pragma unroll 4
    for (int e = tid; e < kBS * kTC; e += kThreads) {
        const int cc = e / kBS, m = e % kBS;
        a.S[(long long)(col0 + cc) * ld + i0 + m] = __float2bfloat16((float)sb[m * kTC + cc]);
    }
}

// ---- phase B, pipelined (tcgen05 variant, P <= 32) ---------------------------------------------------------------
// Same decisions as dense_phase_b_strips, with everything that is not the decision chain moved off it.  Scratch
// that does NOT alias the TMA ring, so the state-independent part of the prologue runs while the GEMM streams:
//   Js   [3][128][16]  the strip's J columns: J[block rows r0.., strip columns]  (cp.async, one strip ahead; three
//                      buffers: the strip being decided, the previous one still being applied, the next one arriving)
//   ul   [2][16][64]   log2 of the strip's local uniforms, ug [2][16][32] of its world-line uniforms
//   ds   [2][16][64]   spin changes of a strip
//   sb   [128][64]     spins of the block, hb [128] fields
// During the decisions of strip r0 (two warps) the other six warps (a) apply the PREVIOUS strip's rank-16 correction
// to the rows beyond the current strip's successor, (b) write the previous strip's final spins back, (c) fetch the
// next strip's J columns and draw its uniforms.  After the decisions only the next strip's 16 rows get this strip's
// correction before the chain continues.
constexpr int kS2Js = 0;                                  // float [3][kBS][kSB]
constexpr int kS2Ul = kS2Js + 3 * kBS * kSB * 4;          // float [2][kSB][kTC]
constexpr int kS2Ug = kS2Ul + 2 * kSB * kTC * 4;          // float [2][kSB][kTC/2]
constexpr int kS2Ds = kS2Ug + 2 * kSB * (kTC / 2) * 4;    // float [2][kSB][kTC]
constexpr int kS2Sb = kS2Ds + 2 * kSB * kTC * 4;          // int8  [kBS][kTC]
constexpr int kS2Hb = kS2Sb + kBS * kTC;                  // float [kBS]
constexpr int kS2Bytes = kS2Hb + kBS * 4;

struct Strips2 {
    float *Js, *ul, *ug, *ds, *hb;
    signed char *sb;
    __device__ explicit Strips2(unsigned char *scr)
        : Js(reinterpret_cast<float *>(scr + kS2Js)), ul(reinterpret_cast<float *>(scr + kS2Ul)),
          ug(reinterpret_cast<float *>(scr + kS2Ug)), ds(reinterpret_cast<float *>(scr + kS2Ds)),
          hb(reinterpret_cast<float *>(scr + kS2Hb)), sb(reinterpret_cast<signed char *>(scr + kS2Sb))
    {
    }
};

// J columns of strip r0 (rows r0 .. 127 of the block, 16 columns): t = this thread's index among nt helpers
__device__ __forceinline__ void strips2_fetch_J(const DensePass &a, const Strips2 &z, int r0, int t, int nt)
{
    float *dst = z.Js + ((r0 / kSB) % 3) * kBS * kSB;
    const long long ld = a.Npad;
    for (int e = t; e < (kBS - r0) * (kSB / 4); e += nt) {
        const int lr = e / (kSB / 4), c4 = (e % (kSB / 4)) * 4;
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + lr * kSB + c4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d),
                     "l"(a.Jf + (long long)(a.i0 + r0 + lr) * ld + a.i0 + r0 + c4)
                     : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// uniforms of strip r0 (same Philox counters as the unpipelined scheme)
template <int PT>
__device__ __forceinline__ void strips2_draw(const DensePass &a, const Strips2 &z, int r0, int col0, int t, int nt)
{
    constexpr int G4 = (PT + 3) / 4, NREP = kTC / PT;
    float *ul = z.ul + ((r0 / kSB) & 1) * kSB * kTC, *ug = z.ug + ((r0 / kSB) & 1) * kSB * (kTC / 2);
    for (int e = t; e < kSB * NREP * G4; e += nt) {
        const int j = e / (NREP * G4), rr = (e / G4) % NREP, g = e % G4;
        const uint32_t rep = a.replica_offset + (uint32_t)((col0 + rr * PT) / PT);
        uint32_t rnd[4];
        mcs_philox4x32_rk(rep, (uint32_t)(a.i0 + r0 + j), a.sweep_lo, (a.sweep_hi << 8) | (uint32_t)g, a.keys, rnd);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (4 * g + q < PT) ul[j * kTC + rr * PT + 4 * g + q] = __log2f((float)rnd[q] + 1.0f) - 32.0f;
    }
    if (a.global_moves) {
        for (int e = t; e < kSB * NREP; e += nt) {
            const int j = e / NREP, rr = e % NREP;
            const uint32_t rep = a.replica_offset + (uint32_t)((col0 + rr * PT) / PT);
            uint32_t rnd[4];
            mcs_philox4x32_rk(rep, (uint32_t)(a.i0 + r0 + j), a.sweep_lo, (a.sweep_hi << 8) | MCS_TAG_GLOBAL, a.keys, rnd);
            ug[j * (kTC / 2) + rr] = __log2f((float)rnd[0] + 1.0f) - 32.0f;
        }
    }
}

// Hb[rows lo .. hi) += J[row, strip rs columns] . deltas of strip rs
__device__ __forceinline__ void strips2_update(const Strips2 &z, float *Hb, int rs, int lo, int hi, int t, int nt)
{
    const float *Js = z.Js + ((rs / kSB) % 3) * kBS * kSB, *ds = z.ds + ((rs / kSB) & 1) * kSB * kTC;
    for (int it = t; it < (hi - lo) * (kTC / 4); it += nt) {
        const int row = lo + it / (kTC / 4), c4 = (it % (kTC / 4)) * 4;
        float4 h = *reinterpret_cast<const float4 *>(Hb + row * kHld + c4);
        const float *jr = Js + (row - rs) * kSB;
#pragma unroll
        for (int j = 0; j < kSB; ++j) {
            const float jv = jr[j];
            const float4 dd = *reinterpret_cast<const float4 *>(ds + j * kTC + c4);
            h.x = fmaf(jv, dd.x, h.x);
            h.y = fmaf(jv, dd.y, h.y);
            h.z = fmaf(jv, dd.z, h.z);
            h.w = fmaf(jv, dd.w, h.w);
        }
        *reinterpret_cast<float4 *>(Hb + row * kHld + c4) = h;
    }
}

// final spins of strip rs -> S
__device__ __forceinline__ void strips2_writeback(const DensePass &a, const Strips2 &z, int rs, int col0, int t, int nt)
{
    const long long ld = a.Npad;
    for (int e = t; e < kSB * kTC; e += nt) {
        const int cc = e / kSB, j = e % kSB;
        a.S[(long long)(col0 + cc) * ld + a.i0 + rs + j] = __float2bfloat16((float)z.sb[(rs + j) * kTC + cc]);
    }
}

// everything that does not depend on the GEMM (run by warps 4-7 while it streams): fields, own-block spins,
// J columns and uniforms of strip 0
template <int PT>
__device__ __forceinline__ void strips2_early(const DensePass &a, unsigned char *scr, int col0, int t, int nt)
{
    const Strips2 z(scr);
    const long long ld = a.Npad;
    strips2_fetch_J(a, z, 0, t, nt);
    for (int e = t; e < kBS; e += nt) z.hb[e] = __ldg(&a.h[a.i0 + e]);
    // spins: item = (column, 8 consecutive sites) = one 16-byte load of bf16, stored transposed as int8.  The own
    // block's spins were written by this column group's step `step - nblocks` (the caller has waited for it).
    for (int e = t; e < kBS * kTC / 8; e += nt) {
        const int cc = e / (kBS / 8), m8 = (e % (kBS / 8)) * 8;
        const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(a.S + (long long)(col0 + cc) * ld + a.i0 + m8));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 8; ++q) z.sb[(m8 + q) * kTC + cc] = ((w[q >> 1] >> (16 * (q & 1) + 15)) & 1u) ? -1 : 1;
    }
    strips2_draw<PT>(a, z, 0, col0, t, nt);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <int PT>
__device__ __forceinline__ void dense_phase_b_strips2(const DensePass &a, float *Hb, unsigned char *scr, int col0)
{
    const Strips2 z(scr);
    const int tid = threadIdx.x;
    const int P = a.P;
    const int c = tid & (kTC - 1);
    const int col = col0 + c;
    const int k = col % PT;
    const int cl = c - k + (k == 0 ? P - 1 : k - 1), cr = k < P ? c - k + (k == P - 1 ? 0 : k + 1) : c;
    const int crep = c / PT;
    const bool valid = col < a.C && k < P;
    const bool oddP = (P & 1) != 0 && P > 1;
    const bool lastodd = oddP && k == P - 1;
    const bool odd = (k & 1) != 0 && !lastodd, even = (k & 1) == 0 && !lastodd;
    const float sc = a.nl2e_over_t, bco = a.bcoef, jp2 = a.jperp2;
    const bool trotter = a.trotter != 0 && P > 1, glob = a.global_moves != 0;
    const unsigned segmask = PT >= 32 ? 0xffffffffu : (((1u << (PT & 31)) - 1u) << ((tid & 31) & ~(PT - 1)));
    constexpr int NH = kThreads - kTC; // helper threads
    dense_stamp(a, 2);
#define MCS_ACCEPT(dE, lg) (valid & (((dE) <= 0.0f) | ((dE) * sc >= (lg))))
#pragma unroll 1
    for (int r0 = 0; r0 < kBS; r0 += kSB) {
        const int b = (r0 / kSB) & 1;
        if (tid < kTC) {
            const float *Js = z.Js + ((r0 / kSB) % 3) * kBS * kSB, *ul = z.ul + b * kSB * kTC, *ug = z.ug + b * kSB * (kTC / 2);
            float *ds = z.ds + b * kSB * kTC;
            float f[kSB], lu[kSB], hb[kSB], s0[kSB], gu[kSB];
#pragma unroll
            for (int j = 0; j < kSB; ++j) {
                f[j] = Hb[(r0 + j) * kHld + c];
                lu[j] = ul[j * kTC + c];
                hb[j] = z.hb[r0 + j];
                s0[j] = (float)z.sb[(r0 + j) * kTC + c];
                gu[j] = glob ? ug[j * (kTC / 2) + crep] : 0.0f;
            }
#pragma unroll
            for (int j = 0; j < kSB; ++j) {
                const float bf = bco * (f[j] + hb[j]); // -2B * local field
                float s = s0[j];
                if (trotter) {
                    float nb = __shfl_sync(0xffffffffu, s, cl & 31) + __shfl_sync(0xffffffffu, s, cr & 31);
                    float dE = s * fmaf(jp2, nb, bf);
                    s = (even & MCS_ACCEPT(dE, lu[j])) ? -s : s;
                    nb = __shfl_sync(0xffffffffu, s, cl & 31) + __shfl_sync(0xffffffffu, s, cr & 31);
                    dE = s * fmaf(jp2, nb, bf);
                    s = (odd & MCS_ACCEPT(dE, lu[j])) ? -s : s;
                    if (oddP) { // warp-uniform
                        nb = __shfl_sync(0xffffffffu, s, cl & 31) + __shfl_sync(0xffffffffu, s, cr & 31);
                        dE = s * fmaf(jp2, nb, bf);
                        s = (lastodd & MCS_ACCEPT(dE, lu[j])) ? -s : s;
                    }
                } else {
                    const float dE = s * bf;
                    s = MCS_ACCEPT(dE, lu[j]) ? -s : s;
                }
                if (glob) {
                    const int part = valid ? __float2int_rn(s * bf * a.gscale) : 0;
                    const float dE = (float)__reduce_add_sync(segmask, part) * a.ginv;
                    s = MCS_ACCEPT(dE, gu[j]) ? -s : s;
                }
                z.sb[(r0 + j) * kTC + c] = (signed char)s;
                const float d = s - s0[j];
#pragma unroll
                for (int j2 = j + 1; j2 < kSB; ++j2) f[j2] = fmaf(Js[j2 * kSB + j], d, f[j2]); // local row j2, column j
                ds[j * kTC + c] = d;
            }
        } else {
            const int t = tid - kTC;
            if (r0 + kSB < kBS) strips2_fetch_J(a, z, r0 + kSB, t, NH); // next strip's J columns
            if (r0 > 0) {
                strips2_update(z, Hb, r0 - kSB, r0 + kSB, kBS, t, NH);      // previous strip -> rows beyond the next one
                strips2_writeback(a, z, r0 - kSB, col0, t, NH);
            }
            if (r0 + kSB < kBS) strips2_draw<PT>(a, z, r0 + kSB, col0, t, NH);
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        if (r0 == 0) dense_stamp(a, 3);
        if (r0 + kSB < kBS) strips2_update(z, Hb, r0, r0 + kSB, r0 + 2 * kSB, tid, kThreads); // this strip -> the next strip's rows
        __syncthreads();
        if (r0 == 0) dense_stamp(a, 4);
    }
#undef MCS_ACCEPT
    dense_stamp(a, 5);
    strips2_writeback(a, z, kBS - kSB, col0, tid, kThreads);
}

__device__ __forceinline__ void dense_phase_b_dispatch(const DensePass &a, float *Hb, unsigned char *scr,
                                                       float *uglob, int col0)
{
    switch (a.PS) {
    case 1: dense_phase_b_strips<1>(a, Hb, scr, uglob, col0); break;
    case 2: dense_phase_b_strips<2>(a, Hb, scr, uglob, col0); break;
    case 4: dense_phase_b_strips<4>(a, Hb, scr, uglob, col0); break;
    case 8: dense_phase_b_strips<8>(a, Hb, scr, uglob, col0); break;
    case 16: dense_phase_b_strips<16>(a, Hb, scr, uglob, col0); break;
    case 32: dense_phase_b_strips<32>(a, Hb, scr, uglob, col0); break;
    default: dense_phase_b_rows<0>(a, Hb, scr, uglob, col0); break;
    }
}

// ---- variant 1: phase A on the legacy tensor path (mma.sync through WMMA), operands read from L2 ----
__global__ void __launch_bounds__(kThreads) dense_block_kernel(const __grid_constant__ DensePass a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *Hb = reinterpret_cast<float *>(smem_raw); // [kBS][kHld]
    unsigned char *scr = smem_raw + kBS * kHld * 4;
    float *uglob = reinterpret_cast<float *>(scr + kScrBytes);
    const int warp = threadIdx.x >> 5;
    const int col0 = blockIdx.x * kTC;
    const int i0 = a.i0;
    const long long ld = a.Npad;
    {
        using namespace nvcuda;
        const int row0 = (warp >> 1) * 32, c0w = (warp & 1) * 32;
        wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[2][2];
#pragma unroll
        for (int ri = 0; ri < 2; ++ri)
#pragma unroll
            for (int ci = 0; ci < 2; ++ci) wmma::fill_fragment(acc[ri][ci], 0.0f);
        const __nv_bfloat16 *Ahi = a.Jhi + (long long)(i0 + row0) * ld;
        const __nv_bfloat16 *Alo = a.Jlo + (long long)(i0 + row0) * ld;
        const __nv_bfloat16 *Bp = a.S + (long long)(col0 + c0w) * ld;
#pragma unroll 2
        for (int k0 = 0; k0 < a.Npad; k0 += 16) {
            wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, wmma::row_major> fhi[2], flo[2];
            wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::col_major> fb[2];
#pragma unroll
            for (int ri = 0; ri < 2; ++ri) {
                wmma::load_matrix_sync(fhi[ri], Ahi + (long long)(16 * ri) * ld + k0, (unsigned)ld);
                wmma::load_matrix_sync(flo[ri], Alo + (long long)(16 * ri) * ld + k0, (unsigned)ld);
            }
#pragma unroll
            for (int ci = 0; ci < 2; ++ci) wmma::load_matrix_sync(fb[ci], Bp + (long long)(16 * ci) * ld + k0, (unsigned)ld);
#pragma unroll
            for (int ri = 0; ri < 2; ++ri)
#pragma unroll
                for (int ci = 0; ci < 2; ++ci) {
                    wmma::mma_sync(acc[ri][ci], fhi[ri], fb[ci], acc[ri][ci]);
                    wmma::mma_sync(acc[ri][ci], flo[ri], fb[ci], acc[ri][ci]);
                }
        }
#pragma unroll
        for (int ri = 0; ri < 2; ++ri)
#pragma unroll
            for (int ci = 0; ci < 2; ++ci)
                wmma::store_matrix_sync(Hb + (row0 + 16 * ri) * kHld + c0w + 16 * ci, acc[ri][ci], kHld,
                                        wmma::mem_row_major);
    }
    __syncthreads();
    dense_phase_b_dispatch(a, Hb, scr, uglob, col0);
}

// ---- variant 2 (default): phase A on tcgen05 -------------------------------------------------------
// TMA (cp.async.bulk.tensor, 128-byte swizzle) streams [128 x 64] tiles of Jhi / Jlo and a [64 x 64] tile of
// S per stage into a 3-stage shared-memory ring; one thread issues tcgen05.mma (M = 128, N = 64, K = 16,
// kind::f16, bf16 inputs) for the hi and the lo tile into ONE fp32 accumulator in tensor memory (64
// columns); tcgen05.commit hands stages back to the producer and finally signals the epilogue warps, which
// read the accumulator with tcgen05.ld (lane = block row) into the shared field tile Hb.  Phase B then
// reuses the ring's shared memory.
constexpr int kStages = 3;
constexpr int kBK = 64;                                   // K elements per stage = one 128-byte swizzle row
constexpr int kStageA = kBS * kBK * 2;                    // 16 KB
constexpr int kStageB = kTC * kBK * 2;                    //  8 KB
constexpr int kStageBytes = 2 * kStageA + kStageB;        // 40 KB
constexpr int kRingBytes = kStages * kStageBytes;         // 120 KB
constexpr int kHbBytes = kBS * kHld * 4;
constexpr int kTmemCols = 64;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const void *tmap, int x, int y, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}
// K-major, 128-byte-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused (= 1),
// descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

struct DenseMaps {
    alignas(64) unsigned char jhi[128];
    alignas(64) unsigned char jlo[128];
    alignas(64) unsigned char s[128];
};

__global__ void __launch_bounds__(kThreads, 1)
    dense_block_kernel_tc(const __grid_constant__ DensePass a, const __grid_constant__ DenseMaps maps)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    mcs_pdl_launch_dependents(); // the next block's launch, barrier set-up and TMEM allocation overlap this block's tail
    dense_stamp(a, 0);
    // carve: [ring 120 KB | Hb 34 KB | barriers], ring 1024-byte aligned (128-byte swizzle atoms are 1024 B);
    // phase B scratch aliases the ring
    unsigned char *ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float *Hb = reinterpret_cast<float *>(ring + kRingBytes);
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + kRingBytes + kHbBytes);
    uint64_t *empty = full + kStages;
    uint64_t *accum_full = empty + kStages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum_full + 1);
    unsigned char *scr2 = ring + kRingBytes + kHbBytes + 128 + kUglobBytes; // pipelined phase B (16-byte aligned)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int col0 = blockIdx.x * kTC;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(accum_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const int nkb = a.Npad / kBK;

    // K order: the block's field needs every spin of its columns, but the spins of block j were last written by
    // the step that visited block j.  Cyclic order starting at the own block = oldest writer first: everything up to
    // the last few blocks is long final, and only the tiles of the block just before this one have to wait for the
    // previous step of this column group -- which is running on another SM right now (the kernels of consecutive
    // blocks overlap: no griddepcontrol.wait), so this step's GEMM hides behind that step's decision chain.
    const int b0 = a.i0 / kBS;
    const int tpb = kBS / kBK; // K tiles per block
    if (warp == 0 && lane == 0) {
        // ---- TMA producer
        unsigned seen = 0;
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % kStages;
            const int q = kb / tpb;                         // cyclic position of the tile's block
            const int kbe = (b0 * tpb + kb) % nkb;          // the tile
            const long long need = (long long)a.step - (a.nblocks - 1 - q); // steps that must be complete
            if (need > (long long)seen) {
                unsigned spins = 0;
                do {
                    asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(seen) : "l"(a.ver + blockIdx.x) : "memory");
                    if ((long long)seen >= need) break;
                    __nanosleep(100);
                } while (++spins < (1u << 24));
                if ((long long)seen < need) atomicExch(a.ver + gridDim.x, 1u); // cannot happen: the producer runs
                asm volatile("fence.proxy.async;" ::: "memory"); // generic-proxy writes -> the TMA reads below
            }
            mbar_wait(&empty[s], ((kb / kStages) & 1) ^ 1);
            unsigned char *st = ring + s * kStageBytes;
            mbar_expect_tx(&full[s], kStageBytes);
            tma_load_2d(st, maps.jhi, kbe * kBK, a.i0, &full[s]);
            tma_load_2d(st + kStageA, maps.jlo, kbe * kBK, a.i0, &full[s]);
            tma_load_2d(st + 2 * kStageA, maps.s, kbe * kBK, col0, &full[s]);
        }
    } else if (warp == 1 && lane == 0) {
        // ---- MMA issuer: D[128 x 64] (+)= Ahi . B^T + Alo . B^T
        // instruction descriptor: D = F32, A = B = BF16, both K-major, N = 64, M = 128
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTC >> 3) << 17) |
                               ((uint32_t)(kBS >> 4) << 24);
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % kStages;
            mbar_wait(&full[s], (kb / kStages) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = smem_u32(ring + s * kStageBytes);
            const uint64_t dhi = umma_desc_sw128(sa), dlo = umma_desc_sw128(sa + kStageA),
                           db = umma_desc_sw128(sa + 2 * kStageA);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) // 16 bf16 = 32 bytes = 2 descriptor units along K
                umma_f16(tmem_base, dhi + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) umma_f16(tmem_base, dlo + 2 * k, db + 2 * k, idesc, 1u);
            umma_commit(&empty[s]); // stage reusable once these MMAs have read it
        }
        umma_commit(accum_full);
    } else if (warp >= 4 && a.PS <= 32) {
        // ---- warps 4-7: the part of phase B's prologue that does not need the fields, under the GEMM.  The own
        // block's spins were last written by this column group's step `step - nblocks`.
        if (tid == 128) {
            const long long need = (long long)a.step - a.nblocks + 1;
            unsigned seen = 0, spins = 0;
            while (need > 0) {
                asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(seen) : "l"(a.ver + blockIdx.x) : "memory");
                if ((long long)seen >= need || ++spins >= (1u << 24)) break;
                __nanosleep(100);
            }
            if (need > 0 && (long long)seen < need) atomicExch(a.ver + gridDim.x, 1u);
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");
        switch (a.PS) {
        case 1: strips2_early<1>(a, scr2, col0, tid - 128, 128); break;
        case 2: strips2_early<2>(a, scr2, col0, tid - 128, 128); break;
        case 4: strips2_early<4>(a, scr2, col0, tid - 128, 128); break;
        case 8: strips2_early<8>(a, scr2, col0, tid - 128, 128); break;
        case 16: strips2_early<16>(a, scr2, col0, tid - 128, 128); break;
        default: strips2_early<32>(a, scr2, col0, tid - 128, 128); break;
        }
    }
    __syncwarp();
    if (warp < 4) {
        // ---- epilogue: TMEM lane = block row, 64 fp32 columns -> Hb
        mbar_wait(accum_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t v[64];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
            "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
            "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]),
              "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]),
              "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]),
              "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]),
              "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float *hrow = Hb + (warp * 32 + lane) * kHld;
#pragma unroll
        for (int q = 0; q < 64; ++q) hrow[q] = __uint_as_float(v[q]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
    }
    dense_stamp(a, 1);
    if (a.PS <= 32) { // pipelined phase B on its own scratch (its early part ran under the GEMM)
        switch (a.PS) {
        case 1: dense_phase_b_strips2<1>(a, Hb, scr2, col0); break;
        case 2: dense_phase_b_strips2<2>(a, Hb, scr2, col0); break;
        case 4: dense_phase_b_strips2<4>(a, Hb, scr2, col0); break;
        case 8: dense_phase_b_strips2<8>(a, Hb, scr2, col0); break;
        case 16: dense_phase_b_strips2<16>(a, Hb, scr2, col0); break;
        default: dense_phase_b_strips2<32>(a, Hb, scr2, col0); break;
        }
    } else { // P > 32: row-by-row scheme; its scratch lives in the (now idle) ring
        dense_phase_b_dispatch(a, Hb, ring, reinterpret_cast<float *>(ring + kRingBytes + kHbBytes + 128), col0);
    }
    dense_stamp(a, 6);
    // this column group has completed step a.step: its spins are written back (all threads), publish
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        asm volatile("st.release.gpu.u32 [%0], %1;" ::"l"(a.ver + blockIdx.x), "r"(a.step + 1u) : "memory");
    }
}

// W[N][Rpad] (bit k = slice k) -> S[(r P + k)][site]
__global__ void dense_expand_piqmc_kernel(const uint64_t *__restrict__ W, __nv_bfloat16 *__restrict__ S, int N,
                                          int Npad, long long R, long long Rpad, int P, int PS, long long Cpad)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Cpad * Npad) return;
    const long long col = t / Npad;
    const int i = (int)(t % Npad);
    const long long r = col / PS;
    const int k = (int)(col % PS);
    float v = 1.0f;
    if (i < N && r < R && k < P) v = ((W[(long long)i * Rpad + r] >> k) & 1ull) ? -1.0f : 1.0f;
    S[t] = __float2bfloat16(v);
}

__global__ void dense_compress_piqmc_kernel(const __nv_bfloat16 *__restrict__ S, uint64_t *__restrict__ W, int N,
                                            int Npad, long long R, long long Rpad, int P, int PS)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)N * R) return;
    const long long r = t / N;
    const int i = (int)(t % N);
    uint64_t w = 0;
    for (int k = 0; k < P; ++k)
        w |= (uint64_t)(__bfloat162float(S[(r * PS + k) * Npad + i]) < 0.0f) << k;
    W[(long long)i * Rpad + r] = w;
}

// V[N][G] (bit b of word g = restart 32 g + b) -> S[restart][site]
__global__ void dense_expand_sa_kernel(const uint32_t *__restrict__ V, __nv_bfloat16 *__restrict__ S, int N, int Npad,
                                       long long R, long long G, long long Cpad)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Cpad * Npad) return;
    const long long r = t / Npad;
    const int i = (int)(t % Npad);
    float v = 1.0f;
    if (i < N && r < R) v = ((V[(long long)i * G + (r >> 5)] >> (r & 31)) & 1u) ? -1.0f : 1.0f;
    S[t] = __float2bfloat16(v);
}

__global__ void dense_compress_sa_kernel(const __nv_bfloat16 *__restrict__ S, uint32_t *__restrict__ V, int N, int Npad,
                                         long long R, long long G)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)N * G) return;
    const long long g = t / N;
    const int i = (int)(t % N);
    uint32_t v = 0;
    for (int b = 0; b < 32; ++b) {
        const long long r = g * 32 + b;
        if (r < R) v |= (uint32_t)(__bfloat162float(S[r * Npad + i]) < 0.0f) << b;
    }
    V[(long long)i * G + g] = v;
}

constexpr size_t kSmemBytes = kBS * kHld * 4 + kScrBytes + kUglobBytes;
constexpr size_t kSmemBytesTc = kRingBytes + kHbBytes + 128 + kUglobBytes + kS2Bytes + 1024; // + barriers + world-line
                                                                 // uniforms (P > 32) + pipelined phase B + alignment slack
static_assert(kSmemBytesTc <= 227 * 1024, "dense_block_kernel_tc: shared memory");
static_assert(kScrBytes <= kRingBytes, "phase B scratch must fit the ring");

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map_bf16_2d(void *out128, const void *base, uint64_t inner, uint64_t rows, uint32_t box_inner,
                     uint32_t box_rows)
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        cudaDriverEntryPointQueryResult q;
        void *p = nullptr;
        MCS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        MCS_REQUIRE(p && q == cudaDriverEntryPointSuccess, MCS_ENODEVICE, "cuTensorMapEncodeTiled not available");
        fn = (EncodeTiledFn)p;
    }
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap is 128 bytes");
    const cuuint64_t dims[2] = {inner, rows};
    const cuuint64_t strides[1] = {inner * 2};
    const cuuint32_t box[2] = {box_inner, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn((CUtensorMap *)out128, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MCS_REQUIRE(r == CUDA_SUCCESS, MCS_ENODEVICE, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return MCS_OK;
}

} // namespace

static void launch_dense_tc(unsigned grid, cudaStream_t stream, const DensePass &a, const DenseMaps &maps)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytesTc;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    mcs_note_launch(cudaLaunchKernelEx(&cfg, dense_block_kernel_tc, a, maps));
}

bool mcs_dense_supported(const mcs_instance *inst, int P)
{
    if (!inst->dense || inst->nsteps != 1) return false;
    return P >= 1 && P <= kTC; // any number of slices: a replica takes the next power of two (or 64) columns
}

// kind: MCS_KIND_PIQMC (A, B schedules) or MCS_KIND_SA (A = temperature schedule, B unused)
int mcs_launch_dense_sweeps(mcs_state *st, int kind, const double *A, const double *B, int64_t S, int mcsteps,
                            float temp, int global_moves, uint64_t seed, uint64_t replica_offset,
                            uint64_t sweep_offset)
{
    mcs_instance *inst = st->inst;
    MCS_CUDA(cudaSetDevice(inst->device));
    const int P = (int)st->P;
    int PS = 1; // columns per replica
    while (PS < P) PS *= 2;
    const long long C = st->R * PS;
    const long long Cpad = (C + kTC - 1) / kTC * kTC;
    const int Npad = (int)inst->Npad;
    if (!st->d_S16 || st->S16_cols != Cpad) {
        if (st->d_S16) MCS_CUDA(cudaFree(st->d_S16));
        st->d_S16 = nullptr;
        MCS_CUDA(cudaMalloc(&st->d_S16, (size_t)Cpad * Npad * sizeof(__nv_bfloat16)));
        st->S16_cols = Cpad;
    }
    __nv_bfloat16 *S16 = (__nv_bfloat16 *)st->d_S16;
    // opt-in shared-memory sizes are a per-device property of the function: set them once per device
    static bool attr_set[64] = {};
    const int dev = inst->device;
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        MCS_CUDA(cudaFuncSetAttribute(dense_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)kSmemBytes));
        MCS_CUDA(cudaFuncSetAttribute(dense_block_kernel_tc, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)kSmemBytesTc));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    // MCS_DENSE_IMPL=wmma selects the legacy-tensor-path variant (cross-check); default: tcgen05 + TMA
    const char *impl = getenv("MCS_DENSE_IMPL");
    const bool use_tc = !(impl && impl[0] == 'w');
    DenseMaps maps;
    if (use_tc) {
        MCS_TRY(make_map_bf16_2d(maps.jhi, inst->d_Jhi, (uint64_t)Npad, (uint64_t)Npad, kBK, kBS));
        MCS_TRY(make_map_bf16_2d(maps.jlo, inst->d_Jlo, (uint64_t)Npad, (uint64_t)Npad, kBK, kBS));
        MCS_TRY(make_map_bf16_2d(maps.s, S16, (uint64_t)Npad, (uint64_t)Cpad, kBK, kTC));
    }
    const long long nexp = Cpad * Npad;
    if (kind == MCS_KIND_PIQMC)
        dense_expand_piqmc_kernel<<<(unsigned)((nexp + 255) / 256), 256, 0, inst->stream>>>(
            st->d_W, S16, (int)inst->N, Npad, st->R, st->Rpad, P, PS, Cpad);
    else
        dense_expand_sa_kernel<<<(unsigned)((nexp + 255) / 256), 256, 0, inst->stream>>>(st->d_V, S16, (int)inst->N,
                                                                                        Npad, st->R, st->G, Cpad);
    inst->launches++;

    DensePass a;
    a.Jhi = (const __nv_bfloat16 *)inst->d_Jhi;
    a.Jlo = (const __nv_bfloat16 *)inst->d_Jlo;
    a.Jf = inst->d_Jf;
    a.h = inst->d_hpad;
    a.S = S16;
    a.Npad = Npad;
    a.N = (int)inst->N;
    a.C = (int)C;
    a.P = P;
    a.PS = PS;
    a.trotter = kind == MCS_KIND_PIQMC ? 1 : 0;
    a.keys = mcs_philox_expand(seed);
    a.replica_offset = (uint32_t)replica_offset;
    a.global_moves = (kind == MCS_KIND_PIQMC && global_moves) ? 1 : 0;
    a.gscale = a.ginv = 1.0f;
    a.trace = nullptr;
    a.ver = nullptr;
    a.step = 0;
    a.nblocks = (int)((inst->N + kBS - 1) / kBS);
    const unsigned groups = (unsigned)(Cpad / kTC);
    if (use_tc) {
        MCS_CUDA(cudaMallocAsync((void **)&a.ver, (groups + 1) * sizeof(unsigned), inst->stream));
        MCS_CUDA(cudaMemsetAsync(a.ver, 0, (groups + 1) * sizeof(unsigned), inst->stream));
    }
    if (getenv("MCS_DENSE_TRACE")) MCS_CUDA(cudaMalloc((void **)&a.trace, 8 * sizeof(unsigned long long)));
    const double teff = (double)temp * (double)P;
    uint64_t sweep = sweep_offset;
    for (int64_t f = 0; f < S; ++f) {
        if (kind == MCS_KIND_PIQMC) {
            const double jperp = -0.5 * teff * log(tanh(A[f] / teff)); // qmc.pyx:95
            a.bcoef = (float)(-2.0 * B[f]);
            a.jperp2 = (float)(2.0 * jperp);
            a.nl2e_over_t = (float)(-1.4426950408889634 / teff);
            const double dEmax = std::max(1e-30, std::fabs(2.0 * B[f]) * inst->field_bound * (double)P);
            const int K = std::max(-60, std::min(60, (int)std::floor(std::log2(1073741824.0 / dEmax))));
            a.gscale = (float)std::ldexp(1.0, K);
            a.ginv = (float)std::ldexp(1.0, -K);
        } else {
            a.bcoef = -2.0f;
            a.jperp2 = 0.0f;
            a.nl2e_over_t = (float)(-1.4426950408889634 / A[f]);
        }
        for (int step = 0; step < mcsteps; ++step, ++sweep) {
            a.sweep_lo = (uint32_t)sweep;
            a.sweep_hi = (uint32_t)(sweep >> 32);
            for (int i0 = 0; i0 < (int)inst->N; i0 += kBS) {
                a.i0 = i0;
                if (use_tc) {
                    launch_dense_tc(groups, inst->stream, a, maps);
                    a.step++;
                }
                else
                    dense_block_kernel<<<(unsigned)(Cpad / kTC), kThreads, kSmemBytes, inst->stream>>>(a);
                inst->launches++;
            }
        }
    }
    if (kind == MCS_KIND_PIQMC) {
        const long long n = inst->N * st->R;
        dense_compress_piqmc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, inst->stream>>>(
            S16, st->d_W, (int)inst->N, Npad, st->R, st->Rpad, P, PS);
    } else {
        const long long n = inst->N * st->G;
        dense_compress_sa_kernel<<<(unsigned)((n + 255) / 256), 256, 0, inst->stream>>>(S16, st->d_V, (int)inst->N, Npad,
                                                                                       st->R, st->G);
    }
    inst->launches++;
    MCS_CUDA(mcs_take_launch_error());
    MCS_CUDA(cudaGetLastError());
    if (a.ver) {
        if (getenv("MCS_DENSE_CHECK")) { // tests: surface the (impossible) wait time-out
            unsigned flag = 0;
            MCS_CUDA(cudaMemcpyAsync(&flag, a.ver + groups, sizeof(flag), cudaMemcpyDeviceToHost, inst->stream));
            MCS_CUDA(cudaStreamSynchronize(inst->stream));
            MCS_CUDA(cudaFreeAsync(a.ver, inst->stream));
            MCS_REQUIRE(flag == 0, MCS_ENODEVICE, "dense sweeps: a wait on the previous block's step timed out");
        } else {
            MCS_CUDA(cudaFreeAsync(a.ver, inst->stream));
        }
    }
    if (a.trace) { // timestamps of the last block launch: start, GEMM done, prologue done, first strip decided,
                   // first strip applied, all strips done, spins written back
        unsigned long long t[8];
        MCS_CUDA(cudaStreamSynchronize(inst->stream));
        MCS_CUDA(cudaMemcpy(t, a.trace, sizeof(t), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[mcs dense] CTA 0 of the last block: GEMM %.1f us, prologue %.1f, strip 0 decisions %.1f, strip 0 "
                        "update %.1f, all strips %.1f, write-back %.1f\n",
                (t[1] - t[0]) * 1e-3, (t[2] - t[1]) * 1e-3, (t[3] - t[2]) * 1e-3, (t[4] - t[3]) * 1e-3,
                (t[5] - t[2]) * 1e-3, (t[6] - t[5]) * 1e-3);
        cudaFree(a.trace);
    }
    return MCS_OK;
}
