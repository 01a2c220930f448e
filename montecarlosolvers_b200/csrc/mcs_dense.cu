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
#include <cuda_bf16.h>
#include <mma.h>

#include <algorithm>
#include <cmath>

#include "mcs_common.cuh"

namespace {

constexpr int kBS = 128;      // sites per block
constexpr int kTC = 64;       // columns per CTA (a multiple of every supported P)
constexpr int kThreads = 256; // 8 warps
constexpr int kHld = kTC + 4; // leading dimension of the field tile (multiple of 4 floats for WMMA stores)
constexpr int kJld = kBS + 1;

struct DensePass {
    const __nv_bfloat16 *Jhi, *Jlo; // [Npad][Npad] row i = couplings of site i (bf16 split of J)
    const float *Jf;                // [Npad][Npad] fp32
    const float *h;                 // [Npad]
    __nv_bfloat16 *S;               // [Cpad][Npad]
    int Npad, N, C, P, i0;
    int trotter;      // 1: PIQMC (columns of a replica form a ring of P slices), 0: SA
    float bcoef;      // -2 B (PIQMC, qmc.pyx:96) or -2 (SA, sa.pyx:91-94)
    float jperp2;     // 2 J_perp
    float nl2e_over_t;
    mcs_philox_keys keys;
    uint32_t sweep_lo, sweep_hi, replica_offset;
    int global_moves;
};

__device__ __forceinline__ void bar_decide() { asm volatile("bar.sync 1, %0;" ::"n"(kTC)); }

__global__ void __launch_bounds__(kThreads) dense_block_kernel(const __grid_constant__ DensePass a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *Hb = reinterpret_cast<float *>(smem_raw);               // [kBS][kHld]
    float *Jd = Hb + kBS * kHld;                                   // [kBS][kJld]
    float *delta = Jd + kBS * kJld;                                // [kTC]
    float *gterm = delta + kTC;                                    // [kTC]
    signed char *sb = reinterpret_cast<signed char *>(gterm + kTC); // [kBS][kTC]
    const int tid = threadIdx.x, warp = tid >> 5;
    const int col0 = blockIdx.x * kTC;
    const int i0 = a.i0;
    const long long ld = a.Npad;

    // ---- A. Hb = (Jhi + Jlo)[i0:i0+128, :] * S[:, col0:col0+64]  (tensor cores, fp32 accumulate) ----
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
    // diagonal block of J (fp32) and the block's spins of this CTA's columns
    for (int e = tid; e < kBS * kBS; e += kThreads) {
        const int r = e / kBS, c = e % kBS;
        Jd[r * kJld + c] = __ldg(&a.Jf[(long long)(i0 + r) * ld + i0 + c]);
    }
    for (int e = tid; e < kBS * kTC; e += kThreads) {
        const int c = e / kBS, m = e % kBS; // consecutive threads -> consecutive sites of one column
        sb[m * kTC + c] = __bfloat162float(a.S[(long long)(col0 + c) * ld + i0 + m]) < 0.0f ? -1 : 1;
    }
    __syncthreads();

    // ---- B. site after site ------------------------------------------------------------------
    const int c = tid;                 // decision threads: tid < kTC
    const int col = col0 + c;
    const int P = a.P;
    const int k = col % P;             // slice
    const int cl = c - k + (k == 0 ? P - 1 : k - 1), cr = c - k + (k == P - 1 ? 0 : k + 1);
    const uint32_t rep = a.replica_offset + (uint32_t)(col / P);
    const int mend = min(kBS, a.N - i0);
    for (int m = 0; m < mend; ++m) {
        int flipped = 0;
        if (tid < kTC) {
            const int site = i0 + m;
            const float field = Hb[m * kHld + c] + __ldg(&a.h[site]);
            const int s_init = sb[m * kTC + c];
            uint32_t rnd[4];
            mcs_philox4x32_10_rk(rep, (uint32_t)site, a.sweep_lo, (a.sweep_hi << 8) | (uint32_t)(k >> 2), a.keys, rnd);
            const uint32_t u = (k & 3) == 0 ? rnd[0] : (k & 3) == 1 ? rnd[1] : (k & 3) == 2 ? rnd[2] : rnd[3];
            int s = s_init;
#pragma unroll 1
            for (int parity = 0; parity < 2; ++parity) { // even slices, then odd slices (P is even or 1)
                if ((k & 1) == parity) {
                    float dE = a.bcoef * (float)s * field;
                    if (a.trotter) dE += a.jperp2 * (float)(s * (sb[m * kTC + cl] + sb[m * kTC + cr]));
                    if (col < a.C && u <= mcs_accept_threshold(dE, a.nl2e_over_t)) {
                        s = -s;
                        sb[m * kTC + c] = (signed char)s;
                    }
                }
                if (!a.trotter) break;
                bar_decide();
            }
            if (a.global_moves) { // world-line move: all P slices of the replica (qmc.pyx:405-438)
                gterm[c] = a.bcoef * (float)s * field;
                bar_decide();
                float dE = 0.0f;
                for (int q = 0; q < P; ++q) dE += gterm[c - k + q];
                mcs_philox4x32_10_rk(rep, (uint32_t)site, a.sweep_lo, (a.sweep_hi << 8) | MCS_TAG_GLOBAL, a.keys, rnd);
                if (col < a.C && rnd[0] <= mcs_accept_threshold(dE, a.nl2e_over_t)) {
                    s = -s;
                    sb[m * kTC + c] = (signed char)s;
                }
            }
            delta[c] = (float)(s - s_init);
            flipped = s != s_init;
        }
        if (!__syncthreads_or(flipped)) continue; // nobody flipped: fields unchanged
        {
            const int row = tid >> 1, cbase = (tid & 1) * (kTC / 2);
            if (row > m) {
                const float jv = Jd[row * kJld + m];
                if (jv != 0.0f) {
                    float *hrow = Hb + row * kHld + cbase;
#pragma unroll 8
                    for (int q = 0; q < kTC / 2; ++q) hrow[q] = fmaf(jv, delta[cbase + q], hrow[q]);
                }
            }
        }
        __syncthreads();
    }
    __syncthreads();
    // write the block's spins back
    for (int e = tid; e < kBS * kTC; e += kThreads) {
        const int cc = e / kBS, m = e % kBS;
        a.S[(long long)(col0 + cc) * ld + i0 + m] = __float2bfloat16((float)sb[m * kTC + cc]);
    }
}

// W[N][Rpad] (bit k = slice k) -> S[(r P + k)][site]
__global__ void dense_expand_piqmc_kernel(const uint64_t *__restrict__ W, __nv_bfloat16 *__restrict__ S, int N,
                                          int Npad, long long R, long long Rpad, int P, long long Cpad)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Cpad * Npad) return;
    const long long col = t / Npad;
    const int i = (int)(t % Npad);
    const long long r = col / P;
    const int k = (int)(col % P);
    float v = 1.0f;
    if (i < N && r < R) v = ((W[(long long)i * Rpad + r] >> k) & 1ull) ? -1.0f : 1.0f;
    S[t] = __float2bfloat16(v);
}

__global__ void dense_compress_piqmc_kernel(const __nv_bfloat16 *__restrict__ S, uint64_t *__restrict__ W, int N,
                                            int Npad, long long R, long long Rpad, int P)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)N * R) return;
    const long long r = t / N;
    const int i = (int)(t % N);
    uint64_t w = 0;
    for (int k = 0; k < P; ++k)
        w |= (uint64_t)(__bfloat162float(S[(r * P + k) * Npad + i]) < 0.0f) << k;
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

constexpr size_t kSmemBytes = sizeof(float) * (kBS * kHld + kBS * kJld + 2 * kTC) + kBS * kTC;

} // namespace

bool mcs_dense_supported(const mcs_instance *inst, int P)
{
    if (!inst->dense || inst->nsteps != 1) return false;
    return P == 1 || (P <= kTC && (kTC % P) == 0);
}

// kind: MCS_KIND_PIQMC (A, B schedules) or MCS_KIND_SA (A = temperature schedule, B unused)
int mcs_launch_dense_sweeps(mcs_state *st, int kind, const double *A, const double *B, int64_t S, int mcsteps,
                            float temp, int global_moves, uint64_t seed, uint64_t replica_offset,
                            uint64_t sweep_offset)
{
    mcs_instance *inst = st->inst;
    MCS_CUDA(cudaSetDevice(inst->device));
    const int P = (int)st->P;
    const long long C = st->R * P;
    const long long Cpad = (C + kTC - 1) / kTC * kTC;
    const int Npad = (int)inst->Npad;
    if (!st->d_S16 || st->S16_cols != Cpad) {
        if (st->d_S16) MCS_CUDA(cudaFree(st->d_S16));
        st->d_S16 = nullptr;
        MCS_CUDA(cudaMalloc(&st->d_S16, (size_t)Cpad * Npad * sizeof(__nv_bfloat16)));
        st->S16_cols = Cpad;
    }
    __nv_bfloat16 *S16 = (__nv_bfloat16 *)st->d_S16;
    static bool attr_set = false;
    if (!attr_set) {
        MCS_CUDA(cudaFuncSetAttribute(dense_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)kSmemBytes));
        attr_set = true;
    }
    const long long nexp = Cpad * Npad;
    if (kind == MCS_KIND_PIQMC)
        dense_expand_piqmc_kernel<<<(unsigned)((nexp + 255) / 256), 256, 0, inst->stream>>>(
            st->d_W, S16, (int)inst->N, Npad, st->R, st->Rpad, P, Cpad);
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
    a.trotter = kind == MCS_KIND_PIQMC ? 1 : 0;
    a.keys = mcs_philox_expand(seed);
    a.replica_offset = (uint32_t)replica_offset;
    a.global_moves = (kind == MCS_KIND_PIQMC && global_moves) ? 1 : 0;
    const double teff = (double)temp * (double)P;
    uint64_t sweep = sweep_offset;
    for (int64_t f = 0; f < S; ++f) {
        if (kind == MCS_KIND_PIQMC) {
            const double jperp = -0.5 * teff * log(tanh(A[f] / teff)); // qmc.pyx:95
            a.bcoef = (float)(-2.0 * B[f]);
            a.jperp2 = (float)(2.0 * jperp);
            a.nl2e_over_t = (float)(-1.4426950408889634 / teff);
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
                dense_block_kernel<<<(unsigned)(Cpad / kTC), kThreads, kSmemBytes, inst->stream>>>(a);
                inst->launches++;
            }
        }
    }
    if (kind == MCS_KIND_PIQMC) {
        const long long n = inst->N * st->R;
        dense_compress_piqmc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, inst->stream>>>(
            S16, st->d_W, (int)inst->N, Npad, st->R, st->Rpad, P);
    } else {
        const long long n = inst->N * st->G;
        dense_compress_sa_kernel<<<(unsigned)((n + 255) / 256), 256, 0, inst->stream>>>(S16, st->d_V, (int)inst->N, Npad,
                                                                                       st->R, st->G);
    }
    inst->launches++;
    MCS_CUDA(cudaGetLastError());
    return MCS_OK;
}
