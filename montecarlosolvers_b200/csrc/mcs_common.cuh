// mcs_common.cuh -- internal declarations shared by the translation units of libmcs_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/mcs_b200.h"

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
void mcs_set_error(const char *fmt, ...);
int mcs_cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define MCS_CUDA(call)                                                         \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) return mcs_cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define MCS_REQUIRE(cond, code, ...)    \
    do {                                \
        if (!(cond)) {                  \
            mcs_set_error(__VA_ARGS__); \
            return (code);              \
        }                               \
    } while (0)

#define MCS_TRY(expr)              \
    do {                           \
        int rc__ = (expr);         \
        if (rc__ != MCS_OK) return rc__; \
    } while (0)

// ------------------------------------------------------------------------------------------
// compiled instance (see mcs_instance.cu)
// ------------------------------------------------------------------------------------------
struct mcs_instance {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr; // copy streams of the pipelined one-shot calls (lazy)
    cudaStream_t s_aux[3] = {nullptr, nullptr, nullptr}; // extra launch streams: replica chunks pass-interleaved (lazy)
    cudaEvent_t ev_aux0 = nullptr, ev_aux1[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_up[16] = {}, ev_done[16] = {}, ev_t0 = nullptr;
    int64_t N = 0, maxnb = 0;
    int64_t nsteps = 1;  // > 1: time-dependent couplings, one table per schedule step (Noisy* functions)
    int ncolors = 0;
    int maxdeg = 0;      // max number of quadratic neighbours of a site (fields excluded)
    int dpad = 1;        // row length of the ELL tables (>= 1)
    bool has_field = false;
    bool lut_ok = false; // (maxdeg + has_field + 2) <= 10 planes: the threshold-table PIQMC kernels apply
    bool dense = false;  // near-complete graph: blocked tensor-core sweeps (mcs_dense.cu)
    int64_t Npad = 0;    // N rounded up to the dense block size (128)
    int64_t launches = 0;
    int dynamics = 0;    // MCS_DYN_COLORED / MCS_DYN_REFERENCE: what *_sweeps and the one-shot calls run
    // The one-shot host-buffer calls (mcs_*_anneal*, mcs_exact_*) share this instance's scratch batches, staging
    // buffer, window fields and stream: they hold this lock for their whole duration, so concurrent callers (the
    // reference released the GIL in its loops, qmc.pyx:92) are serialised per instance instead of corrupting each
    // other.  Calls on resident mcs_state batches are the caller's to order.
    std::recursive_mutex call_mutex;
    std::vector<struct mcs_state *> states; // live replica batches (orphaned if the instance dies first)
    struct mcs_state *scratch[4] = {nullptr, nullptr, nullptr, nullptr}; // per-kind batch reused by the
                                                                         // one-shot host-buffer calls

    // host copies
    std::vector<int32_t> color;       // [N]
    std::vector<int32_t> order;       // [N] sites sorted by colour (stable in site index)
    std::vector<int32_t> color_start; // [ncolors + 1] offsets into order

    // device tables
    int32_t *d_tab_idx = nullptr; // [nsteps][N][maxnb]  reference table, neighbour index column (int(nbs[..,0]))
    double *d_tab_J = nullptr;    // [nsteps][N][maxnb]  reference table, coupling column (fp64, row order kept)
    int32_t *d_ell_idx = nullptr; // [N][dpad]           quadratic neighbours, padded with the site itself
    float *d_ell_J = nullptr;     // [nsteps][N][dpad]   fp32 couplings, padded with 0
    float *d_h = nullptr;         // [nsteps][N]         fp32 local fields

    // tables of schedule step f (step 0 for static instances)
    const float *ell_J_at(int64_t f) const { return d_ell_J + (nsteps > 1 ? (size_t)f * N * dpad : 0); }
    const float *h_at(int64_t f) const { return d_h + (nsteps > 1 ? (size_t)f * N : 0); }
    const int32_t *tab_idx_at(int64_t f) const { return d_tab_idx + (nsteps > 1 ? (size_t)f * N * maxnb : 0); }
    const double *tab_J_at(int64_t f) const { return d_tab_J + (nsteps > 1 ? (size_t)f * N * maxnb : 0); }
    int32_t *d_order = nullptr;   // [N]
    int32_t *d_pos = nullptr;     // [N] inverse of order: position of a site in the colour-sorted list
    int max_offdiag = 0;          // most table entries with j != i in one row (padding included)
    double *d_etab = nullptr;     // [N][16] fixed-order energy terms per neighbour sign pattern (lazy, mcs_energy_tables)
    int32_t *d_etab_j = nullptr;  // [N][4]  the row's off-diagonal neighbours in table order, -1 padded
    void *d_Jhi = nullptr, *d_Jlo = nullptr; // dense only: [Npad][Npad] bf16 split J = hi + lo
    float *d_Jf = nullptr;                   // dense only: [Npad][Npad] fp32
    float *d_hpad = nullptr;                 // dense only: [Npad]
    double field_bound = 0.0;                // dense only: max_i (sum_j |J_ij| + |h_i|)
};

struct mcs_state {
    mcs_instance *inst = nullptr;
    int kind = 0;
    int64_t R = 0, P = 1;
    int64_t Rpad = 0;        // PIQMC / SVMC: R rounded up to a multiple of 32 (lanes = replicas)
    int64_t G = 0;           // SA: number of 32-replica words per site
    uint64_t *d_W = nullptr; // PIQMC  [N][Rpad]  bit k = slice k, bit set <=> spin -1
    uint32_t *d_V = nullptr; // SA     [N][G]     bit b of word g = replica 32 g + b, bit set <=> spin -1
    float *d_theta = nullptr; // SVMC  [N][Rpad]
    float *d_cosz = nullptr;  // SVMC  [N][Rpad]  cos(theta)
    void *d_stage = nullptr;  // staging buffer for host <-> device conversion
    size_t stage_bytes = 0;
    void *d_S16 = nullptr;    // dense sweeps: spins as bf16 +-1, [Cpad][Npad], column = (replica, slice)
    long long S16_cols = 0;
    uint64_t *d_Wpk = nullptr; // PIQMC packed mode: the packed working words of a sweep call, [N][32 x group warps] (kept)
    size_t Wpk_bytes = 0;
    double *d_eout = nullptr;    // per-replica energies of mcs_state_energies (kept: no cudaMalloc per call)
    size_t eout_bytes = 0;
    void *d_best = nullptr;      // best-slice results: ebest f64[R] | kbest i32[R] | conf i8[R][N]
    size_t best_bytes = 0;
    int32_t *d_labels = nullptr; // cluster moves: union-find parents [(N P + 1)][replicas]
    size_t labels_bytes = 0;
    // active replica window [v0, v0 + vR) of a PIQMC batch (vR < 0: everything).  Replicas are independent, so
    // the one-shot call pipelines chunks: upload(c+1) and download(c-1) overlap the sweeps of chunk c.
    long long v0 = 0, vR = -1;
    long long win_lo() const { return vR < 0 ? 0 : v0; }
    long long win_pad() const { return vR < 0 ? Rpad : vR; }                  // multiple of 32
    long long win_valid() const { return vR < 0 ? R : std::max(0ll, std::min(R - v0, vR)); }
};

int mcs_state_reserve_stage(mcs_state *st, size_t bytes);
// batch of the given shape owned by the instance and reused across one-shot calls (no cudaMalloc per call)
int mcs_instance_scratch_state(mcs_instance *inst, int kind, int64_t R, int64_t P, mcs_state **out);

// ------------------------------------------------------------------------------------------
// Philox4x32 counter-based RNG (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11) with
// MCS_PHILOX_ROUNDS = 7 rounds: the fewest for which the authors report a clean BigCrush ("Philox4x32-7 is
// Crush-resistant", their Table 2); ten is their default with a safety margin.  The generator is the binding
// cost of the sweep kernels (45 % of the PIQMC pass at ten rounds), so the margin is spent here:
// -DMCS_PHILOX_ROUNDS=10 restores it.  Key = the user's 64-bit seed
// (uniform over the launch); counter = (replica or word index, site, sweep, call tag), so a
// replica's stream does not depend on launch geometry or on how replicas are sharded over GPUs.
// ------------------------------------------------------------------------------------------
#ifndef MCS_BYTES_BY_PRMT
#define MCS_BYTES_BY_PRMT 0 // index bytes of a decision call: 0 = all eight through the LSU, 1 = four, 2 = none (PRMT)
#endif
#ifndef MCS_PHILOX_ROUNDS
#define MCS_PHILOX_ROUNDS 7
#endif
#define MCS_PHILOX_M0 0xD2511F53u
#define MCS_PHILOX_M1 0xCD9E8D57u
#define MCS_PHILOX_W0 0x9E3779B9u
#define MCS_PHILOX_W1 0xBB67AE85u

__host__ __device__ __forceinline__ void mcs_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                           uint32_t k0, uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < MCS_PHILOX_ROUNDS; ++r) {
#ifdef __CUDA_ARCH__
        uint32_t hi0 = __umulhi(MCS_PHILOX_M0, c0), lo0 = MCS_PHILOX_M0 * c0;
        uint32_t hi1 = __umulhi(MCS_PHILOX_M1, c2), lo1 = MCS_PHILOX_M1 * c2;
#else
        uint64_t p0 = (uint64_t)MCS_PHILOX_M0 * c0, p1 = (uint64_t)MCS_PHILOX_M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += MCS_PHILOX_W0;
        k1 += MCS_PHILOX_W1;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

// Same generator with the round keys precomputed on the host (rk[2r] = k0 + r W0,
// rk[2r+1] = k1 + r W1).  The keys then sit in the kernel-parameter constant bank and feed the
// 3-input XOR directly: 4 instructions per round (2 IMAD.WIDE + 2 LOP3) instead of 6.
struct mcs_philox_keys {
    uint32_t rk[2 * MCS_PHILOX_ROUNDS];
};

inline mcs_philox_keys mcs_philox_expand(uint64_t seed)
{
    mcs_philox_keys k;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < MCS_PHILOX_ROUNDS; ++r) {
        k.rk[2 * r] = k0;
        k.rk[2 * r + 1] = k1;
        k0 += MCS_PHILOX_W0;
        k1 += MCS_PHILOX_W1;
    }
    return k;
}

__device__ __forceinline__ void mcs_philox4x32_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                     const mcs_philox_keys &k, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < MCS_PHILOX_ROUNDS; ++r) {
        const uint32_t hi0 = __umulhi(MCS_PHILOX_M0, c0), lo0 = MCS_PHILOX_M0 * c0;
        const uint32_t hi1 = __umulhi(MCS_PHILOX_M1, c2), lo1 = MCS_PHILOX_M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k.rk[2 * r];
        const uint32_t n2 = hi0 ^ c3 ^ k.rk[2 * r + 1];
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

// Powers of two kept in the kernel-parameter constant bank: shifts done as IMAD / IMAD.HI by a
// multiplier ptxas cannot see through stay on the FMA pipe (an immediate 2^k would be strength-reduced to
// SHF on the ALU pipe, which is the binding unit of the sweep kernels).
//   left shift by s  (0 <= s <= 31):  x * up[s]
//   right shift by r (1 <= r <= 8):   __umulhi(x, down[r])        down[r] = 2^(32 - r)
struct mcs_pow2_table {
    uint32_t up[32];
    uint32_t down[9];
};

inline mcs_pow2_table mcs_pow2_make()
{
    mcs_pow2_table t;
    for (int i = 0; i < 32; ++i) t.up[i] = 1u << i;
    t.down[0] = 0;
    for (int r = 1; r <= 8; ++r) t.down[r] = 1u << (32 - r);
    return t;
}

// call tags (counter word 3): which draw inside one (replica, site, sweep)
enum {
    MCS_TAG_GROUP0 = 0,      // 0..15: PIQMC slice groups / SA bit groups
    MCS_TAG_LAST_SLICE = 16, // PIQMC odd-P closing slice
    MCS_TAG_GLOBAL = 17,     // PIQMC world-line move
    MCS_TAG_SVMC = 18,       // SVMC proposal + acceptance
    MCS_TAG_GLOBAL2 = 19,    // PIQMC world-line move, members 4.. of a packed group
    MCS_TAG_LAST_SLICE2 = 20, // PIQMC odd-P closing slice, members 4.. of a packed group
    MCS_TAG_REFINE = 0x80,   // OR-ed into a group tag: second half of the lazily refined uniforms
    MCS_TAG_INIT = 0x40000000u
};

// ---- programmatic dependent launch (sm_90+) -----------------------------------------------------------
// Consecutive colour passes depend on each other through the state only.  A pass kernel signals
// launch_dependents at once (the next pass may start occupying SMs as soon as every CTA of this one is resident)
// and executes griddepcontrol.wait right before its first read of the state: its launch latency, its parameter
// / coupling loads and its threshold-table build overlap the tail of the previous pass.  Matters for small
// batches, where a pass takes only a few microseconds.
__device__ __forceinline__ void mcs_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void mcs_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// First launch error since the last mcs_take_launch_error() on this thread: the pass launchers sit in loops of
// thousands of launches; they record a failure here and the sweep function reports it when the loop ends.
extern thread_local cudaError_t mcs_launch_error;
inline cudaError_t mcs_note_launch(cudaError_t e)
{
    if (e != cudaSuccess && mcs_launch_error == cudaSuccess) mcs_launch_error = e;
    return e;
}
inline cudaError_t mcs_take_launch_error()
{
    const cudaError_t e = mcs_launch_error;
    mcs_launch_error = cudaSuccess;
    return e;
}

template <typename Kernel, typename Args>
inline cudaError_t mcs_launch_pdl(Kernel kernel, dim3 grid, dim3 block, cudaStream_t stream, const Args &args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return mcs_note_launch(cudaLaunchKernelEx(&cfg, kernel, args));
}

// u <= T with T == 0 meaning NEVER: mcs_accept_threshold returns 0 exactly when exp(-dE/teff) 2^32 < 1 (underflow,
// T = 0 in the schedule, NaN energies) -- the reference never accepts there (it compares 0 > rand()/RAND_MAX,
// qmc.pyx:142).  In the table kernels the pair (u, T) = (0, 0) always looks like a tie, so the rule lives in the
// refinement path at no cost to the hot loop.
__device__ __forceinline__ bool mcs_accepts(uint32_t u, uint32_t T) { return (u <= T) & (T != 0u); }

// ---- building blocks shared by the bit-packed sweep kernels (mcs_piqmc.cu, mcs_sa.cu) ---------------
// Instruction budget, measured on B200 (benchmarks/micro/pipe_rates.cu): ALU-pipe instructions (LOP3, PRMT,
// IADD3, SHF, ISETP, VIADDMNMX) and FMA-pipe IMAD take 2 issue cycles per warp each and overlap with each
// other; IMAD.WIDE / IMAD.HI hold the FMA pipe for 4 cycles AND the ALU pipe for 2.  A Philox4x32-10 call is
// 20 IMAD.WIDE + 20 LOP3 -- both pipes saturated for ~65 cycles -- so (1) one call decides eight attempts
// (lazily refined uniforms, below), (2) left shifts and the shifting-in of decision bits are IMAD / IMAD.X by
// constant-bank powers of two (FMA pipe), (3) right shifts stay SHF (IMAD.HI would cost both pipes), (4) the
// bytes of an index word are extracted by the LSU (shared-memory bounce) rather than by PRMT.

// x shifted left by DELTA bits (right if negative): left on the FMA pipe, right on the ALU pipe
template <int DELTA>
__device__ __forceinline__ uint32_t mcs_plane_shift(uint32_t x, const mcs_pow2_table &t)
{
    if (DELTA == 0) return x;
    if (DELTA > 0) return x * t.up[DELTA > 0 ? DELTA : 0];
    return x >> (DELTA < 0 ? -DELTA : 0);
}

// byte i of v, zero extended (one PRMT; the second operand supplies the zero bytes)
__device__ __forceinline__ uint32_t mcs_prmt_byte(uint32_t v, int i)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(v), "r"(0u), "r"(0x4440u | (uint32_t)i));
    return r;
}

// threshold-table entry; with SH == 2 `off` already is the byte offset (pattern index * 4)
template <int SH>
__device__ __forceinline__ uint32_t mcs_lut_at(const uint32_t *lut, uint32_t off)
{
    return SH == 2 ? *(const uint32_t *)((const char *)lut + off) : lut[off];
}

// Decisions.  The tables hold ~T, so the carry of ~T + u is "u > T" = reject.  IMAD.X shifts it into a Horner
// accumulator (acc * mul + carry, mul = 256 from the constant bank so that ptxas cannot turn it into an ALU
// shift-add): visiting bytes 3,2,1,0 leaves the reject bit of byte i at bit 8 i.
//
// Lazily refined uniforms.  The uniform of an attempt is the 32-bit number u = (v << 16) | r, v = 16 bits of
// the group pair's Philox call, r = 16 bits of a SECOND call (tag | MCS_TAG_REFINE) that is evaluated only when
// it can matter: u <= T is decided by v alone unless v == T >> 16 (probability 2^-16), so one call serves
// eight attempts instead of four.  Fast path: group A compares the word x itself (v = x >> 16; the low half
// of x stands in for r and cannot change a decided comparison), group B compares x << 16 (v = x & 0xffff,
// r = 0).  The sum s = ~T + u lies within 2^16 of a wrap whenever a comparison is undecided (the test
// s + 2^16 < 2^17 mod 2^32 is conservative; one VIADDMNMX per attempt keeps the minimum in smin); the call is
// then flagged and redone with both halves (mcs_refine_call) after the hot loop.  The outcome is bit-identical
// to always evaluating both calls; tie_thr = 0xffffffff does exactly that (MCS_ALWAYS_REFINE=1, tested).
__device__ __forceinline__ uint32_t mcs_horner_reject(uint32_t acc, uint32_t mul, uint32_t nT, uint32_t u,
                                                      uint32_t &smin)
{
    uint32_t out, s;
    asm("add.cc.u32 %1, %2, %3;\n\tmadc.lo.u32 %0, %4, %5, 0;"
        : "=r"(out), "=r"(s)
        : "r"(nT), "r"(u), "r"(acc), "r"(mul));
    smin = min(smin, s + 0x10000u);
    return out;
}

// flags = flags * 2 + (smin < tie_thr): the "call undecided" bit, again as carry + IMAD.X
__device__ __forceinline__ void mcs_horner_flag(uint32_t &flags, uint32_t smin, uint32_t tie_thr, uint32_t two)
{
    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\tmadc.lo.u32 %0, %0, %3, 0;\n\t}"
        : "+r"(flags)
        : "r"(~smin), "r"(tie_thr), "r"(two));
}

// One Philox call decides the eight attempts whose pattern-index bytes are accA (group A) and accB (group B);
// returns the Horner accumulators (reject bit of byte i at bit 8 i) and updates the undecided flags.  The index
// bytes go through this thread's private 8-byte shared-memory slot (one 64-bit store, eight byte loads:
// conflict-free, a warp's slots are consecutive) -- inline PTX so that the store is not forwarded into shifts.
template <int SH>
__device__ __forceinline__ void mcs_decide_call(uint32_t &chA, uint32_t &chB, uint32_t &flags, uint32_t accA,
                                                uint32_t accB, const uint32_t *lut, uint32_t c0, uint32_t c1,
                                                uint32_t c2, uint32_t c3, const mcs_philox_keys &keys,
                                                const mcs_pow2_table &pow2, uint32_t tie_thr, uint2 *slot)
{
    uint32_t x[4];
    mcs_philox4x32_rk(c0, c1, c2, c3, keys, x);
    chA = 0;
    chB = 0;
    uint32_t smin = 0xFFFFFFFFu;
    const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(slot);
#if MCS_BYTES_BY_PRMT == 1
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(saddr), "r"(accA) : "memory");
#elif MCS_BYTES_BY_PRMT == 0
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(saddr), "r"(accA), "r"(accB) : "memory");
#endif
#pragma unroll
    for (int i = 3; i >= 0; --i) {
        uint32_t oA, oB;
#if MCS_BYTES_BY_PRMT == 2
        oA = mcs_prmt_byte(accA, i);
        oB = mcs_prmt_byte(accB, i);
#elif MCS_BYTES_BY_PRMT == 1
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(oA) : "r"(saddr + i) : "memory");
        oB = mcs_prmt_byte(accB, i);
#else
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(oA) : "r"(saddr + i) : "memory");
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(oB) : "r"(saddr + 4 + i) : "memory");
#endif
        chA = mcs_horner_reject(chA, pow2.up[8], mcs_lut_at<SH>(lut, oA), x[i], smin);
        chB = mcs_horner_reject(chB, pow2.up[8], mcs_lut_at<SH>(lut, oB), x[i] * pow2.up[16], smin);
    }
    mcs_horner_flag(flags, smin, tie_thr, pow2.up[1]);
}

// Slow path of the lazily refined uniforms: the same eight attempts with both Philox halves,
// u = (v << 16) | r.  Out of line so that this cold code costs the hot path no registers.
template <int SH>
__device__ __noinline__ uint2 mcs_refine_call(uint32_t accA, uint32_t accB, const uint32_t *lut, uint32_t c0,
                                              uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
    uint32_t x[4], f[4];
    mcs_philox4x32(c0, c1, c2, c3, k0, k1, x);
    mcs_philox4x32(c0, c1, c2, c3 | MCS_TAG_REFINE, k0, k1, f);
    uint32_t chA = 0, chB = 0; // reject bit of byte i at bit 8 i
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t TA = ~mcs_lut_at<SH>(lut, (accA >> (8 * i)) & 0xFFu);
        const uint32_t TB = ~mcs_lut_at<SH>(lut, (accB >> (8 * i)) & 0xFFu);
        const uint32_t uA = (x[i] & 0xFFFF0000u) | (f[i] >> 16);
        const uint32_t uB = (x[i] << 16) | (f[i] & 0xFFFFu);
        chA |= (mcs_accepts(uA, TA) ? 0u : 1u) << (8 * i);
        chB |= (mcs_accepts(uB, TB) ? 0u : 1u) << (8 * i);
    }
    return make_uint2(chA, chB);
}

// ---- column tables of the cluster-resident kernels (mcs_sa.cu: sa_cluster_kernel) -------------------------
// There the 32 lanes of a warp are 32 different SITES, each with its own threshold table.  Entry e of local site
// l sits at word e * stride + l (stride a multiple of 32): the bank of a lookup is the lane's site, whatever the
// pattern, so the random lookups of a warp never conflict.  `col` points at the site's column, `off` is the
// pattern index * 4 as in mcs_lut_at<2>: the address is one IMAD (FMA pipe) instead of an add on the ALU pipe.
__device__ __forceinline__ uint32_t mcs_lut_col(const uint32_t *col, uint32_t off, uint32_t stride)
{
    return *(const uint32_t *)((const char *)col + off * stride);
}

// mcs_decide_call / mcs_refine_call on a column table: same Philox call, same uniforms, same decisions
__device__ __forceinline__ void mcs_decide_call_col(uint32_t &chA, uint32_t &chB, uint32_t &flags, uint32_t accA,
                                                    uint32_t accB, const uint32_t *col, uint32_t stride, uint32_t c0,
                                                    uint32_t c1, uint32_t c2, uint32_t c3,
                                                    const mcs_philox_keys &keys, const mcs_pow2_table &pow2,
                                                    uint32_t tie_thr, uint2 *slot)
{
    uint32_t x[4];
    mcs_philox4x32_rk(c0, c1, c2, c3, keys, x);
    chA = 0;
    chB = 0;
    uint32_t smin = 0xFFFFFFFFu;
    const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(slot);
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(saddr), "r"(accA), "r"(accB) : "memory");
#pragma unroll
    for (int i = 3; i >= 0; --i) {
        uint32_t oA, oB;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(oA) : "r"(saddr + i) : "memory");
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(oB) : "r"(saddr + 4 + i) : "memory");
        chA = mcs_horner_reject(chA, pow2.up[8], mcs_lut_col(col, oA, stride), x[i], smin);
        chB = mcs_horner_reject(chB, pow2.up[8], mcs_lut_col(col, oB, stride), x[i] * pow2.up[16], smin);
    }
    mcs_horner_flag(flags, smin, tie_thr, pow2.up[1]);
}

static __device__ __noinline__ uint2 mcs_refine_call_col(uint32_t accA, uint32_t accB, const uint32_t *col,
                                                         uint32_t stride, uint32_t c0, uint32_t c1, uint32_t c2,
                                                         uint32_t c3, uint32_t k0, uint32_t k1)
{
    uint32_t x[4], f[4];
    mcs_philox4x32(c0, c1, c2, c3, k0, k1, x);
    mcs_philox4x32(c0, c1, c2, c3 | MCS_TAG_REFINE, k0, k1, f);
    uint32_t chA = 0, chB = 0; // reject bit of byte i at bit 8 i
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t TA = ~mcs_lut_col(col, (accA >> (8 * i)) & 0xFFu, stride);
        const uint32_t TB = ~mcs_lut_col(col, (accB >> (8 * i)) & 0xFFu, stride);
        const uint32_t uA = (x[i] & 0xFFFF0000u) | (f[i] >> 16);
        const uint32_t uB = (x[i] << 16) | (f[i] & 0xFFFFu);
        chA |= (mcs_accepts(uA, TA) ? 0u : 1u) << (8 * i);
        chB |= (mcs_accepts(uB, TB) ? 0u : 1u) << (8 * i);
    }
    return make_uint2(chA, chB);
}

// ---- thread-block clusters: distributed shared memory and the cluster barrier (sm_90+) --------------------
__device__ __forceinline__ uint32_t mcs_cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// shared::cluster address of `saddr` (a shared::cta address of this CTA's layout) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mcs_mapa(uint32_t saddr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ uint32_t mcs_ld_cluster_u32(uint32_t caddr)
{
    uint32_t v;
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(caddr) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t mcs_ld_cluster_u64(uint32_t caddr)
{
    uint64_t v;
    asm volatile("ld.shared::cluster.u64 %0, [%1];" : "=l"(v) : "r"(caddr) : "memory");
    return v;
}
// split barrier: every thread of every CTA of the cluster arrives (release) and later waits (acquire)
__device__ __forceinline__ void mcs_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void mcs_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// ---- wide index fields (9 or 10 planes: degree + field of 7 or 8 in the PIQMC kernel) --------------------
// The pattern index no longer fits a byte, so an index word holds TWO 16-bit fields (index * 4) and one Philox
// call serves four index words: A0, A1 (compare x[1], x[0] and x[3], x[2] themselves) and B0, B1 (the same words
// shifted left by 16).  Otherwise identical to mcs_decide_call / mcs_refine_call: reject bit of field i ends at
// bit 16 i of the Horner accumulator.
__device__ __forceinline__ void mcs_decide_call16(uint32_t (&ch)[4], uint32_t &flags, const uint32_t (&acc)[4],
                                                  const uint32_t *lut, uint32_t c0, uint32_t c1, uint32_t c2,
                                                  uint32_t c3, const mcs_philox_keys &keys,
                                                  const mcs_pow2_table &pow2, uint32_t tie_thr, uint4 *slot)
{
    uint32_t x[4];
    mcs_philox4x32_rk(c0, c1, c2, c3, keys, x);
    uint32_t smin = 0xFFFFFFFFu;
    const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(slot);
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(acc[0]), "r"(acc[1]), "r"(acc[2]),
                 "r"(acc[3])
                 : "memory");
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        ch[w] = 0;
#pragma unroll
        for (int i = 1; i >= 0; --i) {
            uint32_t off;
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(off) : "r"(saddr + 4 * w + 2 * i) : "memory");
            const uint32_t xi = x[2 * (w & 1) + i];
            ch[w] = mcs_horner_reject(ch[w], pow2.up[16], mcs_lut_at<2>(lut, off), w < 2 ? xi : xi * pow2.up[16], smin);
        }
    }
    mcs_horner_flag(flags, smin, tie_thr, pow2.up[1]);
}

static __device__ __noinline__ uint4 mcs_refine_call16(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, const uint32_t *lut,
                                                uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                uint32_t k1)
{
    uint32_t x[4], f[4];
    mcs_philox4x32(c0, c1, c2, c3, k0, k1, x);
    mcs_philox4x32(c0, c1, c2, c3 | MCS_TAG_REFINE, k0, k1, f);
    const uint32_t acc[4] = {a0, a1, a2, a3};
    uint32_t ch[4] = {0u, 0u, 0u, 0u}; // reject bit of field i at bit 16 i
#pragma unroll
    for (int w = 0; w < 4; ++w)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const uint32_t T = ~mcs_lut_at<2>(lut, (acc[w] >> (16 * i)) & 0xFFFFu);
            const uint32_t xi = x[2 * (w & 1) + i], fi = f[2 * (w & 1) + i];
            const uint32_t u = w < 2 ? ((xi & 0xFFFF0000u) | (fi >> 16)) : ((xi << 16) | (fi & 0xFFFFu));
            ch[w] |= (mcs_accepts(u, T) ? 0u : 1u) << (16 * i);
        }
    return make_uint4(ch[0], ch[1], ch[2], ch[3]);
}

inline uint32_t mcs_tie_threshold()
{
    const char *e = getenv("MCS_ALWAYS_REFINE");
    return (e && e[0] == '1') ? 0xFFFFFFFFu : 0x20000u;
}

// Metropolis acceptance threshold: the move is accepted iff a uniform 32-bit draw u satisfies
// u <= T (mcs_accepts).  dE <= 0 -> always (qmc.pyx:140-141); otherwise P(accept) = ceil(p 2^32)/2^32 with
// p = exp(-dE/teff) (qmc.pyx:142), floored at 2^-32 (the reference's own floor is 2^-31: it
// compares exp(..) > rand()/RAND_MAX and rand() returns 0 once in 2^31 draws).
// nl2e_over_t = -log2(e)/teff.  NaN dE (inf - inf at A = 0) -> never, like the reference.
__device__ __forceinline__ uint32_t mcs_accept_threshold(float dE, float nl2e_over_t)
{
    if (dE <= 0.0f) return 0xFFFFFFFFu;
    float t = ceilf(exp2f(dE * nl2e_over_t) * 4294967296.0f);
    if (!(t >= 1.0f)) return 0u;
    if (t >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)t - 1u;
}

// Fixed-order energy by table (rows with at most four off-diagonal entries): builds the instance's tables on first
// use (mcs_piqmc.cu).  Returns false when the instance does not qualify.
bool mcs_energy_tables(mcs_instance *inst);

// PIQMC packed mode: floor(64 / P) (at most 6) world lines per working word -- every P from 2 to 32
inline bool mcs_piqmc_packs(int P) { return P >= 2 && P <= 32; }

// kernels launchers implemented in the other translation units
int mcs_launch_piqmc_sweeps(mcs_state *st, const double *A, const double *B, int64_t S, int mcsteps, float temp,
                            int global_moves, uint64_t seed, uint64_t replica_offset, uint64_t sweep_offset,
                            const double *lookuptable /* nullptr: no Ohmic bath */);
// reference dynamics (mcs_refdyn.cu): random-permutation sequential sweeps executed as dependency waves
int mcs_launch_refdyn_sweeps(mcs_state *st, int kind, const double *A, const double *B, int64_t S, int mcsteps,
                             float temp, int variant /* PIQMC: global_moves; SVMC: tf */, uint64_t seed,
                             uint64_t replica_offset, uint64_t sweep_offset);
int mcs_launch_cluster_moves(mcs_state *st, double coef_a, double coef_b, double temp, const double *lookuptable,
                             int nmoves, uint64_t seed, uint64_t replica_offset, uint64_t sweep_offset);
bool mcs_dense_supported(const mcs_instance *inst, int P);
int mcs_launch_dense_sweeps(mcs_state *st, int kind, const double *A, const double *B, int64_t S, int mcsteps,
                            float temp, int global_moves, uint64_t seed, uint64_t replica_offset,
                            uint64_t sweep_offset);
int mcs_launch_sa_sweeps(mcs_state *st, const double *sched, int64_t S, int mcsteps, uint64_t seed,
                         uint64_t replica_offset, uint64_t sweep_offset);
int mcs_launch_svmc_sweeps(mcs_state *st, const double *A, const double *B, int64_t S, int mcsteps, float temp,
                           int tf, uint64_t seed, uint64_t replica_offset, uint64_t sweep_offset);
