// mcs_piqmc.cu -- path-integral QMC sweeps on bit-packed world lines (sm_100a).
//
// Replaces the loop nest of qmc.QuantumAnneal (reference qmc.pyx:93-143) and
// qmc.QuantumAnnealGlobal (qmc.pyx:358-438).
//
// Data layout in HBM:  W[site][replica] : uint64, bit k = Trotter slice k of that site's world
// line in that replica, bit set <=> spin -1.  Replica is the fastest axis, so the 32 lanes of a
// warp (= 32 consecutive replicas of ONE site) read and write 256 contiguous bytes, every
// coupling of the site is warp-uniform, and the Trotter neighbours of slice k are bits k+-1 of
// the same register (a rotate).
//
// One launch = one colour class of one sweep.  A warp owns (site, 32 replicas); each lane
//   1. XORs its word with its <= 6 neighbour words: plane j bit k = "slice k is anti-aligned with
//      neighbour j" (+ the word itself as the field plane, + two rotated XORs for slices k-1, k+1);
//   2. the warp has precomputed, for its site and this schedule step, the Metropolis acceptance
//      threshold of every one of the 2^(planes) sign patterns into shared memory (all lanes of a
//      warp share it -- that is why replicas, not sites, sit on the lanes);
//   3. even slices, then odd slices (then slice P-1 alone when P is odd -- the ring is not
//      2-colourable): the planes are transposed into one byte-wide pattern index per slice, one LDS
//      fetches the threshold, one Philox4x32-10 call decides eight slices (lazily refined uniforms).
// In-plane neighbours belong to other colour classes and are frozen during the launch, so every
// attempt sees exactly the state a sequential sweep would (detailed balance per attempt).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "mcs_common.cuh"

namespace {

constexpr int kWarps = 4; // warps per CTA

struct PiqmcPass {
    uint64_t *W;
    const int32_t *ell_idx;
    const float *ell_J;
    const float *h;
    const int32_t *sites; // the colour class
    int nsites;
    int dpad;
    int nq;      // quadratic planes in use (= maxdeg)
    int field;   // 1 if the field plane is in use
    int G;       // warps per site = Rpad / 32
    long long Rpad;
    int P;
    float bcoef;       // -2 B         (qmc.pyx:96)
    float jperp2;      // 2 J_perp     (qmc.pyx:95,137-138)
    float nl2e_over_t; // -log2(e)/teff
    mcs_philox_keys keys; // ten Philox round keys, read straight from the parameter constant bank
    mcs_pow2_table pow2;  // 2^0 .. 2^31 (multipliers of the plane transposition, see phase())
    uint32_t sweep_lo, sweep_hi;
    uint32_t replica_offset;
    int global_moves;
    uint32_t tie_thr; // 0x20000; 0xffffffff evaluates the refinement call for every attempt (test hook)
    long long half;   // fused mode (P <= 32, two replicas per thread): replica r is paired with r + half
    // packed mode (even P <= 20): a thread owns `pk` replicas as consecutive P-bit segments of its working word
    // (pk P <= 64 bits: P = 20 fills 60 of them).  Replicas are grouped on their GLOBAL index: block b = 32 pk
    // consecutive replicas, lane l of the block's warp owns members b 32 pk + 32 m + l, m = 0 .. pk - 1 (every
    // member load of a warp is one coalesced 256-byte line)
    int pk;                // 0: off
    long long group0;      // global index of the block warp 0 owns
    long long nvalid;      // replicas of this window that exist (members outside [0, nvalid) are skipped)
    uint64_t seg_lsb;      // bit 0 of every segment
    long long gw_lo, gw_n; // packed mode: this launch covers the group warps [gw_lo, gw_lo + gw_n) of the window
    // MODE_PACKN: the packed working words themselves, [N][gp] (gp = 32 x group warps of the window), built once per
    // sweep call from W and unpacked at its end: one 64-bit load per row instead of pk guarded loads and shifts
    uint64_t *Wp;
    long long gp;
    int chunks;       // launches of this colour pass in flight together (replica chunks on several streams)
    int wpt;          // plain mode: words per thread ...
    long long wstep;  // ... replica r0 + k wstep is the thread's k-th word (wstep = replicas of the launch / wpt)
};

// float(v) for |v| <= 64 without the conversion unit (I2F runs at a quarter of the FMA rate): 2^23 + 64 + v is an
// integer below 2^24, so its float bit pattern is 0x4B000000 + 64 + v and one exact subtraction gives v
__device__ __forceinline__ float small_int_to_float(int v)
{
    return __int_as_float(0x4B000040 + v) - 8388672.0f;
}

__device__ __forceinline__ uint64_t rotl_ring(uint64_t w, int P, uint64_t mask)
{
    return ((w << 1) | (w >> (P - 1))) & mask; // bit k <- bit k-1 (ring of P slices)
}
__device__ __forceinline__ uint64_t rotr_ring(uint64_t w, int P, uint64_t mask)
{
    return ((w >> 1) | (w << (P - 1))) & mask; // bit k <- bit k+1
}

__device__ __forceinline__ uint64_t b0_shift(int P) { return 0x0000000100000001ull << (P - 1); }

// Fused mode: two world lines of P <= 32 slices in one 64-bit word (replica A in the low, replica B in the high
// half); the Trotter ring closes inside each half.
__device__ __forceinline__ uint64_t rotl_ring2(uint64_t w, int P, uint64_t pm2)
{
    const uint64_t b0 = 0x0000000100000001ull;
    return ((w << 1) & pm2 & ~b0) | ((w >> (P - 1)) & b0);
}
__device__ __forceinline__ uint64_t rotr_ring2(uint64_t w, int P, uint64_t pm2)
{
    const uint64_t bp = b0_shift(P);
    return ((w >> 1) & pm2 & ~bp) | ((w << (P - 1)) & bp);
}

// Packed mode: pk world lines of P slices each at bit offsets 0, P, 2 P, ...; the ring closes inside each segment
// (lsb = bit 0 of every segment, msb = lsb << (P - 1)).
__device__ __forceinline__ uint64_t rotl_seg(uint64_t w, int P, uint64_t lsb)
{
    return ((w << 1) & ~lsb) | ((w >> (P - 1)) & lsb);
}
__device__ __forceinline__ uint64_t rotr_seg(uint64_t w, int P, uint64_t lsb)
{
    const uint64_t msb = lsb << (P - 1);
    return ((w >> 1) & ~msb) | ((w << (P - 1)) & msb);
}

// how the working word of a thread is made up
enum { MODE_PLAIN = 0, MODE_FUSE = 1, MODE_PACK = 2, MODE_PACKN = 3 }; // PACKN: packed words resident in HBM

// Pattern-index bits live at bit positions SH .. SH+NPL+1 of a field of an index word: one BYTE per slice up to
// 8 planes, one HALF WORD per slice for 9 and 10 planes (FW = field width).  Where it fits (up to 6 planes, and
// always in half words) the index is stored pre-multiplied by 4 (SH = 2), so that the extracted field IS the
// shared-memory byte offset of the threshold.
template <int NPL>
struct LutGeom {
    static constexpr int NPP = NPL + 2;          // in-plane planes + the two Trotter planes
    static constexpr int ENT = 1 << NPP;         // sign patterns
    static constexpr int FW = NPP <= 8 ? 8 : 16;
    static constexpr int SH = (NPP <= 6 || FW == 16) ? 2 : 0;
    static constexpr int NPAIR = (NPP + 1) / 2;
};

// Within one Trotter-parity phase only every other bit of a plane is used, so two planes are interleaved into
// one word first (plane 2j lowered to / kept at the even bit, plane 2j+1 one above it): the per-group
// transposition then moves TWO index bits with one shift + one LOP3.
//   PARITY 0 (slices at even bits k): m = (a & 0x5555...) | ((b << 1) & 0xAAAA...)   a at k, b at k+1
//   PARITY 1 (slices at odd  bits k): m = ((a >> 1) & 0x5555...) | (b & 0xAAAA...)   a at k-1, b at k
// Both stay inside the slice's own byte (k+1 <= 7 for even k, k-1 >= 0 for odd k).
template <int PARITY>
__device__ __forceinline__ uint32_t interleave_pair(uint32_t a, uint32_t b, const mcs_pow2_table &pow2)
{
    if (PARITY == 0) return (a & 0x55555555u) | (mcs_plane_shift<1>(b, pow2) & 0xAAAAAAAAu);
    return (mcs_plane_shift<-1>(a, pow2) & 0x55555555u) | (b & 0xAAAAAAAAu);
}

// Index word of group G (slices 8 i + 7 - G of this 32-bit half, i = 0..3; 7 - G has the phase's parity):
// byte i = pattern index of slice 8 i + 7 - G (times 4 if SH == 2).  m[j] = interleaved planes 2j, 2j+1; a
// trailing single plane is m[NPAIR-1] as is.
template <int NPP, int G, int PARITY>
__device__ __forceinline__ uint32_t gather_index(const uint32_t (&m)[(NPP + 1) / 2], const mcs_pow2_table &pow2)
{
    constexpr int SH = (NPP <= 6) ? 2 : 0;
    constexpr int S = 7 - G;                     // bit of the slice inside its byte
    constexpr int LOW = PARITY == 0 ? S : S - 1; // where plane 2j of a pair sits
    uint32_t acc = 0;
#define MCS_PAIRWORD(j)                                                                                       \
    if (2 * (j) + 1 < NPP)                                                                                    \
        acc |= mcs_plane_shift<SH + 2 * (j) - LOW>(m[(j) < (NPP + 1) / 2 ? (j) : 0], pow2) & (0x03030303u << (SH + 2 * (j))); \
    else if (2 * (j) < NPP)                                                                                   \
        acc |= mcs_plane_shift<SH + 2 * (j) - S>(m[(j) < (NPP + 1) / 2 ? (j) : 0], pow2) & (0x01010101u << (SH + 2 * (j)));
    MCS_PAIRWORD(0) MCS_PAIRWORD(1) MCS_PAIRWORD(2) MCS_PAIRWORD(3)
#undef MCS_PAIRWORD
    return acc;
}

// Half-word version (9, 10 planes): field i = pattern index * 4 of slice 16 i + S of this 32-bit half (i = 0, 1).
template <int NPP, int S, int PARITY>
__device__ __forceinline__ uint32_t gather_index16(const uint32_t (&m)[(NPP + 1) / 2], const mcs_pow2_table &pow2)
{
    constexpr int SH = 2;
    constexpr int LOW = PARITY == 0 ? S : S - 1; // where plane 2j of a pair sits
    uint32_t acc = 0;
#define MCS_PAIRWORD(j)                                                                                       \
    if (2 * (j) + 1 < NPP)                                                                                    \
        acc |= mcs_plane_shift<SH + 2 * (j) - LOW>(m[(j) < (NPP + 1) / 2 ? (j) : 0], pow2) & (0x00030003u << (SH + 2 * (j))); \
    else if (2 * (j) < NPP)                                                                                   \
        acc |= mcs_plane_shift<SH + 2 * (j) - S>(m[(j) < (NPP + 1) / 2 ? (j) : 0], pow2) & (0x00010001u << (SH + 2 * (j)));
    MCS_PAIRWORD(0) MCS_PAIRWORD(1) MCS_PAIRWORD(2) MCS_PAIRWORD(3) MCS_PAIRWORD(4)
#undef MCS_PAIRWORD
    return acc;
}

// One call of the half-word version: slices S0 > S1 > S2 > S3 (same parity) of both fields of half HALF;
// S0, S1 are the "A" groups, S2, S3 the "B" groups of mcs_decide_call16.
template <int NPL, int S0, int PARITY>
__device__ __forceinline__ void gather_call16(uint32_t (&acc)[4], const uint32_t (&m)[LutGeom<NPL>::NPAIR],
                                              const mcs_pow2_table &pow2)
{
    constexpr int NPP = LutGeom<NPL>::NPP;
    acc[0] = gather_index16<NPP, S0, PARITY>(m, pow2);
    acc[1] = gather_index16<NPP, S0 - 2, PARITY>(m, pow2);
    acc[2] = gather_index16<NPP, S0 - 4, PARITY>(m, pow2);
    acc[3] = gather_index16<NPP, S0 - 6, PARITY>(m, pow2);
}

template <int NPL, int S0, int PARITY>
__device__ __forceinline__ void decide_call16(uint32_t &rej, uint32_t &flags, const uint32_t (&m)[LutGeom<NPL>::NPAIR],
                                              const uint32_t *lut, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              const mcs_philox_keys &keys, const mcs_pow2_table &pow2,
                                              uint32_t tie_thr, uint4 *slot)
{
    uint32_t acc[4], ch[4];
    gather_call16<NPL, S0, PARITY>(acc, m, pow2);
    mcs_decide_call16(ch, flags, acc, lut, c0, c1, c2, c3, keys, pow2, tie_thr, slot);
    rej = ch[0] * pow2.up[S0] + rej;
    rej = ch[1] * pow2.up[S0 - 2] + rej;
    rej = ch[2] * pow2.up[S0 - 4] + rej;
    rej = ch[3] * pow2.up[S0 - 6] + rej;
}

template <int NPL, int S0, int PARITY>
__device__ __forceinline__ void refine_call16(uint32_t &rej, const uint32_t (&m)[LutGeom<NPL>::NPAIR],
                                              const uint32_t *lut, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              const mcs_philox_keys &keys, const mcs_pow2_table &pow2)
{
    uint32_t acc[4];
    gather_call16<NPL, S0, PARITY>(acc, m, pow2);
    const uint4 ch = mcs_refine_call16(acc[0], acc[1], acc[2], acc[3], lut, c0, c1, c2, c3, keys.rk[0], keys.rk[1]);
    const uint32_t mask = (0x00010001u << S0) | (0x00010001u << (S0 - 2)) | (0x00010001u << (S0 - 4)) |
                          (0x00010001u << (S0 - 6));
    rej = (rej & ~mask) | (ch.x << S0) | (ch.y << (S0 - 2)) | (ch.z << (S0 - 4)) | (ch.w << (S0 - 6));
}

// Slow path (see mcs_common.cuh, "lazily refined uniforms"): redo one flagged call (groups GA, GB of one half)
// with both Philox halves and replace its eight reject bits.  Runs at the end of the phase.
template <int NPL, int GA, int GB, int PARITY>
__device__ __forceinline__ void refine_pair(uint32_t &rej, const uint32_t (&m)[LutGeom<NPL>::NPAIR],
                                            const uint32_t *lut, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                            const mcs_philox_keys &keys, const mcs_pow2_table &pow2)
{
    constexpr int NPP = LutGeom<NPL>::NPP;
    const uint32_t accA = gather_index<NPP, GA, PARITY>(m, pow2);
    const uint32_t accB = gather_index<NPP, GB, PARITY>(m, pow2);
    const uint2 ch = mcs_refine_call<LutGeom<NPL>::SH>(accA, accB, lut, c0, c1, c2, c3, keys.rk[0], keys.rk[1]);
    rej = (rej & ~((0x01010101u << (7 - GA)) | (0x01010101u << (7 - GB)))) | (ch.x << (7 - GA)) | (ch.y << (7 - GB));
}

// groups GA and GB (same half, same parity) share one Philox call; rej accumulates REJECT bits
template <int NPL, int GA, int GB>
__device__ __forceinline__ void decide_pair(uint32_t &rej, uint32_t &flags, uint32_t accA, uint32_t accB,
                                            const uint32_t *lut, uint32_t c0, uint32_t c1, uint32_t c2,
                                            uint32_t c3, const mcs_philox_keys &keys, const mcs_pow2_table &pow2,
                                            uint32_t tie_thr, uint2 *slot)
{
    uint32_t chA, chB;
    mcs_decide_call<LutGeom<NPL>::SH>(chA, chB, flags, accA, accB, lut, c0, c1, c2, c3, keys, pow2, tie_thr, slot);
    rej = chA * pow2.up[7 - GA] + rej;
    rej = chB * pow2.up[7 - GB] + rej;
}

// One Trotter-parity phase of a word: attempts every slice k with k % 2 == PARITY that is in `allowed`,
// against the (complemented) thresholds in lut[].  Returns the flip mask.
// FUSE: the two 32-bit halves are two replicas (counters c0h[0], c0h[1]) of P <= 32 slices each; every half then
// uses the tags and the skip rule of half 0, so a replica gets exactly the decisions it would get alone.
// PACK: pk replicas of one group in consecutive P-bit segments, ONE counter (the global group index): the word is
// treated like a single world line of pk P slices (tags and skip rule of the plain mode with P -> `bits`), only
// the Trotter ring closes per segment.
template <int NPL, int PARITY, bool FULL, int MODE>
__device__ __forceinline__ uint64_t phase(const uint64_t (&pl)[NPL], uint64_t w, int P, uint64_t pmask,
                                          uint64_t allowed, const uint32_t *lut, const uint32_t (&c0h)[2], uint32_t c1,
                                          uint32_t c2, uint32_t c3hi, const mcs_philox_keys &keys,
                                          const mcs_pow2_table &pow2, uint32_t tie_thr, uint2 *bounce, int nthreads,
                                          uint64_t seg_lsb, int bits)
{
    constexpr int NPP = LutGeom<NPL>::NPP, NPAIR = LutGeom<NPL>::NPAIR;
    constexpr bool FUSE = MODE == MODE_FUSE;
    // bit k: slice k anti-aligned with slice k-1 / k+1
    const uint64_t tl = w ^ (MODE >= MODE_PACK ? rotl_seg(w, P, seg_lsb)
                                               : (FUSE ? rotl_ring2(w, P, pmask) : rotl_ring(w, P, pmask)));
    const uint64_t tr = w ^ (MODE >= MODE_PACK ? rotr_seg(w, P, seg_lsb)
                                               : (FUSE ? rotr_ring2(w, P, pmask) : rotr_ring(w, P, pmask)));
    const uint32_t c0 = c0h[0];
    uint32_t m[2][NPAIR];
#pragma unroll
    for (int H = 0; H < 2; ++H)
#pragma unroll
        for (int j = 0; j < NPAIR; ++j) {
            const int pa = 2 * j, pb = 2 * j + 1;
            const uint64_t A = pa < NPL ? pl[pa < NPL ? pa : 0] : (pa == NPL ? tl : tr);
            const uint64_t B = pb < NPL ? pl[pb < NPL ? pb : 0] : (pb == NPL ? tl : tr);
            const uint32_t a = (uint32_t)(A >> (32 * H)), b = (uint32_t)(B >> (32 * H));
            m[H][j] = pb < NPP ? interleave_pair<PARITY>(a, b, pow2) : a;
        }
    uint32_t rej[2] = {0u, 0u}, flags = 0u;
    if (LutGeom<NPL>::FW == 16) {
        // half-word index fields: call q of half H covers slices 32 H + 16 i + s, s = S0, S0-2, S0-4, S0-6,
        // S0 = 14 + PARITY (q = 0) or 6 + PARITY (q = 1); tag = 8 H + 4 PARITY + q
        uint4 *slot16 = reinterpret_cast<uint4 *>(bounce);
#define MCS_CALL16(H, Q, S0)                                                                                 \
    if (FULL || 32 * H + (S0) - 6 < bits) {                                                                  \
        decide_call16<NPL, (S0), PARITY>(rej[H], flags, m[H], lut, c0, c1, c2,                               \
                                         c3hi | (uint32_t)(8 * H + 4 * PARITY + (Q)), keys, pow2, tie_thr,   \
                                         slot16 + (2 * H + (Q)) * nthreads);                                 \
    } else {                                                                                                 \
        flags *= 2u;                                                                                         \
    }
        MCS_CALL16(0, 0, 14 + PARITY) MCS_CALL16(0, 1, 6 + PARITY)
        MCS_CALL16(1, 0, 14 + PARITY) MCS_CALL16(1, 1, 6 + PARITY)
#undef MCS_CALL16
        if (flags) {
#define MCS_REFINE16(H, Q, S0, BIT)                                                                          \
    if (flags & (BIT))                                                                                       \
        refine_call16<NPL, (S0), PARITY>(rej[H], m[H], lut, c0, c1, c2,                                      \
                                         c3hi | (uint32_t)(8 * H + 4 * PARITY + (Q)), keys, pow2);
            MCS_REFINE16(0, 0, 14 + PARITY, 8u) MCS_REFINE16(0, 1, 6 + PARITY, 4u)
            MCS_REFINE16(1, 0, 14 + PARITY, 2u) MCS_REFINE16(1, 1, 6 + PARITY, 1u)
#undef MCS_REFINE16
        }
        return ~(((uint64_t)rej[1] << 32) | rej[0]) & allowed;
    }
    // group G of half H holds slices 32 H + 8 i + 7 - G; parity of 7 - G == PARITY  <=>  G = 1-PARITY, 3-PARITY, ...
    // pair q = 2 H + (GA > 3) is flags bit 3 - q after the four Horner steps; a pair entirely beyond the last
    // slice is skipped (warp-uniform branch) but still shifts the flags
#define MCS_PAIR(H, GA, GB)                                                                                  \
    if (FULL || (FUSE ? 0 : 32 * H) + 7 - (GB) < bits) {                                                     \
        const uint32_t accA = gather_index<NPP, (GA), PARITY>(m[H], pow2);                                   \
        const uint32_t accB = gather_index<NPP, (GB), PARITY>(m[H], pow2);                                   \
        decide_pair<NPL, (GA), (GB)>(rej[H], flags, accA, accB, lut, c0h[FUSE ? H : 0], c1, c2,              \
                                     c3hi | (uint32_t)((FUSE ? 0 : H) * 8 + (GA)), keys, pow2, tie_thr,      \
                                     bounce + (2 * H + (GA) / 4) * nthreads);                                \
    } else {                                                                                                 \
        flags *= 2u;                                                                                         \
    }
    MCS_PAIR(0, 1 - PARITY, 3 - PARITY) MCS_PAIR(0, 5 - PARITY, 7 - PARITY)
    MCS_PAIR(1, 1 - PARITY, 3 - PARITY) MCS_PAIR(1, 5 - PARITY, 7 - PARITY)
#undef MCS_PAIR
    if (flags) { // rare (about 3 % of the warp-phases): Horner order, pair 0 ended at bit 3 ... pair 3 at bit 0
#define MCS_REFINE(H, GA, GB, BIT)                                                                           \
    if (flags & (BIT))                                                                                       \
        refine_pair<NPL, (GA), (GB), PARITY>(rej[H], m[H], lut, c0h[FUSE ? H : 0], c1, c2,                   \
                                             c3hi | (uint32_t)((FUSE ? 0 : H) * 8 + (GA)), keys, pow2);
        MCS_REFINE(0, 1 - PARITY, 3 - PARITY, 8u) MCS_REFINE(0, 5 - PARITY, 7 - PARITY, 4u)
        MCS_REFINE(1, 1 - PARITY, 3 - PARITY, 2u) MCS_REFINE(1, 5 - PARITY, 7 - PARITY, 1u)
#undef MCS_REFINE
    }
    return ~(((uint64_t)rej[1] << 32) | rej[0]) & allowed;
}

// WARPS warps per CTA, all working on the SAME site (WARPS*32 consecutive replicas), so the threshold table
// is built once per CTA at a compile-time shared-memory address.  grid = (CTAs per site, sites of the colour).
// FULL: P == 64 (every group exists: no per-pair branch, one basic block).
// FLD:  the instance has (1) / has no (0) field plane; the in-plane planes are then j < NPL - FLD, all
//       compile-time (rows shorter than maxdeg are padded with the site itself and J = 0: a zero plane).
// register cap: the big basic blocks otherwise tempt ptxas into 80+ registers for no gain (measured); 64 is
// spill-free up to 4 in-plane planes, the 5- and 6-plane kernels get 80, the half-word-index kernels 128
#ifndef MCS_LUT_MINBLOCKS
#define MCS_LUT_MINBLOCKS(NPL) ((NPL) >= 7 ? 4 : (NPL) >= 5 ? 6 : 8)
#endif
// MODE_FUSE: P <= 32 and the thread owns TWO replicas, r and r + a.half, as the low and the high half of one working
//       word -- a word's worth of Philox calls, transpositions and decisions then serves 2 P instead of P
//       attempts (P = 20: 6.5e11 -> 1.0e12 attempts/s).  The decisions of a replica do not depend on the mode.
// MODE_PACK: even P <= 20 and the thread owns a.pk = min(floor(64 / P), 6) replicas as consecutive P-bit segments:
//       P = 20 uses 60 of the 64 attempt slots of a word instead of 40.  Groups are defined on GLOBAL replica
//       indices (see PiqmcPass::pk) and the Philox counter is the group's index, so results do not depend on how
//       replicas are sharded over GPUs, windows or calls; members that lie outside this window are neither read
//       nor written.
#ifndef MCS_LUT_MULTI_MINBLOCKS
#define MCS_LUT_MULTI_MINBLOCKS 8
#endif
template <int NPL, int WARPS, bool FULL, int FLD, int MODE, bool MULTI = false>
__global__ void __launch_bounds__(WARPS * 32, (MULTI && WARPS == 1) ? MCS_LUT_MULTI_MINBLOCKS : MCS_LUT_MINBLOCKS(NPL)) piqmc_lut_pass_kernel(const __grid_constant__ PiqmcPass a)
{
    static_assert(!(FULL && MODE != MODE_PLAIN), "fused / packed modes are for P <= 32");
    constexpr bool FUSE = MODE == MODE_FUSE, PACK = MODE >= MODE_PACK, NATIVE = MODE == MODE_PACKN;
    constexpr int ENT = LutGeom<NPL>::ENT, NQ = NPL - FLD;
    __shared__ uint32_t s_lut[ENT];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    mcs_pdl_launch_dependents();
    const unsigned si = blockIdx.y + 65535u * blockIdx.z;
    if (si >= (unsigned)a.nsites) return; // only when the colour class has more than 65535 sites (CTA-uniform)
    const int site = __ldg(&a.sites[si]);
    const long long r0 = ((long long)blockIdx.x * WARPS + warp) * 32 + lane; // first replica (PACK: group) of this thread

    // ---- per-site coefficients (CTA-uniform) and the acceptance-threshold table ---------------
    float c[NPL];
    int nb[NPL];
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
        if (j < NQ) {
            nb[j] = __ldg(&a.ell_idx[(uint64_t)(uint32_t)site * (uint32_t)a.dpad + j]);
            c[j] = a.bcoef * __ldg(&a.ell_J[(uint64_t)(uint32_t)site * (uint32_t)a.dpad + j]);
        } else {
            nb[j] = site;
            c[j] = a.bcoef * __ldg(&a.h[site]);
        }
    }
    for (int e = threadIdx.x; e < ENT; e += WARPS * 32) {
        float dE = 0.0f;
#pragma unroll
        for (int j = 0; j < NPL; ++j) dE += ((e >> j) & 1) ? -c[j] : c[j];
        const int anti = ((e >> NPL) & 1) + ((e >> (NPL + 1)) & 1); // anti-aligned Trotter neighbours
        dE += a.jperp2 * (float)(2 - 2 * anti);
        s_lut[e] = ~mcs_accept_threshold(dE, a.nl2e_over_t);
    }

    // ---- this lane's world line(s) and the in-plane anti-alignment planes ----------------------
    const int P = FULL ? 64 : a.P;
    const int pk = PACK ? a.pk : 1, bits = PACK ? pk * P : P;
    const uint64_t pm1 = (FULL || P == 64) ? ~0ull : ((1ull << P) - 1ull); // one world line
    const uint64_t pmask = FUSE ? (pm1 | (pm1 << 32)) : (PACK ? (bits == 64 ? ~0ull : ((1ull << bits) - 1ull)) : pm1);
    // row offsets as one IMAD.WIDE.U32 each (site indices and Rpad are below 2^32)
    const uint32_t rpad = (uint32_t)a.Rpad;
    // PACK: member m of this thread is local replica first + 32 m (global index (group0 + warp) 32 pk + 32 m + lane)
    const long long gwarp0 = (PACK ? a.gw_lo : 0) + (long long)blockIdx.x * WARPS + warp;
    const bool gw_ok0 = !PACK || gwarp0 < a.gw_lo + a.gw_n; // the last CTA of a packed launch may have spare warps
    const long long first0 = PACK ? (a.group0 + gwarp0) * 32 * pk + lane - (long long)a.replica_offset : r0;
    const uint32_t gp = NATIVE ? (uint32_t)a.gp : 0u;
    uint32_t present = 0; // PACK: members that exist in this window
    if (PACK && !NATIVE && gw_ok0)
        for (int m = 0; m < pk; ++m)
            if (first0 + 32 * m >= 0 && first0 + 32 * m < a.nvalid) present |= 1u << m;
    mcs_pdl_wait(); // everything above depends on the instance and the schedule only
    // Plain mode: a thread takes a.wpt words (replicas r0, r0 + wstep, ...) one after the other, so that the site's
    // coefficients, neighbour indices and threshold table -- a quarter of the instructions of a one-word thread --
    // are set up once for all of them.
    // (MULTI is a template parameter: the one-word kernel keeps its straight-line code)
    // Resident packed words (MODE_PACKN) likewise: the thread's k-th word belongs to group warp gwarp0 + k wstep.
    int wpt = ((MODE == MODE_PLAIN || NATIVE) && MULTI) ? a.wpt : 1;
    // a short last slab: fewer words for its threads (warp-uniform; decided here, not by a test inside the loop)
    if (MODE == MODE_PLAIN && MULTI) wpt = (int)min((long long)wpt, ((long long)a.G * 32 - r0 + a.wstep - 1) / a.wstep);
    for (int kw = 0; kw < wpt; ++kw) {
    const long long r = r0 + ((MODE == MODE_PLAIN && MULTI) ? kw * a.wstep : 0), first = PACK ? first0 : r;
    const long long gwarp = gwarp0 + ((NATIVE && MULTI) ? kw * a.wstep : 0);
    const bool gw_ok = (NATIVE && MULTI) ? gwarp < a.gw_lo + a.gw_n : gw_ok0;
    if (NATIVE && MULTI && !gw_ok) break; // the last slab may be short (warp-uniform; never the first word)
    uint64_t *Wn = NATIVE ? a.Wp + (gw_ok ? gwarp * 32 + lane : 0) : nullptr;
    const uint64_t *Wr = a.W + first, *Wr2 = Wr + (FUSE ? a.half : 0);
    auto load = [&](int row) -> uint64_t {
        if (NATIVE) return gw_ok ? Wn[(uint64_t)(uint32_t)row * gp] : 0ull;
        if (PACK) {
            uint64_t v = 0;
#pragma unroll
            for (int m = 0; m < 6; ++m)
                if ((present >> m) & 1u) v |= Wr[(uint64_t)(uint32_t)row * rpad + 32 * m] << (m * P);
            return v;
        }
        const uint64_t lo = Wr[(uint64_t)(uint32_t)row * rpad];
        return FUSE ? (lo | (Wr2[(uint64_t)(uint32_t)row * rpad] << 32)) : lo;
    };
    uint64_t w = load(site);
    const uint64_t w0 = w;
    uint64_t pl[NPL];
#pragma unroll
    for (int j = 0; j < NPL; ++j) pl[j] = j < NQ ? (w ^ load(nb[j])) & pmask : w; // field plane: bit set <=> s = -1
    if (kw == 0) { // the table is published once; the first word's loads were issued before the barrier
        if (WARPS == 1)
            __syncwarp();
        else
            __syncthreads();
    }
    const uint32_t *lut = s_lut;
    const uint32_t c0h[2] = {PACK ? (uint32_t)((a.group0 + gwarp) * 32 + lane) : a.replica_offset + (uint32_t)r,
                             a.replica_offset + (uint32_t)(r + a.half)};
    const uint32_t c1 = (uint32_t)site, c2 = a.sweep_lo;
    const uint32_t c3hi = a.sweep_hi << 8;
    const bool oddP = !FULL && (P & 1) != 0;
    // slice P-1 of every world line in the word
    const uint64_t last = FUSE ? b0_shift(P) : (PACK ? a.seg_lsb << (P - 1) : (1ull << (P - 1)));
    uint64_t even_allowed = 0x5555555555555555ull & pmask, odd_allowed = 0xAAAAAAAAAAAAAAAAull & pmask;
    // An odd ring is not 2-colourable: slice P-1 neighbours slice 0 and is handled alone below.  In a packed word the
    // segments start at bits of either parity (a phase then takes the even slices of one segment and the odd slices
    // of the next: alternate slices of every ring, which is all that matters), so both masks lose the last slices.
    if (oddP) even_allowed &= ~last, odd_allowed &= ~last;

    // [phase][call][thread]: private slots for the index fields (8 bytes per call, 16 with half-word fields)
    __shared__ __align__(16) uint2 s_bounce[(LutGeom<NPL>::FW == 16 ? 16 : 8) * WARPS * 32];
    constexpr int kSlot = LutGeom<NPL>::FW == 16 ? 2 : 1; // uint2 per thread and call
    uint2 *bounce = s_bounce + kSlot * threadIdx.x;
    w ^= phase<NPL, 0, FULL, MODE>(pl, w, P, pmask, even_allowed, lut, c0h, c1, c2, c3hi, a.keys, a.pow2, a.tie_thr,
                                   bounce, WARPS * 32, a.seg_lsb, bits);
    w ^= phase<NPL, 1, FULL, MODE>(pl, w, P, pmask, odd_allowed, lut, c0h, c1, c2, c3hi, a.keys, a.pow2, a.tie_thr,
                                   bounce + 4 * kSlot * WARPS * 32, WARPS * 32, a.seg_lsb, bits);
    if (oddP && PACK) { // the closing slices of the pk rings: independent of each other, one Philox call per four
        const uint64_t tl = w ^ rotl_seg(w, P, a.seg_lsb), tr = w ^ rotr_seg(w, P, a.seg_lsb);
        uint32_t rnd[4];
        uint64_t fl = 0;
        for (int m = 0; m < pk; ++m) {
            if ((m & 3) == 0)
                mcs_philox4x32_rk(c0h[0], c1, c2, c3hi | (m == 0 ? MCS_TAG_LAST_SLICE : MCS_TAG_LAST_SLICE2), a.keys, rnd);
            const int k = m * P + P - 1;
            uint32_t idx = 0;
#pragma unroll
            for (int j = 0; j < NPL; ++j) idx |= (uint32_t)((pl[j] >> k) & 1ull) << j;
            idx |= (uint32_t)((tl >> k) & 1ull) << NPL;
            idx |= (uint32_t)((tr >> k) & 1ull) << (NPL + 1);
            const uint32_t u = (m & 3) == 0 ? rnd[0] : (m & 3) == 1 ? rnd[1] : (m & 3) == 2 ? rnd[2] : rnd[3];
            if (mcs_accepts(u, ~lut[idx])) fl |= 1ull << k;
        }
        w ^= fl;
    } else if (oddP) {
        const uint64_t tl = w ^ (FUSE ? rotl_ring2(w, P, pmask) : rotl_ring(w, P, pmask));
        const uint64_t tr = w ^ (FUSE ? rotr_ring2(w, P, pmask) : rotr_ring(w, P, pmask));
#pragma unroll
        for (int hh = 0; hh < (FUSE ? 2 : 1); ++hh) {
            const int k = P - 1 + 32 * hh;
            uint32_t idx = 0;
#pragma unroll
            for (int j = 0; j < NPL; ++j) idx |= (uint32_t)((pl[j] >> k) & 1ull) << j;
            idx |= (uint32_t)((tl >> k) & 1ull) << NPL;
            idx |= (uint32_t)((tr >> k) & 1ull) << (NPL + 1);
            uint32_t rnd[4];
            mcs_philox4x32_rk(c0h[hh], c1, c2, c3hi | MCS_TAG_LAST_SLICE, a.keys, rnd);
            if (mcs_accepts(rnd[0], ~lut[idx])) w ^= 1ull << k;
        }
    }

    // ---- world-line move: flip all P slices of this site (qmc.pyx:405-438) ---------------------
    // The neighbours belong to other colour classes and have not moved during this launch: their anti-alignment
    // with the UPDATED word is the old plane XOR this pass's flips (no second read of the neighbour words).
    if (a.global_moves) {
        const uint64_t flips = w ^ w0;
        if (PACK) { // one decision per member; one Philox call serves four members
            uint32_t rnd[4];
            for (int m = 0; m < pk; ++m) {
                if ((m & 3) == 0)
                    mcs_philox4x32_rk(c0h[0], c1, c2, c3hi | (m == 0 ? MCS_TAG_GLOBAL : MCS_TAG_GLOBAL2), a.keys, rnd);
                const uint64_t seg = pm1 << (m * P);
                float dE = 0.0f;
#pragma unroll
                for (int j = 0; j < NPL; ++j)
                    dE += c[j] * small_int_to_float(P - 2 * __popcll((j < NQ ? pl[j] ^ flips : w) & seg));
                const uint32_t u = (m & 3) == 0 ? rnd[0] : (m & 3) == 1 ? rnd[1] : (m & 3) == 2 ? rnd[2] : rnd[3];
                if (mcs_accepts(u, mcs_accept_threshold(dE, a.nl2e_over_t))) w ^= seg;
            }
        } else {
            float dE[2] = {0.0f, 0.0f};
#pragma unroll
            for (int j = 0; j < NPL; ++j) {
                const uint64_t x = (j < NQ ? pl[j] ^ flips : w) & pmask;
                if (FUSE) {
                    dE[0] += c[j] * small_int_to_float(P - 2 * __popc((uint32_t)x));
                    dE[1] += c[j] * small_int_to_float(P - 2 * __popc((uint32_t)(x >> 32)));
                } else {
                    dE[0] += c[j] * small_int_to_float(P - 2 * __popcll(x));
                }
            }
#pragma unroll
            for (int hh = 0; hh < (FUSE ? 2 : 1); ++hh) {
                uint32_t rnd[4];
                mcs_philox4x32_rk(c0h[hh], c1, c2, c3hi | MCS_TAG_GLOBAL, a.keys, rnd);
                if (mcs_accepts(rnd[0], mcs_accept_threshold(dE[hh], a.nl2e_over_t))) w ^= FUSE ? (pm1 << (32 * hh)) : pm1;
            }
        }
    }
    if (NATIVE) {
        if (gw_ok) Wn[(uint64_t)(uint32_t)site * gp] = w;
    } else if (PACK) {
#pragma unroll
        for (int m = 0; m < 6; ++m)
            if ((present >> m) & 1u) a.W[(uint64_t)(uint32_t)site * rpad + first + 32 * m] = (w >> (m * P)) & pm1;
    } else if (FUSE) {
        a.W[(uint64_t)(uint32_t)site * rpad + r] = w & 0xFFFFFFFFull;
        a.W[(uint64_t)(uint32_t)site * rpad + r + a.half] = w >> 32;
    } else {
        a.W[(uint64_t)(uint32_t)site * rpad + r] = w;
    }
    } // words of this thread
}

// ------------------------------------------------------------------------------------------
// General-degree pass (any sparse graph): same layout and phases, but the energy difference of
// each slice is accumulated directly over the ELL row instead of looked up.  Used when a site has
// more sign patterns than the 1024-entry table (maxdeg + field > 8).
// ------------------------------------------------------------------------------------------
template <int PARITY>
__device__ __forceinline__ uint64_t phase_direct(const PiqmcPass &a, int site, long long r, uint64_t w, int P,
                                                 uint64_t pmask, uint64_t allowed, uint32_t c0, uint32_t c1,
                                                 uint32_t c2, uint32_t c3hi)
{
    const uint64_t tl = w ^ rotl_ring(w, P, pmask);
    const uint64_t tr = w ^ rotr_ring(w, P, pmask);
    const float hc = a.field ? a.bcoef * __ldg(&a.h[site]) : 0.0f;
    float e[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) {
        const int k = 2 * q + PARITY;
        const int anti = (int)((tl >> k) & 1ull) + (int)((tr >> k) & 1ull);
        e[q] = a.jperp2 * (float)(2 - 2 * anti) + (((w >> k) & 1ull) ? -hc : hc);
    }
    for (int j = 0; j < a.dpad; ++j) {
        const float cj = a.bcoef * __ldg(&a.ell_J[(long long)site * a.dpad + j]);
        if (cj == 0.0f) continue; // padding (warp-uniform: same site on every lane)
        const int nbj = __ldg(&a.ell_idx[(long long)site * a.dpad + j]);
        const uint64_t x = w ^ a.W[(long long)nbj * a.Rpad + r];
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            const int k = 2 * q + PARITY;
            const uint32_t sgn = (uint32_t)((x >> k) & 1ull) << 31;
            e[q] += __uint_as_float(__float_as_uint(cj) ^ sgn);
        }
    }
    uint64_t flip = 0;
#pragma unroll
    for (int q4 = 0; q4 < 8; ++q4) {
        if (8 * q4 + PARITY >= P) continue;
        uint32_t rnd[4];
        mcs_philox4x32_rk(c0, c1, c2, c3hi | (uint32_t)(PARITY * 8 + q4), a.keys, rnd);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int q = 4 * q4 + i;
            if (mcs_accepts(rnd[i], mcs_accept_threshold(e[q], a.nl2e_over_t))) flip |= 1ull << (2 * q + PARITY);
        }
    }
    return flip & allowed;
}

__global__ void __launch_bounds__(kWarps * 32) piqmc_direct_pass_kernel(const __grid_constant__ PiqmcPass a)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * kWarps + warp;
    if (item >= (long long)a.nsites * a.G) return;
    const int site = a.sites[item / a.G];
    const long long r = (item % a.G) * 32 + lane;
    const int P = a.P;
    const uint64_t pmask = P == 64 ? ~0ull : ((1ull << P) - 1ull);
    uint64_t w = a.W[(long long)site * a.Rpad + r];
    const uint32_t c0 = a.replica_offset + (uint32_t)r, c1 = (uint32_t)site, c2 = a.sweep_lo;
    const uint32_t c3hi = a.sweep_hi << 8;
    const bool oddP = (P & 1) != 0;
    uint64_t even_allowed = 0x5555555555555555ull & pmask;
    if (oddP) even_allowed &= ~(1ull << (P - 1));
    const uint64_t odd_allowed = 0xAAAAAAAAAAAAAAAAull & pmask;
    w ^= phase_direct<0>(a, site, r, w, P, pmask, even_allowed, c0, c1, c2, c3hi);
    w ^= phase_direct<1>(a, site, r, w, P, pmask, odd_allowed, c0, c1, c2, c3hi);
    const float hc = a.field ? a.bcoef * __ldg(&a.h[site]) : 0.0f;
    if (oddP) {
        const int k = P - 1;
        const uint64_t tl = w ^ rotl_ring(w, P, pmask), tr = w ^ rotr_ring(w, P, pmask);
        const int anti = (int)((tl >> k) & 1ull) + (int)((tr >> k) & 1ull);
        float dE = a.jperp2 * (float)(2 - 2 * anti) + (((w >> k) & 1ull) ? -hc : hc);
        for (int j = 0; j < a.dpad; ++j) {
            const float cj = a.bcoef * __ldg(&a.ell_J[(long long)site * a.dpad + j]);
            const int nbj = __ldg(&a.ell_idx[(long long)site * a.dpad + j]);
            const uint64_t x = w ^ a.W[(long long)nbj * a.Rpad + r];
            dE += ((x >> k) & 1ull) ? -cj : cj;
        }
        uint32_t rnd[4];
        mcs_philox4x32_rk(c0, c1, c2, c3hi | MCS_TAG_LAST_SLICE, a.keys, rnd);
        if (mcs_accepts(rnd[0], mcs_accept_threshold(dE, a.nl2e_over_t))) w ^= 1ull << k;
    }
    if (a.global_moves) {
        float dE = hc * small_int_to_float(P - 2 * __popcll(w & pmask));
        for (int j = 0; j < a.dpad; ++j) {
            const float cj = a.bcoef * __ldg(&a.ell_J[(long long)site * a.dpad + j]);
            const int nbj = __ldg(&a.ell_idx[(long long)site * a.dpad + j]);
            const uint64_t x = (w ^ a.W[(long long)nbj * a.Rpad + r]) & pmask;
            dE += cj * small_int_to_float(P - 2 * __popcll(x));
        }
        uint32_t rnd[4];
        mcs_philox4x32_rk(c0, c1, c2, c3hi | MCS_TAG_GLOBAL, a.keys, rnd);
        if (mcs_accepts(rnd[0], mcs_accept_threshold(dE, a.nl2e_over_t))) w ^= pmask;
    }
    a.W[(long long)site * a.Rpad + r] = w;
}

// ------------------------------------------------------------------------------------------
// Ohmic-bath pass (qmc.DissipativeQuantumAnneal[Global], reference qmc.pyx:223-278, 523-609).
// The bath couples slice k to EVERY other slice of the same world line,
//     dE_bath(k) = sum_{d=1}^{P-1} 2 teff (s_k s_{k+d}) lookuptable[d-1]                (qmc.pyx:268-273),
// so the slices of a word are visited one after another (no Trotter parity classes); sites of a colour
// class and replicas stay parallel.  With y = rotr_ring(w, k) ^ (s_k ? ~0 : 0), bit d of y says
// "slice k+d is anti-aligned with slice k", and the sum is C0 - sum_d 4 teff lut[d-1] y_d: a weighted
// popcount evaluated byte by byte from a shared-memory table (8 x 256 floats, built once per CTA).
// ------------------------------------------------------------------------------------------
struct BathArgs {
    const float *lut4; // [64]: 4 teff lookuptable[d-1] at index d (0 for d = 0 and d >= P)
    float c0;          // 2 teff sum_d lookuptable[d-1]
};

// NPL > 0: the site's coupling planes (degree + field <= NPL) live in registers.  NPL == 0: any degree -- the
// in-plane term of every slice walks the ELL row and re-reads the neighbours' words (L1 hits: the lines are this
// warp's own, replicas on lanes).
template <int NPL>
__global__ void __launch_bounds__(kWarps * 32) piqmc_bath_pass_kernel(const __grid_constant__ PiqmcPass a,
                                                                      const BathArgs bath)
{
    __shared__ float s_tab[8][256];
    for (int e = threadIdx.x; e < 8 * 256; e += kWarps * 32) {
        const int b = e >> 8, v = e & 255;
        float acc = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if ((v >> j) & 1) acc += __ldg(&bath.lut4[8 * b + j]);
        s_tab[b][v] = acc;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * kWarps + warp;
    if (item >= (long long)a.nsites * a.G) return;
    const int site = a.sites[item / a.G];
    const long long r = (item % a.G) * 32 + lane;
    constexpr int NREG = NPL > 0 ? NPL : 1;

    float c[NREG];
    int nb[NREG];
    const int P = a.P;
    const uint64_t pmask = P == 64 ? ~0ull : ((1ull << P) - 1ull);
    uint64_t w = a.W[(long long)site * a.Rpad + r];
    uint64_t pl[NREG];
    if (NPL > 0) {
#pragma unroll
        for (int j = 0; j < NREG; ++j) {
            if (j < a.nq) {
                nb[j] = __ldg(&a.ell_idx[(long long)site * a.dpad + j]);
                c[j] = a.bcoef * __ldg(&a.ell_J[(long long)site * a.dpad + j]);
            } else {
                nb[j] = site;
                c[j] = (a.field && j == a.nq) ? a.bcoef * __ldg(&a.h[site]) : 0.0f;
            }
        }
#pragma unroll
        for (int j = 0; j < NREG; ++j) {
            if (j < a.nq)
                pl[j] = (w ^ a.W[(long long)nb[j] * a.Rpad + r]) & pmask;
            else
                pl[j] = (a.field && j == a.nq) ? w : 0ull;
        }
    }
    const float hc = a.field ? a.bcoef * __ldg(&a.h[site]) : 0.0f; // NPL == 0 only
    const uint32_t c0 = a.replica_offset + (uint32_t)r, c1 = (uint32_t)site, c2 = a.sweep_lo;
    const uint32_t c3hi = a.sweep_hi << 8;
    const int nbytes = (P + 7) >> 3;
    uint32_t rnd[4];
    for (int k = 0; k < P; ++k) {
        if ((k & 3) == 0) mcs_philox4x32_rk(c0, c1, c2, c3hi | (uint32_t)(k >> 2), a.keys, rnd);
        const uint32_t sk = (uint32_t)(w >> k) & 1u;
        float dE = 0.0f;
        if (NPL > 0) {
#pragma unroll
            for (int j = 0; j < NREG; ++j) dE += ((pl[j] >> k) & 1ull) ? -c[j] : c[j];
        } else {
            dE = sk ? -hc : hc;
            for (int j = 0; j < a.dpad; ++j) { // padding entries: J = 0, index = the site itself
                const float cj = a.bcoef * __ldg(&a.ell_J[(long long)site * a.dpad + j]);
                const int nbj = __ldg(&a.ell_idx[(long long)site * a.dpad + j]);
                const uint32_t sj = (uint32_t)(a.W[(long long)nbj * a.Rpad + r] >> k) & 1u;
                dE += (sj ^ sk) ? -cj : cj;
            }
        }
        const int kl = k == 0 ? P - 1 : k - 1, kr = k == P - 1 ? 0 : k + 1;
        const int anti = (int)(((uint32_t)(w >> kl) & 1u) ^ sk) + (int)(((uint32_t)(w >> kr) & 1u) ^ sk);
        dE += a.jperp2 * (float)(2 - 2 * anti);
        uint64_t y = k == 0 ? w : (((w >> k) | (w << (P - k))) & pmask);
        if (sk) y ^= pmask; // bit 0 becomes 0 either way: lut4[0] == 0
        float wsum = 0.0f;
        for (int b = 0; b < nbytes; ++b) wsum += s_tab[b][(uint32_t)(y >> (8 * b)) & 255u];
        dE += bath.c0 - wsum;
        const uint32_t u = (k & 3) == 0 ? rnd[0] : (k & 3) == 1 ? rnd[1] : (k & 3) == 2 ? rnd[2] : rnd[3];
        if (mcs_accepts(u, mcs_accept_threshold(dE, a.nl2e_over_t))) w ^= 1ull << k;
    }
    if (a.global_moves) { // a world-line flip leaves every s_k s_k' invariant: no bath term (qmc.pyx:575-609)
        float dE = 0.0f;
        if (NPL > 0) {
#pragma unroll
            for (int j = 0; j < NREG; ++j) {
                uint64_t x;
                if (j < a.nq)
                    x = (w ^ a.W[(long long)nb[j] * a.Rpad + r]) & pmask;
                else
                    x = (a.field && j == a.nq) ? w : 0ull;
                dE += c[j] * small_int_to_float(P - 2 * __popcll(x));
            }
        } else {
            dE = hc * small_int_to_float(P - 2 * __popcll(w & pmask));
            for (int j = 0; j < a.dpad; ++j) {
                const float cj = a.bcoef * __ldg(&a.ell_J[(long long)site * a.dpad + j]);
                const int nbj = __ldg(&a.ell_idx[(long long)site * a.dpad + j]);
                if (nbj == site) continue; // padding
                dE += cj * small_int_to_float(P - 2 * __popcll((w ^ a.W[(long long)nbj * a.Rpad + r]) & pmask));
            }
        }
        mcs_philox4x32_rk(c0, c1, c2, c3hi | MCS_TAG_GLOBAL, a.keys, rnd);
        if (mcs_accepts(rnd[0], mcs_accept_threshold(dE, a.nl2e_over_t))) w ^= pmask;
    }
    a.W[(long long)site * a.Rpad + r] = w;
}

// ------------------------------------------------------------------------------------------
// host <-> packed conversion, initialisation, energies
// ------------------------------------------------------------------------------------------
// int8 [R][N][P] (host order) <-> W[N][Rpad].  A 32x32 tile of (replica, site) is transposed through
// shared memory so that both sides are coalesced: the spin side reads/writes 32 sites * P contiguous
// bytes per warp, the packed side 32 replicas * 8 contiguous bytes per warp.
// sign bits of 4 spin bytes -> 4-bit nibble: bits 7,15,23,31 are moved to 28..31 by one multiply.
__device__ __forceinline__ uint32_t sign_nibble(uint32_t x) { return ((x & 0x80808080u) * 0x00204081u) >> 28; }
// 4-bit nibble -> 4 spin bytes (+1 = 0x01, -1 = 0xFF)
__device__ __forceinline__ uint32_t nibble_spins(uint32_t n)
{
    const uint32_t b = (n * 0x00204081u) & 0x01010101u;
    return 0x01010101u | (b * 0xFEu);
}

// W <-> packed working words (MODE_PACKN): thread = (site, group warp, lane); member m of the word is the window's
// column first + 32 m, first = (group0 + g) 32 pk + lane - replica_offset, where it exists
__global__ void piqmc_words_pack_kernel(const uint64_t *__restrict__ W, uint64_t *__restrict__ Wp, long long N,
                                        long long rpad, long long gw, long long group0, long long roff,
                                        long long nvalid, int pk, int P, int unpack)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * gw * 32) return;
    const long long i = t / (gw * 32), x = t % (gw * 32), g = x >> 5;
    const int lane = (int)(x & 31);
    const long long first = (group0 + g) * 32 * pk + lane - roff;
    const uint64_t pm1 = (1ull << P) - 1ull;
    uint64_t *Wi = const_cast<uint64_t *>(W) + i * rpad;
    if (unpack) {
        const uint64_t v = Wp[t];
        for (int m = 0; m < pk; ++m)
            if (first + 32 * m >= 0 && first + 32 * m < nvalid) Wi[first + 32 * m] = (v >> (m * P)) & pm1;
    } else {
        uint64_t v = 0;
        for (int m = 0; m < pk; ++m)
            if (first + 32 * m >= 0 && first + 32 * m < nvalid) v |= (Wi[first + 32 * m] & pm1) << (m * P);
        Wp[t] = v;
    }
}

__global__ void __launch_bounds__(1024) piqmc_pack_kernel(const int8_t *__restrict__ in, uint64_t *__restrict__ W,
                                                          long long N, long long R, long long Rpad, int P)
{
    __shared__ uint64_t tile[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const long long i0 = (long long)blockIdx.x * 32, r0 = (long long)blockIdx.y * 32;
    {
        const long long r = r0 + ty, i = i0 + tx;
        uint64_t w = 0;
        if (r < R && i < N) {
            const int8_t *src = in + (r * N + i) * P;
            if ((P & 15) == 0) {
                for (int q = 0; q < P / 16; ++q) {
                    const uint4 x = __ldg(reinterpret_cast<const uint4 *>(src) + q);
                    const uint32_t nib = sign_nibble(x.x) | (sign_nibble(x.y) << 4) | (sign_nibble(x.z) << 8) |
                                         (sign_nibble(x.w) << 12);
                    w |= (uint64_t)nib << (16 * q);
                }
            } else if ((P & 3) == 0) {
                for (int q = 0; q < P / 4; ++q)
                    w |= (uint64_t)sign_nibble(__ldg(reinterpret_cast<const uint32_t *>(src) + q)) << (4 * q);
            } else {
                for (int k = 0; k < P; ++k) w |= (uint64_t)(src[k] < 0) << k;
            }
        }
        tile[ty][tx] = w;
    }
    __syncthreads();
    {
        const long long i = i0 + ty, r = r0 + tx;
        if (i < N) W[i * Rpad + r] = tile[tx][ty]; // r < 32 gridDim.y = width of the replica window
    }
}

__global__ void __launch_bounds__(1024) piqmc_unpack_kernel(const uint64_t *__restrict__ W, int8_t *__restrict__ out,
                                                            long long N, long long R, long long Rpad, int P)
{
    __shared__ uint64_t tile[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const long long i0 = (long long)blockIdx.x * 32, r0 = (long long)blockIdx.y * 32;
    {
        const long long i = i0 + ty, r = r0 + tx;
        tile[ty][tx] = i < N ? W[i * Rpad + r] : 0ull;
    }
    __syncthreads();
    const long long r = r0 + ty, i = i0 + tx;
    if (r >= R || i >= N) return;
    const uint64_t w = tile[tx][ty];
    int8_t *dst = out + (r * N + i) * P;
    if ((P & 15) == 0) {
        for (int q = 0; q < P / 16; ++q) {
            const uint32_t h = (uint32_t)(w >> (16 * q));
            uint4 x;
            x.x = nibble_spins(h & 0xFu);
            x.y = nibble_spins((h >> 4) & 0xFu);
            x.z = nibble_spins((h >> 8) & 0xFu);
            x.w = nibble_spins((h >> 12) & 0xFu);
            reinterpret_cast<uint4 *>(dst)[q] = x;
        }
    } else if ((P & 3) == 0) {
        for (int q = 0; q < P / 4; ++q)
            reinterpret_cast<uint32_t *>(dst)[q] = nibble_spins((uint32_t)(w >> (4 * q)) & 0xFu);
    } else {
        for (int k = 0; k < P; ++k) dst[k] = ((w >> k) & 1ull) ? -1 : 1;
    }
}

// int8 [R][N] -> W: every slice of a world line starts from that spin, i.e. confs = np.tile(state, (P, 1)).T of
// the example (santoro80.py:286) done on the device: 1/P of the host traffic of the full [R][N][P] upload.
__global__ void __launch_bounds__(1024) piqmc_tile_kernel(const int8_t *__restrict__ in, uint64_t *__restrict__ W,
                                                          long long N, long long R, long long Rpad, int P)
{
    __shared__ int8_t tile[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const long long i0 = (long long)blockIdx.x * 32, r0 = (long long)blockIdx.y * 32;
    {
        const long long r = r0 + ty, i = i0 + tx;
        tile[ty][tx] = (r < R && i < N) ? in[r * N + i] : (int8_t)1;
    }
    __syncthreads();
    const long long i = i0 + ty, r = r0 + tx;
    const uint64_t pmask = P == 64 ? ~0ull : ((1ull << P) - 1ull);
    if (i < N) W[i * Rpad + r] = tile[tx][ty] < 0 ? pmask : 0ull;
}

// Best slice of every anneal (santoro80.py:290-296 only ever uses the min-over-slices energy): arg-min over the
// fixed-order fp64 energies [R][P] (first minimum, like np.argmin) ...
__global__ void piqmc_argmin_kernel(const double *__restrict__ E, double *__restrict__ ebest, int32_t *__restrict__ kbest,
                                    long long R, int P)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    double best = E[r * P];
    int kb = 0;
    for (int k = 1; k < P; ++k) {
        const double e = E[r * P + k];
        if (e < best) {
            best = e;
            kb = k;
        }
    }
    ebest[r] = best;
    kbest[r] = kb;
}

// ... and that slice's spins as int8 [R][N] (N bytes per anneal instead of N P)
__global__ void __launch_bounds__(1024) piqmc_extract_kernel(const uint64_t *__restrict__ W,
                                                             const int32_t *__restrict__ kbest, int8_t *__restrict__ out,
                                                             long long N, long long R, long long Rpad)
{
    __shared__ int8_t tile[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const long long i0 = (long long)blockIdx.x * 32, r0 = (long long)blockIdx.y * 32;
    {
        const long long i = i0 + ty, r = r0 + tx;
        int8_t v = 1;
        if (i < N && r < R) v = ((W[i * Rpad + r] >> kbest[r]) & 1ull) ? -1 : 1;
        tile[ty][tx] = v;
    }
    __syncthreads();
    const long long r = r0 + ty, i = i0 + tx;
    if (r < R && i < N) out[r * N + i] = tile[tx][ty];
}

__global__ void piqmc_init_kernel(uint64_t *W, long long N, long long R, long long Rpad, int P, uint32_t key0,
                                  uint32_t key1, uint32_t replica_offset)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * Rpad) return;
    const long long i = t / Rpad, r = t % Rpad;
    uint32_t rnd[4];
    mcs_philox4x32(replica_offset + (uint32_t)r, (uint32_t)i, 0u, MCS_TAG_INIT, key0, key1, rnd);
    const uint64_t pmask = P == 64 ? ~0ull : ((1ull << P) - 1ull);
    W[i * Rpad + r] = (r < R && (rnd[0] & 1u)) ? pmask : 0ull;
}

// Fixed-order fp64 classical energy per (replica, slice), bit-identical to the oracle's
// mcs_oracle_ising_energy (definition of tools.ClassicalIsingEnergy, tools.pyx:99-118, with the
// BLAS-order ambiguity removed): rows in site order, entries in table order, no FMA.
// MB > 0: rows of at most MB entries; the row of site i + 1 and the MB neighbour words of site i are requested
// together, ahead of the arithmetic (the loop is a chain of dependent L2 loads otherwise: 1.2 us per site).
template <int MB>
__global__ void piqmc_energy_kernel(const uint64_t *__restrict__ W, const int32_t *__restrict__ tab_idx,
                                    const double *__restrict__ tab_J, double *__restrict__ out, long long N,
                                    int maxnb, long long R, long long Rpad, int P)
{
    // thread = (replica, group of 8 consecutive slices): every word is loaded once for 8 accumulators; each
    // accumulator still adds the sites in order 0..N-1 with the row's entries in table order (bit-exactness)
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int k0 = (blockIdx.y * blockDim.y + threadIdx.y) * 8;
    if (r >= R || k0 >= P) return;
    double e[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) e[q] = 0.0;
    if (MB > 0) {
        // software pipeline, two sites deep: while site i is summed, the words of site i + 1 and the table row of
        // site i + 2 are in flight (the batch does not fit the L2: a word costs a DRAM round trip)
        constexpr int M = MB > 0 ? MB : 1;
        int j1[M], j2[M];      // rows of sites i + 1, i + 2
        double v1[M], v2[M];
        uint32_t b1[M], w1;    // words of site i + 1
        auto load_row = [&](long long i, int (&j)[M], double (&v)[M]) {
#pragma unroll
            for (int s = 0; s < M; ++s) {
                const bool in = s < maxnb && i < N;
                j[s] = in ? __ldg(&tab_idx[i * maxnb + s]) : 0;
                v[s] = in ? __ldg(&tab_J[i * maxnb + s]) : 0.0;
            }
        };
        auto load_words = [&](long long i, const int (&j)[M], uint32_t (&b)[M], uint32_t &w) {
#pragma unroll
            for (int s = 0; s < M; ++s) b[s] = (uint32_t)(__ldg(&W[(long long)j[s] * Rpad + r]) >> k0);
            w = (uint32_t)(__ldg(&W[(i < N ? i : 0) * Rpad + r]) >> k0);
        };
        load_row(0, j1, v1);
        load_words(0, j1, b1, w1);
        load_row(1, j2, v2);
        for (long long i = 0; i < N; ++i) {
            int j[M];
            double jv[M];
            uint32_t bits[M];
            const uint32_t wi = w1;
#pragma unroll
            for (int s = 0; s < M; ++s) j[s] = j1[s], jv[s] = v1[s], bits[s] = b1[s];
#pragma unroll
            for (int s = 0; s < M; ++s) j1[s] = j2[s], v1[s] = v2[s];
            load_words(i + 1, j1, b1, w1);
            load_row(i + 2, j2, v2);
            double pair[8], field = 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q) pair[q] = 0.0;
#pragma unroll
            for (int s = 0; s < M; ++s) {
                if (s >= maxnb) break;
                if (j[s] == (int)i) {
                    field = __dadd_rn(field, jv[s]);
                } else {
                    const double njv = -jv[s];
#pragma unroll
                    for (int q = 0; q < 8; ++q) pair[q] = __dadd_rn(pair[q], ((bits[s] >> q) & 1u) ? njv : jv[s]);
                }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const double t = __dadd_rn(__dmul_rn(0.5, pair[q]), field);
                e[q] = __dadd_rn(e[q], ((wi >> q) & 1u) ? -t : t); // s_i * (0.5 pair + field)
            }
        }
    } else {
        for (long long i = 0; i < N; ++i) {
            double pair[8], field = 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q) pair[q] = 0.0;
            for (int s = 0; s < maxnb; ++s) {
                const int j = __ldg(&tab_idx[i * maxnb + s]);
                const double jv = __ldg(&tab_J[i * maxnb + s]);
                if (j == i) {
                    field = __dadd_rn(field, jv);
                } else {
                    const uint32_t bits = (uint32_t)(W[(long long)j * Rpad + r] >> k0);
                    const double njv = -jv;
#pragma unroll
                    for (int q = 0; q < 8; ++q) pair[q] = __dadd_rn(pair[q], ((bits >> q) & 1u) ? njv : jv); // jv * s_j
                }
            }
            const uint32_t wi = (uint32_t)(W[i * Rpad + r] >> k0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const double t = __dadd_rn(__dmul_rn(0.5, pair[q]), field);
                e[q] = __dadd_rn(e[q], ((wi >> q) & 1u) ? -t : t); // s_i * (0.5 pair + field)
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q)
        if (k0 + q < P) out[r * P + k0 + q] = e[q];
}


// ---- fixed-order energy by table ---------------------------------------------------------------------------
// The chain kernel above spends seven fp64 operations per (replica, slice, site) on a chain of dependent loads:
// 3.7 ms at 4096 anneals and 2.4 ms however small the batch (the fixed cost that capped the 8-GPU end-to-end
// efficiency).  For a row with at most four off-diagonal entries the term s_i (0.5 pair + field) takes 2 x 16 values: the
// sixteen t0[b] = 0.5 pair(b) + field, pair(b) accumulated in table order with the signs of pattern b, are computed
// ONCE per instance by exactly the operations of the chain kernel; an accumulator then looks its term up and spends
// one addition per site.  Same values, same order of additions: bit-identical (tests).
__global__ void energy_table_kernel(const int32_t *__restrict__ tab_idx, const double *__restrict__ tab_J,
                                    double *__restrict__ etab, int32_t *__restrict__ etab_j, long long N, int maxnb)
{
    const long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= N * 16) return;
    const long long i = x >> 4;
    const int b = (int)(x & 15);
    double pair = 0.0, field = 0.0;
    int slot = 0;
    for (int s = 0; s < maxnb; ++s) {
        const int j = tab_idx[i * maxnb + s];
        const double jv = tab_J[i * maxnb + s];
        if (j == i) {
            field = __dadd_rn(field, jv);
        } else {
            pair = __dadd_rn(pair, ((b >> slot) & 1) ? -jv : jv); // jv * s_j
            if (b == 0) etab_j[i * 4 + slot] = j;
            ++slot;
        }
    }
    if (b == 0)
        for (; slot < 4; ++slot) etab_j[i * 4 + slot] = (int32_t)i; // unused pattern bits are never looked at
    etab[x] = __dadd_rn(__dmul_rn(0.5, pair), field);
}

// bit q of a byte -> bit 4 q
__host__ __device__ constexpr uint32_t spread_nibbles(uint32_t v)
{
    uint32_t r = 0;
    for (int q = 0; q < 8; ++q) r |= ((v >> q) & 1u) << (4 * q);
    return r;
}

constexpr int kLTile = 128;
// thread = (replica, group of 8 consecutive slices); CTA = 32 replicas x blockDim.y slice groups
template <int CH>
__global__ void __launch_bounds__(256) piqmc_energy_lut_kernel(const uint64_t *__restrict__ W,
                                                               const double *__restrict__ etab,
                                                               const int32_t *__restrict__ etab_j,
                                                               double *__restrict__ out, long long N, long long R,
                                                               long long Rpad, int P)
{
    __shared__ __align__(16) double s_t[kLTile][16];
    __shared__ __align__(16) int32_t s_j[kLTile][4];
    __shared__ uint32_t s_spread[256];
    const int tid = threadIdx.y * 32 + threadIdx.x, nthr = blockDim.y * 32;
    const long long r = (long long)blockIdx.x * 32 + threadIdx.x; // < Rpad
    const int k0 = (blockIdx.y * blockDim.y + threadIdx.y) * 8;
    // the eight slices of this thread sit in one 32-bit half of the word; raw halves are kept in registers and
    // shifted at use, and every load is unconditional (a conditional load is consumed inside its branch: the loads
    // of a chunk would then wait for each other)
    const uint32_t *Wr = reinterpret_cast<const uint32_t *>(W + r) + (k0 >> 5);
    const int sh = k0 & 31;
    const long long rstride = 2 * Rpad;
    for (int v = tid; v < 256; v += nthr) s_spread[v] = spread_nibbles((uint32_t)v);
    double e[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) e[q] = 0.0;
    for (long long tile0 = 0; tile0 < N; tile0 += kLTile) {
        const int nt = (int)min((long long)kLTile, N - tile0);
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
        auto load_chunk = [&](int c0, uint32_t (&bb)[CH][4], uint32_t (&ww)[CH]) {
#pragma unroll
            for (int u = 0; u < CH; ++u) {
                const int i = min(c0 + u, nt - 1); // past the tile: a valid row, never summed
                const int4 j = *reinterpret_cast<const int4 *>(s_j[i]);
                bb[u][0] = __ldg(&Wr[(long long)j.x * rstride]);
                bb[u][1] = __ldg(&Wr[(long long)j.y * rstride]);
                bb[u][2] = __ldg(&Wr[(long long)j.z * rstride]);
                bb[u][3] = __ldg(&Wr[(long long)j.w * rstride]);
                ww[u] = __ldg(&Wr[(tile0 + i) * rstride]);
            }
        };
        auto sum_chunk = [&](int c0, const uint32_t (&bb)[CH][4], const uint32_t (&ww)[CH]) {
#pragma unroll
            for (int u = 0; u < CH; ++u) {
                const int i = c0 + u;
                if (i < nt) {
                    // anti-alignment pattern of slice k0 + q in nibble q (bit s: neighbour s is -1)
                    const uint32_t packed = s_spread[(bb[u][0] >> sh) & 0xFFu] | (s_spread[(bb[u][1] >> sh) & 0xFFu] << 1) |
                                            (s_spread[(bb[u][2] >> sh) & 0xFFu] << 2) |
                                            (s_spread[(bb[u][3] >> sh) & 0xFFu] << 3);
                    const uint32_t wi = ww[u] >> sh;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const double t = s_t[i][(packed >> (4 * q)) & 15u];
                        e[q] = __dadd_rn(e[q], ((wi >> q) & 1u) ? -t : t); // s_i * (0.5 pair + field)
                    }
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
    if (r < R) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (k0 + q < P) out[r * P + k0 + q] = e[q];
    }
}

} // namespace

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
template <int NPL, int FLD>
static void launch_lut_wf(int warps, const PiqmcPass &a0, cudaStream_t s)
{
    // a.G = warps of work per site; `warps` divides it, so a CTA never straddles two sites
    PiqmcPass a = a0;
    const unsigned ny = (unsigned)std::min(a.nsites, 65535), nz = (unsigned)((a.nsites + 65534) / 65535);
    const bool no_fuse = getenv("MCS_NO_FUSE") != nullptr; // tests: one replica per thread for every P
    const bool no_pack = getenv("MCS_NO_PACK") != nullptr; // tests: at most two replicas per thread
    if (LutGeom<NPL>::FW == 8 && mcs_piqmc_packs(a.P) && !no_fuse && !no_pack) { // pk >= 3
        // pk = floor(64 / P) replicas per thread (at most 6: loads per thread), blocks of 32 pk GLOBAL replicas per warp
        a.pk = std::min(64 / a.P, 6);
        const long long blk = 32ll * a.pk; // replicas per warp
        a.group0 = (long long)a.replica_offset / blk;
        long long gw = ((long long)a.replica_offset + a.nvalid - 1) / blk - a.group0 + 1; // warps
        if (a.gw_n > 0)
            gw = a.gw_n; // a chunk of the window's group warps (two-stream sweeps)
        else
            a.gw_lo = 0, a.gw_n = gw;
        // warps per CTA (they share the site's table): the largest of 4, 2, 1 that leaves at most a tenth of the
        // launched warps without a group
        int wf = 1;
        for (int cand = 4; cand >= 2; cand /= 2)
            if (gw >= cand && ((gw + cand - 1) / cand * cand - gw) * 10 <= gw) {
                wf = cand;
                break;
            }
        a.seg_lsb = 0;
        for (int m = 0; m < a.pk; ++m) a.seg_lsb |= 1ull << (m * a.P);
        a.half = 0;
        const dim3 grid((unsigned)((gw + wf - 1) / wf), ny, nz);
        constexpr int MP = LutGeom<NPL>::FW == 8 ? MODE_PACK : MODE_PLAIN, MN = LutGeom<NPL>::FW == 8 ? MODE_PACKN : MODE_PLAIN;
        if (a.Wp && gw * std::max(1, a.chunks) >= 8 && (a.pk <= 4 || getenv("MCS_WPT")) && !getenv("MCS_PACK_ONE_WORD")) {
            // one-warp CTAs, several packed words per thread (see the plain mode below): the thread's k-th word is in
            // group warp gw_lo + blockIdx.x + k wstep.  Up to four world lines per word (P >= 13); with five or six the
            // per-member world-line work dominates and one word per thread measured faster (P = 10: 1.04 against 1.01e12)
            int want = 64;
            if (const char *e = getenv("MCS_WPT")) want = atoi(e);
            a.wpt = 1;
            while (2 * a.wpt <= want && gw / (2 * a.wpt) >= 1) a.wpt *= 2;
            a.wstep = (gw + a.wpt - 1) / a.wpt; // group warps per slab; the last slab may be short (guarded)
            const dim3 g1((unsigned)a.wstep, ny, nz);
            mcs_launch_pdl(piqmc_lut_pass_kernel<NPL, 1, false, FLD, MN, true>, g1, dim3(32), s, a);
        } else if (a.Wp) {
            if (wf == 4)
                mcs_launch_pdl(piqmc_lut_pass_kernel<NPL, 4, false, FLD, MN>, grid, dim3(128), s, a);
            else if (wf == 2)
                mcs_launch_pdl(piqmc_lut_pass_kernel<NPL, 2, false, FLD, MN>, grid, dim3(64), s, a);
            else
                mcs_launch_pdl(piqmc_lut_pass_kernel<NPL, 1, false, FLD, MN>, grid, dim3(32), s, a);
        } else if (wf == 4)
            mcs_launch_pdl(piqmc_lut_pass_kernel<NPL, 4, false, FLD, MP>, grid, dim3(128), s, a);
        else if (wf == 2)
            mcs_launch_pdl(piqmc_lut_pass_kernel<NPL, 2, false, FLD, MP>, grid, dim3(64), s, a);
        else
            mcs_launch_pdl(piqmc_lut_pass_kernel<NPL, 1, false, FLD, MP>, grid, dim3(32), s, a);
        return;
    }
    a.pk = 0;
    if (LutGeom<NPL>::FW == 8 && a.P <= 32 && a.G % 2 == 0 && !no_fuse) {
        // two replicas per thread: r and r + half, half = the window's replicas / 2
        const int gf = a.G / 2, wf = (gf % 4 == 0) ? 4 : (gf % 2 == 0) ? 2 : 1;
        a.half = (long long)gf * 32;
        const dim3 grid((unsigned)(gf / wf), ny, nz);
        if (wf == 4)
            mcs_launch_pdl(piqmc_lut_pass_kernel<NPL, 4, false, FLD, LutGeom<NPL>::FW == 8 ? MODE_FUSE : MODE_PLAIN>, grid, dim3(128), s, a);
        else if (wf == 2)
            mcs_launch_pdl(piqmc_lut_pass_kernel<NPL, 2, false, FLD, LutGeom<NPL>::FW == 8 ? MODE_FUSE : MODE_PLAIN>, grid, dim3(64), s, a);
        else
            mcs_launch_pdl(piqmc_lut_pass_kernel<NPL, 1, false, FLD, LutGeom<NPL>::FW == 8 ? MODE_FUSE : MODE_PLAIN>, grid, dim3(32), s, a);
        return;
    }
    a.half = 0;
    // Words per thread (P = 64): a quarter of a one-word thread's instructions is set-up that depends on the site only, so
    // a thread takes several replicas one after the other.  Measured at cfg3 (profiles/r02_wpt.log), four-warp CTAs:
    // 1.94e12 (one word) -> 2.00e12 (four) -> 2.035e12 attempts/s (sixteen); ONE-warp CTAs (no CTA barrier, the warp
    // builds its own table) with up to 64 words per thread: 2.08e12 at 4096 anneals, 2.09e12 at 1024, 2.06e12 at 512,
    // 2.01e12 at 256 -- as long as the launches in flight keep a warp per site and chunk (at 128 anneals the one-word
    // CTAs win: 1.68e12 against 1.55e12).  P < 64 (per-pair branches, fewer attempts per word) does not gain.
    a.wpt = 1;
    int cw = warps; // warps per CTA
    {
        int want = 64, force_w = 0;
        if (const char *e = getenv("MCS_WPT")) want = atoi(e);
        if (const char *e = getenv("MCS_WPT_WARPS")) force_w = atoi(e);
        const bool narrow = force_w ? force_w == 1 : (long long)a.G * std::max(1, a.chunks) >= 8;
        if (narrow) { // any group count: slabs of ceil(G / wpt) groups, the last one may be short (guarded in the kernel)
            cw = 1;
            a.wpt = std::max(1, std::min(std::min(want, 64), a.G));
            const int slab = (a.G + a.wpt - 1) / a.wpt;
            a.wpt = (a.G + slab - 1) / slab;
            a.wstep = (long long)slab * 32;
        } else if (warps == 4 && a.P == 64) {
            for (int cand = 16; cand >= 2; cand /= 2)
                if (cand <= want && (a.G / 4) % cand == 0) {
                    a.wpt = cand;
                    break;
                }
        }
    }
    if (cw != 1 || a.wpt == 1) a.wstep = (long long)(a.G / a.wpt) * 32;
    const dim3 grid((unsigned)(cw == 1 && a.wpt > 1 ? a.wstep / 32 : a.G / cw / a.wpt), ny, nz);
    if (a.P == 64 && cw == 1 && a.wpt > 1)
        mcs_launch_pdl(piqmc_lut_pass_kernel<NPL, 1, true, FLD, MODE_PLAIN, true>, grid, dim3(32), s, a);
    else if (cw == 1 && a.wpt > 1)
        mcs_launch_pdl(piqmc_lut_pass_kernel<NPL, 1, false, FLD, MODE_PLAIN, true>, grid, dim3(32), s, a);
    else if (cw == 1)
        mcs_launch_pdl(piqmc_lut_pass_kernel<NPL, 1, false, FLD, MODE_PLAIN>, grid, dim3(32), s, a);
    else if (a.P == 64 && warps == 4 && a.wpt > 1)
        mcs_launch_pdl(piqmc_lut_pass_kernel<NPL, 4, true, FLD, MODE_PLAIN, true>, grid, dim3(128), s, a);
    else if (a.P == 64 && warps == 4)
        mcs_launch_pdl(piqmc_lut_pass_kernel<NPL, 4, true, FLD, MODE_PLAIN>, grid, dim3(128), s, a);
    else if (warps == 4)
        mcs_launch_pdl(piqmc_lut_pass_kernel<NPL, 4, false, FLD, MODE_PLAIN>, grid, dim3(128), s, a);
    else if (warps == 2)
        mcs_launch_pdl(piqmc_lut_pass_kernel<NPL, 2, false, FLD, MODE_PLAIN>, grid, dim3(64), s, a);
    else
        mcs_launch_pdl(piqmc_lut_pass_kernel<NPL, 1, false, FLD, MODE_PLAIN>, grid, dim3(32), s, a);
}

template <int NPL>
static void launch_lut_w(int warps, const PiqmcPass &a, cudaStream_t s)
{
    if (a.field)
        launch_lut_wf<NPL, 1>(warps, a, s);
    else
        launch_lut_wf<NPL, 0>(warps, a, s);
}

static void launch_lut(int npl, int warps, cudaStream_t s, const PiqmcPass &a)
{
    switch (npl) {
    case 1: launch_lut_w<1>(warps, a, s); break;
    case 2: launch_lut_w<2>(warps, a, s); break;
    case 3: launch_lut_w<3>(warps, a, s); break;
    case 4: launch_lut_w<4>(warps, a, s); break;
    case 5: launch_lut_w<5>(warps, a, s); break;
    case 6: launch_lut_w<6>(warps, a, s); break;
    case 7: launch_lut_w<7>(warps, a, s); break;
    default: launch_lut_w<8>(warps, a, s); break;
    }
}

static void launch_bath(int npl, unsigned grid, cudaStream_t s, const PiqmcPass &a, const BathArgs &b)
{
    switch (npl) {
    case 1: piqmc_bath_pass_kernel<1><<<grid, kWarps * 32, 0, s>>>(a, b); break;
    case 2: piqmc_bath_pass_kernel<2><<<grid, kWarps * 32, 0, s>>>(a, b); break;
    case 3: piqmc_bath_pass_kernel<3><<<grid, kWarps * 32, 0, s>>>(a, b); break;
    case 4: piqmc_bath_pass_kernel<4><<<grid, kWarps * 32, 0, s>>>(a, b); break;
    case 5: piqmc_bath_pass_kernel<5><<<grid, kWarps * 32, 0, s>>>(a, b); break;
    case 6: piqmc_bath_pass_kernel<6><<<grid, kWarps * 32, 0, s>>>(a, b); break;
    case 7: piqmc_bath_pass_kernel<7><<<grid, kWarps * 32, 0, s>>>(a, b); break; // Chimera with local fields
    case 8: piqmc_bath_pass_kernel<8><<<grid, kWarps * 32, 0, s>>>(a, b); break;
    default: piqmc_bath_pass_kernel<0><<<grid, kWarps * 32, 0, s>>>(a, b); break; // any degree: row walk
    }
}

int mcs_launch_piqmc_sweeps(mcs_state *st, const double *A, const double *B, int64_t S, int mcsteps, float temp,
                            int global_moves, uint64_t seed, uint64_t replica_offset, uint64_t sweep_offset,
                            const double *lookuptable)
{
    mcs_instance *inst = st->inst;
    const int P = (int)st->P;
    const double teff = (double)temp * (double)P; // qmc.pyx:85: temp is a C float
    MCS_REQUIRE(teff != 0.0 || S == 0, MCS_EZERODIV, "float division");
    MCS_CUDA(cudaSetDevice(inst->device));
    if (inst->dynamics == MCS_DYN_REFERENCE) {
        MCS_REQUIRE(!lookuptable, MCS_EUNSUPPORTED,
                    "reference dynamics: the Ohmic-bath sweeps are served by the coloured kernel or the exact replay");
        return mcs_launch_refdyn_sweeps(st, MCS_KIND_PIQMC, A, B, S, mcsteps, temp, global_moves, seed, replica_offset,
                                        sweep_offset);
    }
    if (!lookuptable && mcs_dense_supported(inst, P))
        return mcs_launch_dense_sweeps(st, MCS_KIND_PIQMC, A, B, S, mcsteps, temp, global_moves, seed, replica_offset,
                                       sweep_offset);
    BathArgs bath;
    bath.lut4 = nullptr;
    bath.c0 = 0.0f;
    float *d_lut4 = nullptr;
    if (lookuptable) {
        float h4[64];
        double sum = 0.0;
        for (int d = 0; d < 64; ++d) h4[d] = 0.0f;
        for (int d = 1; d < P; ++d) {
            h4[d] = (float)(4.0 * teff * lookuptable[d - 1]);
            sum += lookuptable[d - 1];
        }
        bath.c0 = (float)(2.0 * teff * sum);
        MCS_CUDA(cudaMallocAsync((void **)&d_lut4, sizeof(h4), inst->stream));
        MCS_CUDA(cudaMemcpyAsync(d_lut4, h4, sizeof(h4), cudaMemcpyHostToDevice, inst->stream));
        MCS_CUDA(cudaStreamSynchronize(inst->stream)); // h4 is on this stack frame
        bath.lut4 = d_lut4;
    }
    PiqmcPass a;
    a.W = st->d_W + st->win_lo();
    a.ell_idx = inst->d_ell_idx;
    a.ell_J = inst->d_ell_J;
    a.h = inst->d_h;
    a.dpad = inst->dpad;
    a.nq = inst->maxdeg;
    a.field = inst->has_field ? 1 : 0;
    a.G = (int)(st->win_pad() / 32);
    a.Rpad = st->Rpad;
    a.P = P;
    a.keys = mcs_philox_expand(seed);
    a.pow2 = mcs_pow2_make();
    a.replica_offset = (uint32_t)(replica_offset + (uint64_t)st->win_lo());
    a.global_moves = global_moves ? 1 : 0;
    a.tie_thr = mcs_tie_threshold();
    a.half = 0;
    a.pk = 0;
    a.group0 = 0;
    a.nvalid = st->win_valid();
    a.seg_lsb = 0;
    a.gw_lo = 0;
    a.gw_n = 0;
    a.Wp = nullptr;
    a.gp = 0;
    a.wpt = 1;
    a.wstep = 0;
    a.chunks = 1;
    const int npl = std::max(1, inst->maxdeg + (inst->has_field ? 1 : 0));
    const int warps = (a.G % 4 == 0) ? 4 : (a.G % 2 == 0) ? 2 : 1;
    uint64_t sweep = sweep_offset;
    const bool force_direct = getenv("MCS_FORCE_DIRECT") != nullptr; // tests: general-degree kernel on any instance
    MCS_REQUIRE(inst->nsteps == 1 || S <= inst->nsteps, MCS_EINVAL,
                "time-dependent instance has %lld tables but the schedule has %lld steps", (long long)inst->nsteps,
                (long long)S);
    // Mid-size batches (512 anneals per GPU when cfg3 is spread over eight): a colour pass is ten waves of CTAs and
    // its last, partial wave plus the kernel boundary leave SMs idle (56.0 us per pass against 8 x 54.6 for eight
    // times the batch).  Replicas are independent, so the window is cut into two chunks of whole 128- / 256-replica blocks
    // whose passes alternate on two streams: the tail of one chunk's pass runs under the body of the other's.
    // Windows already guarantee that results do not depend on how replicas are grouped (tests); in the packed mode a
    // chunk is a range of the window's group warps.
    const long long G0 = a.G;
    const bool packed_mode = mcs_piqmc_packs(P) && !getenv("MCS_NO_FUSE") && !getenv("MCS_NO_PACK");
    // Packed mode (even P <= 20): the working words (pk world lines each) are built ONCE per call into a scratch array
    // and the passes run on them -- one 64-bit load per table row instead of pk guarded loads and shifts -- then
    // unpacked (MCS_PACK_GATHER=1: gather the members in every pass, the round-1/2 kernel; same decisions)
    uint64_t *d_Wp = nullptr;
    long long pk_gw = 0, pk_group0 = 0;
    int pk_n = 0;
    const bool lut_path = !lookuptable && inst->lut_ok && !force_direct;
    if (packed_mode && lut_path && npl + 2 <= 8 && a.nvalid > 0) {
        pk_n = std::min(64 / P, 6);
        const long long blk = 32ll * pk_n;
        pk_group0 = (long long)a.replica_offset / blk;
        pk_gw = ((long long)a.replica_offset + a.nvalid - 1) / blk - pk_group0 + 1;
        if (!getenv("MCS_PACK_GATHER") && (long long)S * mcsteps * inst->ncolors >= 4) {
            const long long n = inst->N * pk_gw * 32;
            if (st->Wpk_bytes < (size_t)n * sizeof(uint64_t)) { // kept with the batch: no allocation per call
                MCS_CUDA(cudaStreamSynchronize(inst->stream));
                if (st->d_Wpk) MCS_CUDA(cudaFree(st->d_Wpk));
                st->d_Wpk = nullptr;
                st->Wpk_bytes = 0;
                MCS_CUDA(cudaMalloc((void **)&st->d_Wpk, (size_t)n * sizeof(uint64_t)));
                st->Wpk_bytes = (size_t)n * sizeof(uint64_t);
            }
            d_Wp = st->d_Wpk;
            piqmc_words_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, inst->stream>>>(
                a.W, d_Wp, inst->N, st->Rpad, pk_gw, pk_group0, (long long)a.replica_offset, a.nvalid, pk_n, P, 0);
            inst->launches++;
            a.Wp = d_Wp;
            a.gp = pk_gw * 32;
        }
    }
    long long max_sites = 0;
    for (int c = 0; c < inst->ncolors; ++c)
        max_sites = std::max(max_sites, (long long)(inst->color_start[c + 1] - inst->color_start[c]));
    // measured (profiles/r02_piqmc_streams.log): 2.1 us per pass saved at every size from 512 anneals up (56.1 -> 54.0 us
    // at 512: the per-GPU rate of the 8-GPU run equals the 1-GPU rate), nothing more with three streams
    // groups of 32 replicas per block: any for the one-warp CTAs of the plain mode, four for four-warp CTAs, eight with
    // two replicas per thread
    const char *force_warps = getenv("MCS_WPT_WARPS");
    const long long gran = P <= 32 ? 8 : ((force_warps && atoi(force_warps) == 4) ? 4 : 1);
    int nchunk = 1;
    const bool can_chunk = lut_path && (packed_mode ? (d_Wp != nullptr && pk_gw >= 8) : G0 % gran == 0);
    if (can_chunk && (packed_mode || G0 >= std::max(2 * gran, 8ll)) && G0 * max_sites <= (long long)1 << 20 &&
        !getenv("MCS_ONE_STREAM"))
        nchunk = 2;
    if (const char *e = getenv("MCS_STREAMS"))
        if (can_chunk)
            nchunk = (int)std::max(1ll, std::min(std::min(4ll, packed_mode ? pk_gw / 4 : G0 / std::max(gran, 4ll)), atoll(e)));
    a.chunks = nchunk;
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
    uint64_t *const W0 = a.W;
    const uint32_t roff0 = a.replica_offset;
    const long long valid0 = a.nvalid;
    for (int64_t f = 0; f < S; ++f) {
        a.ell_J = inst->ell_J_at(f);
        a.h = inst->h_at(f);
        const double jperp = -0.5 * teff * log(tanh(A[f] / teff)); // qmc.pyx:95
        a.bcoef = (float)(-2.0 * B[f]);                            // qmc.pyx:96
        a.jperp2 = (float)(2.0 * jperp);
        a.nl2e_over_t = (float)(-1.4426950408889634 / teff);
        for (int step = 0; step < mcsteps; ++step, ++sweep) {
            a.sweep_lo = (uint32_t)sweep;
            a.sweep_hi = (uint32_t)(sweep >> 32);
            for (int c = 0; c < inst->ncolors; ++c) {
                a.sites = inst->d_order + inst->color_start[c];
                a.nsites = inst->color_start[c + 1] - inst->color_start[c];
                if (a.nsites == 0) continue;
                if (nchunk > 1 && packed_mode) { // chunks of the window's group warps (whole four-warp CTAs)
                    for (int q = 0; q < nchunk; ++q) {
                        const long long g0 = (pk_gw / 4 * q / nchunk) * 4;
                        const long long g1 = q + 1 == nchunk ? pk_gw : (pk_gw / 4 * (q + 1) / nchunk) * 4;
                        a.gw_lo = g0;
                        a.gw_n = g1 - g0;
                        launch_lut(npl, 4, q ? inst->s_aux[q - 1] : inst->stream, a);
                        inst->launches++;
                    }
                    a.gw_lo = 0;
                    a.gw_n = 0;
                    continue;
                }
                if (nchunk > 1) {
                    for (int q = 0; q < nchunk; ++q) {
                        const long long g0 = (G0 / gran * q / nchunk) * gran, g1 = (G0 / gran * (q + 1) / nchunk) * gran;
                        a.W = W0 + 32 * g0; // replica is the fastest axis: a chunk is a column offset
                        a.G = (int)(g1 - g0);
                        a.replica_offset = roff0 + (uint32_t)(32 * g0);
                        a.nvalid = std::max(0ll, std::min(valid0 - 32 * g0, 32 * (g1 - g0)));
                        launch_lut(npl, 4, q ? inst->s_aux[q - 1] : inst->stream, a);
                        inst->launches++;
                    }
                    continue;
                }
                const long long items = (long long)a.nsites * a.G;
                if (lookuptable)
                    launch_bath(npl, (unsigned)((items + kWarps - 1) / kWarps), inst->stream, a, bath);
                else if (inst->lut_ok && !force_direct)
                    launch_lut(npl, warps, inst->stream, a);
                else
                    piqmc_direct_pass_kernel<<<(unsigned)((items + kWarps - 1) / kWarps), kWarps * 32, 0,
                                               inst->stream>>>(a);
                inst->launches++;
            }
        }
    }
    for (int q = 0; q + 1 < nchunk; ++q) {
        MCS_CUDA(cudaEventRecord(inst->ev_aux1[q], inst->s_aux[q]));
        MCS_CUDA(cudaStreamWaitEvent(inst->stream, inst->ev_aux1[q], 0));
    }
    if (d_Wp) {
        const long long n = inst->N * pk_gw * 32;
        piqmc_words_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, inst->stream>>>(
            W0, d_Wp, inst->N, st->Rpad, pk_gw, pk_group0, (long long)roff0, valid0, pk_n, P, 1);
        inst->launches++;
    }
    if (d_lut4) MCS_CUDA(cudaFreeAsync(d_lut4, inst->stream));
    MCS_CUDA(mcs_take_launch_error());
    MCS_CUDA(cudaGetLastError());
    return MCS_OK;
}

int mcs_piqmc_pack(mcs_state *st, const int8_t *d_in)
{
    mcs_instance *inst = st->inst;
    dim3 grid((unsigned)((inst->N + 31) / 32), (unsigned)(st->win_pad() / 32));
    piqmc_pack_kernel<<<grid, dim3(32, 32), 0, inst->stream>>>(d_in + st->win_lo() * inst->N * st->P,
                                                               st->d_W + st->win_lo(), inst->N, st->win_valid(),
                                                               st->Rpad, (int)st->P);
    inst->launches++;
    MCS_CUDA(cudaGetLastError());
    return MCS_OK;
}

int mcs_piqmc_unpack(mcs_state *st, int8_t *d_out)
{
    mcs_instance *inst = st->inst;
    dim3 grid((unsigned)((inst->N + 31) / 32), (unsigned)(st->win_pad() / 32));
    piqmc_unpack_kernel<<<grid, dim3(32, 32), 0, inst->stream>>>(st->d_W + st->win_lo(),
                                                                 d_out + st->win_lo() * inst->N * st->P, inst->N,
                                                                 st->win_valid(), st->Rpad, (int)st->P);
    inst->launches++;
    MCS_CUDA(cudaGetLastError());
    return MCS_OK;
}

int mcs_piqmc_tile(mcs_state *st, const int8_t *d_in)
{
    mcs_instance *inst = st->inst;
    dim3 grid((unsigned)((inst->N + 31) / 32), (unsigned)(st->Rpad / 32));
    piqmc_tile_kernel<<<grid, dim3(32, 32), 0, inst->stream>>>(d_in, st->d_W, inst->N, st->R, st->Rpad, (int)st->P);
    inst->launches++;
    MCS_CUDA(cudaGetLastError());
    return MCS_OK;
}

// d_E: energies [R][P] (mcs_piqmc_energy); outputs on the device
int mcs_piqmc_best(mcs_state *st, const double *d_E, double *d_ebest, int32_t *d_kbest, int8_t *d_conf)
{
    mcs_instance *inst = st->inst;
    piqmc_argmin_kernel<<<(unsigned)((st->R + 127) / 128), 128, 0, inst->stream>>>(d_E, d_ebest, d_kbest, st->R,
                                                                                  (int)st->P);
    inst->launches++;
    if (d_conf) {
        dim3 grid((unsigned)((inst->N + 31) / 32), (unsigned)(st->Rpad / 32));
        piqmc_extract_kernel<<<grid, dim3(32, 32), 0, inst->stream>>>(st->d_W, d_kbest, d_conf, inst->N, st->R,
                                                                      st->Rpad);
        inst->launches++;
    }
    MCS_CUDA(cudaGetLastError());
    return MCS_OK;
}

int mcs_piqmc_init(mcs_state *st, uint64_t seed, uint64_t replica_offset)
{
    mcs_instance *inst = st->inst;
    const long long n = inst->N * st->Rpad;
    piqmc_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, inst->stream>>>(
        st->d_W, inst->N, st->R, st->Rpad, (int)st->P, (uint32_t)seed, (uint32_t)(seed >> 32),
        (uint32_t)replica_offset);
    inst->launches++;
    MCS_CUDA(cudaGetLastError());
    return MCS_OK;
}

bool mcs_energy_tables(mcs_instance *inst)
{
    if (inst->max_offdiag > 4 || inst->N <= 0) return false;
    if (inst->d_etab) return true;
    double *et = nullptr;
    int32_t *ej = nullptr;
    if (cudaMalloc((void **)&et, (size_t)inst->N * 16 * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void **)&ej, (size_t)inst->N * 4 * sizeof(int32_t)) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(et);
        return false;
    }
    const long long n = inst->N * 16;
    energy_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, inst->stream>>>(
        inst->tab_idx_at(inst->nsteps - 1), inst->tab_J_at(inst->nsteps - 1), et, ej, inst->N, (int)inst->maxnb);
    inst->launches++;
    inst->d_etab = et;
    inst->d_etab_j = ej;
    return true;
}

int mcs_piqmc_energy(mcs_state *st, double *d_out)
{
    mcs_instance *inst = st->inst;
    const int kgroups = (int)((st->P + 7) / 8);
    dim3 grid((unsigned)((st->R + 31) / 32), (unsigned)((kgroups + 3) / 4));
    const int32_t *ti = inst->tab_idx_at(inst->nsteps - 1);
    const double *tj = inst->tab_J_at(inst->nsteps - 1);
    const char *chain = getenv("MCS_ENERGY_CHAIN"); // tests: the chain kernel
    if (!chain && mcs_energy_tables(inst)) {
        // enough CTAs for every SM when the batch is small: a CTA takes as few slice groups as that needs
        const long long warps = st->Rpad / 32 * kgroups;
        int wpc = kgroups;
        while (wpc > 1 && (wpc % 2 == 0) && warps / wpc < 296) wpc /= 2;
        if (kgroups % wpc != 0) wpc = 1;
        const dim3 lgrid((unsigned)(st->Rpad / 32), (unsigned)(kgroups / wpc)), lblock(32, (unsigned)wpc);
        piqmc_energy_lut_kernel<8><<<lgrid, lblock, 0, inst->stream>>>(st->d_W, inst->d_etab, inst->d_etab_j, d_out,
                                                                      inst->N, st->R, st->Rpad, (int)st->P);
        inst->launches++;
        MCS_CUDA(cudaGetLastError());
        return MCS_OK;
    }
    if (inst->maxnb <= 4)
        piqmc_energy_kernel<4><<<grid, dim3(32, 4), 0, inst->stream>>>(st->d_W, ti, tj, d_out, inst->N, (int)inst->maxnb,
                                                                        st->R, st->Rpad, (int)st->P);
    else if (inst->maxnb <= 8)
        piqmc_energy_kernel<8><<<grid, dim3(32, 4), 0, inst->stream>>>(st->d_W, ti, tj, d_out, inst->N, (int)inst->maxnb,
                                                                        st->R, st->Rpad, (int)st->P);
    else
        piqmc_energy_kernel<0><<<grid, dim3(32, 4), 0, inst->stream>>>(st->d_W, ti, tj, d_out, inst->N, (int)inst->maxnb,
                                                                        st->R, st->Rpad, (int)st->P);
    inst->launches++;
    MCS_CUDA(cudaGetLastError());
    return MCS_OK;
}
