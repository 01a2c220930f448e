// mcs_cluster.cu -- Swendsen-Wang cluster moves on the (space x Trotter) lattice with GPU union-find (sm_100a).
//
// Status of the reference for this row: its README advertises "Wolff and Swendsen-Yang
// cluster updates" (README.md:4), the code holds experimental single-cluster Wolff variants
// (qmc.pyx:620-1621: non-standard r*p growth rule, wrong-dtype buffers that raise on Linux) and NO
// Swendsen-Wang.  There is no trajectory oracle; parity for this file is equilibrium statistics against
// exact enumeration and against the single-spin kernels (tests/test_gpu_cluster.py) -- "parity unpinned".
//
// The PIQMC action in units of teff is  S/teff = sum_bonds -K_b s s'  with
//   in-plane bond (i,j) of slice k :  K = -B J_ij / teff   (qmc.pyx:114-125; J > 0 is antiferromagnetic)
//   Trotter bond (i,k)-(i,k+1)     :  K = +J_perp / teff   (qmc.pyx:95,127-138)
//   field on (i,k)                 :  bond to a fixed ghost spin +1 with K = -B h_i / teff.
// One move: every SATISFIED bond (K s s' > 0) is activated with probability 1 - exp(-2|K|); connected
// components are labelled by a lock-free union-find (atomic hooking of the larger root under the smaller,
// path halving); every component not containing the ghost is flipped with probability 1/2.
// Parallel over everything: one thread per (site, replica) walks the word's slices; each replica has its own
// forest in L[node][replica] (replica fastest, like W).
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "mcs_common.cuh"

namespace {

struct ClusterArgs {
    uint64_t *W;   // PIQMC words [N][Rpad]   (SA: nullptr)
    uint32_t *V;   // SA words    [N][G]      (PIQMC: nullptr)
    int32_t *L;    // labels [(N P + 1)][Rl]  (last node = ghost)
    const int32_t *ell_idx;
    const float *ell_J;
    const float *h;
    long long N, R, Rpad, G, Rl; // R: replicas of this chunk, Rl: label stride (replicas per chunk, padded)
    long long r0;                // first replica of the chunk (the label array holds one chunk at a time)
    int P, dpad;
    float kin;   // in-plane factor:  K_ij = kin * J_ij   (= -B/teff, or -1/T for SA)
    float kperp; // Trotter coupling  J_perp / teff
    int nbath;       // Ohmic bath: number of distances d = 1 .. nbath (= P / 2) that carry a bond, 0 = no bath
    float kbath[32]; // K_d = lookuptable[d-1] (qmc.pyx:268-273 in units of teff; symmetric: table[d-1] == table[P-d-1])
    mcs_philox_keys keys;
    uint32_t sweep_lo, sweep_hi, replica_offset;
};

enum { TAG_CL_BOND = 64, TAG_CL_TROTTER = 64 + 16, TAG_CL_GHOST = 64 + 32, TAG_CL_FLIP = 64 + 48, TAG_CL_BATH = 128 };

__device__ __forceinline__ int32_t uf_find(int32_t *L, long long stride, long long r, int32_t v)
{
    int32_t p = L[(long long)v * stride + r];
    while (p != v) {
        const int32_t gp = L[(long long)p * stride + r];
        if (gp != p) L[(long long)v * stride + r] = gp; // path halving (benign race)
        v = p;
        p = gp;
    }
    return v;
}

__device__ __forceinline__ void uf_union(int32_t *L, long long stride, long long r, int32_t a, int32_t b)
{
    for (;;) {
        a = uf_find(L, stride, r, a);
        b = uf_find(L, stride, r, b);
        if (a == b) return;
        if (a < b) {
            const int32_t t = a;
            a = b;
            b = t;
        } // hook the larger root a under the smaller root b
        const int32_t old = atomicCAS(&L[(long long)a * stride + r], a, b);
        if (old == a) return;
    }
}

// spin word of (site, replica) as a 64-bit mask over slices (SA: one bit)
__device__ __forceinline__ uint64_t load_word(const ClusterArgs &a, long long i, long long r)
{
    r += a.r0;
    if (a.W) return a.W[i * a.Rpad + r];
    return (uint64_t)((a.V[i * a.G + (r >> 5)] >> (r & 31)) & 1u);
}

// bit k set with probability p (p_thr = p * 2^32 as an integer threshold), independent per (slice, replica)
__device__ __forceinline__ uint64_t bernoulli_mask(const ClusterArgs &a, uint32_t c0, uint32_t c1, uint32_t tagbase,
                                                   uint32_t p_thr, uint64_t want)
{
    uint64_t m = 0;
    for (int g = 0; g < (a.P + 3) / 4; ++g) {
        if (((want >> (4 * g)) & 0xFull) == 0) continue; // no candidate bond among these four slices
        uint32_t rnd[4];
        mcs_philox4x32_rk(c0, c1, a.sweep_lo, (a.sweep_hi << 8) | (tagbase + (uint32_t)g), a.keys, rnd);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (rnd[j] < p_thr) m |= 1ull << (4 * g + j);
    }
    return m & want;
}

__device__ __forceinline__ uint32_t prob_threshold(float k2abs) // 1 - exp(-2|K|) as a 32-bit threshold
{
    const float p = 1.0f - __expf(-k2abs);
    const float t = p * 4294967296.0f;
    return t >= 4294967040.0f ? 0xFFFFFFFFu : (uint32_t)t;
}

__global__ void cluster_init_kernel(int32_t *L, long long nodes, long long Rl)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nodes * Rl) return;
    L[t] = (int32_t)(t / Rl);
}

__global__ void cluster_union_kernel(const __grid_constant__ ClusterArgs a)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.N * a.Rl) return;
    const long long i = t / a.Rl, r = t % a.Rl;
    if (r >= a.R) return;
    const int P = a.P;
    const uint64_t pmask = P == 64 ? ~0ull : ((1ull << P) - 1ull);
    const uint64_t w = load_word(a, i, r);
    const uint32_t c0 = a.replica_offset + (uint32_t)(a.r0 + r);
    const int32_t ghost = (int32_t)(a.N * P);
    // in-plane bonds, each taken once from the row of its smaller endpoint
    for (int s = 0; s < a.dpad; ++s) {
        const int j = __ldg(&a.ell_idx[i * a.dpad + s]);
        const float jv = __ldg(&a.ell_J[i * a.dpad + s]);
        if (j <= i || jv == 0.0f) continue;
        const float K = a.kin * jv;
        const uint64_t x = (w ^ load_word(a, j, r)) & pmask;    // anti-aligned slices
        const uint64_t sat = (K > 0.0f ? ~x : x) & pmask;        // K s s' > 0
        const uint64_t act = bernoulli_mask(a, c0, (uint32_t)(i * a.dpad + s), TAG_CL_BOND, prob_threshold(2.0f * fabsf(K)), sat);
        for (uint64_t m = act; m; m &= m - 1) {
            const int k = __ffsll((long long)m) - 1;
            uf_union(a.L, a.Rl, r, (int32_t)(i * P + k), (int32_t)((long long)j * P + k));
        }
    }
    // Trotter bonds (i,k)-(i,k+1), ferromagnetic; for P == 2 the two slices are joined by TWO bonds (the
    // reference adds both neighbours, qmc.pyx:137-138) = one bond of strength 2 K_perp
    if (P > 1 && a.kperp != 0.0f) {
        const uint64_t nxt = ((w >> 1) | (w << (P - 1))) & pmask; // bit k = slice k+1 (ring)
        uint64_t sat = ~(w ^ nxt) & pmask;
        float K = a.kperp;
        if (P == 2) {
            sat &= 1ull;
            K *= 2.0f;
        }
        const uint64_t act = bernoulli_mask(a, c0, (uint32_t)i, TAG_CL_TROTTER, prob_threshold(2.0f * fabsf(K)), sat);
        for (uint64_t m = act; m; m &= m - 1) {
            const int k = __ffsll((long long)m) - 1;
            uf_union(a.L, a.Rl, r, (int32_t)(i * P + k), (int32_t)(i * P + (k + 1 == P ? 0 : k + 1)));
        }
    }
    // Ohmic bath (qmc.pyx:268-273): dE = sum_d 2 teff s_k s_{k+d} table[d-1], i.e. every pair of slices of a world
    // line at ring distance d carries the bond K_d = table[d-1] in units of teff (ferromagnetic for table > 0);
    // pairs (k, k+d), each once: all k for d < P/2, k < P/2 for d = P/2
    for (int d = 1; d <= a.nbath; ++d) {
        const float K = a.kbath[d - 1];
        if (K == 0.0f) continue;
        const uint64_t far = ((w >> d) | (w << (P - d))) & pmask; // bit k = slice (k + d) mod P
        uint64_t sat = (K > 0.0f ? ~(w ^ far) : (w ^ far)) & pmask;
        if (2 * d == P) sat &= (1ull << d) - 1ull;
        const uint64_t act = bernoulli_mask(a, c0, (uint32_t)(i * 32 + (d - 1)), TAG_CL_BATH, prob_threshold(2.0f * fabsf(K)), sat);
        for (uint64_t m = act; m; m &= m - 1) {
            const int k = __ffsll((long long)m) - 1;
            uf_union(a.L, a.Rl, r, (int32_t)(i * P + k), (int32_t)(i * P + (k + d >= P ? k + d - P : k + d)));
        }
    }
    // field: bond to the ghost spin (+1)
    const float hv = __ldg(&a.h[i]);
    if (hv != 0.0f) {
        const float K = a.kin * hv;
        const uint64_t sat = (K > 0.0f ? ~w : w) & pmask; // K s > 0  (bit set <=> s = -1)
        const uint64_t act = bernoulli_mask(a, c0, (uint32_t)i, TAG_CL_GHOST, prob_threshold(2.0f * fabsf(K)), sat);
        for (uint64_t m = act; m; m &= m - 1) {
            const int k = __ffsll((long long)m) - 1;
            uf_union(a.L, a.Rl, r, (int32_t)(i * P + k), ghost);
        }
    }
}

__global__ void cluster_flip_kernel(const __grid_constant__ ClusterArgs a)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.N * a.Rl) return;
    const long long i = t / a.Rl, r = t % a.Rl;
    if (r >= a.R) return;
    const int P = a.P;
    const int32_t groot = uf_find(a.L, a.Rl, r, (int32_t)(a.N * P));
    const uint32_t c0 = a.replica_offset + (uint32_t)(a.r0 + r);
    uint64_t flip = 0;
    for (int k = 0; k < P; ++k) {
        const int32_t root = uf_find(a.L, a.Rl, r, (int32_t)(i * P + k));
        if (root == groot) continue; // tied to the field: stays
        uint32_t rnd[4];
        mcs_philox4x32_rk(c0, (uint32_t)root, a.sweep_lo, (a.sweep_hi << 8) | TAG_CL_FLIP, a.keys, rnd);
        if (rnd[0] & 1u) flip |= 1ull << k;
    }
    if (!flip) return;
    const long long rg = a.r0 + r;
    if (a.W) {
        a.W[i * a.Rpad + rg] ^= flip;
    } else {
        atomicXor(&a.V[i * a.G + (rg >> 5)], 1u << (rg & 31)); // 32 replicas share a word
    }
}

} // namespace

// kind: MCS_KIND_PIQMC (coef_a = Gamma, coef_b = B, temp = T) or MCS_KIND_SA (temp = T);
// lookuptable != nullptr (PIQMC only): Ohmic-bath bonds of the Dissipative solvers, float64 [P-1]
int mcs_launch_cluster_moves(mcs_state *st, double coef_a, double coef_b, double temp, const double *lookuptable,
                             int nmoves, uint64_t seed, uint64_t replica_offset, uint64_t sweep_offset)
{
    mcs_instance *inst = st->inst;
    MCS_REQUIRE(inst->nsteps == 1, MCS_EUNSUPPORTED, "cluster moves need a static coupling table");
    MCS_CUDA(cudaSetDevice(inst->device));
    const int P = (int)st->P;
    const long long nodes = inst->N * P + 1;
    MCS_REQUIRE(nodes < (1ll << 31), MCS_EUNSUPPORTED, "cluster moves: N*P too large for 32-bit labels");
    // labels [(N P + 1)][Rl]: one forest per replica.  The batch is labelled in chunks of Rl replicas so that the
    // array stays below 512 MB whatever the batch (80x80, P = 64: 1.6 MB per replica -> 320 replicas per chunk;
    // all 4096 at once would be 6.7 GB)
    const long long Rall = st->kind == MCS_KIND_PIQMC ? st->Rpad : st->R;
    long long cap = std::max(32ll, ((512ll << 20) / (nodes * 4)) / 32 * 32);
    if (const char *e = getenv("MCS_CLUSTER_CHUNK")) cap = std::max(32ll, atoll(e) / 32 * 32); // tests
    const long long Rl = std::min(Rall, cap);
    const size_t bytes = (size_t)nodes * Rl * sizeof(int32_t);
    if (st->labels_bytes < bytes) {
        if (st->d_labels) MCS_CUDA(cudaFree(st->d_labels));
        st->d_labels = nullptr;
        st->labels_bytes = 0;
        MCS_CUDA(cudaMalloc((void **)&st->d_labels, bytes));
        st->labels_bytes = bytes;
    }
    ClusterArgs a;
    a.W = st->kind == MCS_KIND_PIQMC ? st->d_W : nullptr;
    a.V = st->kind == MCS_KIND_SA ? st->d_V : nullptr;
    a.L = st->d_labels;
    a.ell_idx = inst->d_ell_idx;
    a.ell_J = inst->d_ell_J;
    a.h = inst->d_h;
    a.N = inst->N;
    a.R = st->R;
    a.r0 = 0;
    a.Rpad = st->Rpad;
    a.G = st->G;
    a.Rl = Rl;
    a.P = P;
    a.dpad = inst->dpad;
    if (st->kind == MCS_KIND_PIQMC) {
        const double teff = (double)(float)temp * (double)P;
        MCS_REQUIRE(teff != 0.0, MCS_EZERODIV, "float division");
        a.kin = (float)(-coef_b / teff);
        a.kperp = (float)((-0.5 * teff * log(tanh(coef_a / teff))) / teff);
    } else {
        MCS_REQUIRE(temp > 0.0, MCS_EINVAL, "cluster moves need T > 0");
        a.kin = (float)(-1.0 / temp);
        a.kperp = 0.0f;
    }
    a.nbath = 0;
    for (float &kb : a.kbath) kb = 0.0f;
    if (lookuptable) {
        MCS_REQUIRE(st->kind == MCS_KIND_PIQMC, MCS_EINVAL, "cluster moves: the bath couples Trotter slices (PIQMC only)");
        // a pair of slices at ring distance d is reached as d from one end and as P - d from the other
        // (qmc.pyx:268-273): the table defines an energy only if it is symmetric, like the documented kernel
        // (pi / (P sin(pi d / P)))^2 (qmc.pyx:162-163)
        for (int d = 1; d < P; ++d) {
            const double x = lookuptable[d - 1], y = lookuptable[P - d - 1];
            MCS_REQUIRE(fabs(x - y) <= 1e-9 * std::max(1.0, std::max(fabs(x), fabs(y))), MCS_EUNSUPPORTED,
                        "cluster moves: lookuptable[%d] != lookuptable[%d]: an asymmetric bath table is not an energy "
                        "function, no cluster move can sample it", d - 1, P - d - 1);
        }
        a.nbath = P / 2;
        for (int d = 1; d <= a.nbath; ++d) a.kbath[d - 1] = (float)lookuptable[d - 1];
    }
    a.keys = mcs_philox_expand(seed);
    a.replica_offset = (uint32_t)replica_offset;
    const long long nthreads = inst->N * Rl;
    for (int mv = 0; mv < nmoves; ++mv) {
        const uint64_t sweep = sweep_offset + (uint64_t)mv;
        a.sweep_lo = (uint32_t)sweep;
        a.sweep_hi = (uint32_t)(sweep >> 32);
        for (long long r0 = 0; r0 < st->R; r0 += Rl) { // replicas are independent: chunk after chunk
            a.r0 = r0;
            a.R = std::min(Rl, st->R - r0);
            cluster_init_kernel<<<(unsigned)((nodes * Rl + 255) / 256), 256, 0, inst->stream>>>(a.L, nodes, Rl);
            cluster_union_kernel<<<(unsigned)((nthreads + 127) / 128), 128, 0, inst->stream>>>(a);
            cluster_flip_kernel<<<(unsigned)((nthreads + 127) / 128), 128, 0, inst->stream>>>(a);
            inst->launches += 3;
        }
    }
    MCS_CUDA(cudaGetLastError());
    return MCS_OK;
}
