// mcs_instance.cu -- library plumbing, the instance compiler (neighbour table -> coloured fp32
// ELL tables + fp64 reference-order table) and resident replica-batch bookkeeping.
//
// Replaces the consumer side of tools.GenerateNeighbors' table (reference tools.pyx:28-96; the
// solvers read it at qmc.pyx:114-125, sa.pyx:84-94, svmc.pyx:98-108).
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <numeric>
#include <queue>

#include <cuda_bf16.h>

#include "mcs_common.cuh"

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
thread_local cudaError_t mcs_launch_error = cudaSuccess;

void mcs_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int mcs_cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    mcs_set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    if (e == cudaErrorMemoryAllocation) return MCS_ENOMEM;
    return MCS_ENODEVICE;
}

extern "C" int mcs_abi_version(void) { return MCS_ABI_VERSION; }
extern "C" const char *mcs_last_error(void) { return g_err; }

extern "C" int mcs_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" void *mcs_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        mcs_set_error("mcs_host_alloc: cudaHostAlloc(%zu) failed (no CUDA device?)", bytes);
        return nullptr;
    }
    return p;
}

extern "C" void mcs_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

// ------------------------------------------------------------------------------------------
// instance compiler
// ------------------------------------------------------------------------------------------
namespace {

struct Adj {
    std::vector<std::vector<std::pair<int32_t, double>>> nb; // merged quadratic neighbours
    std::vector<double> h;
};

// Colour the (symmetrised) interaction graph.  Bipartite graphs (even tori, Chimera) get the
// exact 2-colouring by BFS -- on the Santoro lattice that is the checkerboard (row+col) mod 2 --
// everything else greedy in largest-degree-first order.
int color_graph(const std::vector<std::vector<int32_t>> &adj, std::vector<int32_t> &color)
{
    const int64_t n = (int64_t)adj.size();
    color.assign(n, -1);
    bool bip = true;
    for (int64_t s = 0; s < n && bip; ++s) {
        if (color[s] >= 0) continue;
        color[s] = 0;
        std::queue<int32_t> q;
        q.push((int32_t)s);
        while (!q.empty() && bip) {
            int32_t u = q.front();
            q.pop();
            for (int32_t v : adj[u]) {
                if (color[v] < 0) {
                    color[v] = color[u] ^ 1;
                    q.push(v);
                } else if (color[v] == color[u]) {
                    bip = false;
                    break;
                }
            }
        }
    }
    if (bip) {
        int nc = 0;
        for (int64_t i = 0; i < n; ++i) nc = std::max(nc, color[i] + 1);
        return std::max(nc, 1);
    }
    color.assign(n, -1);
    std::vector<int32_t> ord(n);
    std::iota(ord.begin(), ord.end(), 0);
    std::stable_sort(ord.begin(), ord.end(),
                     [&](int32_t a, int32_t b) { return adj[a].size() > adj[b].size(); });
    int nc = 0;
    std::vector<int32_t> mark;
    for (int32_t u : ord) {
        mark.assign(nc + 1, 0);
        for (int32_t v : adj[u])
            if (color[v] >= 0) mark[color[v]] = 1;
        int c = 0;
        while (c < nc && mark[c]) ++c;
        color[u] = c;
        nc = std::max(nc, c + 1);
    }
    return nc;
}

template <typename T>
int upload(T **dst, const std::vector<T> &src)
{
    size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
    MCS_CUDA(cudaMalloc((void **)dst, bytes));
    if (!src.empty()) MCS_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return MCS_OK;
}

} // namespace

// nsteps == 1: the static table nbs[N][maxnb][2].  nsteps > 1: the time-dependent table
// nbs[nsteps][N][maxnb][2] of sa.NoisyAnneal / svmc.NoisySVMC[TF] (sa.pyx:291-378, svmc.pyx:236-448):
// one colouring / ELL structure from the union of all steps' neighbours, one fp32 coupling row and field
// per step (a neighbour absent at some step gets J = 0 there), the fp64 tables of every step verbatim.
extern "C" int mcs_instance_create_steps(const double *nbs, int64_t nsteps, int64_t nspins, int64_t maxnb,
                                         int device, mcs_instance **out)
{
    MCS_REQUIRE(out != nullptr, MCS_EINVAL, "mcs_instance_create: out is NULL");
    *out = nullptr;
    MCS_REQUIRE(nbs != nullptr && nspins > 0 && maxnb > 0 && nsteps > 0, MCS_EINVAL,
                "mcs_instance_create: need nbs != NULL, nsteps > 0, nspins > 0, maxnb > 0 (got %lld, %lld, %lld)",
                (long long)nsteps, (long long)nspins, (long long)maxnb);
    MCS_REQUIRE(nspins < (1ll << 31), MCS_EINVAL, "mcs_instance_create: nspins too large");
    int ndev = mcs_device_count();
    MCS_REQUIRE(ndev > 0, MCS_ENODEVICE,
                "mcs_instance_create: no CUDA device visible; libmcs_b200 has no CPU fallback");
    MCS_REQUIRE(device >= 0 && device < ndev, MCS_EINVAL, "mcs_instance_create: device %d out of range [0,%d)",
                device, ndev);

    // ---- parse the reference table(s) ------------------------------------------------------
    const size_t tab_n = (size_t)nspins * maxnb;
    std::vector<int32_t> tab_idx(tab_n * nsteps);
    std::vector<double> tab_J(tab_n * nsteps);
    std::vector<std::vector<std::map<int32_t, double>>> quad(nsteps, std::vector<std::map<int32_t, double>>(nspins));
    std::vector<double> h((size_t)nspins * nsteps, 0.0);
    bool has_field = false;
    int max_offdiag = 0;
    for (int64_t t = 0; t < nsteps; ++t) {
        const double *tab = nbs + (size_t)t * tab_n * 2;
        for (int64_t i = 0; i < nspins; ++i) {
            int offdiag = 0;
            for (int64_t s = 0; s < maxnb; ++s) offdiag += (int64_t)tab[(i * maxnb + s) * 2] != i;
            max_offdiag = std::max(max_offdiag, offdiag);
            for (int64_t s = 0; s < maxnb; ++s) {
                double fi = tab[(i * maxnb + s) * 2];
                double jv = tab[(i * maxnb + s) * 2 + 1];
                MCS_REQUIRE(fi >= 0.0 && fi < (double)nspins, MCS_EINVAL,
                            "mcs_instance_create: neighbour index %g of spin %lld out of range", fi, (long long)i);
                int32_t j = (int32_t)fi; // int(nbs[i, si, 0]), qmc.pyx:116
                tab_idx[t * tab_n + i * maxnb + s] = j;
                tab_J[t * tab_n + i * maxnb + s] = jv;
                if (jv == 0.0) continue; // zero padding (tools.pyx:52-59) or a null coupling: contributes +-0
                if (j == (int32_t)i) {
                    h[t * nspins + i] += jv;
                    has_field = true;
                } else {
                    quad[t][i][j] += jv;
                }
            }
        }
    }
    // union structure over the steps; symmetrised adjacency for colouring (a one-sided table entry still
    // makes the two sites conflict)
    std::vector<std::vector<int32_t>> nbr(nspins), adj(nspins);
    for (int64_t t = 0; t < nsteps; ++t)
        for (int64_t i = 0; i < nspins; ++i)
            for (auto &kv : quad[t][i]) {
                nbr[i].push_back(kv.first);
                adj[i].push_back(kv.first);
                adj[kv.first].push_back((int32_t)i);
            }
    for (auto *v : {&nbr, &adj})
        for (auto &a : *v) {
            std::sort(a.begin(), a.end());
            a.erase(std::unique(a.begin(), a.end()), a.end());
        }

    mcs_instance *inst = new mcs_instance();
    inst->max_offdiag = max_offdiag;
    inst->device = device;
    inst->N = nspins;
    inst->maxnb = maxnb;
    inst->nsteps = nsteps;
    inst->has_field = has_field;
    inst->ncolors = color_graph(adj, inst->color);
    int maxdeg = 0;
    for (int64_t i = 0; i < nspins; ++i) maxdeg = std::max<int>(maxdeg, (int)nbr[i].size());
    inst->maxdeg = maxdeg;
    inst->dpad = std::max(maxdeg, 1);
    inst->lut_ok = (maxdeg + (has_field ? 1 : 0) + 2) <= 10; // threshold-table PIQMC kernels: up to 8 in-plane planes

    inst->order.resize(nspins);
    std::iota(inst->order.begin(), inst->order.end(), 0);
    std::stable_sort(inst->order.begin(), inst->order.end(),
                     [&](int32_t a, int32_t b) { return inst->color[a] < inst->color[b]; });
    inst->color_start.assign(inst->ncolors + 1, 0);
    for (int64_t i = 0; i < nspins; ++i) inst->color_start[inst->color[i] + 1]++;
    for (int c = 0; c < inst->ncolors; ++c) inst->color_start[c + 1] += inst->color_start[c];

    const size_t ell_n = (size_t)nspins * inst->dpad;
    std::vector<int32_t> ell_idx(ell_n);
    std::vector<float> ell_J(ell_n * nsteps, 0.0f);
    std::vector<float> hf((size_t)nspins * nsteps);
    for (int64_t i = 0; i < nspins; ++i) {
        int s = 0;
        for (int32_t j : nbr[i]) {
            ell_idx[i * inst->dpad + s] = j;
            for (int64_t t = 0; t < nsteps; ++t) {
                auto it = quad[t][i].find(j);
                if (it != quad[t][i].end()) ell_J[t * ell_n + i * inst->dpad + s] = (float)it->second;
            }
            ++s;
        }
        for (; s < inst->dpad; ++s) ell_idx[i * inst->dpad + s] = (int32_t)i;
        for (int64_t t = 0; t < nsteps; ++t) hf[t * nspins + i] = (float)h[t * nspins + i];
    }

    int rc = MCS_OK;
    auto fail = [&](int code) {
        mcs_instance_destroy(inst);
        return code;
    };
    if (cudaSetDevice(device) != cudaSuccess) {
        mcs_cuda_fail(cudaGetLastError(), "cudaSetDevice", __FILE__, __LINE__);
        return fail(MCS_ENODEVICE);
    }
    if (cudaStreamCreateWithFlags(&inst->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&inst->ev0) != cudaSuccess || cudaEventCreate(&inst->ev1) != cudaSuccess) {
        mcs_cuda_fail(cudaGetLastError(), "stream/event create", __FILE__, __LINE__);
        return fail(MCS_ENODEVICE);
    }
    if ((rc = upload(&inst->d_tab_idx, tab_idx)) || (rc = upload(&inst->d_tab_J, tab_J)) ||
        (rc = upload(&inst->d_ell_idx, ell_idx)) || (rc = upload(&inst->d_ell_J, ell_J)) ||
        (rc = upload(&inst->d_h, hf)) || (rc = upload(&inst->d_order, inst->order)))
        return fail(rc);
    {
        std::vector<int32_t> pos((size_t)nspins);
        for (int64_t q = 0; q < nspins; ++q) pos[(size_t)inst->order[(size_t)q]] = (int32_t)q;
        if ((rc = upload(&inst->d_pos, pos))) return fail(rc);
    }
    // dense instances (SK-like): the coupling matrix itself, fp32 and split into two bf16 halves.  Decided from the
    // DENSITY of the graph (at least a quarter of all pairs coupled), not from one hub: a sparse graph with a few
    // high-degree sites stays with the coloured kernels.  The blocked path needs 8 Npad^2 bytes on the device.
    int64_t nnz = 0;
    for (int64_t i = 0; i < nspins; ++i) nnz += (int64_t)quad[0][i].size(); // both orientations of every bond
    inst->dense = nsteps == 1 && maxdeg >= 48 && nspins <= 32768 && 4 * nnz >= nspins * (nspins - 1);
    if (inst->dense) {
        size_t free_b = 0, total_b = 0;
        const size_t need = (size_t)((nspins + 127) / 128 * 128) * (size_t)((nspins + 127) / 128 * 128) * 8u;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || need > free_b / 2) inst->dense = false;
    }
    if (inst->dense) {
        const int64_t Np = (nspins + 127) / 128 * 128;
        inst->Npad = Np;
        std::vector<float> Jf((size_t)Np * Np, 0.0f), hp((size_t)Np, 0.0f);
        std::vector<__nv_bfloat16> Jhi((size_t)Np * Np), Jlo((size_t)Np * Np);
        inst->field_bound = 0.0;
        for (int64_t i = 0; i < nspins; ++i) {
            double rowsum = std::fabs(h[i]);
            for (auto &kv : quad[0][i]) {
                Jf[(size_t)i * Np + kv.first] = (float)kv.second;
                rowsum += std::fabs(kv.second);
            }
            hp[i] = (float)h[i];
            inst->field_bound = std::max(inst->field_bound, rowsum);
        }
        for (size_t e = 0; e < Jf.size(); ++e) {
            const __nv_bfloat16 hi = __float2bfloat16(Jf[e]);
            Jhi[e] = hi;
            Jlo[e] = __float2bfloat16(Jf[e] - __bfloat162float(hi));
        }
        __nv_bfloat16 *dhi = nullptr, *dlo = nullptr;
        if ((rc = upload(&dhi, Jhi)) || (rc = upload(&dlo, Jlo)) || (rc = upload(&inst->d_Jf, Jf)) ||
            (rc = upload(&inst->d_hpad, hp))) {
            cudaFree(dhi);
            cudaFree(dlo);
            return fail(rc);
        }
        inst->d_Jhi = dhi;
        inst->d_Jlo = dlo;
    }
    *out = inst;
    return MCS_OK;
}

extern "C" int mcs_instance_create(const double *nbs, int64_t nspins, int64_t maxnb, int device,
                                   mcs_instance **out)
{
    return mcs_instance_create_steps(nbs, 1, nspins, maxnb, device, out);
}

static void state_release_device(mcs_state *st)
{
    cudaFree(st->d_W);
    cudaFree(st->d_V);
    cudaFree(st->d_theta);
    cudaFree(st->d_cosz);
    cudaFree(st->d_stage);
    cudaFree(st->d_S16);
    cudaFree(st->d_Wpk);
    cudaFree(st->d_eout);
    st->d_eout = nullptr;
    st->eout_bytes = 0;
    cudaFree(st->d_best);
    st->d_best = nullptr;
    st->best_bytes = 0;
    st->d_S16 = nullptr;
    st->d_Wpk = nullptr;
    st->Wpk_bytes = 0;
    st->S16_cols = 0;
    cudaFree(st->d_labels);
    st->d_labels = nullptr;
    st->labels_bytes = 0;
    st->d_W = nullptr;
    st->d_V = nullptr;
    st->d_theta = st->d_cosz = nullptr;
    st->d_stage = nullptr;
    st->stage_bytes = 0;
}

extern "C" void mcs_instance_destroy(mcs_instance *inst)
{
    if (!inst) return;
    cudaSetDevice(inst->device);
    if (inst->stream) cudaStreamSynchronize(inst->stream);
    for (int k = 0; k < 4; ++k) {
        if (inst->scratch[k]) mcs_state_destroy(inst->scratch[k]);
        inst->scratch[k] = nullptr;
    }
    for (mcs_state *st : inst->states) { // batches that outlive their instance become inert shells
        state_release_device(st);
        st->inst = nullptr;
    }
    inst->states.clear();
    cudaFree(inst->d_tab_idx);
    cudaFree(inst->d_tab_J);
    cudaFree(inst->d_ell_idx);
    cudaFree(inst->d_ell_J);
    cudaFree(inst->d_h);
    cudaFree(inst->d_order);
    cudaFree(inst->d_pos);
    cudaFree(inst->d_etab);
    cudaFree(inst->d_etab_j);
    cudaFree(inst->d_Jhi);
    cudaFree(inst->d_Jlo);
    cudaFree(inst->d_Jf);
    cudaFree(inst->d_hpad);
    for (int q = 0; q < 16; ++q) {
        if (inst->ev_up[q]) cudaEventDestroy(inst->ev_up[q]);
        if (inst->ev_done[q]) cudaEventDestroy(inst->ev_done[q]);
    }
    if (inst->s_in) cudaStreamDestroy(inst->s_in);
    if (inst->s_out) cudaStreamDestroy(inst->s_out);
    for (int q = 0; q < 3; ++q) {
        if (inst->s_aux[q]) cudaStreamDestroy(inst->s_aux[q]);
        if (inst->ev_aux1[q]) cudaEventDestroy(inst->ev_aux1[q]);
    }
    if (inst->ev_aux0) cudaEventDestroy(inst->ev_aux0);
    if (inst->ev0) cudaEventDestroy(inst->ev0);
    if (inst->ev1) cudaEventDestroy(inst->ev1);
    if (inst->stream) cudaStreamDestroy(inst->stream);
    delete inst;
}

extern "C" int mcs_instance_info(const mcs_instance *inst, int64_t info[8])
{
    MCS_REQUIRE(inst && info, MCS_EINVAL, "mcs_instance_info: NULL argument");
    info[0] = inst->N;
    info[1] = inst->maxnb;
    info[2] = inst->ncolors;
    info[3] = inst->maxdeg;
    info[4] = inst->has_field ? 1 : 0;
    info[5] = inst->device;
    info[6] = inst->lut_ok ? 1 : 0;
    info[7] = inst->nsteps + (inst->dense ? (1ll << 32) : 0); // bit 32: dense (blocked tensor-core sweeps)
    return MCS_OK;
}

extern "C" int mcs_instance_set_dense(mcs_instance *inst, int enable)
{
    MCS_REQUIRE(inst, MCS_EINVAL, "mcs_instance_set_dense: NULL instance");
    MCS_REQUIRE(!enable || inst->d_Jf, MCS_EINVAL, "mcs_instance_set_dense: instance was not compiled as dense");
    inst->dense = enable != 0;
    return MCS_OK;
}

extern "C" int mcs_instance_set_dynamics(mcs_instance *inst, int dynamics)
{
    MCS_REQUIRE(inst, MCS_EINVAL, "mcs_instance_set_dynamics: NULL instance");
    MCS_REQUIRE(dynamics == MCS_DYN_COLORED || dynamics == MCS_DYN_REFERENCE, MCS_EINVAL,
                "mcs_instance_set_dynamics: unknown mode %d", dynamics);
    inst->dynamics = dynamics;
    return MCS_OK;
}

extern "C" int mcs_instance_colors(const mcs_instance *inst, int32_t *color)
{
    MCS_REQUIRE(inst && color, MCS_EINVAL, "mcs_instance_colors: NULL argument");
    memcpy(color, inst->color.data(), sizeof(int32_t) * inst->N);
    return MCS_OK;
}

extern "C" int mcs_timer_start(mcs_instance *inst)
{
    MCS_REQUIRE(inst, MCS_EINVAL, "mcs_timer_start: NULL instance");
    MCS_CUDA(cudaSetDevice(inst->device));
    MCS_CUDA(cudaEventRecord(inst->ev0, inst->stream));
    return MCS_OK;
}

extern "C" int mcs_timer_stop(mcs_instance *inst, double *ms)
{
    MCS_REQUIRE(inst && ms, MCS_EINVAL, "mcs_timer_stop: NULL argument");
    MCS_CUDA(cudaSetDevice(inst->device));
    MCS_CUDA(cudaEventRecord(inst->ev1, inst->stream));
    MCS_CUDA(cudaEventSynchronize(inst->ev1));
    float f = 0.f;
    MCS_CUDA(cudaEventElapsedTime(&f, inst->ev0, inst->ev1));
    *ms = (double)f;
    return MCS_OK;
}

extern "C" int mcs_synchronize(mcs_instance *inst)
{
    MCS_REQUIRE(inst, MCS_EINVAL, "mcs_synchronize: NULL instance");
    MCS_CUDA(cudaSetDevice(inst->device));
    MCS_CUDA(cudaStreamSynchronize(inst->stream));
    return MCS_OK;
}

extern "C" int64_t mcs_launch_count(const mcs_instance *inst) { return inst ? inst->launches : 0; }

// ------------------------------------------------------------------------------------------
// resident replica batches
// ------------------------------------------------------------------------------------------
extern "C" int mcs_state_create(mcs_instance *inst, int kind, int64_t R, int64_t P, mcs_state **out)
{
    MCS_REQUIRE(out != nullptr, MCS_EINVAL, "mcs_state_create: out is NULL");
    *out = nullptr;
    MCS_REQUIRE(inst != nullptr, MCS_EINVAL, "mcs_state_create: NULL instance");
    MCS_REQUIRE(R > 0, MCS_EINVAL, "mcs_state_create: need at least one replica (R=%lld)", (long long)R);
    MCS_REQUIRE(kind == MCS_KIND_PIQMC || kind == MCS_KIND_SA || kind == MCS_KIND_SVMC, MCS_EINVAL,
                "mcs_state_create: unknown kind %d", kind);
    if (kind == MCS_KIND_PIQMC) {
        MCS_REQUIRE(P >= 2, MCS_EINVAL,
                    "mcs_state_create: PIQMC needs P >= 2 Trotter slices (P=1 reads out of bounds in the "
                    "reference, qmc.pyx:127-129)");
        MCS_REQUIRE(P <= 64, MCS_EUNSUPPORTED,
                    "mcs_state_create: the bit-packed PIQMC kernels hold one site's slices in a 64-bit word; "
                    "P=%lld > 64 is only served by the exact kernel (mcs_exact_qmc)", (long long)P);
    } else {
        MCS_REQUIRE(P == 1, MCS_EINVAL, "mcs_state_create: P must be 1 for SA / SVMC states");
    }
    MCS_CUDA(cudaSetDevice(inst->device));
    mcs_state *st = new mcs_state();
    st->inst = inst;
    st->kind = kind;
    st->R = R;
    st->P = P;
    st->Rpad = kind == MCS_KIND_SVMC ? (R + 127) / 128 * 128 : (R + 31) / 32 * 32; // SVMC lanes own 4 replicas
    st->G = (R + 31) / 32;
    cudaError_t e = cudaSuccess;
    if (kind == MCS_KIND_PIQMC) {
        size_t bytes = (size_t)inst->N * st->Rpad * sizeof(uint64_t);
        e = cudaMalloc((void **)&st->d_W, bytes);
        if (e == cudaSuccess) e = cudaMemsetAsync(st->d_W, 0, bytes, inst->stream);
    } else if (kind == MCS_KIND_SA) {
        size_t bytes = (size_t)inst->N * st->G * sizeof(uint32_t);
        e = cudaMalloc((void **)&st->d_V, bytes);
        if (e == cudaSuccess) e = cudaMemsetAsync(st->d_V, 0, bytes, inst->stream);
    } else {
        size_t bytes = (size_t)inst->N * st->Rpad * sizeof(float);
        e = cudaMalloc((void **)&st->d_theta, bytes);
        if (e == cudaSuccess) e = cudaMalloc((void **)&st->d_cosz, bytes);
        if (e == cudaSuccess) e = cudaMemsetAsync(st->d_theta, 0, bytes, inst->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(st->d_cosz, 0, bytes, inst->stream);
    }
    if (e != cudaSuccess) {
        int rc = mcs_cuda_fail(e, "state allocation", __FILE__, __LINE__);
        mcs_state_destroy(st);
        return rc;
    }
    inst->states.push_back(st);
    *out = st;
    return MCS_OK;
}

extern "C" void mcs_state_destroy(mcs_state *st)
{
    if (!st) return;
    if (st->inst) {
        cudaSetDevice(st->inst->device);
        cudaStreamSynchronize(st->inst->stream);
        auto &v = st->inst->states;
        v.erase(std::remove(v.begin(), v.end(), st), v.end());
        state_release_device(st);
    }
    delete st;
}

int mcs_state_reserve_stage(mcs_state *st, size_t bytes)
{
    if (st->stage_bytes >= bytes) return MCS_OK;
    MCS_CUDA(cudaStreamSynchronize(st->inst->stream));
    if (st->d_stage) cudaFree(st->d_stage);
    st->d_stage = nullptr;
    st->stage_bytes = 0;
    MCS_CUDA(cudaMalloc(&st->d_stage, bytes));
    st->stage_bytes = bytes;
    return MCS_OK;
}

int mcs_instance_scratch_state(mcs_instance *inst, int kind, int64_t R, int64_t P, mcs_state **out)
{
    MCS_REQUIRE(inst && out && kind >= 1 && kind <= 3, MCS_EINVAL, "mcs_instance_scratch_state: bad argument");
    mcs_state *st = inst->scratch[kind];
    if (st && (st->R != R || st->P != P)) {
        mcs_state_destroy(st);
        inst->scratch[kind] = st = nullptr;
    }
    if (!st) {
        MCS_TRY(mcs_state_create(inst, kind, R, P, &st));
        inst->scratch[kind] = st;
    }
    *out = st;
    return MCS_OK;
}

extern "C" int mcs_instance_trim(mcs_instance *inst)
{
    MCS_REQUIRE(inst, MCS_EINVAL, "mcs_instance_trim: NULL instance");
    for (int k = 0; k < 4; ++k) {
        if (inst->scratch[k]) mcs_state_destroy(inst->scratch[k]);
        inst->scratch[k] = nullptr;
    }
    return MCS_OK;
}
