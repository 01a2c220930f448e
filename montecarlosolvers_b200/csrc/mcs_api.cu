// mcs_api.cu -- extern "C" entry points over resident replica batches (see include/mcs_b200.h).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "mcs_common.cuh"

// implemented next to their kernels
int mcs_piqmc_pack(mcs_state *st, const int8_t *d_in);
int mcs_piqmc_unpack(mcs_state *st, int8_t *d_out);
int mcs_piqmc_init(mcs_state *st, uint64_t seed, uint64_t replica_offset);
int mcs_piqmc_energy(mcs_state *st, double *d_out);
int mcs_piqmc_tile(mcs_state *st, const int8_t *d_in);
int mcs_piqmc_best(mcs_state *st, const double *d_E, double *d_ebest, int32_t *d_kbest, int8_t *d_conf);
int mcs_sa_pack(mcs_state *st, const int8_t *d_in);
int mcs_sa_unpack(mcs_state *st, int8_t *d_out);
int mcs_sa_init(mcs_state *st, uint64_t seed, uint64_t replica_offset);
int mcs_sa_energy(mcs_state *st, double *d_out);
int mcs_svmc_pack(mcs_state *st, const double *d_in);
int mcs_svmc_unpack(mcs_state *st, double *d_out);
int mcs_svmc_init(mcs_state *st);
int mcs_svmc_energy(mcs_state *st, double a, double b, double *d_out);

static size_t spin_bytes(const mcs_state *st) { return (size_t)st->R * st->inst->N * st->P; }

extern "C" int mcs_state_upload_spins(mcs_state *st, const int8_t *host)
{
    MCS_REQUIRE(st && st->inst && host, MCS_EINVAL, "mcs_state_upload_spins: NULL argument");
    MCS_REQUIRE(st->kind == MCS_KIND_PIQMC || st->kind == MCS_KIND_SA, MCS_EINVAL,
                "mcs_state_upload_spins: state holds angles, not spins");
    MCS_CUDA(cudaSetDevice(st->inst->device));
    const size_t bytes = spin_bytes(st);
    MCS_TRY(mcs_state_reserve_stage(st, bytes));
    MCS_CUDA(cudaMemcpyAsync(st->d_stage, host, bytes, cudaMemcpyHostToDevice, st->inst->stream));
    return st->kind == MCS_KIND_PIQMC ? mcs_piqmc_pack(st, (const int8_t *)st->d_stage)
                                      : mcs_sa_pack(st, (const int8_t *)st->d_stage);
}

extern "C" int mcs_state_download_spins(mcs_state *st, int8_t *host)
{
    MCS_REQUIRE(st && st->inst && host, MCS_EINVAL, "mcs_state_download_spins: NULL argument");
    MCS_REQUIRE(st->kind == MCS_KIND_PIQMC || st->kind == MCS_KIND_SA, MCS_EINVAL,
                "mcs_state_download_spins: state holds angles, not spins");
    MCS_CUDA(cudaSetDevice(st->inst->device));
    const size_t bytes = spin_bytes(st);
    MCS_TRY(mcs_state_reserve_stage(st, bytes));
    MCS_TRY(st->kind == MCS_KIND_PIQMC ? mcs_piqmc_unpack(st, (int8_t *)st->d_stage)
                                       : mcs_sa_unpack(st, (int8_t *)st->d_stage));
    MCS_CUDA(cudaMemcpyAsync(host, st->d_stage, bytes, cudaMemcpyDeviceToHost, st->inst->stream));
    MCS_CUDA(cudaStreamSynchronize(st->inst->stream));
    return MCS_OK;
}

extern "C" int mcs_state_upload_angles(mcs_state *st, const double *host)
{
    MCS_REQUIRE(st && st->inst && host, MCS_EINVAL, "mcs_state_upload_angles: NULL argument");
    MCS_REQUIRE(st->kind == MCS_KIND_SVMC, MCS_EINVAL, "mcs_state_upload_angles: not an SVMC state");
    MCS_CUDA(cudaSetDevice(st->inst->device));
    const size_t bytes = (size_t)st->R * st->inst->N * sizeof(double);
    MCS_TRY(mcs_state_reserve_stage(st, bytes));
    MCS_CUDA(cudaMemcpyAsync(st->d_stage, host, bytes, cudaMemcpyHostToDevice, st->inst->stream));
    return mcs_svmc_pack(st, (const double *)st->d_stage);
}

extern "C" int mcs_state_download_angles(mcs_state *st, double *host)
{
    MCS_REQUIRE(st && st->inst && host, MCS_EINVAL, "mcs_state_download_angles: NULL argument");
    MCS_REQUIRE(st->kind == MCS_KIND_SVMC, MCS_EINVAL, "mcs_state_download_angles: not an SVMC state");
    MCS_CUDA(cudaSetDevice(st->inst->device));
    const size_t bytes = (size_t)st->R * st->inst->N * sizeof(double);
    MCS_TRY(mcs_state_reserve_stage(st, bytes));
    MCS_TRY(mcs_svmc_unpack(st, (double *)st->d_stage));
    MCS_CUDA(cudaMemcpyAsync(host, st->d_stage, bytes, cudaMemcpyDeviceToHost, st->inst->stream));
    MCS_CUDA(cudaStreamSynchronize(st->inst->stream));
    return MCS_OK;
}

extern "C" int mcs_state_init_random(mcs_state *st, uint64_t seed, uint64_t replica_offset)
{
    MCS_REQUIRE(st && st->inst, MCS_EINVAL, "mcs_state_init_random: NULL or orphaned state");
    MCS_CUDA(cudaSetDevice(st->inst->device));
    if (st->kind == MCS_KIND_PIQMC) return mcs_piqmc_init(st, seed, replica_offset);
    if (st->kind == MCS_KIND_SA) return mcs_sa_init(st, seed, replica_offset);
    return mcs_svmc_init(st);
}

// Result buffer of the energy calls, owned by the batch.  (A cudaMalloc / cudaFree pair per call cost the one-shot
// PIQMC call anywhere between 0 and 470 ms on the B200 boxes while 2000 launches were still in flight.)
static int state_energy_buffer(mcs_state *st, size_t bytes, double **out)
{
    if (st->eout_bytes < bytes) {
        if (st->d_eout) MCS_CUDA(cudaFree(st->d_eout));
        st->d_eout = nullptr;
        st->eout_bytes = 0;
        MCS_CUDA(cudaMalloc((void **)&st->d_eout, bytes));
        st->eout_bytes = bytes;
    }
    *out = st->d_eout;
    return MCS_OK;
}

extern "C" int mcs_state_energies(mcs_state *st, double *host_out)
{
    MCS_REQUIRE(st && st->inst && host_out, MCS_EINVAL, "mcs_state_energies: NULL argument");
    MCS_REQUIRE(st->kind == MCS_KIND_PIQMC || st->kind == MCS_KIND_SA, MCS_EINVAL,
                "mcs_state_energies: use mcs_state_svmc_energies for SVMC states");
    MCS_CUDA(cudaSetDevice(st->inst->device));
    const size_t bytes = (size_t)st->R * st->P * sizeof(double);
    double *d_out = nullptr;
    MCS_TRY(state_energy_buffer(st, bytes, &d_out));
    MCS_TRY(st->kind == MCS_KIND_PIQMC ? mcs_piqmc_energy(st, d_out) : mcs_sa_energy(st, d_out));
    MCS_CUDA(cudaMemcpyAsync(host_out, d_out, bytes, cudaMemcpyDeviceToHost, st->inst->stream));
    MCS_CUDA(cudaStreamSynchronize(st->inst->stream));
    return MCS_OK;
}

// Best slice per anneal, computed on the device (fixed-order fp64 energies -> arg-min -> that slice's spins).
// on_device != 0: the three output pointers are DEVICE pointers on this instance's GPU (tensor handoff); the work is
// queued on the instance's stream and the caller synchronises (mcs_synchronize).  Otherwise host pointers, synchronous.
extern "C" int mcs_state_best(mcs_state *st, double *best_energy, int32_t *best_slice, int8_t *best_conf, int on_device)
{
    MCS_REQUIRE(st && st->inst && st->kind == MCS_KIND_PIQMC, MCS_EINVAL, "mcs_state_best: needs a PIQMC state");
    MCS_CUDA(cudaSetDevice(st->inst->device));
    const size_t R = (size_t)st->R, N = (size_t)st->inst->N;
    double *d_E = nullptr;
    MCS_TRY(state_energy_buffer(st, R * st->P * sizeof(double), &d_E));
    MCS_TRY(mcs_piqmc_energy(st, d_E));
    const size_t need = R * (sizeof(double) + sizeof(int32_t)) + R * N;
    if (st->best_bytes < need) {
        MCS_CUDA(cudaStreamSynchronize(st->inst->stream));
        if (st->d_best) MCS_CUDA(cudaFree(st->d_best));
        st->d_best = nullptr;
        st->best_bytes = 0;
        MCS_CUDA(cudaMalloc(&st->d_best, need));
        st->best_bytes = need;
    }
    double *d_eb = (double *)st->d_best;
    int32_t *d_kb = (int32_t *)(d_eb + R);
    int8_t *d_cf = (int8_t *)(d_kb + R);
    cudaStream_t s = st->inst->stream;
    if (on_device) {
        // arg-min always lands in the batch's own buffer (the extraction needs it); copies are device-to-device
        MCS_TRY(mcs_piqmc_best(st, d_E, d_eb, d_kb, best_conf ? best_conf : nullptr));
        if (best_energy) MCS_CUDA(cudaMemcpyAsync(best_energy, d_eb, R * sizeof(double), cudaMemcpyDeviceToDevice, s));
        if (best_slice) MCS_CUDA(cudaMemcpyAsync(best_slice, d_kb, R * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
        return MCS_OK;
    }
    MCS_TRY(mcs_piqmc_best(st, d_E, d_eb, d_kb, best_conf ? d_cf : nullptr));
    if (best_energy) MCS_CUDA(cudaMemcpyAsync(best_energy, d_eb, R * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (best_slice) MCS_CUDA(cudaMemcpyAsync(best_slice, d_kb, R * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (best_conf) MCS_CUDA(cudaMemcpyAsync(best_conf, d_cf, R * N, cudaMemcpyDeviceToHost, s));
    MCS_CUDA(cudaStreamSynchronize(s));
    return MCS_OK;
}

extern "C" int mcs_state_svmc_energies(mcs_state *st, double a, double b, double *host_out)
{
    MCS_REQUIRE(st && st->inst && host_out, MCS_EINVAL, "mcs_state_svmc_energies: NULL argument");
    MCS_REQUIRE(st->kind == MCS_KIND_SVMC, MCS_EINVAL, "mcs_state_svmc_energies: not an SVMC state");
    MCS_CUDA(cudaSetDevice(st->inst->device));
    const size_t bytes = (size_t)st->R * sizeof(double);
    double *d_out = nullptr;
    MCS_TRY(state_energy_buffer(st, bytes, &d_out));
    MCS_TRY(mcs_svmc_energy(st, a, b, d_out));
    MCS_CUDA(cudaMemcpyAsync(host_out, d_out, bytes, cudaMemcpyDeviceToHost, st->inst->stream));
    MCS_CUDA(cudaStreamSynchronize(st->inst->stream));
    return MCS_OK;
}

extern "C" int mcs_piqmc_sweeps(mcs_state *st, const double *A, const double *B, int64_t S, int mcsteps, float temp,
                                int global_moves, uint64_t seed, uint64_t replica_offset, uint64_t sweep_offset)
{
    MCS_REQUIRE(st && st->inst && st->kind == MCS_KIND_PIQMC, MCS_EINVAL, "mcs_piqmc_sweeps: not a PIQMC state");
    MCS_REQUIRE(S >= 0 && mcsteps >= 0 && (S == 0 || (A && B)), MCS_EINVAL, "mcs_piqmc_sweeps: bad schedule");
    return mcs_launch_piqmc_sweeps(st, A, B, S, mcsteps, temp, global_moves, seed, replica_offset, sweep_offset,
                                   nullptr);
}

extern "C" int mcs_piqmc_sweeps_dissipative(mcs_state *st, const double *A, const double *B, int64_t S, int mcsteps,
                                            float temp, const double *lookuptable, int global_moves, uint64_t seed,
                                            uint64_t replica_offset, uint64_t sweep_offset)
{
    MCS_REQUIRE(st && st->inst && st->kind == MCS_KIND_PIQMC, MCS_EINVAL,
                "mcs_piqmc_sweeps_dissipative: not a PIQMC state");
    MCS_REQUIRE(S >= 0 && mcsteps >= 0 && (S == 0 || (A && B)) && lookuptable, MCS_EINVAL,
                "mcs_piqmc_sweeps_dissipative: bad schedule or NULL lookuptable");
    return mcs_launch_piqmc_sweeps(st, A, B, S, mcsteps, temp, global_moves, seed, replica_offset, sweep_offset,
                                   lookuptable);
}

extern "C" int mcs_sa_sweeps(mcs_state *st, const double *sched, int64_t S, int mcsteps, uint64_t seed,
                             uint64_t replica_offset, uint64_t sweep_offset)
{
    MCS_REQUIRE(st && st->inst && st->kind == MCS_KIND_SA, MCS_EINVAL, "mcs_sa_sweeps: not an SA state");
    MCS_REQUIRE(S >= 0 && mcsteps >= 0 && (S == 0 || sched), MCS_EINVAL, "mcs_sa_sweeps: bad schedule");
    return mcs_launch_sa_sweeps(st, sched, S, mcsteps, seed, replica_offset, sweep_offset);
}

extern "C" int mcs_svmc_sweeps(mcs_state *st, const double *A, const double *B, int64_t S, int mcsteps, float temp,
                               int tf, uint64_t seed, uint64_t replica_offset, uint64_t sweep_offset)
{
    MCS_REQUIRE(st && st->inst && st->kind == MCS_KIND_SVMC, MCS_EINVAL, "mcs_svmc_sweeps: not an SVMC state");
    MCS_REQUIRE(S >= 0 && mcsteps >= 0 && (S == 0 || (A && B)), MCS_EINVAL, "mcs_svmc_sweeps: bad schedule");
    return mcs_launch_svmc_sweeps(st, A, B, S, mcsteps, temp, tf, seed, replica_offset, sweep_offset);
}

extern "C" int mcs_cluster_moves(mcs_state *st, double a, double b, float temp, int nmoves, uint64_t seed,
                                 uint64_t replica_offset, uint64_t sweep_offset)
{
    MCS_REQUIRE(st && st->inst && (st->kind == MCS_KIND_PIQMC || st->kind == MCS_KIND_SA), MCS_EINVAL,
                "mcs_cluster_moves: needs a PIQMC or SA state");
    MCS_REQUIRE(nmoves >= 0, MCS_EINVAL, "mcs_cluster_moves: nmoves < 0");
    return mcs_launch_cluster_moves(st, a, b, (double)temp, nullptr, nmoves, seed, replica_offset, sweep_offset);
}

extern "C" int mcs_cluster_moves_dissipative(mcs_state *st, double a, double b, float temp, const double *lookuptable,
                                             int nmoves, uint64_t seed, uint64_t replica_offset, uint64_t sweep_offset)
{
    MCS_REQUIRE(st && st->kind == MCS_KIND_PIQMC, MCS_EINVAL, "mcs_cluster_moves_dissipative: needs a PIQMC state");
    MCS_REQUIRE(lookuptable, MCS_EINVAL, "mcs_cluster_moves_dissipative: lookuptable is NULL");
    MCS_REQUIRE(nmoves >= 0, MCS_EINVAL, "mcs_cluster_moves_dissipative: nmoves < 0");
    return mcs_launch_cluster_moves(st, a, b, (double)temp, lookuptable, nmoves, seed, replica_offset, sweep_offset);
}

// ---- one-shot host-buffer forms ----------------------------------------------------------------
// The device batch (and its staging buffer) lives in the instance and is reused by the next call of the
// same shape: no cudaMalloc / cudaFree on the call path after the first call.
// Replicas are independent, so a large batch is cut into chunks that flow through a three-stage pipeline:
// H2D(c+1) on one copy stream and D2H(c-1) on another overlap pack + sweeps + unpack of chunk c on the
// compute stream.  Only the first upload and the last download stay exposed.
extern "C" int mcs_piqmc_anneal(mcs_instance *inst, const double *A, const double *B, int64_t S, int mcsteps,
                                float temp, int8_t *confs, int64_t R, int64_t P, int global_moves, uint64_t seed,
                                uint64_t replica_offset, double *energies_out)
{
    MCS_REQUIRE(inst && confs, MCS_EINVAL, "mcs_piqmc_anneal: NULL argument");
    std::lock_guard<std::recursive_mutex> one_call_at_a_time(inst->call_mutex);
    const auto host_t0 = std::chrono::steady_clock::now();
    auto host_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count(); };
    MCS_REQUIRE((double)temp * (double)P != 0.0 || S == 0, MCS_EZERODIV, "float division");
    MCS_REQUIRE(S >= 0 && mcsteps >= 0 && (S == 0 || (A && B)), MCS_EINVAL, "mcs_piqmc_anneal: bad schedule");
    mcs_state *st = nullptr;
    MCS_TRY(mcs_instance_scratch_state(inst, MCS_KIND_PIQMC, R, P, &st));
    MCS_CUDA(cudaSetDevice(inst->device));
    const size_t per_replica = (size_t)inst->N * P;
    MCS_TRY(mcs_state_reserve_stage(st, (size_t)R * per_replica));
    int8_t *stage = (int8_t *)st->d_stage;
    // Replicas are independent, so the batch is cut into windows and the H2D copy of window c+1 / the D2H copy of
    // window c-1 overlap the sweeps of window c.  The first window is small (its upload is exposed in any case)
    // and its upload is TIMED: the host link of the B200 boxes was seen anywhere between 1.5 and 55 GB/s from
    // pinned memory (benchmarks/pcie_diag.py; the cfg3 anneal needs 1.6 GB/s each way).  On a fast link the rest
    // is one large window and a small last one (every window costs a launch tail per colour pass: 128-replica
    // windows sweep 34 % slower per replica than 2048-replica ones); on a slow link it is a ramp 2 : 4 : 4 ... 2 : 1
    // so that every copy hides behind the sweeps of its neighbour.  cfg3, fast link: 3.5 % over the resident time.
    // Dense instances expand the whole batch per call: no windows there.
    constexpr int kMaxWin = 16;
    long long win[kMaxWin];
    int nwin = 0;
    // (reference dynamics: one CTA per replica for the whole schedule -- windows would only underfill the GPU)
    const bool windows = st->Rpad >= 384 && !mcs_dense_supported(inst, (int)P) && inst->dynamics == MCS_DYN_COLORED;
    if (windows && !inst->s_in) {
        MCS_CUDA(cudaStreamCreateWithFlags(&inst->s_in, cudaStreamNonBlocking));
        MCS_CUDA(cudaStreamCreateWithFlags(&inst->s_out, cudaStreamNonBlocking));
        MCS_CUDA(cudaEventCreate(&inst->ev_t0));
        for (int q = 0; q < kMaxWin; ++q) {
            MCS_CUDA(cudaEventCreate(&inst->ev_up[q]));
            MCS_CUDA(cudaEventCreateWithFlags(&inst->ev_done[q], cudaEventDisableTiming));
        }
    }
    cudaStream_t s_in = windows ? inst->s_in : inst->stream;
    cudaStream_t s_out = windows ? inst->s_out : inst->stream;
    // MCS_TRACE_E2E=1: per-window timeline (upload done / sweeps start / sweeps done / download done), ms from call start
    const bool trace = getenv("MCS_TRACE_E2E") != nullptr;
    cudaEvent_t tr0 = nullptr, tr_up[kMaxWin], tr_c0[kMaxWin], tr_c1[kMaxWin], tr_dn[kMaxWin];
    if (trace) {
        cudaEventCreate(&tr0);
        for (int q = 0; q < kMaxWin; ++q) {
            cudaEventCreate(&tr_up[q]);
            cudaEventCreate(&tr_c0[q]);
            cudaEventCreate(&tr_c1[q]);
            cudaEventCreate(&tr_dn[q]);
        }
        cudaEventRecord(tr0, inst->stream);
    }
    if (!windows) {
        win[nwin++] = st->Rpad;
        MCS_CUDA(cudaMemcpyAsync(stage, confs, (size_t)R * per_replica, cudaMemcpyHostToDevice, s_in));
        if (trace) cudaEventRecord(tr_up[0], s_in);
    } else {
        const long long e = 128 * std::max<long long>(1, st->Rpad / 4096);
        // window 0, timed
        MCS_CUDA(cudaEventRecord(inst->ev_t0, s_in));
        MCS_CUDA(cudaMemcpyAsync(stage, confs, (size_t)e * per_replica, cudaMemcpyHostToDevice, s_in));
        MCS_CUDA(cudaEventRecord(inst->ev_up[0], s_in));
        if (trace) cudaEventRecord(tr_up[0], s_in);
        MCS_CUDA(cudaEventSynchronize(inst->ev_up[0]));
        float ms0 = 0.0f;
        MCS_CUDA(cudaEventElapsedTime(&ms0, inst->ev_t0, inst->ev_up[0]));
        double gbs = (double)e * per_replica / (std::max(ms0, 1e-3f) * 1e6);
        if (const char *g = getenv("MCS_ASSUME_LINK_GBS")) gbs = atof(g); // tests: force the slow- / fast-link plan
        win[nwin++] = e;
        long long left = st->Rpad - e;
        if (const char *env = getenv("MCS_WINDOWS")) { // experiments: sizes of the windows after the first one
            for (const char *q = env; *q && nwin < kMaxWin - 1;) {
                char *end;
                const long long v = strtoll(q, &end, 10);
                if (end == q || v <= 0 || v % 128 || v >= left) break;
                win[nwin++] = v;
                left -= v;
                q = *end == ',' ? end + 1 : end;
            }
            win[nwin++] = left;
        } else if (left < 12 * e) { // small batch: one more window, or [rest - e, e] behind a slow link
            if (gbs < 16.0 && left >= 3 * e) {
                win[nwin++] = left - e;
                win[nwin++] = e;
            } else {
                win[nwin++] = left;
            }
        } else if (gbs >= 16.0) { // fast link: [e, big, 2e]
            win[nwin++] = left - 2 * e;
            win[nwin++] = 2 * e;
        } else { // slow link: e, 2e, 4e, 4e, ..., 2e, e
            win[nwin++] = 2 * e;
            left -= 2 * e + 3 * e; // this one and the closing 2e + e
            while (left > 0 && nwin < kMaxWin - 3) {
                const long long v = (left <= 6 * e || nwin == kMaxWin - 4) ? left : 4 * e;
                win[nwin++] = v;
                left -= v;
            }
            win[nwin++] = 2 * e;
            win[nwin++] = e;
        }
        if (trace) fprintf(stderr, "[mcs e2e] first upload %.1f GB/s -> %d windows (host %.1f ms)\n", gbs, nwin, host_ms());
        long long r0 = e;
        for (int c = 1; c < nwin && r0 < R; r0 += win[c], ++c) { // the other uploads, back to back on the H2D stream
            const long long nvalid = std::min<long long>(win[c], R - r0);
            MCS_CUDA(cudaMemcpyAsync(stage + r0 * per_replica, confs + r0 * per_replica,
                                     (size_t)nvalid * per_replica, cudaMemcpyHostToDevice, s_in));
            MCS_CUDA(cudaEventRecord(inst->ev_up[c], s_in));
            if (trace) cudaEventRecord(tr_up[c], s_in);
        }
    }
    int rc = MCS_OK;
    long long r0 = 0;
    for (int c = 0; c < nwin && r0 < R && rc == MCS_OK; r0 += win[c], ++c) {
        const long long nvalid = std::min<long long>(win[c], R - r0);
        st->v0 = r0;
        st->vR = std::min<long long>(win[c], st->Rpad - r0);
        if (windows) MCS_CUDA(cudaStreamWaitEvent(inst->stream, inst->ev_up[c], 0));
        if (trace) cudaEventRecord(tr_c0[c], inst->stream);
        rc = mcs_piqmc_pack(st, stage);
        if (rc == MCS_OK)
            rc = mcs_launch_piqmc_sweeps(st, A, B, S, mcsteps, temp, global_moves, seed, replica_offset, 0, nullptr);
        if (rc == MCS_OK) rc = mcs_piqmc_unpack(st, stage);
        if (rc != MCS_OK) break;
        if (trace) cudaEventRecord(tr_c1[c], inst->stream);
        if (windows) {
            MCS_CUDA(cudaEventRecord(inst->ev_done[c], inst->stream));
            MCS_CUDA(cudaStreamWaitEvent(s_out, inst->ev_done[c], 0));
        }
        MCS_CUDA(cudaMemcpyAsync(confs + r0 * per_replica, stage + r0 * per_replica, (size_t)nvalid * per_replica,
                                 cudaMemcpyDeviceToHost, s_out));
        if (trace) cudaEventRecord(tr_dn[c], s_out);
    }
    st->v0 = 0;
    st->vR = -1;
    if (rc != MCS_OK) {
        cudaDeviceSynchronize();
        return rc;
    }
    if (trace) fprintf(stderr, "[mcs e2e] host: everything enqueued at %.1f ms\n", host_ms());
    if (energies_out) MCS_TRY(mcs_state_energies(st, energies_out)); // overlaps the last download
    if (trace) fprintf(stderr, "[mcs e2e] host: energies back at %.1f ms\n", host_ms());
    if (windows) MCS_CUDA(cudaStreamSynchronize(s_out));
    MCS_CUDA(cudaStreamSynchronize(inst->stream));
    if (trace) {
        fprintf(stderr, "[mcs e2e] host: streams drained at %.1f ms\n", host_ms());
        cudaDeviceSynchronize();
        for (int c = 0; c < nwin; ++c) {
            float up = 0, c0 = 0, c1 = 0, dn = 0;
            cudaEventElapsedTime(&up, tr0, tr_up[c]);
            cudaEventElapsedTime(&c0, tr0, tr_c0[c]);
            cudaEventElapsedTime(&c1, tr0, tr_c1[c]);
            cudaEventElapsedTime(&dn, tr0, tr_dn[c]);
            fprintf(stderr, "[mcs e2e] window %d (%lld replicas): uploaded %.1f, sweeps %.1f .. %.1f, downloaded %.1f ms\n",
                    c, win[c], up, c0, c1, dn);
        }
        cudaEventDestroy(tr0);
        for (int q = 0; q < kMaxWin; ++q) {
            cudaEventDestroy(tr_up[q]);
            cudaEventDestroy(tr_c0[q]);
            cudaEventDestroy(tr_c1[q]);
            cudaEventDestroy(tr_dn[q]);
        }
    }
    return MCS_OK;
}

// The example's protocol fused into one call (santoro80.py:286-296): tile the start configuration over the slices,
// anneal, evaluate every slice, return the best one.  Host traffic: R N bytes in, R (N + 8 P + 12) bytes out
// instead of 2 R N P.
extern "C" int mcs_piqmc_anneal_best(mcs_instance *inst, const double *A, const double *B, int64_t S, int mcsteps,
                                     float temp, const int8_t *spins_in, int input_tiled, int64_t R, int64_t P,
                                     int global_moves, uint64_t seed, uint64_t replica_offset, double *energies_out,
                                     double *best_energy, int32_t *best_slice, int8_t *best_conf)
{
    MCS_REQUIRE(inst && spins_in, MCS_EINVAL, "mcs_piqmc_anneal_best: NULL argument");
    std::lock_guard<std::recursive_mutex> one_call_at_a_time(inst->call_mutex);
    MCS_REQUIRE((double)temp * (double)P != 0.0 || S == 0, MCS_EZERODIV, "float division");
    MCS_REQUIRE(S >= 0 && mcsteps >= 0 && (S == 0 || (A && B)), MCS_EINVAL, "mcs_piqmc_anneal_best: bad schedule");
    mcs_state *st = nullptr;
    MCS_TRY(mcs_instance_scratch_state(inst, MCS_KIND_PIQMC, R, P, &st));
    MCS_CUDA(cudaSetDevice(inst->device));
    if (input_tiled) {
        const size_t bytes = (size_t)R * inst->N;
        MCS_TRY(mcs_state_reserve_stage(st, bytes));
        MCS_CUDA(cudaMemcpyAsync(st->d_stage, spins_in, bytes, cudaMemcpyHostToDevice, inst->stream));
        MCS_TRY(mcs_piqmc_tile(st, (const int8_t *)st->d_stage));
    } else {
        MCS_TRY(mcs_state_upload_spins(st, spins_in));
    }
    MCS_TRY(mcs_launch_piqmc_sweeps(st, A, B, S, mcsteps, temp, global_moves, seed, replica_offset, 0, nullptr));
    MCS_TRY(mcs_state_best(st, best_energy, best_slice, best_conf, 0));
    if (energies_out) { // mcs_state_best left the per-slice energies in the batch's buffer
        MCS_CUDA(cudaMemcpyAsync(energies_out, st->d_eout, (size_t)R * P * sizeof(double), cudaMemcpyDeviceToHost,
                                 inst->stream));
        MCS_CUDA(cudaStreamSynchronize(inst->stream));
    }
    return MCS_OK;
}

extern "C" int mcs_sa_anneal(mcs_instance *inst, const double *sched, int64_t S, int mcsteps, int8_t *svec, int64_t R,
                             uint64_t seed, uint64_t replica_offset, double *energies_out)
{
    MCS_REQUIRE(inst && svec, MCS_EINVAL, "mcs_sa_anneal: NULL argument");
    std::lock_guard<std::recursive_mutex> one_call_at_a_time(inst->call_mutex);
    mcs_state *st = nullptr;
    MCS_TRY(mcs_instance_scratch_state(inst, MCS_KIND_SA, R, 1, &st));
    MCS_TRY(mcs_state_upload_spins(st, svec));
    MCS_TRY(mcs_sa_sweeps(st, sched, S, mcsteps, seed, replica_offset, 0));
    MCS_TRY(mcs_state_download_spins(st, svec));
    if (energies_out) MCS_TRY(mcs_state_energies(st, energies_out));
    return MCS_OK;
}

extern "C" int mcs_svmc_anneal(mcs_instance *inst, const double *A, const double *B, int64_t S, int mcsteps,
                               float temp, double *svec, int64_t R, int tf, uint64_t seed, uint64_t replica_offset)
{
    MCS_REQUIRE(inst && svec, MCS_EINVAL, "mcs_svmc_anneal: NULL argument");
    std::lock_guard<std::recursive_mutex> one_call_at_a_time(inst->call_mutex);
    mcs_state *st = nullptr;
    MCS_TRY(mcs_instance_scratch_state(inst, MCS_KIND_SVMC, R, 1, &st));
    MCS_TRY(mcs_state_upload_angles(st, svec));
    MCS_TRY(mcs_svmc_sweeps(st, A, B, S, mcsteps, temp, tf, seed, replica_offset, 0));
    MCS_TRY(mcs_state_download_angles(st, svec));
    return MCS_OK;
}
