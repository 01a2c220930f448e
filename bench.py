#!/usr/bin/env python
"""bench.py -- spin-flip attempts/s of the annealing-sweep hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (GPU)
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's CPU path

Workload (BASELINE.json configs[2]): 80x80 PIQMC, P = 64 Trotter slices, 4096
independent anneals in total (sharded contiguously over the N ranks: strong scaling, no data-path
collective), A = linspace(3, 1e-8, 1000), B = 1, mcsteps = 1, T = 1/P, each anneal started from
Philox(seed, global anneal index) spins identical across slices.  One "step" = one full anneal of
the rank's shard = 1000 sweeps = 2000 kernel launches (one per checkerboard colour per sweep).
`value` counts local single-spin attempts only: R * 1000 * 64 * 6400 per step over all ranks.

Timing: W warm-up steps, then K steps bracketed by barrier + device sync, CUDA events recorded on the
stream the kernels are launched on (mcs_timer_start/stop), max over ranks.  L2 is flushed between timed
steps (256 MiB memset) -- and at N = 1 the 210 MB state is larger than L2 anyway.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SIDE, NSPINS, P_SLICES, SCHED = 80, 6400, 64, 1000
METRIC = "spin-flip attempts/s, PIQMC 80x80 P=64"
UNIT = "attempts/s"


def load_instance():
    """Neighbour table of the reference's shipped 80x80 instance (fixture tests/golden/santoro80.npz, sign
    flipped as in santoro80.py:244); synthetic J ~ U(-2, 2) torus of the same structure if absent."""
    from montecarlosolvers_b200 import tools
    import scipy.sparse as sps
    path = os.path.join(ROOT, "tests", "golden", "santoro80.npz")
    J = sps.dok_matrix((NSPINS, NSPINS))
    if os.path.isfile(path):
        d = np.load(path)
        for i, j, v in zip(d["i"], d["j"], d["J_file"]):
            J[int(i), int(j)] = -1.0 * v
        name = "santoro_80x80 couplings"
    else:
        rng = np.random.default_rng(0)
        for r in range(N_SIDE):
            for c in range(N_SIDE):
                i = r * N_SIDE + c
                J[i, r * N_SIDE + (c + 1) % N_SIDE] = rng.uniform(-2, 2)
                J[i, ((r + 1) % N_SIDE) * N_SIDE + c] = rng.uniform(-2, 2)
        name = "synthetic 80x80 torus J~U(-2,2)"
    return tools.GenerateNeighbors(NSPINS, J, 4), name


def roofline(achieved, peak, have_peaks, R, ms_per_launch, bytes_per_launch, n_pass, args):
    """The base contract's HBM roofline of the pass kernel, plus what ncu says actually binds it.  The ncu
    numbers come from the committed capture profiles/r01_piqmc_lut_pass_ncu.json (profiles/capture.sh +
    profiles/summarize_ncu.py), taken at 4096 replicas per GPU; traffic scales linearly with the replicas."""
    prof, src = None, os.path.join("profiles", "r01_piqmc_lut_pass_ncu.json")
    try:
        prof = json.load(open(os.path.join(ROOT, src)))
    except (OSError, ValueError):
        pass
    attempts_per_launch = bytes_per_launch / 0.25
    out = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
           "traffic": args.traffic, "kernel": "piqmc_lut_pass_kernel<4,4,true,0,false>", "ms_per_launch": ms_per_launch,
           "algorithmic_bytes_per_launch": bytes_per_launch,
           "algorithmic_bytes_per_attempt": 0.25,
           "peak_source": "MEASURED_PEAKS.json hbm_gbs" if have_peaks else "fallback 6650 GB/s",
           "note": "the sweep is instruction bound (Philox multiplies and the per-attempt threshold compare on the "
                   "ALU / FMA pipes), not HBM bound: see binding_unit, profiles/ and DESIGN.md section 4"}
    if prof:
        scale = R / 4096.0
        if args.traffic is None:
            out["traffic"] = (prof["dram_bytes_read"] + prof["dram_bytes_write"]) * scale
        out["traffic_source"] = "ncu dram__bytes_read.sum + dram__bytes_write.sum, " + src
        out["binding_unit"] = {
            "unit": "SM ALU pipe (+ FMA pipe held by IMAD.WIDE): instruction issue",
            "alu_pipe_inst_pct_of_peak": prof["alu_pipe_inst_pct_of_peak"],
            "fma_pipe_cycles_active_pct": prof["fma_pipe_cycles_active_pct"],
            "issue_slots_busy_pct": prof["issue_slots_busy_pct"],
            "instructions_per_attempt": prof["warp_instructions"] * 32.0 / (attempts_per_launch / scale),
            "dram_pct_of_peak": prof["dram_pct_of_peak"], "source": "ncu --set full, " + src}
    return out


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                power.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                   r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# the reference's CPU path (oracle/_ref compiled reference if it travelled here, else the C port)
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    kind, nbs, sweeps, seed = args
    A = np.linspace(3.0, 1e-8, SCHED)[:sweeps].copy()
    B = np.ones(sweeps)
    s0 = (2 * np.random.RandomState(seed).randint(2, size=NSPINS) - 1).astype(np.int64)
    confs = np.tile(s0, (P_SLICES, 1)).T.copy(order="F")  # as the example passes it (santoro80.py:286)
    if kind == "reference":
        import ctypes
        import importlib
        from oracle import build_ref
        build_ref.import_ref()
        fn = importlib.import_module("solvers.qmc").QuantumAnneal
        ctypes.CDLL(None).srand(1000 + seed)
        t0 = time.perf_counter()
        fn(A, B, 1, 1.0 / P_SLICES, confs, nbs, 1)
        return time.perf_counter() - t0
    from oracle import oracle as orc
    rng = orc.LibcRand(1000 + seed)
    t0 = time.perf_counter()
    orc.QuantumAnneal(A, B, 1, 1.0 / P_SLICES, confs, nbs, 1, rng=rng)
    return time.perf_counter() - t0


def cpu_arm(nbs, sweeps, cores=None, want="auto"):
    """Aggregate attempts/s of `cores` independent single-threaded reference anneals (the reference is
    single-threaded: its OpenMP flags are commented out, setup.py:10-11), `sweeps` sweeps each."""
    import multiprocessing as mp
    from oracle import build_ref
    from oracle import oracle as orc
    kind = "port"
    if want in ("auto", "reference"):
        build_ref.build(verbose=False)
        if build_ref.import_ref() is not None:
            kind = "reference"
    if kind == "port":
        orc.build()
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        times = pool.map(_cpu_worker, [(kind, nbs, sweeps, s) for s in range(cores)])
        wall = time.perf_counter() - t0
    attempts = cores * sweeps * P_SLICES * NSPINS
    one = sweeps * P_SLICES * NSPINS / float(np.median(times))
    return {"value": attempts / wall, "unit": UNIT, "cores": cores, "kind": kind,
            "value_1core": one,
            "sample": "%d independent anneals (one per core), %d sweeps of the 80x80 P=64 schedule each, "
                      "qmc.QuantumAnneal" % (cores, sweeps)}


def cpu_baseline_subprocess(sweeps):
    """The reference arm of this file in its own process (one bounded step); returns its cpu_baseline object."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0",
           "--cpu-sweeps", str(sweeps)]
    try:
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
        return json.loads(p.stdout.strip().splitlines()[-1])["cpu_baseline"]
    except Exception as e:  # noqa: BLE001 -- the GPU line must still be printed
        return {"error": "cpu baseline failed: %r" % (e,)}


def run_reference(args, out):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    nbs, name = load_instance()
    sweeps = args.cpu_sweeps
    vals = []
    for _ in range(args.warmup):
        cpu_arm(nbs, max(1, sweeps // 8))
    t0 = time.perf_counter()
    last = None
    for _ in range(args.steps):
        last = cpu_arm(nbs, sweeps)
        vals.append(last["value"])
    wall = time.perf_counter() - t0
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": name,
            "config": {"workload": "80x80 PIQMC P=64, reference CPU path on the host cores; each step = a bounded "
                                   "sample: one anneal per core, %d sweeps each" % sweeps},
            "cpu_baseline": dict(last, value=v),
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    out["line"] = line
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, out):
    import torch
    import montecarlosolvers_b200 as mcs
    from montecarlosolvers_b200 import parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    nbs, name = load_instance()
    inst = mcs.Instance(nbs, device=local)
    R_total = args.anneals
    lo, hi = parallel.shard(R_total, rank, world)
    R = hi - lo
    S = args.sched
    A = np.linspace(3.0, 1e-8, S)
    B = np.ones(S)
    temp = 1.0 / P_SLICES
    seed = 20261018
    st = mcs.State(inst, mcs._lib.KIND_PIQMC, R, P_SLICES)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        st.init_random(seed, replica_offset=lo)
        st.piqmc_sweeps(A, B, 1, temp, global_moves=False, seed=seed, replica_offset=lo)

    for _ in range(args.warmup):
        step()
    inst.synchronize()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = inst.launches
    ms_dev = 0.0
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        inst.timer_start()
        step()
        ms_dev += inst.timer_stop()
    barrier()
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches = inst.launches - launches0
    t = torch.tensor([ms_dev], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    attempts_step_all = float(R_total) * S * P_SLICES * NSPINS
    value = attempts_step_all * args.steps / (ms_total * 1e-3)

    def post():
        # final energies + the collective the path has: gather per-anneal best-slice energies, broadcast the best
        e = st.energies()
        best_local = e.min(axis=1)
        conf = st.download_spins()
        kbest = e.argmin(axis=1)
        best_conf = np.ascontiguousarray(conf[np.arange(R), :, kbest])
        del conf
        return parallel.gather_best(best_local, best_conf, lo, R_total, device=dev)

    if not os.environ.get("BENCH_POST_LAST"):
        energies, best, _ = post()

    # ---- e2e: host buffers through the one-shot C-ABI call (H2D + pack + sweeps + unpack + D2H + energies)
    e2e = None
    if args.e2e_steps > 0:
        host = mcs.empty_pinned((R, NSPINS, P_SLICES), np.int8)
        e_host = mcs.empty_pinned((R, P_SLICES), np.float64)
        rs = np.random.RandomState(rank)
        s0 = (2 * rs.randint(2, size=(R, NSPINS, 1)) - 1).astype(np.int8)
        L = mcs._lib.load()
        times = []
        for it in range(args.e2e_steps + 1):  # first one is warm-up
            host[...] = s0  # fresh anneal: broadcast over slices (not timed: input preparation)
            barrier()
            t0 = time.perf_counter()
            mcs._lib.check(L.mcs_piqmc_anneal(inst._h, mcs._lib.dptr(A), mcs._lib.dptr(B), S, 1, temp,
                                              host.ctypes.data, R, P_SLICES, 0, seed + it, lo,
                                              mcs._lib.dptr(e_host)))
            barrier()
            times.append(time.perf_counter() - t0)
        # median of the timed calls: the host link of these boxes is occasionally slow for a whole call (every call is
        # listed in ms_each_rank0); host_link_gbs gives the pinned-copy rate seen right after, for context
        tt = torch.tensor([float(np.median(times[1:]))], dtype=torch.float64, device=dev)
        link = {}
        try:
            hp = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
            dp = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
            for key, dst, src in (("h2d", dp, hp), ("d2h", hp, dp)):
                dst.copy_(src, non_blocking=True)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                dst.copy_(src, non_blocking=True)
                torch.cuda.synchronize()
                link[key] = round((256 << 20) / (time.perf_counter() - t0) / 1e9, 1)
            del hp, dp
        except Exception:  # noqa: BLE001
            pass
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": attempts_step_all / float(tt.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(R) * NSPINS * P_SLICES,
               "d2h_bytes_per_step": int(R) * NSPINS * P_SLICES + int(R) * P_SLICES * 8,
               "ms_per_step": 1e3 * float(tt.item()), "ms_each_rank0": [round(1e3 * x, 1) for x in times[1:]],
               "host_link_gbs": link,
               "api": "mcs_piqmc_anneal (C ABI one-shot: pinned int8 [R,N,P] in/out + float64 energies out)"}

    if os.environ.get("BENCH_POST_LAST"):
        energies, best, _ = post()
    if rank == 0:
        peaks = {}
        ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.isfile(ppath):
            with open(ppath) as f:
                peaks = json.load(f)
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # dominant kernel: piqmc_lut_pass_kernel, one launch per colour class per sweep.
        # algorithmic bytes: 0.25 B per attempt (read + write of one bit-packed spin), DESIGN.md section 4
        n_pass = launches - args.steps  # minus the init kernel of each step
        ms_per_launch = ms_dev / max(n_pass, 1)
        bytes_per_launch = 0.25 * (NSPINS / 2) * R * P_SLICES
        achieved = bytes_per_launch / (ms_per_launch * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u64",
            "dtype_detail": "bit-packed spins in u64 words (bit k = Trotter slice k); u32 acceptance thresholds from f32 energies; u32 Philox",
            "data": name + ", Philox-initialised spins",
            "config": {"workload": "80x80 PIQMC P=64, %d anneals total (%d per GPU), A=linspace(3,1e-8,%d), B=1, "
                                   "mcsteps=1, T=1/64 (BASELINE configs[2])" % (R_total, R, S),
                       "l2": "flushed between timed steps (256 MiB memset); state %.0f MB per GPU" % (
                           R * NSPINS * 8 / 1e6),
                       "timer": "CUDA events on the launch stream (mcs_timer_*), max over ranks",
                       "wall_s_timed_region": wall},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline(achieved, peak, bool(peaks), R, ms_per_launch, bytes_per_launch, n_pass, args),
            "e2e": e2e,
            "result": {"best_residual_energy_per_spin": None, "mean_best_slice_energy": float(np.mean(energies)),
                       "best_anneal": int(best)},
        }
        gs = os.path.join(ROOT, "tests", "golden", "santoro80.npz")
        if os.path.isfile(gs):
            egs = float(np.load(gs)["e_gs_per_spin"])
            line["result"]["best_residual_energy_per_spin"] = float(np.min(energies)) / NSPINS - egs
            line["result"]["mean_residual_energy_per_spin"] = float(np.mean(energies)) / NSPINS - egs
        # CPU baseline (rank 0, N = 1 only), AFTER every GPU measurement and in a separate interpreter: loading all
        # host cores first left the pinned buffers of the e2e leg on slower pages (e2e +15 % when it ran first)
        line["cpu_baseline"] = cpu_baseline_subprocess(args.cpu_sweeps) if (world == 1 and args.cpu_sweeps > 0) else None
        out["line"] = line
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


class _StdoutToStderr(object):
    """Everything libraries print to fd 1 while the bench runs (e.g. NCCL's version banner) goes to stderr,
    so that stdout carries exactly ONE line: the JSON result."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--anneals", type=int, default=4096, help="total anneals over all ranks")
    ap.add_argument("--sched", type=int, default=SCHED, help="schedule length (sweeps per step)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sweeps", type=int, default=40, help="sweeps per core for the CPU baseline sample")
    ap.add_argument("--traffic", type=float, default=None,
                    help="ncu dram__bytes_read+write per launch of the dominant kernel; default: the committed "
                         "capture profiles/r01_piqmc_lut_pass_ncu_full.txt scaled to this run's replicas")
    args = ap.parse_args()
    out = {}
    with _StdoutToStderr():
        rc = run_reference(args, out) if args.impl == "reference" else run_ours(args, out)
    if "line" in out:
        print(json.dumps(out["line"]), flush=True)
    return rc


if __name__ == "__main__":
    sys.exit(main())
