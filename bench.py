#!/usr/bin/env python
"""bench.py -- spin-flip attempts/s of the annealing-sweep hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (GPU)
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's CPU path

Workload (BASELINE.json configs[2]): 80x80 PIQMC, P = 64 Trotter slices, 4096
independent anneals in total (sharded contiguously over the N ranks: strong scaling, no data-path
collective), A = linspace(3, 1e-8, 1000), B = 1, mcsteps = 1, T = 1/P, each anneal started from
Philox(seed, global anneal index) spins identical across slices.  One "step" = one full anneal of
the rank's shard = 1000 sweeps = 2000 colour passes, each two launches (one per replica chunk, on two streams).
`value` counts local single-spin attempts only: R * 1000 * 64 * 6400 per step over all ranks.

Timing: W warm-up steps, then K steps bracketed by barrier + device sync, CUDA events recorded on the
stream the kernels are launched on (mcs_timer_start/stop), max over ranks.  L2 is flushed between timed
steps (256 MiB memset) -- and at N = 1 the 210 MB state is larger than L2 anyway.

`e2e` = the same anneal through the one-shot C-ABI call mcs_piqmc_anneal_best with HOST buffers: the example's
per-anneal protocol (santoro80.py:286-296: tile the start state over the slices, anneal, evaluate every slice, keep
the best) -- start states int8 [R, N] in, per-slice energies + best energy / slice / configuration out; at N > 1 the
path's only collective (gather the per-anneal best energies, broadcast the winner) runs inside the timed region.
`e2e_full_confs` is the drop-in shaped call (mcs_piqmc_anneal: int8 [R, N, P] world lines both ways).

`configs` (rank 0, N = 1): the other BASELINE configs and the two reference-order modes, each with its own value,
dominant kernel, roofline object and -- from the reference arm run as a subprocess -- the matching reference
function timed on the host cores (bounded samples).
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


def workload_name(R_total, world, S):
    return ("80x80 PIQMC P=64, %d anneals total (%d per GPU), A=linspace(3,1e-8,%d), B=1, mcsteps=1, T=1/64 "
            "(BASELINE configs[2])" % (R_total, R_total // max(world, 1), S))


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


def chimera_instance(m=16, seed=0):
    """cfg4: Chimera C_m, m x m cells of K_{4,4}, J = +-1 from default_rng(seed), no fields."""
    from montecarlosolvers_b200 import tools
    import scipy.sparse as sps
    rng = np.random.default_rng(seed)
    n = 8 * m * m
    J = sps.dok_matrix((n, n))

    def q(r, c, side, k):
        return ((r * m + c) * 2 + side) * 4 + k

    for r in range(m):
        for c in range(m):
            for a in range(4):
                for b in range(4):
                    J[q(r, c, 0, a), q(r, c, 1, b)] = float(rng.choice([-1.0, 1.0]))
                if r + 1 < m:
                    J[q(r, c, 0, a), q(r + 1, c, 0, a)] = float(rng.choice([-1.0, 1.0]))
                if c + 1 < m:
                    J[q(r, c, 1, a), q(r, c + 1, 1, a)] = float(rng.choice([-1.0, 1.0]))
    return tools.GenerateNeighbors(n, J, 6)


def sk_instance(n=2048, seed=0):
    """cfg5: dense SK couplings J_ij ~ N(0, 1)/sqrt(n), reference table float64 [n, n-1, 2]."""
    rng = np.random.default_rng(seed)
    Jm = np.triu(rng.normal(size=(n, n)) / np.sqrt(n), 1)
    full = Jm + Jm.T
    nb = np.zeros((n, n - 1, 2))
    ar = np.arange(n)
    for i in range(n):
        idx = np.delete(ar, i)
        nb[i, :, 0] = idx
        nb[i, :, 1] = full[i, idx]
    return nb


def load_peaks():
    ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(ppath):
        with open(ppath) as f:
            return json.load(f)
    return {}


def hbm_roofline(attempts_per_s, bytes_per_attempt, peaks, kernel, ms_per_launch=None, note=None):
    peak = float(peaks.get("hbm_gbs", 6650.0))
    ach = attempts_per_s * bytes_per_attempt / 1e9
    out = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
           "kernel": kernel, "algorithmic_bytes_per_attempt": bytes_per_attempt,
           "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"}
    if ms_per_launch is not None:
        out["ms_per_launch"] = ms_per_launch
    if note:
        out["note"] = note
    return out


def roofline(achieved, peak, have_peaks, R, ms_per_launch, bytes_per_launch, n_pass, args):
    """The base contract's HBM roofline of the pass kernel, plus what ncu says actually binds it.  The ncu numbers
    come from the newest committed capture profiles/r0*_piqmc_lut_pass_ncu.json (profiles/capture.sh +
    profiles/summarize_ncu.py); traffic scales linearly with the replicas a launch covers."""
    prof, src = None, None
    for rnd in ("r02", "r01"):
        cand = os.path.join("profiles", "%s_piqmc_lut_pass_ncu.json" % rnd)
        try:
            prof, src = json.load(open(os.path.join(ROOT, cand))), cand
            break
        except (OSError, ValueError):
            continue
    attempts_per_launch = bytes_per_launch / 0.25
    out = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
           "traffic": args.traffic, "kernel": "piqmc_lut_pass_kernel<4,1,true,0,MODE_PLAIN,MULTI> (one-warp CTAs, up to 64 world lines per thread)", "ms_per_launch": ms_per_launch,
           "algorithmic_bytes_per_launch": bytes_per_launch,
           "algorithmic_bytes_per_attempt": 0.25,
           "peak_source": "MEASURED_PEAKS.json hbm_gbs" if have_peaks else "fallback 6650 GB/s",
           "note": "the sweep is instruction bound (Philox multiplies and the per-attempt threshold compare on the "
                   "ALU / FMA pipes), not HBM bound: see binding_unit, profiles/ and DESIGN.md section 4"}
    if prof:
        # the profiled launch (profiles/summarize_ncu.py records what it covered; older captures: grid_size CTAs of
        # 128 threads, one 64-slice word per thread)
        prof_attempts = float(prof.get("attempts_in_launch", float(prof.get("grid_size", 102400)) * 128 * 64))
        scale = attempts_per_launch / prof_attempts
        if args.traffic is None:
            out["traffic"] = (prof["dram_bytes_read"] + prof["dram_bytes_write"]) * scale
        out["traffic_source"] = "ncu dram__bytes_read.sum + dram__bytes_write.sum, " + src
        out["binding_unit"] = {
            "unit": "SM ALU pipe (+ FMA pipe held by IMAD.WIDE): instruction issue",
            "alu_pipe_inst_pct_of_peak": prof["alu_pipe_inst_pct_of_peak"],
            "fma_pipe_cycles_active_pct": prof["fma_pipe_cycles_active_pct"],
            "issue_slots_busy_pct": prof["issue_slots_busy_pct"],
            "instructions_per_attempt": prof["warp_instructions"] * 32.0 / prof_attempts,
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
_CPU = {}  # tables inherited by the forked workers


def _strided(n_total, n):
    """n schedule steps spread evenly over a schedule of n_total steps (hot AND cold end)."""
    n = max(1, min(n, n_total))
    return np.unique(np.linspace(0, n_total - 1, n).round().astype(int))


def _cpu_call(kind, what, seed):
    """One reference call on this core; returns (attempts, callable)."""
    tab = _CPU[what["table"]]
    if kind == "reference":
        import importlib
        qmc, sa, svmc = (importlib.import_module("solvers." + m) for m in ("qmc", "sa", "svmc"))
    else:
        from oracle import oracle as orc
        qmc = sa = svmc = orc
    rs = np.random.RandomState(seed)
    solver, n = what["solver"], tab.shape[0]
    kw = {} if kind == "reference" else {"rng": 1000 + seed}
    if solver in ("QuantumAnneal", "QuantumAnnealGlobal"):
        P, idx = what["P"], _strided(what["sched"], what["sweeps"])
        A = np.linspace(3.0, 1e-8, what["sched"])[idx].copy()
        B = np.ones(A.size)
        s0 = (2 * rs.randint(2, size=n) - 1).astype(np.int64)
        confs = np.tile(s0, (P, 1)).T.copy(order="F")  # as the example passes it (santoro80.py:286)
        fn = getattr(qmc, solver)
        return A.size * P * n, lambda: fn(A, B, 1, 1.0 / P, confs, tab, 1, **kw)
    if solver == "Anneal":
        sched = np.linspace(3.0, 0.0, what["sched"])[_strided(what["sched"], what["sweeps"])].copy()
        s0 = (2 * rs.randint(2, size=n) - 1).astype(np.int64)
        return sched.size * n, lambda: sa.Anneal(sched, 1, s0, tab, **kw)
    if solver == "SpinVectorMonteCarloCompact":
        idx = _strided(what["sched"], what["sweeps"])
        s = np.linspace(1e-3, 1.0, what["sched"])[idx]
        A, B = (3.0 * (1 - s)).copy(), s.copy()
        v = np.full((what["reads"], n), np.pi / 2)
        np.random.seed(seed)
        return A.size * what["reads"] * n, lambda: svmc.SpinVectorMonteCarloCompact(A, B, 1, 0.1, v, tab, **kw)
    raise ValueError(solver)


def _cpu_worker(args):
    kind, what, seed, barrier = args
    if kind == "reference":
        import ctypes
        ctypes.CDLL(None).srand(1000 + seed)
    attempts, call = _cpu_call(kind, what, seed)
    barrier.wait()  # every worker is forked, imported and has its inputs: only the reference call is timed
    t0 = time.time()
    call()
    t1 = time.time()
    return attempts, t0, t1


def cpu_arm(what, cores=None, want="auto"):
    """Aggregate attempts/s of `cores` independent single-threaded reference calls (the reference is
    single-threaded: its OpenMP flags are commented out, setup.py:10-11).  Timed from the first worker's start to
    the last worker's end of the reference call itself (pool start-up, imports and input preparation excluded)."""
    import multiprocessing as mp
    from oracle import build_ref
    from oracle import oracle as orc
    kind = "port"
    if want in ("auto", "reference"):
        build_ref.build(verbose=False)
        if build_ref.import_ref() is not None:
            kind = "reference"
            import importlib
            for m in ("qmc", "sa", "svmc"):  # mapped in the PARENT too: visible to a dlopen audit of this process
                importlib.import_module("solvers." + m)
    if kind == "port":
        orc.build()
        orc.lib()
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    barrier = ctx.Manager().Barrier(cores)
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(kind, what, s, barrier) for s in range(cores)], chunksize=1)
    attempts = sum(r[0] for r in res)
    wall = max(r[2] for r in res) - min(r[1] for r in res)
    one = float(np.median([r[0] / (r[2] - r[1]) for r in res]))
    return {"value": attempts / wall, "unit": UNIT, "cores": cores, "kind": kind, "value_1core": one,
            "cpu_model": _cpu_model(),
            "sample": "%d independent %s calls (one per core), %s" % (cores, what["solver"], what["desc"])}


def _cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_workloads(cpu_sweeps):
    """The reference function matching every bench entry, as bounded samples of the same workloads."""
    return {
        "cfg3": {"table": "santoro", "solver": "QuantumAnneal", "P": 64, "sched": SCHED, "sweeps": cpu_sweeps,
                 "desc": "80x80 P=64, %d sweeps spread evenly over the 1000-step schedule" % cpu_sweeps},
        "cfg1": {"table": "santoro", "solver": "QuantumAnnealGlobal", "P": 20, "sched": 354, "sweeps": 2 * cpu_sweeps,
                 "desc": "80x80 P=20 (examples/santoro80.py), %d sweeps spread over the 354-step schedule" % (
                     2 * cpu_sweeps)},
        "cfg2": {"table": "santoro", "solver": "Anneal", "sched": 1000, "sweeps": 1000,
                 "desc": "80x80, the full 1000-temperature schedule, one restart per core"},
        "cfg4": {"table": "chimera", "solver": "SpinVectorMonteCarloCompact", "sched": 1000, "sweeps": 400, "reads": 1,
                 "desc": "Chimera C16 (2048 rotors), 1 read, 400 sweeps spread over the 1000-step schedule"},
        "cfg5": {"table": "sk", "solver": "QuantumAnneal", "P": 32, "sched": 200, "sweeps": 2,
                 "desc": "dense SK N=2048 P=32 on the reference's [2048, 2047, 2] table, 2 sweeps"},
    }


def cpu_tables(names):
    for n in names:
        if n in _CPU:
            continue
        if n == "santoro":
            _CPU[n] = load_instance()[0]
        elif n == "chimera":
            _CPU[n] = chimera_instance(16)
        elif n == "sk":
            _CPU[n] = sk_instance(2048)


def cpu_baseline_subprocess(sweeps, configs):
    """The reference arm of this file in its own process; returns {"cfg3": cpu_baseline, "cfg1": ..., ...}."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0",
           "--cpu-sweeps", str(sweeps), "--cpu-configs", ",".join(configs)]
    try:
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
        line = json.loads(p.stdout.strip().splitlines()[-1])
        out = dict(line.get("cpu_configs") or {})
        out["cfg3"] = line["cpu_baseline"]
        return out
    except Exception as e:  # noqa: BLE001 -- the GPU line must still be printed
        return {"cfg3": {"error": "cpu baseline failed: %r" % (e,)}}


def run_reference(args, out):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = cpu_workloads(args.cpu_sweeps)
    cpu_tables(["santoro"])
    name = load_instance()[1]
    vals = []
    for _ in range(args.warmup):
        cpu_arm(dict(wl["cfg3"], sweeps=max(1, args.cpu_sweeps // 8)))
    t0 = time.perf_counter()
    last = None
    for _ in range(args.steps):
        last = cpu_arm(wl["cfg3"])
        vals.append(last["value"])
    wall = time.perf_counter() - t0
    v = float(np.mean(vals))
    extra = {}
    for c in [c for c in (args.cpu_configs or "").split(",") if c and c != "cfg3"]:
        try:
            cpu_tables([wl[c]["table"]])
            extra[c] = cpu_arm(wl[c])
        except Exception as e:  # noqa: BLE001
            extra[c] = {"error": repr(e)}
    world = max(1, args.gpus)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": name,
            "config": {"workload": workload_name(args.anneals, world, args.sched)},
            "sampling": "reference CPU path on the host cores; each step = a bounded sample of that workload: one "
                        "qmc.QuantumAnneal call per core, %d sweeps spread evenly over the schedule" % args.cpu_sweeps,
            "cpu_baseline": dict(last, value=v),
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if extra:
        line["cpu_configs"] = extra
    out["line"] = line
    return 0


# ------------------------------------------------------------------------------------------------
# the other BASELINE configs + the reference-order modes (rank 0, one GPU)
# ------------------------------------------------------------------------------------------------
def timed(inst, fn, reps=3):
    fn()
    inst.synchronize()
    best = 1e30
    for _ in range(reps):
        inst.timer_start()
        fn()
        best = min(best, inst.timer_stop())
    return best


def run_configs(mcs, inst, peaks, which):
    out = {}
    K = mcs._lib
    if "cfg1" in which:  # examples/santoro80.py: P = 20, world-line moves, tau = 354
        P, tau, R = 20, 354, 4096
        A, B = np.linspace(3.0, 1e-8, tau), np.ones(tau)
        st = mcs.State(inst, K.KIND_PIQMC, R, P)
        st.init_random(1)
        l0 = inst.launches
        ms = timed(inst, lambda: st.piqmc_sweeps(A, B, 1, 1.0 / P, global_moves=True, seed=2), reps=2)
        v = R * tau * P * NSPINS / (ms * 1e-3)
        out["cfg1"] = {"workload": "examples/santoro80.py protocol: 80x80 PIQMC P=20 with world-line moves "
                                   "(QuantumAnnealGlobal), tau=354, %d anneals" % R, "value": v, "unit": UNIT,
                       "ms": ms, "gpu_launches": (inst.launches - l0) // 3,
                       "roofline": hbm_roofline(v, 0.25, peaks, "piqmc_lut_pass_kernel<4,1,false,0,MODE_PACKN,MULTI> (three "
                                                "20-slice world lines per working word, packed words resident in HBM)", ms / (2 * tau))}
        st.close()
    if "cfg2" in which:  # sa.Anneal, 1024 restarts
        tau, R = 1000, 1024
        sched = np.linspace(3.0, 0.0, tau)
        st = mcs.State(inst, K.KIND_SA, R, 1)
        st.init_random(1)
        ms = timed(inst, lambda: st.sa_sweeps(sched, 1, seed=2), reps=3)
        v = R * tau * NSPINS / (ms * 1e-3)
        out["cfg2"] = {"workload": "sa.Anneal on 80x80 Santoro, %d restarts, linspace(3,0,1000), 1 sweep each" % R,
                       "value": v, "unit": UNIT, "ms": ms,
                       "roofline": hbm_roofline(v, 0.25, peaks, "sa_lut_pass_kernel<4,*,0>", ms / (2 * tau),
                                                "latency bound at this batch size: a colour pass is 3200 warps (one "
                                                "wave): 2.1 us of SM-bound work + 2.3 us kernel-boundary latency per "
                                                "pass; the cluster-resident kernel (sa_cluster_kernel: state and "
                                                "threshold tables in distributed shared memory, hardware cluster "
                                                "barrier between passes) ties at this size because only 7 x 16 SMs "
                                                "hold resident clusters, and wins below it "
                                                "(profiles/r02_sa_cluster.log)")}
        st.close()
        # the latency regime proper: 224 restarts, whole schedule in one launch of 7 clusters x 16 CTAs
        Rs = 224
        st = mcs.State(inst, K.KIND_SA, Rs, 1)
        st.init_random(1)
        l0 = inst.launches
        ms_c = timed(inst, lambda: st.sa_sweeps(sched, 1, seed=2), reps=3)
        nl = (inst.launches - l0) // 4
        os.environ["MCS_CLUSTER"] = "0"
        try:
            ms_m = timed(inst, lambda: st.sa_sweeps(sched, 1, seed=2), reps=3)
        finally:
            os.environ.pop("MCS_CLUSTER", None)
        out["cfg2"]["small_batch"] = {
            "workload": "the same anneal with %d restarts (latency regime)" % Rs, "launches_per_anneal": nl,
            "cluster_resident": {"value": Rs * tau * NSPINS / (ms_c * 1e-3), "us_per_colour_pass": 1e3 * ms_c / (2 * tau)},
            "one_launch_per_pass": {"value": Rs * tau * NSPINS / (ms_m * 1e-3), "us_per_colour_pass": 1e3 * ms_m / (2 * tau)},
            "unit": UNIT}
        st.close()
    if "refdyn" in which:  # cfg3 shape, the reference's own visiting order in distribution
        P, R, S = 64, 4096, 4
        inst.set_dynamics("reference")
        try:
            A, B = np.linspace(3.0, 1e-8, SCHED)[_strided(SCHED, S)].copy(), np.ones(S)
            st = mcs.State(inst, K.KIND_PIQMC, R, P)
            st.init_random(1)
            ms = timed(inst, lambda: st.piqmc_sweeps(A, B, 1, 1.0 / P, seed=2), reps=2)
            v = R * S * P * NSPINS / (ms * 1e-3)
            out["cfg3_reference_dynamics"] = {
                "workload": "80x80 PIQMC P=64, %d anneals, %d sweeps spread over the schedule, dynamics=reference "
                            "(fresh random permutation per slice, sequential visits, slices in order; parity tier c "
                            "two-sided vs the reference: tests/test_gpu_refdyn.py)" % (R, S),
                "value": v, "unit": UNIT, "ms": ms,
                "roofline": hbm_roofline(v, 0.25, peaks, "refdyn_ising_kernel<u64,false,4>", ms,
                                         "latency / barrier bound: dependency waves of ~550 sites, one CTA per anneal")}
            st.close()
        finally:
            inst.set_dynamics("colored")
    if "exact" in which:  # bit-exact sequential replay of the reference (glibc rand stream, fp64)
        P, R, S = 64, 4096, 4
        confs = np.repeat((2 * np.random.RandomState(0).randint(2, size=(R, NSPINS, 1)) - 1).astype(np.int8), P, axis=2)
        A, B = np.linspace(3.0, 1e-8, SCHED)[_strided(SCHED, S)].copy(), np.ones(S)
        dt = 1e30
        for rep in range(2):  # the first call also pays for the page faults of the 1.7 GB host buffer
            t0 = time.perf_counter()
            mcs.qmc.QuantumAnneal(A, B, 1, 1.0 / P, confs, inst, 1, exact=True, libc_seed=1000)
            dt = min(dt, time.perf_counter() - t0)
        out["cfg3_exact_replay"] = {
            "workload": "80x80 PIQMC P=64, %d anneals, %d sweeps, exact=True: bit-exact replay of the reference's "
                        "trajectories (Fisher-Yates from glibc rand(), sequential fp64 visits) -- one warp per anneal, "
                        "state in shared memory, rand() stream 31 values per step, shuffle iterations and visits in "
                        "conflict-free windows; wall clock of the C-ABI call incl. 1.7 GB of host copies each way "
                        "(the kernel alone: 7.8e9 attempts/s, profiles/r02_exact_probe.log)" % (R, S),
            "value": R * S * P * NSPINS / dt, "unit": UNIT, "ms": dt * 1e3,
            "roofline": {"bound": "latency", "note": "one sequential chain per anneal by construction (the rand() "
                                                     "stream and the visiting order are serial): about 490 "
                                                     "dependent windows of ~1700 cycles per slice sweep, 3 warps "
                                                     "per SM (shared memory); parity tier (b), not a throughput path"}}
        del confs
    if "cfg4" in which:  # SVMC on Chimera C16, 2048 reads
        cn = chimera_instance(16)
        ci = mcs.Instance(cn, device=inst.device)
        s = np.linspace(1e-3, 1.0, 1000)
        A4, B4 = 3.0 * (1 - s), s
        st = mcs.State(ci, K.KIND_SVMC, 2048, 1)
        st.init_random(0)
        ms = timed(ci, lambda: st.svmc_sweeps(A4, B4, 1, 0.1, tf=False, seed=5), reps=3)
        v = 2048 * 1000 * 2048 / (ms * 1e-3)
        ci.set_dynamics("reference")  # the reference's visiting order in distribution (svmc.pyx:83-91)
        try:
            st.init_random(0)
            sel = _strided(1000, 50)
            ms_rd = timed(ci, lambda: st.svmc_sweeps(A4[sel].copy(), B4[sel].copy(), 1, 0.1, tf=False, seed=5), reps=2)
            v_rd = 2048 * 50 * 2048 / (ms_rd * 1e-3)
        finally:
            ci.set_dynamics("colored")
        out["cfg4"] = {"workload": "svmc.SpinVectorMonteCarloCompact on Chimera C16 (2048 rotors, %d colours), 2048 "
                                   "reads, A=3(1-s), B=s, s=linspace(1e-3,1,1000), T=0.1" % ci.ncolors,
                       "value": v, "unit": UNIT, "ms": ms,
                       "reference_dynamics": {"value": v_rd, "unit": UNIT, "ms": ms_rd, "sweeps": 50,
                                              "kernel": "refdyn_svmc_kernel (dependency waves, one CTA per read)"},
                       "roofline": hbm_roofline(v, 8.0, peaks, "svmc_pass_kernel", ms / (1000 * ci.ncolors),
                                                "SURVEY 8d: 8 B per attempt (theta and cos theta, fp32, read + write); "
                                                "the 33.5 MB state is L2 resident, the kernel is bound by the latency "
                                                "of the dependent neighbour loads, not by HBM")}
        st.close()
        ci.close()
    if "cfg5" in which:  # dense SK, tensor-core local fields + sequential decisions; Swendsen-Wang on the same state
        n, P5, R5, S5 = 2048, 32, 128, 20
        di = mcs.Instance(sk_instance(n), device=inst.device)
        A5, B5 = np.linspace(3.0, 1e-8, 200)[_strided(200, S5)].copy(), np.ones(S5)
        st = mcs.State(di, K.KIND_PIQMC, R5, P5)
        st.init_random(1)
        ms = timed(di, lambda: st.piqmc_sweeps(A5, B5, 1, 1.0 / P5, global_moves=True, seed=2), reps=2)
        cols = R5 * P5
        flops = S5 * 2.0 * n * n * cols  # 2 N^2 flop per replica-slice column and sweep
        tpeak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        ach = flops / (ms * 1e-3) / 1e12
        o = {"workload": "dense SK N=2048 PIQMC P=32 with world-line moves, %d replicas (%d columns), %d sweeps: "
                         "local fields by tcgen05 GEMM, sequential in-block decisions" % (R5, cols, S5),
             "value": S5 * n * cols / (ms * 1e-3), "unit": UNIT, "ms": ms, "ms_per_sweep": ms / S5,
             "roofline": {"bound": "tensor", "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak,
                          "traffic": None, "kernel": "dense_block_kernel_tc",
                          "algorithmic_flops_per_sweep": 2.0 * n * n * cols,
                          "executed_bf16_tflops": 2 * ach,
                          "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1400",
                          "note": "J is split hi + lo into two bf16 operands (16 mantissa bits): the tensor pipe "
                                  "executes twice the algorithmic flops; the sweep is bound by the 2048 sequential "
                                  "site decisions, not by the GEMM"}}
        # the same sweeps with 256 replicas: 128 column groups of 64 = 128 of the 148 SMs busy instead of 64
        st2 = mcs.State(di, K.KIND_PIQMC, 2 * R5, P5)
        st2.init_random(1)
        ms2 = timed(di, lambda: st2.piqmc_sweeps(A5, B5, 1, 1.0 / P5, global_moves=True, seed=2), reps=2)
        o["replicas_256"] = {"value": S5 * n * 2 * cols / (ms2 * 1e-3), "unit": UNIT, "ms_per_sweep": ms2 / S5,
                             "tensor_frac": 2 * flops / (ms2 * 1e-3) / 1e12 / tpeak}
        st2.close()
        try:
            st.cluster_moves(1.0, 1.0, 1.0 / P5, nmoves=1, seed=3)
            ms_sw = timed(di, lambda: st.cluster_moves(1.0, 1.0, 1.0 / P5, nmoves=2, seed=3, sweep_offset=1), reps=1) / 2
            o["swendsen_wang"] = {"ms_per_move": ms_sw, "cluster_sites_per_s": n * cols / (ms_sw * 1e-3),
                                  "kernel": "cluster_{init,union,flip}_kernel (GPU union-find)"}
        except Exception as e:  # noqa: BLE001
            o["swendsen_wang"] = {"error": repr(e)}
        out["cfg5"] = o
        st.close()
        di.close()
    return out


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

    # final result of the resident batch: best slice per anneal on the device (fixed-order fp64 energies -> arg-min
    # -> that slice's spins), then the collective the path has -- all device buffers, no world-line download
    e_dev = torch.empty(R, dtype=torch.float64, device=dev)
    k_dev = torch.empty(R, dtype=torch.int32, device=dev)
    c_dev = torch.empty((R, NSPINS), dtype=torch.int8, device=dev)
    st.best_into(e_dev.data_ptr(), k_dev.data_ptr(), c_dev.data_ptr())
    inst.synchronize()
    e_all, best, best_conf = parallel.gather_best_device(e_dev, c_dev, lo, R_total)
    energies = e_all.cpu().numpy()

    # ---- e2e: host buffers through the one-shot C-ABI call (H2D + tile + sweeps + best slice + D2H) + the gather
    e2e, e2e_full = None, None
    L = mcs._lib.load()
    if args.e2e_steps > 0:
        rs = np.random.RandomState(rank)
        h_in = mcs.empty_pinned((R, NSPINS), np.int8)
        h_in[...] = (2 * rs.randint(2, size=(R, NSPINS)) - 1).astype(np.int8)
        h_e = mcs.empty_pinned((R, P_SLICES), np.float64)
        h_eb = mcs.empty_pinned((R,), np.float64)
        h_kb = mcs.empty_pinned((R,), np.int32)
        h_cf = mcs.empty_pinned((R, NSPINS), np.int8)
        times = []
        for it in range(args.e2e_steps + 1):  # first one is warm-up
            barrier()
            t0 = time.perf_counter()
            mcs._lib.check(L.mcs_piqmc_anneal_best(
                inst._h, mcs._lib.dptr(A), mcs._lib.dptr(B), S, 1, temp, h_in.ctypes.data, 1, R, P_SLICES, 0,
                seed + it, lo, mcs._lib.dptr(h_e), mcs._lib.dptr(h_eb), h_kb.ctypes.data_as(mcs._lib.c_i32p),
                h_cf.ctypes.data))
            if dist is not None:  # the path's one exchange step, inside the timed region
                parallel.gather_best(h_eb, h_cf, lo, R_total, device=dev)
            barrier()
            times.append(time.perf_counter() - t0)
        tt = torch.tensor([float(np.median(times[1:]))], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": attempts_step_all / float(tt.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(R) * NSPINS,
               "d2h_bytes_per_step": int(R) * (NSPINS + P_SLICES * 8 + 12),
               "ms_per_step": 1e3 * float(tt.item()), "ms_each_rank0": [round(1e3 * x, 1) for x in times[1:]],
               "api": "mcs_piqmc_anneal_best (C ABI one-shot, the example's protocol santoro80.py:286-296: pinned int8 "
                      "[R,N] start states in -> tiled over the slices on the device -> anneal -> float64 energies "
                      "[R,P], best energy / slice [R] and best configuration int8 [R,N] out)"
                      + ("; + parallel.gather_best (NCCL all_gather of the best energies, broadcast of the winner)"
                         if dist is not None else "")}
        del h_in, h_e, h_eb, h_kb, h_cf
        # the drop-in shaped call: full world lines both ways
        if args.e2e_full_steps > 0:
            host = mcs.empty_pinned((R, NSPINS, P_SLICES), np.int8)
            e_host = mcs.empty_pinned((R, P_SLICES), np.float64)
            s0 = (2 * rs.randint(2, size=(R, NSPINS, 1)) - 1).astype(np.int8)
            times = []
            for it in range(args.e2e_full_steps + 1):
                host[...] = s0  # fresh anneal: broadcast over slices (not timed: input preparation)
                barrier()
                t0 = time.perf_counter()
                mcs._lib.check(L.mcs_piqmc_anneal(inst._h, mcs._lib.dptr(A), mcs._lib.dptr(B), S, 1, temp,
                                                  host.ctypes.data, R, P_SLICES, 0, seed + it, lo,
                                                  mcs._lib.dptr(e_host)))
                barrier()
                times.append(time.perf_counter() - t0)
            tt = torch.tensor([float(np.median(times[1:]))], dtype=torch.float64, device=dev)
            if dist is not None:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_full = {"value": attempts_step_all / float(tt.item()), "unit": UNIT,
                        "h2d_bytes_per_step": int(R) * NSPINS * P_SLICES,
                        "d2h_bytes_per_step": int(R) * NSPINS * P_SLICES + int(R) * P_SLICES * 8,
                        "ms_per_step": 1e3 * float(tt.item()),
                        "api": "mcs_piqmc_anneal (drop-in shaped: pinned int8 [R,N,P] world lines in/out + energies)"}
            del host, e_host
        inst.trim()

    if rank == 0:
        peaks = load_peaks()
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # dominant kernel: piqmc_lut_pass_kernel, one launch per colour class per sweep.
        # algorithmic bytes: 0.25 B per attempt (read + write of one bit-packed spin), DESIGN.md section 4
        # (a colour pass of a batch of 512+ anneals is TWO launches, one per replica chunk, alternating on two streams:
        # a launch covers half the batch and consecutive launches overlap; ms_per_launch = step time / launches)
        n_pass = launches - args.steps  # minus the init kernel of each step
        ms_per_launch = ms_dev / max(n_pass, 1)
        bytes_per_launch = 0.25 * NSPINS * R * P_SLICES * S * args.steps / max(n_pass, 1)
        achieved = bytes_per_launch / (ms_per_launch * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u64",
            "dtype_detail": "bit-packed spins in u64 words (bit k = Trotter slice k); u32 acceptance thresholds from f32 energies; u32 Philox",
            "data": name + ", Philox-initialised spins",
            "config": {"workload": workload_name(R_total, world, S)},
            "timing": {"l2": "flushed between timed steps (256 MiB memset); state %.0f MB per GPU" % (
                           R * NSPINS * 8 / 1e6),
                       "timer": "CUDA events on the launch stream (mcs_timer_*), max over ranks",
                       "wall_s_timed_region": wall},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline(achieved, peak, bool(peaks), R, ms_per_launch, bytes_per_launch, n_pass, args),
            "e2e": e2e,
            "e2e_full_confs": e2e_full,
            "result": {"best_residual_energy_per_spin": None, "mean_best_slice_energy": float(np.mean(energies)),
                       "best_anneal": int(best),
                       "best_conf_energy_check": None},
        }
        gs = os.path.join(ROOT, "tests", "golden", "santoro80.npz")
        if os.path.isfile(gs):
            egs = float(np.load(gs)["e_gs_per_spin"])
            line["result"]["best_residual_energy_per_spin"] = float(np.min(energies)) / NSPINS - egs
            line["result"]["mean_residual_energy_per_spin"] = float(np.mean(energies)) / NSPINS - egs
        # the broadcast configuration really has the winning energy (host recomputation, sparse)
        try:
            sc = best_conf.cpu().numpy().astype(np.float64)
            idx, Jc = nbs[:, :, 0].astype(int), nbs[:, :, 1]
            line["result"]["best_conf_energy_check"] = float(0.5 * np.sum(Jc * sc[:, None] * sc[idx])) - float(
                np.min(energies))
        except Exception:  # noqa: BLE001
            pass
        which = [c for c in (args.configs or "").split(",") if c] if world == 1 else []
        if which:
            try:
                line["configs"] = run_configs(mcs, inst, peaks, which)
            except Exception as e:  # noqa: BLE001 -- the headline line must still be printed
                line["configs"] = {"error": repr(e)}
        # CPU baselines (rank 0, N = 1 only), AFTER every GPU measurement and in a separate interpreter: loading all
        # host cores first left the pinned buffers of the e2e leg on slower pages (e2e +15 % when it ran first)
        line["cpu_baseline"] = None
        if world == 1 and args.cpu_sweeps > 0:
            cpu = cpu_baseline_subprocess(args.cpu_sweeps, [c for c in which if c.startswith("cfg") and "_" not in c])
            line["cpu_baseline"] = cpu.get("cfg3")
            for c, v in cpu.items():
                if c != "cfg3" and isinstance(line.get("configs"), dict) and c in line["configs"]:
                    line["configs"][c]["cpu_baseline"] = v
            if isinstance(line.get("configs"), dict):
                for c in ("cfg3_reference_dynamics", "cfg3_exact_replay"):
                    if c in line["configs"]:
                        line["configs"][c]["cpu_baseline"] = cpu.get("cfg3")
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
    ap.add_argument("--e2e-full-steps", type=int, default=2, help="timed calls of the drop-in shaped e2e variant")
    ap.add_argument("--cpu-sweeps", type=int, default=40, help="sweeps per core for the CPU baseline sample")
    ap.add_argument("--configs", default="cfg1,cfg2,refdyn,exact,cfg4,cfg5",
                    help="extra entries measured on rank 0 at N = 1 (empty: none)")
    ap.add_argument("--cpu-configs", default="", help="(reference arm) also time these configs' reference functions")
    ap.add_argument("--traffic", type=float, default=None,
                    help="ncu dram__bytes_read+write per launch of the dominant kernel; default: the committed "
                         "capture profiles/r0*_piqmc_lut_pass_ncu.json scaled to this run's replicas")
    args = ap.parse_args()
    out = {}
    with _StdoutToStderr():
        rc = run_reference(args, out) if args.impl == "reference" else run_ours(args, out)
    if "line" in out:
        print(json.dumps(out["line"]), flush=True)
    return rc


if __name__ == "__main__":
    sys.exit(main())
