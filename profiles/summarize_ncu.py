#!/usr/bin/env python
"""Turn an .ncu-rep (brought back from the GPU box in gpurun_out/) into the small text summary that is
committed under profiles/.   python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/NAME.txt"""
import csv
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__issue_active.avg.per_cycle_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
]


def launch_list(csv_path, out_path, header):
    """ncu --metrics gpu__time_duration.sum --csv log -> compact launch list (id, kernel, grid, block, ns)."""
    rows = [r for r in csv.reader(open(csv_path)) if r and r[0] != "" and not r[0].startswith("==")]
    hdr = rows[0]
    ix = {k: i for i, k in enumerate(hdr)}
    tot, per = 0.0, {}
    with open(out_path, "w") as f:
        for h in header:
            f.write("# %s\n" % h)
        f.write("id,kernel,grid,block,gpu__time_duration_ns\n")
        for r in rows[1:]:
            if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
                continue
            ns = float(r[ix["Metric Value"]].replace(",", ""))
            if r[ix["Metric Unit"]] in ("us", "usecond"):
                ns *= 1e3
            elif r[ix["Metric Unit"]] in ("ms", "msecond"):
                ns *= 1e6
            name = r[ix["Kernel Name"]]
            name = name[:name.index("(")] if "(" in name else name
            f.write('"%s","%s","%s","%s","%d"\n' % (r[ix["ID"]], name, r[ix["Grid Size"]], r[ix["Block Size"]], ns))
            tot += ns
            per[name] = per.get(name, 0.0) + ns
        f.write("# share of the listed GPU time: " + "; ".join(
            "%s %.2f%%" % (k, 100 * v / tot) for k, v in sorted(per.items(), key=lambda kv: -kv[1])) + "\n")


def as_json(path, out_path):
    """Key numbers of the first kernel in the report, for bench.py's roofline / binding_unit fields."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]

    def val(k):
        v = float(r[hdr.index(k)].replace(",", ""))
        u = units[hdr.index(k)]
        return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
    grid = r[hdr.index("launch__grid_size")]
    d = {
        "kernel": r[hdr.index("Kernel Name")],
        "source": os.path.basename(path),
        "grid_size": int(float(grid.replace(",", ""))),
        "duration_us_under_ncu": val("gpu__time_duration.sum") if units[hdr.index("gpu__time_duration.sum")].startswith("us") else None,
        "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
        "dram_pct_of_peak": val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "warp_instructions": val("smsp__inst_executed.sum"),
        "issue_slots_busy_pct": 100.0 * val("smsp__issue_active.avg.per_cycle_active"),
        "alu_pipe_inst_pct_of_peak": val("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "fma_pipe_cycles_active_pct": val("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        "lsu_pipe_inst_pct_of_peak": val("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        "registers_per_thread": int(val("launch__registers_per_thread")),
        "block_size": int(val("launch__block_size")),
        # profiles/capture.sh profiles bench.py's cfg3 workload (4096 anneals on one GPU): a launch of the pass kernel
        # covers one 2048-replica chunk of one colour class = 3200 sites x 2048 world lines x 64 slices
        "attempts_in_launch": 3200 * 2048 * 64,
    }
    json.dump(d, open(out_path, "w"), indent=1)
    return d


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel: %s" % r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                print("  %-78s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        st = []
        for i, k in enumerate(hdr):
            if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k:
                try:
                    st.append((float(r[i].replace(",", "")), k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(v for v, _ in st) or 1.0
        print("  warp-state samples: " + ", ".join("%s %.1f%%" % (k, 100 * v / tot) for v, k in sorted(st, reverse=True)[:8]))
        print()


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--round":
        # python profiles/summarize_ncu.py --round r01 : turn gpurun_out/cap_* (profiles/capture.sh) into the
        # committed summaries
        rnd = sys.argv[2]
        here = os.path.dirname(os.path.abspath(__file__))
        go = os.path.join(os.path.dirname(here), "gpurun_out")
        launch_list(os.path.join(go, "cap_launches.csv"), os.path.join(here, rnd + "_launch_list.csv"),
                    ["ncu --metrics gpu__time_duration.sum --clock-control none -c 200 ; python bench.py --steps 1 "
                     "--warmup 1 --sched 10 --e2e-steps 1 --cpu-sweeps 0",
                     "(10-sweep schedule so that the list is short; a real step has 2000 pass launches + 1 init)"])
        rep = os.path.join(go, "cap_piqmc_pass.ncu-rep")
        with open(os.path.join(here, rnd + "_piqmc_lut_pass_ncu_full.txt"), "w") as f:
            old = sys.stdout
            sys.stdout = f
            print("# ncu --set full --clock-control none --import-source on -k regex:piqmc_lut_pass -s 10 -c 1 ; same command")
            main(rep)
            sys.stdout = old
        print(as_json(rep, os.path.join(here, rnd + "_piqmc_lut_pass_ncu.json")))
    else:
        main(sys.argv[1])
