#!/bin/bash
# Runs ON THE GPU BOX (gpurun -- bash profiles/capture.sh): the plain command first, then the ncu launch list
# and one --set full capture of the dominant kernel for the SAME command; then (round 2) full captures of the dense
# tensor-core block kernel and of the reference-dynamics kernel from their probes.  Outputs land in gpurun_out/ and
# are turned into the committed summaries by `python profiles/summarize_ncu.py --round rNN` on the build box.
set -u
CMD="python bench.py --steps 1 --warmup 1 --sched 10 --e2e-steps 1 --e2e-full-steps 0 --cpu-sweeps 0 --configs="
mkdir -p gpurun_out
$CMD > gpurun_out/cap_plain.json 2> gpurun_out/cap_plain.err || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/cap_launches.csv \
    $CMD > gpurun_out/cap_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:piqmc_lut_pass -s 10 -c 1 -f \
    -o gpurun_out/cap_piqmc_pass $CMD > gpurun_out/cap_ncu2.log 2>&1
R=128 S=2 python benchmarks/dense_probe.py > gpurun_out/cap_dense_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dense_block_kernel_tc -s 40 -c 1 -f \
    -o gpurun_out/cap_dense env R=128 S=2 python benchmarks/dense_probe.py > gpurun_out/cap_ncu3.log 2>&1
python benchmarks/refdyn_probe.py 296 1 > gpurun_out/cap_refdyn_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:refdyn_ising_kernel -s 1 -c 1 -f \
    -o gpurun_out/cap_refdyn python benchmarks/refdyn_probe.py 296 1 > gpurun_out/cap_ncu4.log 2>&1
python benchmarks/sa_cluster_ncu.py > gpurun_out/cap_sa_cluster_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sa_cluster_kernel -s 2 -c 1 -f \
    -o gpurun_out/cap_sa_cluster python benchmarks/sa_cluster_ncu.py > gpurun_out/cap_ncu5.log 2>&1
# summaries on the box (the .ncu-rep files together exceed what gpurun copies back); the reports themselves are dropped
python profiles/summarize_ncu.py --round r02 > gpurun_out/cap_summary.log 2>&1
mkdir -p gpurun_out/profiles_out
cp profiles/r02_launch_list.csv profiles/r02_piqmc_lut_pass_ncu.json profiles/r02_piqmc_lut_pass_ncu_full.txt gpurun_out/profiles_out/
python profiles/summarize_ncu.py gpurun_out/cap_dense.ncu-rep > gpurun_out/profiles_out/r02_dense_kernel_ncu.txt 2>/dev/null
python profiles/summarize_ncu.py gpurun_out/cap_refdyn.ncu-rep > gpurun_out/profiles_out/r02_refdyn_kernel_ncu.txt 2>/dev/null
python profiles/summarize_ncu.py gpurun_out/cap_sa_cluster.ncu-rep > gpurun_out/profiles_out/r02_sa_cluster_kernel_ncu.txt 2>/dev/null
rm -f gpurun_out/cap_dense.ncu-rep gpurun_out/cap_refdyn.ncu-rep gpurun_out/cap_piqmc_pass.ncu-rep
echo "capture done: $(ls gpurun_out | grep cap_ | tr '\n' ' ')"
