#!/bin/bash
# round-2 run 14: ncu evidence for profiles/r02 (launch list of the bench command; --set full of the kernels of C2, C3, C4, C5)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
NCU="ncu --clock-control none"
B="python bench.py --steps 2 --warmup 3 --workloads none --no-cpu"
$B > gpurun_out/r02_ncu_plain_c2.log 2>&1 &&
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02_launches_c2.csv $B > gpurun_out/r02_ncu_c2a.log 2>&1
$B > gpurun_out/r02_ncu_plain_c2b.log 2>&1 &&
$NCU --set full --import-source on -k regex:'k_triage|k_deep|k_scan_gather' -s 9 -c 3 -f -o gpurun_out/r02_prof_c2 $B > gpurun_out/r02_ncu_c2b.log 2>&1
W3="python tools/bench_workloads.py c3 --scale 0.2 --no-cpu --steps 1"
$W3 > gpurun_out/r02_ncu_plain_c3.log 2>&1 &&
$NCU --set full --import-source on -k regex:'k_deep|k_big|k_sort_sets' -s 4 -c 4 -f -o gpurun_out/r02_prof_c3 $W3 > gpurun_out/r02_ncu_c3.log 2>&1
$W3 > gpurun_out/r02_ncu_plain_c3b.log 2>&1 &&
$NCU --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section Occupancy --section LaunchStats -k regex:'k_big' -s 1 -c 1 -f -o gpurun_out/r02_prof_c3_big $W3 > gpurun_out/r02_ncu_c3b.log 2>&1
W5="python tools/bench_workloads.py c5 --scale 0.2 --no-cpu --steps 1"
$W5 > gpurun_out/r02_ncu_plain_c5.log 2>&1 &&
$NCU --set full --import-source on -k regex:'k_deep' -s 2 -c 1 -f -o gpurun_out/r02_prof_c5 $W5 > gpurun_out/r02_ncu_c5.log 2>&1
W4="python tools/bench_workloads.py c4 --no-cpu --steps 1"
$W4 > gpurun_out/r02_ncu_plain_c4.log 2>&1 &&
$NCU --set full --import-source on -k regex:'k_deep' -s 1 -c 1 -f -o gpurun_out/r02_prof_c4 $W4 > gpurun_out/r02_ncu_c4.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -8; tail -2 gpurun_out/r02_ncu_c3.log gpurun_out/r02_ncu_c5.log
