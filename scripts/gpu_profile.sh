#!/bin/bash
# Round profile recipe (run under gpurun): plain bench first, then the ncu launch list of the SAME
# command, then one --set full capture per dominant kernel.  Outputs land in gpurun_out/.
set -u
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-sustained --also map"
$CMD > $OUT/bench_prof_plain.json 2> $OUT/bench_prof_plain.err || { echo "plain run failed"; tail -5 $OUT/bench_prof_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_list.log 2>&1
echo "launch list rc=$?"
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:project_fold -s 3 -c 1 -f -o $OUT/prof_fold $CMD > $OUT/ncu_fold.log 2>&1
echo "fold rc=$?"
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:map_h_kernel -s 1 -c 1 -f -o $OUT/prof_maph $CMD > $OUT/ncu_maph.log 2>&1
echo "map_h rc=$?"
CMD3="python bench.py --workload c3 --no-also --no-cpu --steps 2 --warmup 3 --c3-total 262144"
$CMD3 > $OUT/bench_prof_c3.json 2>/dev/null && ncu --set full --clock-control none --import-source on -k regex:project_fold -s 2 -c 1 -f -o $OUT/prof_c3 $CMD3 > $OUT/ncu_c3.log 2>&1
echo "c3 rc=$?"
python scripts/fused_gather_probe.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:project_fold -s 2 -c 1 -f -o $OUT/prof_foldg python scripts/fused_gather_probe.py > $OUT/ncu_foldg.log 2>&1
echo "fold+gather rc=$?"
ls -la $OUT | grep -E "prof|launches"
