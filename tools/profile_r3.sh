#!/bin/bash
# Round-3 profiles (one GPU).  Every command runs once without ncu first; ncu only if that run exited 0.
set -u
O=gpurun_out
NCU="ncu --clock-control none"
# 1. launch list of the default bench command (shares, not absolutes: cold caches, serialised)
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cfg1 --sampler-tune 3 --sampler-draws 2 --e2e-evals 4"
$B > $O/r3_bench_plain.log 2>&1 && $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/r3_bench_launch_list.csv $B > $O/r3_bench_ncu.log 2>&1
# 2. full captures of the evaluation kernels
for cfg in "cfg3 --C 64" "cfg2 --C 4" "cfg3sh --C 64 --sharded"; do
  set -- $cfg; name=$1; shift
  python tools/run_lik.py "$@" > $O/r3_lik_${name}_plain.log 2>&1 && $NCU --set full --import-source on -k regex:logp_grad -s 3 -c 1 -o $O/prof_r3_lik_${name} -f python tools/run_lik.py "$@" > $O/r3_lik_${name}_ncu.log 2>&1
done
# 3. full captures of the persistent sampler kernels (64 ticks per launch)
for v in 2; do
  BGH_TICK=$v python tools/run_sampler.py --tune 6 > $O/r3_tick${v}_plain.log 2>&1 && BGH_TICK=$v $NCU --set full --import-source on -k regex:nuts_tick -s 3 -c 1 -o $O/prof_r3_tick${v} -f python tools/run_sampler.py --tune 6 > $O/r3_tick${v}_ncu.log 2>&1
done
# the reports are too large to travel back: keep the raw metric page (one row per launch) and drop the reports
for f in $O/prof_r3_*.ncu-rep; do
  b=$(basename $f .ncu-rep)
  ncu -i $f --page raw --csv > $O/${b}_raw.csv 2>/dev/null
  ncu -i $f --page source --csv --print-source sass > $O/${b}_sass.csv 2>/dev/null
  python tools/ncu_top_lines.py $O/${b}_sass.csv 40 > $O/${b}_top_stalls.txt 2>/dev/null
  rm -f $f $O/${b}_sass.csv
done
ls -la $O/prof_r3_* | awk '{print $5, $9}'
