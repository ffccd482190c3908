#!/bin/sh
# The evidence run of a round, on the GPU box (one B200): every file lands in gpurun_out/ with the prefix $1.
# 1 bench line (defaults)  2 reference arm  3 ncu launch list  4 ncu --set full (config 2, config 4)  5 traffic pass
# 6 bench lines of BASELINE configs 1 and 3.  Each step only after the plain command before it has exited 0.
P=${1:-final}
O=gpurun_out
mkdir -p $O
python bench.py > $O/${P}_bench.json 2> $O/${P}_bench.err || { echo "bench failed"; tail -5 $O/${P}_bench.err; exit 1; }
tail -c 400 $O/${P}_bench.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > $O/${P}_ref.json 2> $O/${P}_ref.err || echo "reference arm failed"
for w in config1 config3; do python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-side --no-microbench > $O/${P}_$w.json 2> $O/${P}_$w.err || echo "$w failed"; done
B2="python bench.py --workload config2 --steps 1 --warmup 1 --no-cpu --no-side --no-microbench"
$B2 > $O/${P}_c2_plain.json 2> $O/${P}_c2_plain.err && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/${P}_launches.csv $B2 > $O/${P}_launches.log 2>&1
T2="python tools/traffic_run.py --workload config2 --frames 1"
$T2 > $O/${P}_c2_traffic_plain.json && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shade' --launch-count 6 -f -o $O/${P}_c2_full $T2 > $O/${P}_c2_full.log 2>&1
T4="env RT_MAX_PATHS_MI=128 python tools/traffic_run.py --workload config4 --frames 1"
$T4 > $O/${P}_c4_traffic_plain.json && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shade' --launch-skip 2 --launch-count 4 -f -o $O/${P}_c4_full $T4 > $O/${P}_c4_full.log 2>&1
for w in config2 4k config4; do timeout 600 sh tools/extend_traffic.sh $w 1 2> $O/traffic_$w.err || echo "traffic $w failed"; done
# (compute-sanitizer is closed on this pool: the memcheck pass of round 1 is not repeated)
ls -la $O | tail -25
