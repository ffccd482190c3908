#!/bin/sh
# DRAM bytes per segment of k_extend (roofline.traffic), measured with ncu on one screenshot of a bench workload.
# usage (on the GPU box): tools/extend_traffic.sh config2 1   ->  gpurun_out/traffic_config2.csv / .json
# then here:              python tools/extend_traffic.py gpurun_out/traffic_config2   (merges into profiles/extend_traffic.json)
set -e
W=${1:-config2}; F=${2:-1}
mkdir -p gpurun_out
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    --clock-control none -k regex:'k_extend|k_shade' --csv --log-file gpurun_out/traffic_$W.csv \
    python tools/traffic_run.py --workload $W --frames $F > gpurun_out/traffic_$W.json
tail -1 gpurun_out/traffic_$W.json
