#!/bin/sh
# Strong-scaling lines on N GPUs of one box: the default 4K job (frame split + ncclReduce) and BASELINE config 4
# (10 M triangles, image-tile split + gather).  usage: tools/scale_run.sh N [tag]   -> gpurun_out/<tag>_{4k,c4tiles}_nN.json
N=${1:-2}; TAG=${2:-scale}
run() {  # name, extra bench args
  name=$1; shift
  if [ "$N" = 1 ]; then
    timeout 900 python bench.py --gpus 1 "$@" > gpurun_out/${TAG}_${name}_n$N.json 2> gpurun_out/${TAG}_${name}_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N "$@" > gpurun_out/${TAG}_${name}_n$N.json 2> gpurun_out/${TAG}_${name}_n$N.err
  fi
  echo "$name N=$N rc=$?"; tail -c 300 gpurun_out/${TAG}_${name}_n$N.err
  python - gpurun_out/${TAG}_${name}_n$N.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["config"]["name"], "N", d["n_gpus"], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 1),
          "e2e", round(d["e2e"]["value"], 1), "crc", d.get("checksum"), "gate", d.get("parity_gate"))
except Exception as e:
    print("no line:", e)
PY
}
run 4k --steps 3 --warmup 3 --no-side --no-cpu
run c4tiles --workload config4 --split tiles --steps 2 --warmup 1 --no-side --no-cpu --no-microbench
