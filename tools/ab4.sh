#!/bin/sh
# A/B on BASELINE config 4 (10 M triangles, 4K, 2 of its 16 frames): one bench line per environment.
# usage: tools/ab4.sh "NAME=VAL NAME=VAL" "NAME=VAL" ...   ("" = defaults)
for envs in "$@"; do
  out=$(env $envs python bench.py --workload config4 --frames 2 --steps 2 --warmup 1 --no-cpu --no-side --no-microbench 2>&1 | tail -1)
  python - "$envs" "$out" <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2])
    r = d["roofline"]
    print(f"[config4 {sys.argv[1] or 'default'}] {d['value']:.0f} Mrays/s  {d['ms_per_step']:.1f} ms/step  extend {r['t_measured_ms']:.2f} ms/launch  "
          f"share {r['extend_share_of_step']:.3f}  visits {r['node_visits_per_segment']:.2f} tris {r['tri_tests_per_segment']:.2f} build {d['bvh']['build_ms']:.2f} ms "
          f"crc {d['checksum']} gate {d['parity_gate'].get('screenshot')}")
except Exception as e:
    print(f"[{sys.argv[1]}] FAILED: {sys.argv[2][-400:]}")
PY
done
