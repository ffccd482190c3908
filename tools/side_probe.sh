sh tools/ab4.sh "RT_SHADE_DEFER=0" "RT_SHADE_DEFER=2" "" 2>&1 | tee gpurun_out/ab_r2e_6.txt
for e in "RT_SHADE_DEFER=3" "RT_SHADE_DEFER=2" "RT_SHADE_DEFER=0"; do
  env $e python bench.py --steps 1 --warmup 1 --no-cpu --no-microbench > gpurun_out/side_$e.json 2>gpurun_out/side_$e.err
  python - "$e" gpurun_out/side_$e.json <<'PY' | tee -a gpurun_out/ab_r2e_6.txt
import json,sys
d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
for k in ("config2","config4"):
    c=d[k]; print(f"[side block {k} {sys.argv[1]}] {c['value']:.0f} Mrays/s {c['ms_per_step']:.1f} ms/step shade share {c['shade_roofline']['share_of_step']:.3f}")
print(f"[4k {sys.argv[1]}] {d['value']:.0f}")
PY
done
