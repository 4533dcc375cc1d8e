for c in ${CS:-19 18 21 20}; do
  B200ZK_TABLE_BITS=$c python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c$c.json 2> gpurun_out/c$c.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/c$c.json").read().strip().splitlines()[-1])
print($c, d["value"], d["stages_ms"]["msm"], d["kernels_in_profiled_step"]["msm_accumulate"]["ms"], d["msm"]["ms"])
PY
done
