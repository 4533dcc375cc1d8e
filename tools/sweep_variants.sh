#!/bin/bash
# A/B of libb200zk variants built with tools/build_variant.sh: tools/sweep_variants.sh NAME... ("default" = the shipped library)
for v in "$@"; do
  lib=$PWD/halo2-plonky2-verifier_b200/libb200zk_$v.so
  [ "$v" = default ] && lib=$PWD/halo2-plonky2-verifier_b200/libb200zk.so
  B200ZK_LIB=$lib python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/v_$v.json 2> gpurun_out/v_$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/v_$v.json").read().strip().splitlines()[-1])
print("$v", d["value"], d["stages_ms"]["msm"], d["kernels_in_profiled_step"]["msm_accumulate"]["ms"])
PY
done
