#!/bin/bash
# One GPU visit to validate and A/B-measure the opt-in kernel variants written at the end of round 1
# (profiles/r01/experiments/README.md).  Everything lands in gpurun_out/.   usage: bash tools/validate_variants.sh
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; exit 1; }
run_tests() {   # name, env assignment, -k expression
  env $2 timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "$3" > gpurun_out/variant_$1_kernels.log 2>&1
  echo "$1 kernels exit $?"; tail -3 gpurun_out/variant_$1_kernels.log
  env $2 timeout 300 python -m pytest tests/test_network_gpu.py -m gpu -x -q -p no:cacheprovider -k "not every_activation" > gpurun_out/variant_$1_network.log 2>&1
  echo "$1 network exit $?"; tail -3 gpurun_out/variant_$1_network.log
}
bench() {       # name, env assignment
  env $2 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --detail gpurun_out/variant_$1_detail.json > gpurun_out/variant_$1_bench.json 2> gpurun_out/variant_$1_bench.err
  python - <<PY
import json
d = json.load(open("gpurun_out/variant_$1_bench.json"))
print("$1", round(d["value"]), "img/s", round(d["ms_per_step"], 4), "ms/step, e2e", round(d["e2e"]["value"]))
PY
}
bench baseline "HAWQ_B200_X=0"
run_tests lean "HAWQ_B200_MMA_FAST=1" "requant or retiled"
bench lean "HAWQ_B200_MMA_FAST=1"
run_tests epi16 "HAWQ_B200_EPI16=1" "residual or dual"
bench epi16 "HAWQ_B200_EPI16=1"
run_tests stem "HAWQ_B200_STEM_PERSIST=1" "stem"
bench stem "HAWQ_B200_STEM_PERSIST=1"
bench all "HAWQ_B200_MMA_FAST=1 HAWQ_B200_EPI16=1 HAWQ_B200_STEM_PERSIST=1"
