#!/bin/bash
# One GPU visit: parity tests, smoke, short benches, ncu launch list.  Everything lands in gpurun_out/.
# usage: tools/gpu_check.sh [tests|bench|ncu|all]
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
what=${1:-all}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
if [[ $what == tests || $what == all ]]; then
  timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
fi
if [[ $what == bench || $what == all ]]; then
  for cfg in "resnet50 uniform8" "resnet50 uniform4" "resnet50 bops_0.5" "resnet18 uniform4"; do
    set -- $cfg
    timeout 600 python bench.py --arch $1 --scheme $2 --steps 20 --warmup 5 --detail gpurun_out/detail_$1_$2.json > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err
    echo "bench $1 $2 exit $?"; head -c 1500 gpurun_out/bench_$1_$2.json; echo; tail -3 gpurun_out/bench_$1_$2.err
  done
fi
if [[ $what == quick ]]; then
  timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_kernels.log 2>&1
  echo "pytest kernels exit $?"; tail -12 gpurun_out/pytest_kernels.log
  timeout 900 python -m pytest tests/test_network_gpu.py -m gpu -x -q -p no:cacheprovider -k "not every_activation" > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest network exit $?"; tail -5 gpurun_out/pytest_gpu.log
  for cfg in ${CFGS:-resnet50:uniform8}; do
    set -- ${cfg/:/ }
    timeout 600 python bench.py --arch $1 --scheme $2 --steps 20 --warmup 5 --no-cpu-baseline --detail gpurun_out/detail_$1_$2.json > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err
    echo "bench $1 $2 exit $?"; python - <<PY
import json
d=json.load(open("gpurun_out/bench_$1_$2.json"))
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["tensor_tops"])
PY
    tail -3 gpurun_out/bench_$1_$2.err
  done
fi
if [[ $what == ncu || $what == all ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
     python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench.log 2>&1
  echo "ncu exit $?"; tail -2 gpurun_out/ncu_bench.log | head -c 600
fi
