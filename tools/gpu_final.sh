#!/bin/bash
# Round-end style validation: full -m gpu suite, smoke, headline benches, reference arm, ncu launch list (+DRAM bytes) and one --set full capture.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest full exit $?"; tail -4 gpurun_out/pytest_gpu_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; echo "reference arm exit $?"; head -c 300 gpurun_out/bench_reference_arm.json; echo
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "default bench exit $?"
for cfg in resnet50:uniform8 resnet50:uniform4 resnet50:bops_0.5 resnet18:uniform4 resnet18:uniform8; do
  set -- ${cfg/:/ }
  timeout 300 python bench.py --arch $1 --scheme $2 --steps 30 --warmup 5 --no-cpu-baseline --detail gpurun_out/detail_$1_$2.json > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err
  echo "bench $1 $2 exit $?"; python - <<PY
import json
d=json.load(open("gpurun_out/bench_$1_$2.json"))
print({k:d[k] for k in ("value","ms_per_step")}, "e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["frac"],3), d["clocks"])
PY
done
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list exit $?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_tc -s 49 -c 6 -o gpurun_out/prof_tc_final -f python tools/profile_forward.py --forwards 2 > gpurun_out/prof_final.log 2>&1; echo "ncu full exit $?"
