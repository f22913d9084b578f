#!/bin/bash
# One GPU visit (no rebuild: the in-tree .so travels with the snapshot).  usage: tools/gpu_visit.sh <steps...>
#   probe | halo | kernels | network | bench[:arch:scheme[:batch]] | full | ncu | ncufull:<regex>
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for what in "$@"; do
  case $what in
    probe)
      nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/probe_b200 tools/probe_b200.cu && PROBE_TMA_ONLY=${PROBE_TMA_ONLY:-} timeout 200 /tmp/probe_b200 > gpurun_out/probe_b200.txt 2>&1
      echo "probe exit $?"; tail -30 gpurun_out/probe_b200.txt;;
    c1)
      timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "conv1x1" > gpurun_out/pytest_c1.log 2>&1
      echo "c1 exit $?"; tail -15 gpurun_out/pytest_c1.log;;
    dualk)
      timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "dual" > gpurun_out/pytest_dual.log 2>&1
      echo "dual exit $?"; tail -15 gpurun_out/pytest_dual.log;;
    stem)
      timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "stem" > gpurun_out/pytest_stem.log 2>&1
      echo "stem exit $?"; tail -15 gpurun_out/pytest_stem.log;;
    halo)
      timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "halo" > gpurun_out/pytest_halo.log 2>&1
      echo "halo exit $?"; tail -15 gpurun_out/pytest_halo.log;;
    kernels)
      timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_kernels.log 2>&1
      echo "kernels exit $?"; tail -12 gpurun_out/pytest_kernels.log;;
    network)
      timeout 1200 python -m pytest tests/test_network_gpu.py -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_network.log 2>&1
      echo "network exit $?"; tail -8 gpurun_out/pytest_network.log;;
    full)
      timeout 1800 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu_full.log 2>&1
      echo "pytest full exit $?"; tail -8 gpurun_out/pytest_gpu_full.log
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log;;
    bench*)
      IFS=: read -r _ arch scheme batch a4 <<< "$what"
      arch=${arch:-resnet50}; scheme=${scheme:-uniform8}; batch=${batch:-128}; a4=${a4:-byte}
      tag=${arch}_${scheme}_b${batch}${TAG:-}; [[ $a4 == packed ]] && tag=${tag}_packed
      timeout 600 python bench.py --arch $arch --scheme $scheme --batch $batch --a4-storage $a4 --steps ${STEPS:-50} --warmup 5 --no-cpu-baseline --detail gpurun_out/detail_$tag.json > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
      echo "bench $tag exit $?"; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$tag.json"))
    print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", round(d["e2e"]["value"]), "roof", d.get("roofline", {}).get("frac"), d.get("clocks"))
except Exception as e:
    print("no bench line:", e)
PY
      tail -3 gpurun_out/bench_$tag.err;;
    ncu)
      timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench.log 2>&1
      echo "ncu list exit $?"
      python tools/summarize_ncu.py gpurun_out/launches.csv "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-roofline" --json gpurun_out/ncu_traffic_entry.json > gpurun_out/ncu_launch_summary.txt 2>&1
      tail -12 gpurun_out/ncu_launch_summary.txt;;
    ncufull:*)
      rx=${what#ncufull:}
      timeout 600 ncu --set full --import-source on --clock-control none -k regex:$rx -s ${NCU_SKIP:-8} -c ${NCU_COUNT:-4} -o gpurun_out/prof_$rx -f python tools/profile_forward.py --forwards 2 > gpurun_out/prof_$rx.log 2>&1
      echo "ncu full exit $?";;
  esac
done
