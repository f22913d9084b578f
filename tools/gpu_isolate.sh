#!/bin/bash
# Which kernel family / launch feature breaks the CUDA-graph network tests?  (debugging aid)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {   # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m pytest tests/test_network_gpu.py -m gpu -x -q -p no:cacheprovider -k "${K:-graph_int8 and resnet50-uniform8}" > gpurun_out/iso_$name.log 2>&1
  echo "$name exit $? : $(tail -1 gpurun_out/iso_$name.log)"
}
run base X=0
run no_c1 HAWQ_B200_C1=0
run no_halo HAWQ_B200_HALO=0
run no_pdl HAWQ_B200_PDL=0
run no_c1_no_halo HAWQ_B200_C1=0 HAWQ_B200_HALO=0
K="graph_int8 and resnet50-uniform4" run u4_base X=0
K="graph_int8 and resnet50-uniform4" run u4_no_c1 HAWQ_B200_C1=0
K="graph_int8 and resnet50-uniform4" run u4_no_halo HAWQ_B200_HALO=0
K="graph_int8 and resnet50-uniform4" run u4_no_pdl HAWQ_B200_PDL=0
