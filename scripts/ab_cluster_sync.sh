#!/bin/bash
# A/B on one box of the cluster kernel's synchronisation variants (spectral_cl.cuh):
#   shipped                  plain remote stores + release / acquire cluster barrier after every transposition, split write-after-read barriers at 256^2
#   -DSCL_ASYNC_HANDOVER     st.async + mbarrier (complete_tx) hand-over of the transpositions
#   -DSCL_JOINT_SYNC         write-after-read barriers as one arrive + wait after the line transform
# usage (GPU box): bash scripts/ab_cluster_sync.sh "<flags of the variant>" > gpurun_out/ab_cluster_sync.log
cd "$(dirname "$0")/.."
VAR="${1:--DSCL_ASYNC_HANDOVER}"
run() { for cfg in "256 64" "256 256" "128 64" "128 512"; do timeout 120 python scripts/spectral_time.py $cfg 200 || echo "FAILED $cfg"; done; }
echo "== shipped"; run; run
NV="nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --split-compile=0"
cp build/obj/tu_spectral_cl.o /tmp/tu_spectral_cl.o.keep; cp pyqg_generative_b200/libqgb200.so /tmp/libqgb200.so.keep
$NV $VAR -c -o build/obj/tu_spectral_cl.o pyqg_generative_b200/csrc/tu_spectral_cl.cu 2>/dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o pyqg_generative_b200/libqgb200.so build/obj/*.o
echo "== variant $VAR"; run; run
cp /tmp/tu_spectral_cl.o.keep build/obj/tu_spectral_cl.o; cp /tmp/libqgb200.so.keep pyqg_generative_b200/libqgb200.so
echo "== shipped again"; run
