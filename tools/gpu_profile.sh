#!/bin/bash
# (gpurun brings back at most 64 MiB: a --set full report is ~2.7 MB per kernel, hence the -c limits)
# ncu evidence for profiles/ (run under gpurun, 1 GPU): launch list with DRAM bytes of one ResNet-50 step, then --set full
# captures of the tensor-core kernels and the BatchNorm kernels.  Usage: bash tools/gpu_profile.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off --replay-mode application --csv --log-file gpurun_out/launches_${TAG}.csv \
    python tools/ncu_step.py > gpurun_out/ncu_launches_${TAG}.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'umma_kernel|halo_conv_kernel|wgrad_halo_kernel|stem_dgrad_kernel' -c 8 -f -o gpurun_out/prof_umma_${TAG} \
    python tools/ncu_step.py > gpurun_out/ncu_full_umma_${TAG}.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'bn_|col_reduce' -c 5 -f -o gpurun_out/prof_bn_${TAG} \
    python tools/ncu_step.py > gpurun_out/ncu_full_bn_${TAG}.log 2>&1
ls -la gpurun_out | tail -12
# the stem trio and the max-pool (the layers furthest below their HBM roofline: 3x224x224 -> 64x112x112 -> 64x56x56)
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'stem_fprop_kernel|stem_dgrad_kernel|stem_wgrad_kernel|maxpool_idx|bn_relu_pool' -c 7 -f -o gpurun_out/prof_stem_${TAG} \
    python tools/ncu_step.py > gpurun_out/ncu_full_stem_${TAG}.log 2>&1
ls -la gpurun_out | grep ${TAG}
