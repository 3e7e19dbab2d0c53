#!/bin/bash
# Round-2 ncu evidence for profiles/ (run on the GPU box through gpurun): launch list of two timed HDL-64 scans, full captures
# of the fused stage-C kernel (bench workload and > L2 stream), of the frame-batched k-NN kernel and of the solver step kernel.
mkdir -p gpurun_out
LIO_BENCH_PROFILE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/r2_launches_hdl64.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2_ncu_launch.log 2>&1
LIO_BENCH_PROFILE=1 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:asm_ppp -c 2 \
  -f -o gpurun_out/r2_asm_ppp_hdl64 python bench.py --steps 1 --warmup 3 > gpurun_out/r2_ncu_asm.log 2>&1
LIO_BENCH_PROFILE=1 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:knn_plane -c 1 \
  -f -o gpurun_out/r2_knn_hdl64 python bench.py --steps 1 --warmup 3 > gpurun_out/r2_ncu_knn.log 2>&1
LIO_BENCH_PROFILE=1 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:k_step -s 2 -c 1 \
  -f -o gpurun_out/r2_k_step_hdl64 python bench.py --steps 1 --warmup 3 > gpurun_out/r2_ncu_kstep.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:asm_ppp -c 2 -f -o gpurun_out/r2_asm_stream \
  python -c "
from lio_mapping_b200 import estimator
print(estimator.asm_stream_bench(1<<24, 2))" > gpurun_out/r2_ncu_asm_stream.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail
