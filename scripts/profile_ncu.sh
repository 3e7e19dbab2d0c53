#!/bin/bash
# ncu evidence for profiles/ (run on the GPU box through gpurun): launch list of two timed scans, a full capture of the
# k-NN + plane-fit kernel inside the bench, and a full capture of the fused stage-C kernel on the > L2 stream.
mkdir -p gpurun_out
LIO_BENCH_PROFILE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/r1c_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
LIO_BENCH_PROFILE=1 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:knn_plane -c 3 \
  -f -o gpurun_out/r1c_knn python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_knn.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:asm_ppp -c 3 -f -o gpurun_out/r1c_asm_stream \
  python -c "
from lio_mapping_b200 import estimator
print(estimator.asm_stream_bench(1<<24, 2))" > gpurun_out/ncu_asm.log 2>&1
ls -la gpurun_out | tail
tail -3 gpurun_out/ncu_asm.log
