#!/bin/bash
# one ncu --set full capture of the fused stage-C kernel on the realistic > L2 stream (run through gpurun)
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:asm_ppp -c 2 -f -o gpurun_out/r1c_asm_stream \
  python -c "
from lio_mapping_b200 import estimator
print(estimator.asm_stream_bench(1<<24, 2))" > gpurun_out/ncu_asm.log 2>&1
tail -2 gpurun_out/ncu_asm.log
