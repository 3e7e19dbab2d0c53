#!/bin/bash
# Round-2 final evidence for profiles/ (run on the GPU box through gpurun): launch list of two timed HDL-64 scans, full captures of the
# solver step kernel, the fused stage-C kernel and the frame-batched k-NN kernel, bench records of the three workloads, the device-side
# timeline of one solve and the Cholesky profile.
mkdir -p gpurun_out
LIO_BENCH_TRACE=1 python bench.py --steps 30 --warmup 5 > gpurun_out/r2b_bench_hdl64.json 2> gpurun_out/r2b_trace_hdl64.txt
python bench.py --workload vlp16 --steps 30 --warmup 5 > gpurun_out/r2b_bench_vlp16.json 2> /dev/null
python bench.py --workload stress128 --steps 8 --warmup 3 > gpurun_out/r2b_bench_stress128.json 2> /dev/null
python scripts/chol_profile.py 171 > gpurun_out/r2b_chol_profile.txt 2>&1
LIO_BENCH_PROFILE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/r2b_launches_hdl64.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2b_ncu_launch.log 2>&1
LIO_BENCH_PROFILE=1 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:k_step -s 2 -c 1 \
  -f -o gpurun_out/r2b_k_step_hdl64 python bench.py --steps 1 --warmup 3 > gpurun_out/r2b_ncu_kstep.log 2>&1
LIO_BENCH_PROFILE=1 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:asm_ppp -c 2 \
  -f -o gpurun_out/r2b_asm_ppp_hdl64 python bench.py --steps 1 --warmup 3 > gpurun_out/r2b_ncu_asm.log 2>&1
LIO_BENCH_PROFILE=1 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:knn_plane -c 1 \
  -f -o gpurun_out/r2b_knn_hdl64 python bench.py --steps 1 --warmup 3 > gpurun_out/r2b_ncu_knn.log 2>&1
ls -la gpurun_out/r2b_* | tail -20
