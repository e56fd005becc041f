#!/bin/bash
# Round-2 profiling pass (run through gpurun on one B200); outputs under gpurun_out/, summaries are copied into profiles/ by hand.
set -x
mkdir -p gpurun_out
# 1. launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/r2_bench_under_ncu.log 2>&1
# 2. one --set full capture of the sampler at the bench's launch configuration (8 bases x 1e6 shots)
ncu --set full --import-source on --clock-control none -k regex:sampler_pair_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/r2_sampler \
    python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2_sampler_ncu.log 2>&1
ncu -i gpurun_out/r2_sampler.ncu-rep --page raw --csv > gpurun_out/r2_sampler_raw.csv 2>/dev/null
# 3. one --set full capture of the fused training kernel at batch 8192 and the per-kernel launch list of one training step
ncu --set full --import-source on --clock-control none -k regex:train_fused_kernel --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2_fused \
    python benchmarks/train_launches.py 8192 > gpurun_out/r2_fused_ncu.log 2>&1
ncu -i gpurun_out/r2_fused.ncu-rep --page raw --csv > gpurun_out/r2_fused_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv \
    --log-file gpurun_out/r2_train_launches.csv python benchmarks/train_launches.py 1024 8192 > /dev/null 2>&1
# 4. stand-alone HBM-class kernels and the recon breakdown
python benchmarks/hbm_kernels.py --out gpurun_out/r2_hbm_kernels.json > gpurun_out/r2_hbm_kernels.log 2>&1
python benchmarks/recon_breakdown.py > gpurun_out/r2_recon_breakdown.log 2>&1
python benchmarks/train_fused_stamps.py 8192 > gpurun_out/r2_fused_stamps.log 2>&1
ls -la gpurun_out | tail -20
