# one --set full capture of train_fused_kernel<512> (C4 model, batch 1024) with source correlation
ncu --set full --import-source on --clock-control none -k regex:train_fused_kernel --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2_fused python benchmarks/train_fused_check.py > gpurun_out/r2_fused_ncu.log 2>&1
ncu -i gpurun_out/r2_fused.ncu-rep --page raw --csv > gpurun_out/r2_fused_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_fused.ncu-rep --page source --csv > gpurun_out/r2_fused_source.csv 2>/dev/null
tail -2 gpurun_out/r2_fused_ncu.log
