# one --set full capture of the grouped weight-gradient GEMM (C4 model, batch 8192) with source correlation
ncu --set full --import-source on --clock-control none -k regex:gemm_tc_group_kernel --launch-skip 2 --launch-count 2 -f -o gpurun_out/r2_wgrad python benchmarks/train_launches.py 8192 > gpurun_out/r2_wgrad_ncu.log 2>&1
ncu -i gpurun_out/r2_wgrad.ncu-rep --page raw --csv > gpurun_out/r2_wgrad_raw.csv 2>/dev/null
tail -2 gpurun_out/r2_wgrad_ncu.log
