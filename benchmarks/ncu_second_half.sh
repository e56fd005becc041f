# --set full captures of the kernels added in the second half of round 2 (one B200; outputs under gpurun_out/)
ncu --set full --import-source on --clock-control none -k regex:histogram_copies_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2_hist python benchmarks/hbm_kernels.py --only-hist > gpurun_out/r2_hist_ncu.log 2>&1
ncu -i gpurun_out/r2_hist.ncu-rep --page raw --csv > gpurun_out/r2_hist_raw.csv 2>/dev/null
ncu --set full --import-source on --clock-control none -k regex:jacobi_block_kernel --launch-skip 2 --launch-count 2 -f -o gpurun_out/r2_eigblk python benchmarks/eig_large.py --dims 256 --reps 1 > gpurun_out/r2_eigblk_ncu.log 2>&1
ncu -i gpurun_out/r2_eigblk.ncu-rep --page raw --csv > gpurun_out/r2_eigblk_raw.csv 2>/dev/null
tail -n 2 gpurun_out/r2_hist_ncu.log; tail -n 2 gpurun_out/r2_eigblk_ncu.log
