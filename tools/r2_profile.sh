# round-2 profiling pass (run under gpurun from the repo root): bash tools/r2_profile.sh
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-graph > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_traffic.csv python -m tools.prof_ops 1 vgg > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:"conv_gemm_kernel|wgrad_gemm_kernel" -c 15 -o gpurun_out/r2_gemm_full -f python -m tools.prof_ops 1 > /dev/null 2>&1
ncu -i gpurun_out/r2_gemm_full.ncu-rep --page raw --csv > gpurun_out/r2_gemm_full_raw.csv 2>/dev/null
ls -la gpurun_out/r2_gemm_full.ncu-rep; wc -l gpurun_out/r2_launches.csv gpurun_out/r2_traffic.csv gpurun_out/r2_gemm_full_raw.csv
