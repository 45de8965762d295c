set -x
(time python -m pytest tests -m gpu -x -q) > gpurun_out/f_pytest.log 2>&1; tail -4 gpurun_out/f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/f_bench.log 2> gpurun_out/f_bench.err; tail -2 gpurun_out/f_bench.err; cut -c1-400 gpurun_out/f_bench.log
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_traffic.csv python -m tests.prof_ops 1 vgg > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:"conv_gemm_kernel|wgrad_gemm_kernel|wgrad_epilogue" -c 9 -o gpurun_out/r1_gemm_full -f python -m tests.prof_ops 1 > /dev/null 2>&1
ncu -i gpurun_out/r1_gemm_full.ncu-rep --page raw --csv > gpurun_out/r1_gemm_full_raw.csv 2>/dev/null
ls -la gpurun_out/r1_gemm_full.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-graph > /dev/null 2>&1
wc -l gpurun_out/r1_launches.csv gpurun_out/r1_traffic.csv
