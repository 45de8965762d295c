set -x
python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -2 gpurun_out/r2_final_bench.err; cut -c1-250 gpurun_out/r2_final_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_bench_reference.json 2>> gpurun_out/r2_final_bench.err; cut -c1-400 gpurun_out/r2_final_bench_reference.json
python bench.py --workload vgg16_prune_cycle --steps 40 > gpurun_out/r2_final_prune_cycle.json 2>> gpurun_out/r2_final_bench.err; cut -c1-200 gpurun_out/r2_final_prune_cycle.json
bash tools/r2_profile.sh
