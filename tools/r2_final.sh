# final round-2 GPU pass (run under gpurun from the repo root): bash tools/r2_final.sh
set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
(time python -m pytest tests -m gpu -q) > gpurun_out/r2_final_tests.log 2>&1; tail -6 gpurun_out/r2_final_tests.log
python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -2 gpurun_out/r2_final_bench.err; cut -c1-220 gpurun_out/r2_final_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_bench_reference.json 2>> gpurun_out/r2_final_bench.err; cut -c1-200 gpurun_out/r2_final_bench_reference.json
for w in vgg16_prune_cycle spherenet20 resnet50; do python bench.py --workload $w --steps 20 > gpurun_out/r2_final_$w.json 2>> gpurun_out/r2_final_bench.err; cut -c1-200 gpurun_out/r2_final_$w.json; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-graph > /dev/null 2>&1
wc -l gpurun_out/r2_launches.csv
