# round-2 GPU pass: tests, bench lines, launch list (run under gpurun from the repo root): bash tools/r2_run.sh TAG
TAG=${1:-r2}
set -x
(time python -m pytest tests -m gpu -q) > gpurun_out/${TAG}_tests.log 2>&1; tail -15 gpurun_out/${TAG}_tests.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -2 gpurun_out/${TAG}_bench.err; cut -c1-300 gpurun_out/${TAG}_bench.json
python bench.py --workload spherenet20 --steps 20 > gpurun_out/${TAG}_spherenet20.json 2> gpurun_out/${TAG}_spherenet20.err; cut -c1-300 gpurun_out/${TAG}_spherenet20.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-graph > /dev/null 2>&1
wc -l gpurun_out/${TAG}_launches.csv
