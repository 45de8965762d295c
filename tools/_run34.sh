compute-sanitizer --tool memcheck --error-exitcode 9 python tools/memcheck_r2.py > gpurun_out/r2_memcheck.txt 2>&1; echo rc=$?; tail -6 gpurun_out/r2_memcheck.txt
compute-sanitizer --tool racecheck --error-exitcode 9 python tools/memcheck_r2.py > gpurun_out/r2_racecheck.txt 2>&1; echo rc=$?; tail -6 gpurun_out/r2_racecheck.txt
