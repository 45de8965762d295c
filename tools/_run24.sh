ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_resnet50_launches.csv python tools/prof_workload.py resnet50 task1 2 > /dev/null 2>&1
wc -l gpurun_out/r2_resnet50_launches.csv
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_spherenet20_launches.csv python tools/prof_workload.py spherenet20 task1 2 > /dev/null 2>&1
wc -l gpurun_out/r2_spherenet20_launches.csv
