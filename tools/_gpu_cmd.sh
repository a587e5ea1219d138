timeout -s KILL 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu_w.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu_w.log
timeout -s KILL 300 python bench.py > gpurun_out/r2_bench_w.json 2> gpurun_out/r2_bench_w.err; tail -c 600 gpurun_out/r2_bench_w.json
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:num_sacc2 -c 3 -o gpurun_out/r2_prof_sacc2_s22 python tools/prof_driver.py --scale 22 --phases 18 --slab 9 --reps 1 > gpurun_out/r2_prof_sacc2_s22.log 2>&1; tail -2 gpurun_out/r2_prof_sacc2_s22.log
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:merge2_tma -c 2 -o gpurun_out/r2_prof_merge2_tma python tools/merge_bench.py --scale 17 --reps 1 > gpurun_out/r2_prof_merge2_tma.log 2>&1; tail -2 gpurun_out/r2_prof_merge2_tma.log
timeout -s KILL 200 python tools/er_bench.py > gpurun_out/r2_er_w.log 2>&1; tail -2 gpurun_out/r2_er_w.log
