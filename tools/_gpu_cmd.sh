timeout -s KILL 400 python -m pytest tests/test_spgemm_gpu.py -x -q -m gpu -k "register_sort or er_config or select_max or committed or random" > gpurun_out/r2_pytest_gpu_z1.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu_z1.log
timeout -s KILL 400 python tools/sweep.py --scale 22 --set regsort_packed=1 --set regsort_packed=0 > gpurun_out/r2_sweep_z.log 2>&1; cat gpurun_out/r2_sweep_z.log
timeout -s KILL 200 python tools/er_bench.py > gpurun_out/r2_er_z.log 2>&1; tail -2 gpurun_out/r2_er_z.log
