timeout -s KILL 700 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu_final.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu_final.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_final.log 2>&1; tail -1 gpurun_out/r2_smoke_final.log
timeout -s KILL 300 python bench.py > gpurun_out/r2_bench_final_n1.json 2> gpurun_out/r2_bench_final_n1.err; tail -c 400 gpurun_out/r2_bench_final_n1.json
timeout -s KILL 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_final_ref.json 2> gpurun_out/r2_bench_final_ref.err; tail -c 600 gpurun_out/r2_bench_final_ref.json
