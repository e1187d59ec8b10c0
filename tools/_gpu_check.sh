mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_minkloc.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/s10_tests.log; cat gpurun_out/s10_tests.log
(timeout 300 python tools/microbench_conv.py 2>&1 | tail -7) 
(timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/s10_prof.json 2>&1 | tail -1 | cut -c1-200)
