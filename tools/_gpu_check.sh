mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_minkloc.py -m gpu -x -q 2>&1 | tail -7) > gpurun_out/s16_tests.log; cat gpurun_out/s16_tests.log
(timeout 400 python bench.py --steps 100 --no-cpu-baseline --profile-out gpurun_out/s16_prof.json 2>&1 | tail -1) > gpurun_out/s16_bench.log; cut -c1-200 gpurun_out/s16_bench.log
