mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/s12_tests.log; cat gpurun_out/s12_tests.log
(timeout 400 python bench.py --steps 100 --no-cpu-baseline --profile-out gpurun_out/s12_prof.json 2>&1 | tail -1) > gpurun_out/s12_bench.log; cut -c1-200 gpurun_out/s12_bench.log
