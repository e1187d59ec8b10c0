for PF in 0 1; do
(EGN_PREFETCH=$PF timeout 400 python bench.py --steps 150 --no-cpu-baseline --profile-out gpurun_out/s21_prof$PF.json 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('PREFETCH',$PF, d['value'], d['e2e']['value'], d['ms_per_step'])")
done
