mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/s19_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s19_launches.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_sconv_ts|k_conv0_tc|k_nbr_down|k_level_emit|k_eca_apply|k_pool_partial" --launch-skip 56 --launch-count 56 -o /tmp/s19_full python tools/profile_forward.py --iters 2 > gpurun_out/s19_ncu.log 2>&1; tail -1 gpurun_out/s19_ncu.log
ncu -i /tmp/s19_full.ncu-rep --page raw --csv > gpurun_out/s19_full_raw.csv 2>/dev/null
du -sh gpurun_out
