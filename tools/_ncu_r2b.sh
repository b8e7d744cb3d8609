set -x
python tools/ncu_forward.py 256 fp16 2 > gpurun_out/ncu_fwd_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"umma|csar_tail|rowconv" --launch-skip 25 --launch-count 25 -f -o gpurun_out/r2b_full python tools/ncu_forward.py 256 fp16 2 > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/r2b_full.ncu-rep --page raw --csv > gpurun_out/r2b_full_raw.csv 2>/dev/null
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/b_for_ncu.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
ls -la gpurun_out/r2b_*
