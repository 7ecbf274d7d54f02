set -x
python tools/prof_step.py 148 cfg2 > gpurun_out/r2c_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on --launch-skip 44 --launch-count 10 -f -o gpurun_out/r2c_cfg2_full python tools/prof_step.py 148 cfg2 > gpurun_out/r2c_ncu_cfg2.log 2>&1
python tools/prof_bm.py > gpurun_out/r2c_bm_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on --launch-skip 10 --launch-count 5 -f -o gpurun_out/r2c_bm_full python tools/prof_bm.py > gpurun_out/r2c_ncu_bm.log 2>&1
python bench.py --steps 2 --warmup 3 --configs none --no-cpu-baseline > gpurun_out/r2c_launch_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 2 --warmup 3 --configs none --no-cpu-baseline > gpurun_out/r2c_launch_ncu.log 2>&1
