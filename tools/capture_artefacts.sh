#!/bin/bash
# final round-2 artefacts: tests, bench (both arms), full ncu capture of one cfg-2 step, launch lists of the bench command
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/final_tests.log 2>&1
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python bench.py --impl reference > gpurun_out/final_bench_ref.json 2>> gpurun_out/final_bench.err
python tools/prof_step.py 148 cfg2 > gpurun_out/final_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on --launch-skip 44 --launch-count 10 -f -o gpurun_out/final_cfg2_full python tools/prof_step.py 148 cfg2 > gpurun_out/final_ncu_cfg2.log 2>&1
python bench.py --steps 2 --warmup 3 --configs none --no-cpu-baseline > gpurun_out/final_launch_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --configs none --no-cpu-baseline > gpurun_out/final_launch_ncu.log 2>&1
MVSV_SERIAL=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_serial.csv python bench.py --steps 2 --warmup 3 --configs none --no-cpu-baseline > gpurun_out/final_launch_ncu_serial.log 2>&1
# then, here: ncu -i gpurun_out/final_cfg2_full.ncu-rep --page raw --csv > profiles/r02_ncu_full_final_raw.csv; python tools/ncu_summary.py
# (--launch-skip 44: two warm-up computes of 22 launches each -- 1 prefilter, 7 + 7 chunked cost kernel / first scan, sweep,
#  last scan, median, 4 speckle kernels -- before the profiled step of 10, which runs its kernels one after the other)
