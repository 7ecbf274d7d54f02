#!/bin/bash
# final round-2 artefacts: tests, bench (both arms), full ncu capture of one cfg-2 step, launch lists of the bench command
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r2d_tests.log
python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/r2d_tests.log 2>&1
python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
python bench.py --impl reference > gpurun_out/r2d_bench_ref.json 2>> gpurun_out/r2d_bench.err
python tools/prof_step.py 148 cfg2 > gpurun_out/r2d_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on --launch-skip 44 --launch-count 10 -f -o gpurun_out/r2d_cfg2_full python tools/prof_step.py 148 cfg2 > gpurun_out/r2d_ncu_cfg2.log 2>&1
python bench.py --steps 2 --warmup 3 --configs none --no-cpu-baseline > gpurun_out/r2d_launch_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 2 --warmup 3 --configs none --no-cpu-baseline > gpurun_out/r2d_launch_ncu.log 2>&1
MVSV_SERIAL=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2d_launches_serial.csv python bench.py --steps 2 --warmup 3 --configs none --no-cpu-baseline > gpurun_out/r2d_launch_ncu_serial.log 2>&1
