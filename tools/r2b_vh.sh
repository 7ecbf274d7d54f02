python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chunked" 2>&1 | tail -3
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for k in cfg2 sgbm_yml_d128 cfg4_hh_1080p_d256 cfg5_4k_d256; do python tools/time_step.py $k 10; done
