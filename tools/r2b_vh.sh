echo default; for k in cfg2 sgbm_yml_d128 cfg4_hh_1080p_d256 cfg5_4k_d256; do python tools/time_step.py $k 10; done
for f in 20 21 24; do echo "cfg2 fpc $f"; MVSV_VH_FRAMES=$f python tools/time_step.py cfg2 10; done
python -m pytest tests/test_gpu_parity.py tests/test_property.py -m gpu -x -q -k "sgbm or full_size or byte_form or random or twin or soak or batch" 2>&1 | tail -2
