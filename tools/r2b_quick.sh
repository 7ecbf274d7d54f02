#!/bin/bash
# quick check used while tuning: the SGBM tests, then kernel times of cfg 2 / sgbm.yml as shipped / cfg 4
python -m pytest tests/test_gpu_parity.py tests/test_property.py -m gpu -x -q -k "sgbm or full_size or byte_form or random" 2>&1 | tail -2
python tools/prof_step.py 148 cfg2 2>/dev/null | head -7
python tools/prof_step.py 74 shipped 2>/dev/null | head -6
python tools/prof_step.py 14 cfg4 2>/dev/null | head -6
