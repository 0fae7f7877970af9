timeout 900 python -m pytest tests/test_binning_gpu.py -q 2>&1 | tail -30
