for i in 1 2 3 4 5 6; do timeout 900 python -m pytest tests/test_parity_gpu.py -q -k "full_size_backward_vs_f64 or backward_twice" 2>&1 | grep -E "^E  |passed|failed" | head -5; done
