timeout 900 python -m pytest tests/test_gpu_parity_r2.py -m gpu -x -q -k "move" 2>&1 | tail -30
