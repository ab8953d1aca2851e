timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "input_check or curves or zones" 2>&1 | tail -12
