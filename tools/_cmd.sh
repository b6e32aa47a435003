python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "half_row or ibin or tcgen05" 2>&1 | tail -15
