set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "sampling_modes or nonfinite or corr_lookup or windowed or bilinear or flow_decoder" 2>&1 | tail -30
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8
