mkdir -p gpurun_out
timeout 300 python tools/encoder_check.py > gpurun_out/encoder_check.log 2>&1; grep -E "gemm time|embed batch|      |FAIL|ALL|Error|error" gpurun_out/encoder_check.log
timeout 600 python -m pytest tests/test_features_gpu.py tests/test_encoder_gpu.py -m gpu -q -x 2>&1 | tail -3
