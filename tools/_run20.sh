mkdir -p gpurun_out
timeout 300 python tools/encoder_check.py > gpurun_out/encoder_check.log 2>&1; grep -E "gemm time|embed batch|      |FAIL|ALL|Error|error" gpurun_out/encoder_check.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:encoder_gemm -s 2 -c 1 -f -o gpurun_out/ncu_r2_gemm2 python tools/gemm_prof.py 262144 1152 384 > /dev/null 2>&1; ls -la gpurun_out/ncu_r2_gemm2.ncu-rep
timeout 600 python -m pytest tests/test_encoder_gpu.py -m gpu -q -x 2>&1 | tail -3
