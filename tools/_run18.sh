mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_features_gpu.py tests/test_encoder_gpu.py tests/test_rank_api_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --config C3 --no-extras --steps 20 > gpurun_out/bench_r2_C3_l.json 2>gpurun_out/c3.err; tail -3 gpurun_out/c3.err
python - <<PY
import json
j=json.load(open("gpurun_out/bench_r2_C3_l.json")); print(j["value"], j["ms_per_step"], j["parity"], j["e2e"]["value"])
for k in j["kernels"]: print("   ", k["kernel"], round(k["ms_per_step"]*1000,1))
PY
timeout 300 python tools/encoder_check.py 2>&1 | grep -E "embed batch|FAIL|ALL"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:encoder_gemm -s 2 -c 1 -f -o gpurun_out/ncu_r2_gemm python tools/gemm_prof.py 262144 1152 384 > /dev/null 2>&1; ls -la gpurun_out/ncu_r2_gemm.ncu-rep
