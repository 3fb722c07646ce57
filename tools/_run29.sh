mkdir -p gpurun_out
t0=$(date +%s)
timeout 300 python -m pytest tests/test_gbdt_gpu.py -m gpu -q -x -k "binary_model" 2>&1 | tail -2
cap() {  # name, kernel regex, config, launch skip, env
  env $5 timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $4 -c 1 -f -o gpurun_out/ncu_r2b_$1 python bench.py --config $3 --steps 3 --warmup 3 --no-extras > /dev/null 2>&1
  python tools/ncu_digest.py gpurun_out/ncu_r2b_$1.ncu-rep 8 > gpurun_out/ncu_r2b_$1_digest.txt 2>&1
  echo "== $1"; cat gpurun_out/ncu_r2b_$1_digest.txt
}
cap slim_c2 gbdt_score_slim C2 4 MR_X=1
cap slim_c3 gbdt_score_slim C3 4 MR_NO_SLIM_ALT=1
cap slim_c3_alt gbdt_score_slim C3 4 MR_X=1
echo "total: $(( $(date +%s) - t0 )) s"
