mkdir -p gpurun_out
for bn in 128 192; do echo "MR_GEMM_BN=$bn"; MR_GEMM_BN=$bn timeout 200 python tools/encoder_check.py 2>&1 | grep -E "gemm time 262|gemm time 8192|embed batch 16384|FAIL|ALL"; done
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for c in C3 C4 C5; do timeout 400 python bench.py --config $c --steps 50 > gpurun_out/bench_r2_final_$c.json 2>gpurun_out/bench_$c.err; tail -1 gpurun_out/bench_$c.err; done
python - <<PY
import json
for c in ("C3","C4","C5"):
    try:
        j=json.load(open(f"gpurun_out/bench_r2_final_{c}.json"))
        r=j["roofline"]
        print(c, round(j["value"]/1e6,1), j["ms_per_step"], "roofline", r.get("kernel"), r.get("bound"), r.get("frac"), r.get("lanes_active_of_32"), j.get("roofline_note"))
        if j.get("query_encoder"): print("   encoder:", json.dumps(j["query_encoder"])[:1200])
    except Exception as ex: print(c, "ERR", ex)
PY
