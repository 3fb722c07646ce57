mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_r2_final_n1.json 2>gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err
for c in C3 C4 C5; do timeout 400 python bench.py --config $c --steps 50 > gpurun_out/bench_r2_final_$c.json 2>/dev/null; done
python - <<PY
import json
j=json.load(open("gpurun_out/bench_r2_final_n1.json"))
print("C2", round(j["value"]/1e6,1), j["ms_per_step"], "e2e", round(j["e2e"]["value"]/1e6,1), j["roofline"].get("frac"), j["parity"], j.get("latency"))
for k in j["kernels"]: print("   ", k["kernel"], round(k["ms_per_step"]*1000,1))
for c,e in (j.get("other_configs") or {}).items(): print(c, e.get("value"), e.get("error"))
for c in ("C3","C4","C5"):
    try:
        j=json.load(open(f"gpurun_out/bench_r2_final_{c}.json"))
        print(c, round(j["value"]/1e6,1), j["ms_per_step"], "e2e", round(j["e2e"]["value"]/1e6,1), j["parity"], j.get("latency"))
        for k in j["kernels"]: print("   ", k["kernel"], round(k["ms_per_step"]*1000,1))
        if j.get("query_encoder"): print("   encoder:", json.dumps(j["query_encoder"])[:900])
    except Exception as ex: print(c, "ERR", ex)
PY
