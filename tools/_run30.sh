mkdir -p gpurun_out
t0=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "gpu tests: $(( $(date +%s) - t0 )) s"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_r2c_n1.json 2> gpurun_out/bench_r2c_n1.err; tail -2 gpurun_out/bench_r2c_n1.err
python - <<PY
import json
j=json.load(open("gpurun_out/bench_r2c_n1.json"))
r=j["roofline"]
print("C2", round(j["value"]/1e6,1), "M/s", round(j["ms_per_step"],4), "ms; e2e", round(j["e2e"]["value"]/1e6,1), "serial", round(j["e2e"]["one_call_at_a_time"]/1e6,1), j["e2e"]["parity_ok"], "| roofline", r.get("kernel"), r.get("kernel_ms"), r.get("frac"), "clocks", j.get("clocks"))
print("    ", [(k["kernel"], round(k["ms_per_step"]*1e3,1), k.get("frac_hbm")) for k in j["kernels"]])
print("   churn", j.get("churn"))
for c,x in j.get("other_configs",{}).items():
    if "error" in x: print(c, x); continue
    print(c, round(x["value"]/1e6,1), "M/s", round(x["ms_per_step"],4), "e2e", round(x["e2e"]["value"]/1e6,1), x["e2e"].get("one_call_at_a_time"), x["e2e"]["parity_ok"], x["parity"], x.get("latency"))
    print("    ", [(k["kernel"], round(k["ms_per_step"]*1e3,1)) for k in x["kernels"]])
PY
echo "total: $(( $(date +%s) - t0 )) s"
