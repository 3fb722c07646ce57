mkdir -p gpurun_out
t0=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "gpu tests: $(( $(date +%s) - t0 )) s"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_r2_final2_n1.json 2> gpurun_out/bench_r2_final2_n1.err; tail -2 gpurun_out/bench_r2_final2_n1.err
for c in C3 C4 C5; do timeout 400 python bench.py --config $c --steps 50 > gpurun_out/bench_r2_final2_$c.json 2>gpurun_out/bench_r2_final2_$c.err; tail -1 gpurun_out/bench_r2_final2_$c.err; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2_final2.csv python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/launches_bench2.log 2>&1
tail -3 gpurun_out/launches_r2_final2.csv | cut -c1-300
python - <<PY
import json
def load(f):
    for ln in reversed(open(f).read().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)
j=load("gpurun_out/bench_r2_final2_n1.json")
r=j["roofline"]
print("C2", round(j["value"]/1e6,1), "M/s", round(j["ms_per_step"],4), "ms; e2e", round(j["e2e"]["value"]/1e6,1), "serial", round(j["e2e"]["one_call_at_a_time"]/1e6,1), j["e2e"]["parity_ok"], "| roofline", r.get("kernel"), r.get("bound"), r.get("kernel_ms"), r.get("frac"), r.get("smem",{}).get("frac"), "clocks", j.get("clocks",{}).get("sm_mhz"), j.get("latency"), "cpu", j["cpu_baseline"]["value"], j["cpu_baseline"]["cores"])
print("    ", [(k["kernel"], round(k["ms_per_step"]*1e3,1), k.get("frac_hbm")) for k in j["kernels"]])
print("   churn", j.get("churn"))
for c in ("C3","C4","C5"):
    x=load(f"gpurun_out/bench_r2_final2_{c}.json")
    print(c, "full:", round(x["value"]/1e6,1), "M/s", round(x["ms_per_step"],4), "e2e", round(x["e2e"]["value"]/1e6,1), round(x["e2e"].get("one_call_at_a_time",0)/1e6,1), x["e2e"]["parity_ok"], x.get("latency"), x["roofline"].get("bound"), x["roofline"].get("frac"), "cpu", x["cpu_baseline"]["value"])
    print("    ", [(k["kernel"], round(k["ms_per_step"]*1e3,1)) for k in x["kernels"]])
PY
echo "total: $(( $(date +%s) - t0 )) s"
