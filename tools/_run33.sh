mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python -m pytest tests/test_gbdt_gpu.py tests/test_group_gpu.py tests/test_features_gpu.py -m gpu -q -x 2>&1 | tail -3
echo "tests: $(( $(date +%s) - t0 )) s"
timeout 400 python bench.py --config C5 --steps 50 > gpurun_out/bench_r2e_C5.json 2>gpurun_out/bench_r2e_C5.err; tail -1 gpurun_out/bench_r2e_C5.err
python - <<PY
import json
def load(f):
    for ln in reversed(open(f).read().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)
x=load("gpurun_out/bench_r2e_C5.json")
print("C5", round(x["value"]/1e6,1), "M/s", round(x["ms_per_step"]*1e3,1), "us e2e", round(x["e2e"]["value"]/1e6,1), x["e2e"]["parity_ok"], x["parity"], x.get("latency"))
print("    ", [(k["kernel"], round(k["ms_per_step"]*1e3,1)) for k in x["kernels"]])
PY
echo "total: $(( $(date +%s) - t0 )) s"
