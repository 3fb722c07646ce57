mkdir -p gpurun_out
t0=$(date +%s)
timeout 300 python -m pytest tests/test_features_gpu.py tests/test_group_gpu.py -m gpu -q -x -k "ordering or group or mega" 2>&1 | tail -2
timeout 200 python bench.py --config C5 --steps 50 > gpurun_out/bench_r2f_C5.json 2>gpurun_out/bench_r2f_C5.err; tail -1 gpurun_out/bench_r2f_C5.err
python - <<PY
import json
def load(f):
    for ln in reversed(open(f).read().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)
x=load("gpurun_out/bench_r2f_C5.json")
print("C5", round(x["value"]/1e6,1), "M/s", round(x["ms_per_step"]*1e3,1), "us", x["parity"], x.get("latency"))
print("    ", [(k["kernel"], round(k["ms_per_step"]*1e3,1)) for k in x["kernels"]])
PY
echo "bench done: $(( $(date +%s) - t0 )) s"
timeout 420 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_features_gpu.py tests/test_gbdt_gpu.py -m gpu -q -x -k "ordering_edge or referer or upsert_batch or binary_model" > gpurun_out/sanitizer_memcheck_r2.txt 2>&1
tail -8 gpurun_out/sanitizer_memcheck_r2.txt
echo "total: $(( $(date +%s) - t0 )) s"
