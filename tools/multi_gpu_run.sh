mkdir -p gpurun_out
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 30 > gpurun_out/bench_r2_final2_n$N.json 2> gpurun_out/bench_r2_final2_n$N.err; tail -2 gpurun_out/bench_r2_final2_n$N.err
python - <<PY
import json
def load(f):
    for ln in reversed(open(f).read().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)
j=load("gpurun_out/bench_r2_final2_n$N.json")
print(j["n_gpus"], "C2", round(j["value"]/1e6,1), "M/s", round(j["ms_per_step"]*1e3,1), "us; e2e", round(j["e2e"]["value"]/1e6,1), j["e2e"]["parity_ok"], j["rank_ms_per_step"])
for c,x in j.get("other_configs",{}).items():
    if "error" in x: print("   ", c, x); continue
    print("   ", c, round(x["value"]/1e6,1), "M/s", round(x["ms_per_step"]*1e3,1), "us e2e", round(x["e2e"]["value"]/1e6,1), x["e2e"]["parity_ok"], x["parity"], x.get("latency"))
    print("      ", [(k["kernel"], round(k["ms_per_step"]*1e3,1)) for k in x["kernels"]])
PY
