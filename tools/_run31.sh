mkdir -p gpurun_out
t0=$(date +%s)
nvidia-smi -L | head -4
timeout 300 python -m pytest tests/test_group_gpu.py -m gpu -q -x 2>&1 | tail -2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 50 > gpurun_out/bench_r2c_n2.json 2> gpurun_out/bench_r2c_n2.err; tail -2 gpurun_out/bench_r2c_n2.err
timeout 300 $TR bench.py --gpus 2 --config C5 --steps 50 > gpurun_out/bench_r2c_c5_n2.json 2> gpurun_out/bench_r2c_c5_n2.err; tail -2 gpurun_out/bench_r2c_c5_n2.err
timeout 300 python bench.py --config C5 --steps 50 > gpurun_out/bench_r2c_c5_n1_samebox.json 2> gpurun_out/bench_r2c_c5_n1.err
timeout 300 python bench.py --steps 30 --no-extras > gpurun_out/bench_r2c_n1_samebox.json 2> gpurun_out/bench_r2c_n1_samebox.err
python - <<PY
import json
def show(f):
    try:
        j=json.load(open(f))
    except Exception as ex:
        print(f, "ERR", ex); return
    print(f, j["n_gpus"], j["config"]["workload"][:3], round(j["value"]/1e6,1), "M/s", round(j["ms_per_step"]*1e3,1), "us; e2e", round(j["e2e"]["value"]/1e6,1), "serial/gpu", round(j["e2e"]["one_call_at_a_time"]/1e6,1), j["e2e"]["parity_ok"], j.get("latency"), j["rank_ms_per_step"])
    print("    ", [(k["kernel"], round(k["ms_per_step"]*1e3,1)) for k in j["kernels"]])
    for c,x in j.get("other_configs",{}).items():
        if "error" in x: print("   ", c, x); continue
        print("   ", c, round(x["value"]/1e6,1), "M/s", round(x["ms_per_step"]*1e3,1), "us e2e", round(x["e2e"]["value"]/1e6,1), x["e2e"]["parity_ok"], x.get("latency"))
for f in ("bench_r2c_n2","bench_r2c_c5_n2","bench_r2c_c5_n1_samebox","bench_r2c_n1_samebox"):
    show(f"gpurun_out/{f}.json")
PY
echo "total: $(( $(date +%s) - t0 )) s"
