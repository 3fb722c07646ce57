# 8-GPU box: mega-request (C5) at 8 / 4 / 2 / 1 members, then the default bench at 8 and 2 GPUs
mkdir -p gpurun_out
run() { # n config out extra
  if [ "$1" = "1" ]; then timeout 600 python bench.py --gpus 1 --config $2 $4 > gpurun_out/$3 2>gpurun_out/$3.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --config $2 $4 > gpurun_out/$3 2>gpurun_out/$3.err; fi
  grep '^{"metric"' gpurun_out/$3 | tail -1 > gpurun_out/$3.line; python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/$3.line").read()); print("$2 N=$1", j["value"], j["ms_per_step"], j.get("latency"), j["parity"])
except Exception as ex: print("$2 N=$1 ERR", ex)
PY
}
run 8 C5 bench_r2_c5_n8.json "--steps 30"
run 4 C5 bench_r2_c5_n4.json "--steps 30"
run 2 C5 bench_r2_c5_n2b.json "--steps 30"
run 1 C5 bench_r2_c5_n1b.json "--steps 30 --no-extras"
