mkdir -p gpurun_out
t0=$(date +%s)
timeout 600 python -m pytest tests/test_gbdt_gpu.py -m gpu -q -x 2>&1 | tail -3
echo "gbdt tests: $(( $(date +%s) - t0 )) s"
# A/B: root table + in-loop categorical vs the previous form, same box
for tag in new noroot; do
  if [ $tag = noroot ]; then export MR_NO_ROOT_TAB=1 MR_NO_CAT16=1; fi
  timeout 300 python bench.py --steps 50 --no-extras > gpurun_out/ab_C2_$tag.json 2> gpurun_out/ab_C2_$tag.err; tail -1 gpurun_out/ab_C2_$tag.err
  timeout 300 python bench.py --config C3 --steps 50 > gpurun_out/ab_C3_$tag.json 2> gpurun_out/ab_C3_$tag.err; tail -1 gpurun_out/ab_C3_$tag.err
done
unset MR_NO_ROOT_TAB MR_NO_CAT16
python - <<PY
import json
for c in ("C2","C3"):
  for tag in ("new","noroot"):
    try:
        j=json.load(open(f"gpurun_out/ab_{c}_{tag}.json"))
        r=j["roofline"]
        print(c, tag, round(j["value"]/1e6,1), "M/s", round(j["ms_per_step"],4), "ms; e2e", round(j["e2e"]["value"]/1e6,1), "parity", j["parity"], "| roofline", r.get("kernel"), r.get("kernel_ms"), r.get("frac"), r.get("lanes_active_of_32"))
        print("    ", [(k["kernel"], round(k["ms_per_step"]*1e3,1), k.get("frac_hbm")) for k in j["kernels"]])
    except Exception as ex: print(c, tag, "ERR", ex)
PY
echo "total: $(( $(date +%s) - t0 )) s"
