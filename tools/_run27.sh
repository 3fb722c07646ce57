mkdir -p gpurun_out
t0=$(date +%s)
timeout 600 python -m pytest tests/test_gbdt_gpu.py -m gpu -q -x 2>&1 | tail -3
echo "gbdt tests: $(( $(date +%s) - t0 )) s"
for tag in alt noalt; do
  if [ $tag = noalt ]; then export MR_NO_SLIM_ALT=1; fi
  for c in C3 C4; do
    MR_DEBUG_LAUNCH=1 timeout 300 python bench.py --config $c --steps 50 > gpurun_out/ab2_${c}_$tag.json 2> gpurun_out/ab2_${c}_$tag.err; grep "slim scorer" gpurun_out/ab2_${c}_$tag.err | sort | uniq -c | sort -rn | head -3
  done
done
unset MR_NO_SLIM_ALT
python - <<PY
import json
for c in ("C3","C4"):
  for tag in ("alt","noalt"):
    try:
        j=json.load(open(f"gpurun_out/ab2_{c}_{tag}.json"))
        r=j["roofline"]
        print(c, tag, round(j["value"]/1e6,1), "M/s", round(j["ms_per_step"],4), "ms; e2e", round(j["e2e"]["value"]/1e6,1), "parity", all(v for k,v in j["parity"].items() if k!="checked_items"), "| roofline", r.get("kernel"), r.get("kernel_ms"), r.get("frac"))
        print("    ", [(k["kernel"], round(k["ms_per_step"]*1e3,1)) for k in j["kernels"]])
    except Exception as ex: print(c, tag, "ERR", ex)
PY
echo "total: $(( $(date +%s) - t0 )) s"
