mkdir -p gpurun_out
t0=$(date +%s)
timeout 600 python -m pytest tests/test_gbdt_gpu.py -m gpu -q -x -k "one_wave or slim_scorer" 2>&1 | tail -3
for tag in alt noalt; do
  if [ $tag = noalt ]; then export MR_NO_SLIM_ALT=1; fi
  for c in C3 C4; do
    MR_DEBUG_LAUNCH=1 timeout 300 python bench.py --config $c --steps 50 > gpurun_out/ab3_${c}_$tag.json 2> gpurun_out/ab3_${c}_$tag.err; grep "slim scorer" gpurun_out/ab3_${c}_$tag.err | sort | uniq -c | sort -rn | head -2
  done
done
unset MR_NO_SLIM_ALT
python - <<PY
import json
for c in ("C3","C4"):
  for tag in ("alt","noalt"):
    try:
        j=json.load(open(f"gpurun_out/ab3_{c}_{tag}.json"))
        r=j["roofline"]
        print(c, tag, round(j["value"]/1e6,1), "M/s", round(j["ms_per_step"],4), "ms; parity", all(v for k,v in j["parity"].items() if k!="checked_items"), "| roofline", r.get("kernel"), r.get("kernel_ms"), r.get("frac"))
    except Exception as ex: print(c, tag, "ERR", ex)
PY
echo "total: $(( $(date +%s) - t0 )) s"
