"""e2e items/s of mr_rank (C2 workload, page-locked buffers) as a function of MR_RANK_SLICE_ITEMS."""
import os, subprocess, sys, json
for s in ("65536", "131072", "262144", "524288", "2000000"):
    env = dict(os.environ, MR_RANK_SLICE_ITEMS=s)
    out = subprocess.run([sys.executable, "bench.py", "--steps", "5", "--warmup", "3"], env=env, capture_output=True, text=True).stdout
    try:
        d = json.loads(out.strip().splitlines()[-1])
        print(s, round(d["e2e"]["value"] / 1e6, 1), "M items/s e2e;", round(d["value"] / 1e6, 1), "device", flush=True)
    except Exception as ex:
        print(s, "failed", ex, out[-300:], flush=True)
