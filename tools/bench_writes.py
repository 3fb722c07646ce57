"""Write-path throughput (host-side state arithmetic inside libmrgpu + dirty-row upload):
events/s through mr_state_apply_writes for a ranklens-like stream of click / impression interactions."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import metarank_b200 as mb
from metarank_b200 import features as F

ctx = mb.Context(0)
feats = [dict(name="ctr", type="rate", top="click", bottom="impression", bucket="24h", periods=[7, 30], normalize={"weight": 10}),
         dict(name="clicks", type="interaction_count", interaction="click", scope="item"),
         dict(name="seen", type="interacted_with", interaction="click", field=["item.tags"], scope="session", count=100, duration="24h")]
fm = F.FeatureMapping(ctx, feats, [f["name"] for f in feats])
ds = F.DeviceState(ctx, fm)
rng = np.random.Generator(np.random.PCG64(1))
N = 200_000
writes = []
t = 1_700_000_000_000
for i in range(N):
    t += int(rng.integers(0, 2000))
    item, sess = f"m{int(rng.integers(0, 20000))}", f"s{int(rng.integers(0, 5000))}"
    if rng.random() < 0.1:
        writes += [("pinc", (("item", item), "ctr_click"), t, 1), ("pinc", (("global",), "ctr_click_norm"), t, 1),
                   ("inc", (("item", item), "clicks"), t, 1), ("append", (("session", sess), "seen_interactions"), t, item)]
    else:
        writes += [("pinc", (("item", item), "ctr_impression"), t, 1), ("pinc", (("global",), "ctr_impression_norm"), t, 1)]
blob = F.pack_writes(writes)
import ctypes as C
buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
a, s = C.c_int64(0), C.c_int64(0)
t0 = time.perf_counter()
mb._capi.check(mb._capi.lib().mr_state_apply_writes(ds._h, buf, C.c_size_t(len(blob)), C.byref(a), C.byref(s)))
t1 = time.perf_counter()
ds.flush()
t2 = time.perf_counter()
print(json.dumps({"events": N, "writes": len(writes), "applied": a.value, "apply_s": t1 - t0, "flush_s": t2 - t1,
                  "events_per_s": N / (t2 - t0), "writes_per_s": len(writes) / (t2 - t0),
                  "note": "single host thread; reference docs quote 1-3 k events/s import throughput (doc/performance.md:7)"}))
