"""Host-only throughput of mr_requests_decode on a large array of RankingEvents (no GPU needed).
usage: python tools/bench_decode.py [threads ...]      e.g.  python tools/bench_decode.py 1 4 auto
DESIGN.md quotes: 2000 events x 100 items, 30 `number` features (4.5 MB of JSON)."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metarank_b200 import _capi  # noqa: E402

lib = _capi.lib()
names = [f"f{j}" for j in range(30)]
doc = json.dumps({"features": [dict(name=n, type="number", scope="item", source=f"metadata.{n}") for n in names],
                  "model_features": names}).encode()
h = C.c_void_p()
assert lib.mr_schema_create(None, doc, C.c_size_t(len(doc)), C.byref(h)) == 0
R = 2000
body = json.dumps([dict(event="ranking", id=f"r{r}", timestamp=1622505601000 + r, user=f"u{r % 50}", session=f"s{r % 70}", fields=[],
                        items=[dict(id=f"item{(r * 131 + i * 17) % 200000}") for i in range(100)]) for r in range(R)]).encode()
print(f"{len(body) / 1e6:.2f} MB, {R} events x 100 items")
out = C.c_void_p()
for thr in sys.argv[1:] or ["1", "auto"]:
    if thr == "auto":
        os.environ.pop("MR_DECODE_THREADS", None)
    else:
        os.environ["MR_DECODE_THREADS"] = thr
    ts = []
    for _ in range(7):
        t0 = time.perf_counter()
        st = lib.mr_requests_decode(h, body, C.c_size_t(len(body)), C.byref(out))
        ts.append(time.perf_counter() - t0)
        assert st == 0
        lib.mr_requests_free(out)
    ts.sort()
    print(f"threads {thr}: min {ts[0] * 1e3:.1f} ms, median {ts[3] * 1e3:.1f} ms = {R * 100 / ts[0] / 1e6:.1f} M items/s")
