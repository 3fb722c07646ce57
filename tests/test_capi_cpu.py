"""CPU-side checks of libmrgpu.so: it loads, exports every symbol the header declares,
its host parsers accept/reject what the oracle's parsers accept/reject, and the GPU
entry points fail loudly (no CPU fallback) when no device is present."""
import ctypes as C

import numpy as np
import pytest

import metarank_b200 as mb
from metarank_b200 import _capi, synth
from metarank_b200.booster import inspect_model
from oracle import model_parse


def test_library_exports_every_declared_symbol():
    lib = _capi.lib()
    names = _capi.declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
    assert b"sm_100a" in lib.mr_version()


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mb.MrError) as e:
        mb.Context(0)
    assert e.value.status == 8 and "no CPU fallback" in e.value.message


@pytest.mark.parametrize("kind,blob", [
    (0, synth.lightgbm_model_text(500, 30, seed=1236)),
    (0, synth.lightgbm_model_text(9, 6, cat_features={2: 40}, stump_every=3, zero_missing=True)),
    (1, synth.xgboost_model_json(200, 16, depth=6)),
    (1, synth.xgboost_model_ubj(20, 16, depth=8, full=False)),
])
def test_host_parser_agrees_with_oracle_parser_on_shape(kind, blob):
    inf = inspect_model(kind, blob)
    om = model_parse.parse_lightgbm_text(blob) if kind == 0 else model_parse.parse_xgboost(blob)
    assert inf.n_features == om["n_features"]
    assert inf.n_trees == len(om["trees"])
    if kind == 0:
        assert inf.n_internal_nodes == sum(t["num_leaves"] - 1 for t in om["trees"])
        assert inf.max_leaves == max(t["num_leaves"] for t in om["trees"])
        assert inf.has_categorical == int(any((t["decision_type"] & 1).any() for t in om["trees"]))
    else:
        assert inf.n_internal_nodes == sum(int((t["left"] != -1).sum()) for t in om["trees"])
    assert inf.device_bytes % 16 == 0 and inf.n_chunks >= 1


def test_chunk_budget_controls_chunk_count():
    blob = synth.lightgbm_model_text(500, 30, seed=1)
    assert inspect_model(0, blob, chunk_kb=1024).n_chunks == 1
    assert inspect_model(0, blob, chunk_kb=4).n_chunks > 40


@pytest.mark.parametrize("kind,blob,status", [
    (0, b"", 2), (0, b"tree\nversion=v4\n", 2), (1, b"binf....", 2), (1, b"bs64\tAAAA", 5), (1, b"x" * 400, 5), (1, b'{"learner": {}}', 2), (7, b"x", 5),
    (0, synth.lightgbm_model_text(2, 3).replace(b"num_class=1", b"num_class=3"), 5),
    (0, synth.lightgbm_model_text(2, 3).replace(b"is_linear=0", b"is_linear=1"), 5),
    (0, synth.lightgbm_model_text(2, 3).replace(b"max_feature_idx=2", b"max_feature_idx=0"), 2),
])
def test_bad_models_are_rejected_with_a_status(kind, blob, status):
    with pytest.raises(mb.MrError) as e:
        inspect_model(kind, blob)
    assert e.value.status == status, e.value


def test_truncated_lightgbm_arrays_are_parse_errors():
    blob = synth.lightgbm_model_text(1, 3, seed=5)
    lines = blob.decode().split("\n")
    i = next(k for k, l in enumerate(lines) if l.startswith("threshold="))
    lines[i] = " ".join(lines[i].split(" ")[:-1])
    with pytest.raises(mb.MrError) as e:
        inspect_model(0, "\n".join(lines).encode())
    assert e.value.status == 2 and "threshold" in e.value.message


def test_header_is_plain_c_and_usable_from_c():
    """include/mr_b200.h compiles as C99 and a C program can drive the host-only entry points."""
    import os
    import subprocess
    import tempfile

    import torch

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "abi_smoke")
        subprocess.run([cc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                        os.path.join(root, "tests", "c", "abi_smoke.c"), "-o", exe,
                        "-L", os.path.join(root, "metarank_b200"), "-lmrgpu",
                        "-Wl,-rpath," + os.path.join(root, "metarank_b200")], check=True)
        args = [exe] + ([] if torch.cuda.is_available() else ["--expect-no-gpu"])
        r = subprocess.run(args, capture_output=True, text=True)
        assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
        assert "sm_100a ok" in r.stdout


def test_compact_level_loop_sass_instruction_budget():
    """Performance regression guard (no GPU needed): the per-tree-level loop of the default scorer is
    hand-written PTX that ptxas schedules into 9 SASS instructions (power-of-two code tile: the code
    address is a single LOP3) or 10 (otherwise); a careless edit around it once silently grew the then
    12-instruction loop to 15 (+19 % kernel time, profiles/ncu_r1_summary.md)."""
    import os
    import re
    import shutil
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    obj = os.path.join(root, "metarank_b200", "csrc", "build", "gbdt_binned.o")
    if not os.path.exists(tool) or not os.path.exists(obj):
        pytest.skip("cuobjdump or the object file is not available")
    sass = subprocess.run([tool, "-sass", obj], capture_output=True, text=True).stdout
    for aligned, budget in ((1, 9), (0, 10)):
        m = re.search(r"Function : \S*compact_kernelIdLb0ELb%dEE\S*\n(.*?)(?:Function :|\Z)" % aligned, sass, re.S)
        assert m, "compact kernel <double, no categorical> not found"
        ins = [re.sub(r"/\*.*?\*/", "", ln).strip() for ln in m.group(1).split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", ln)]
        ins = [i for i in ins if i]
        start = next(k for k, i in enumerate(ins) if i.startswith("LDS.64") and "+UR" in i)
        end = next(k for k in range(start, len(ins)) if "BRA" in ins[k])
        body = ins[start:end + 1]
        assert len(body) <= budget, body
        assert sum("LDS" in i for i in body) == 2 and not any(i.startswith(("DSETP", "DADD", "LDG")) for i in body), body


def test_slim_level_loop_sass_instruction_budget():
    """The 4-byte-node scorer's level loop is 8 SASS instructions — LOP3 (code address), LDS.U16, HSETP2 (k compared as a
    binary16 pattern straight from the entry), LOP3 (child | block base), a predicated +4, LDS (next entry), ISETP (leaf =
    sign bit), BRA — with two shared-memory loads and no 64-bit load (an LDS.64 costs two wavefronts once lanes diverge).
    Models with small categorical bitsets pay two more (SHF, LOP3 with a predicate result) and never leave the loop."""
    import os
    import re
    import shutil
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    obj = os.path.join(root, "metarank_b200", "csrc", "build", "gbdt_binned.o")
    if not os.path.exists(tool) or not os.path.exists(obj):
        pytest.skip("cuobjdump or the object file is not available")
    sass = subprocess.run([tool, "-sass", obj], capture_output=True, text=True).stdout
    def level_loops(fn_body):
        """[(instructions of a backward-branch loop)] of one function's SASS, innermost loops of <= 12 instructions."""
        ins = []
        for ln in fn_body.split("\n"):
            a = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
            if a:
                ins.append((int(a.group(1), 16), a.group(2).strip()))
        at = {addr: k for k, (addr, _) in enumerate(ins)}
        loops = []
        for k, (addr, text) in enumerate(ins):
            b = re.search(r"BRA\s+(?:\S+,\s*)?0x([0-9a-f]+)", text)
            if b and int(b.group(1), 16) in at and 0 < k - at[int(b.group(1), 16)] < 12 and "LDS" in " ".join(t for _, t in ins[at[int(b.group(1), 16)]:k]):
                loops.append([t for _, t in ins[at[int(b.group(1), 16)]:k + 1]])
        return loops

    # <Real, T, CAT, NR>: smem root table / parameter root table (level 0 outside the loop), numeric and small-categorical
    for T in (512, 256, 128):
        for cat, nr, n_ins in ((0, 0, 8), (0, 512, 8), (0, 1920, 8), (2, 0, 10), (2, 512, 10)):
            m = re.search(r"Function : \S*slim_kernelIdLi%dELi%dELi%dEE\S*\n(.*?)(?:Function :|\Z)" % (T, cat, nr), sass, re.S)
            assert m, f"slim kernel <double, {T}, {cat}, {nr}> not found"
            loops = [b for b in level_loops(m.group(1)) if any(i.startswith("HSETP2") for i in b)]
            assert loops, (T, cat, nr)
            for body in loops:
                assert len(body) == n_ins, body
                assert body[0].startswith("LOP3") and body[1].startswith("LDS.U16"), body
                assert sum(i.startswith("LDS") for i in body) == 2 and not any("LDS.64" in i or i.startswith(("DADD", "LDG", "LDC")) for i in body), body
                assert any(".H0_H0" in i and ".H1_H1" in i for i in body if i.startswith("HSETP2")), body
                assert (sum(i.startswith("SHF") for i in body) == 1) == (cat == 2), body
            if nr:
                # level 0 comes from the constant bank: indexed LDC loads, no root-table LDS.128
                assert "LDC.64" in m.group(1) and "LDS.128" not in m.group(1), (T, cat, nr)


def test_ctypes_mirrors_match_the_header_layouts():
    """The Python ctypes structures (metarank_b200/features.py RankBatch, StateInfo) must have the header's
    layout: a C program prints sizeof / offsetof straight from include/mr_b200.h."""
    import ctypes
    import os
    import subprocess
    import tempfile

    from metarank_b200 import features as F

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    fields = [n for n, _ in F.RankBatch._fields_]
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "mr_b200.h"', 'int main(void) {',
           '  printf("%zu\\n", sizeof(mr_rank_batch));']
    src += [f'  printf("%zu\\n", offsetof(mr_rank_batch, {n}));' for n in fields]
    src += ['  printf("%zu\\n", sizeof(mr_state_info));', '  return 0; }']
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "layout.c")
        open(c, "w").write("\n".join(src))
        exe = os.path.join(d, "layout")
        subprocess.run([cc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), c, "-o", exe], check=True)
        out = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert out[0] == ctypes.sizeof(F.RankBatch)
    assert out[1:1 + len(fields)] == [getattr(F.RankBatch, n).offset for n in fields]
    assert out[-1] == ctypes.sizeof(F.StateInfo)


@pytest.mark.timeout(120)
def test_malformed_tree_structures_are_rejected_not_walked():
    """A model whose child pointers leave the tree, form a cycle or share a subtree would be walked forever (or
    out of bounds) by the host-side depth pass and by the kernels: the loader must refuse it.  Host only."""
    import json

    from metarank_b200 import _capi, synth
    from metarank_b200 import booster as B

    text = synth.lightgbm_model_text(3, 5, seed=1).decode()
    assert B.inspect_model(0, text.encode()).n_trees == 3

    def sub(prefix, fn):
        out, done = [], False
        for line in text.split("\n"):
            if line.startswith(prefix) and not done:
                k, v = line.split("=", 1)
                line, done = k + "=" + " ".join(fn(v.split(" "))), True
            out.append(line)
        return "\n".join(out).encode()

    for blob, what in [(sub("left_child=", lambda v: ["999"] + v[1:]), "out of range"),
                       (sub("right_child=", lambda v: ["-999"] + v[1:]), "out of range"),
                       (sub("left_child=", lambda v: ["0"] + v[1:]), "reachable twice"),          # self loop
                       (sub("left_child=", lambda v: v[:1] + [v[0]] + v[2:]), "reachable twice"),  # shared subtree
                       (sub("split_feature=", lambda v: ["77"] + v[1:]), "outside")]:
        with pytest.raises(_capi.MrError) as e:
            B.inspect_model(0, blob)
        assert what in str(e.value), (what, str(e.value))
    # XGBoost JSON: same checks after the node arrays are converted
    doc = json.loads(synth.xgboost_model_json(3, 5, depth=3, seed=2))
    tree = doc["learner"]["gradient_booster"]["model"]["trees"][0]
    assert B.inspect_model(1, json.dumps(doc).encode()).n_trees == 3
    for key, val, what in [("left_children", 0, "reachable twice"), ("right_children", 10_000, "out of range")]:
        bad = json.loads(json.dumps(doc))
        bad["learner"]["gradient_booster"]["model"]["trees"][0][key][0] = val
        with pytest.raises(_capi.MrError) as e:
            B.inspect_model(1, json.dumps(bad).encode())
        assert what in str(e.value) or "range" in str(e.value), str(e.value)
    assert len(tree["left_children"]) > 1


@pytest.mark.timeout(300)
def test_parsers_survive_corrupted_models_and_schemas():
    """Host-only fuzz: byte flips / truncations of model blobs (LightGBM text, XGBoost JSON, UBJSON and binary) and of
    schema documents end in an MrError or a clean parse — never a crash, a hang or an out-of-bounds read."""
    import ctypes as C
    import json

    from metarank_b200 import _capi, synth
    from metarank_b200 import booster as B

    rng = np.random.Generator(np.random.PCG64(3))
    blobs = [(0, synth.lightgbm_model_text(6, 5, seed=1, cat_features={2: 8})), (1, synth.xgboost_model_json(5, 5, depth=3, seed=2)),
             (1, synth.xgboost_model_ubj(5, 5, depth=3, seed=2)),
             (1, synth.xgboost_model_binary(5, 5, depth=3, seed=2, deleted=1)), (1, synth.xgboost_model_binary(4, 6, depth=4, seed=3, magic=True))]
    ok = err = 0
    for kind, base in blobs:
        for _ in range(400):
            b = bytearray(base)
            for _ in range(int(rng.integers(1, 6))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
            if rng.random() < 0.3:
                b = b[:int(rng.integers(0, len(b)))]
            try:
                B.inspect_model(kind, bytes(b))
                # what parses must also pack, and its packing must walk like its trees (host-side layout check)
                _, bad = B.selfcheck_model(kind, bytes(b), 4)
                assert bad == 0
                ok += 1
            except _capi.MrError:
                err += 1
    assert ok > 0 and err > 0
    feats, model = synth.ranklens_config()
    base = json.dumps({"features": feats, "model_features": model}).encode()
    lib = _capi.lib()
    ok = err = 0
    for _ in range(600):
        b = bytearray(base)
        for _ in range(int(rng.integers(1, 5))):
            b[int(rng.integers(0, len(b)))] = int(rng.choice([rng.integers(32, 127), ord('"'), ord("0"), ord("-"), ord("["), ord(":")]))
        if rng.random() < 0.2:
            b = b[:int(rng.integers(0, len(b)))]
        h = C.c_void_p()
        if lib.mr_schema_create(None, bytes(b), C.c_size_t(len(b)), C.byref(h)) == 0:
            lib.mr_schema_free(h)
            ok += 1
        else:
            err += 1
    assert ok > 0 and err > 0


def test_real_file_quirks_of_the_booster_formats():
    """No LightGBM / XGBoost build can be had in this image or on the GPU box (profiles/probe_r2_gbdt_libs.txt), so the
    parsers are held to the quirks of the files those libraries really write, from their public formats:
    LightGBM  the bare `average_output` word (rf boosting), objectives that transform the raw score, `tree_sizes` vs the
              blocks present, a lost `end of trees`, out-of-order blocks, feature_names vs max_feature_idx;
    XGBoost   base_score as a string ("5E-1", and "[5E-1]" from 3.1 on), booleans in default_left, tree_param.num_nodes,
              transforming objectives, multi-target, the deprecated binary header."""
    import json

    lgb = synth.lightgbm_model_text(3, 4, seed=2)
    assert inspect_model(0, lgb).n_trees == 3
    assert inspect_model(0, lgb.replace(b"\n", b"\r\n")).n_trees == 3  # CRLF files
    for text, status, word in [
        (lgb.replace(b"objective=lambdarank", b"objective=lambdarank\naverage_output"), 5, "average_output"),
        (lgb.replace(b"objective=lambdarank", b"objective=binary sigmoid:1"), 5, "binary"),
        (lgb.replace(b"objective=lambdarank", b"objective=regression sqrt"), 5, "regression sqrt"),
        (lgb.replace(b"objective=lambdarank", b"objective=poisson"), 5, "poisson"),
        (lgb.replace(b"end of trees", b""), 2, "truncated"),
        (lgb[: lgb.index(b"Tree=2")], 2, "truncated"),
        (lgb.replace(b"Tree=1", b"Tree=7"), 2, "out of order"),
        (lgb.replace(b"feature_names=", b"feature_names=extra "), 2, "feature_names"),
    ]:
        with pytest.raises(mb.MrError) as e:
            inspect_model(0, text)
        assert e.value.status == status and word in e.value.message, (word, e.value)
    # tree_sizes disagreeing with the blocks present (a file cut between two trees and re-terminated)
    cut = lgb[: lgb.index(b"Tree=2")] + b"end of trees\n"
    with pytest.raises(mb.MrError) as e:
        inspect_model(0, cut)
    assert e.value.status == 2 and "tree_sizes" in e.value.message
    for obj in (b"objective=rank_xendcg", b"objective=regression", b"objective=huber"):
        assert inspect_model(0, lgb.replace(b"objective=lambdarank", obj)).n_trees == 3

    doc = json.loads(synth.xgboost_model_json(3, 5, depth=3, seed=4))
    lmp = doc["learner"]["learner_model_param"]

    def with_(mut):
        d = json.loads(json.dumps(doc))
        mut(d)
        return json.dumps(d).encode()

    for bs in ("5E-1", "[5E-1]", " [ 2.5E-1 ] ", "0.5"):
        assert inspect_model(1, with_(lambda d: d["learner"]["learner_model_param"].update(base_score=bs))).n_trees == 3
    def bools(d):
        for t in d["learner"]["gradient_booster"]["model"]["trees"]:
            t["default_left"] = [bool(x) for x in t["default_left"]]
    assert inspect_model(1, with_(bools)).n_trees == 3  # XGBoost 1.0-1.2 JSON wrote booleans
    assert isinstance(lmp["base_score"], str)  # the synthetic writer uses the library's string form too
    for mut, status, word in [
        (lambda d: d["learner"]["learner_model_param"].update(base_score="[5E-1,2E-1]"), 5, "multi-target"),
        (lambda d: d["learner"]["learner_model_param"].update(base_score="abc"), 2, "base_score"),
        (lambda d: d["learner"]["learner_model_param"].update(num_target="2"), 5, "multi-target"),
        (lambda d: d["learner"]["objective"].update(name="binary:logistic"), 5, "binary:logistic"),
        (lambda d: d["learner"]["gradient_booster"].update(name="dart"), 5, "dart"),
        (lambda d: d["learner"]["gradient_booster"]["model"]["trees"][1].setdefault("tree_param", {}).update(num_nodes="3"), 2, "num_nodes"),
        (lambda d: d["learner"]["gradient_booster"]["model"]["trees"][0].update(
            split_type=[1] * len(d["learner"]["gradient_booster"]["model"]["trees"][0]["left_children"])), 5, "categorical"),
    ]:
        with pytest.raises(mb.MrError) as e:
            inspect_model(1, with_(mut))
        assert e.value.status == status and word in e.value.message, (word, e.value)
    # the deprecated binary encoding is read (test_xgboost_deprecated_binary_encoding_is_read); its base64 wrapping is not
    with pytest.raises(mb.MrError) as e:
        inspect_model(1, b"bs64AAAA")
    assert e.value.status == 5 and "base64" in e.value.message
    with pytest.raises(mb.MrError) as e:
        inspect_model(1, b"binf\x00\x00\x00\x00")
    assert e.value.status == 2 and "truncated" in e.value.message


@pytest.mark.parametrize("kind,blob,form", [
    (0, synth.lightgbm_model_text(60, 30, seed=1, stump_every=7), 3),                       # numeric, root table
    (0, synth.lightgbm_model_text(40, 12, num_leaves=2, seed=2), 3),                        # every leaf hangs off the root: dummy splits
    (0, synth.lightgbm_model_text(30, 10, seed=3, cat_features={3: 16, 7: 5}, stump_every=5), 7),   # small categorical: in-loop form
    (0, synth.lightgbm_model_text(30, 10, seed=4, cat_features={3: 40}), 1),                # categories beyond 15: wide form, no root table
    (0, synth.lightgbm_model_text(2000, 8, num_leaves=4, seed=5), 1),                       # > 1920 trees: chunk-resident root tables
    (0, synth.lightgbm_model_text(2000, 8, num_leaves=4, seed=6, cat_features={1: 9}), 5),
    (0, synth.lightgbm_model_text(20, 200, seed=7), 3),                                     # 256-item tiles
    (0, synth.lightgbm_model_text(12, 240, seed=8), 3),                                     # 128-item tiles
    (1, synth.xgboost_model_json(50, 16, depth=6, seed=9), 3),
])
def test_slim_packing_walks_like_the_trees(kind, blob, form, monkeypatch):
    """pack_slim's bytes, read the way gbdt_score_slim_kernel reads them (root table in the parameter space or in the
    chunk, dummy splits under the root, in-entry bitsets of the small-categorical form), reach the leaf the parsed tree
    reaches, on random code vectors incl. NaN / out-of-range categories — checked on the host, so that a packing bug shows
    up without a GPU."""
    from metarank_b200.booster import selfcheck_model

    got_form, bad = selfcheck_model(kind, blob, 48)
    tile = got_form >> 8
    assert got_form & 255 == form and bad == 0 and tile in (512, 256, 128)
    for smaller in (256, 128):   # the forms one-wave batches are scored from (mr_model::pick_slim)
        if smaller < tile:
            f2, bad = selfcheck_model(kind, blob, 16, max_tile=smaller)
            assert bad == 0 and (f2 == 0 or (f2 & 255 == form and f2 >> 8 == smaller))
    monkeypatch.setenv("MR_NO_ROOT_TAB", "1")
    got_form, bad = selfcheck_model(kind, blob, 16)
    assert got_form & 255 == (form & ~2) and bad == 0
    monkeypatch.setenv("MR_NO_CAT16", "1")
    got_form, bad = selfcheck_model(kind, blob, 16)
    assert got_form & 255 == (form & ~6) and bad == 0


def test_xgboost_deprecated_binary_encoding_is_read():
    """Booster.toByteArray() of xgboost4j up to 2.0 — what a Metarank-trained XGBoost model holds — is the deprecated binary
    encoding.  The same synthetic model written as JSON and as binary parses to the same shape on the host (and, in
    tests/test_oracle_gbdt.py / test_gbdt_gpu.py, to the same scores); the `binf` prefix, pruned nodes left in the arrays
    and a missing attribute block change nothing; every truncation point is a parse error, never a shorter model."""
    j = synth.xgboost_model_json(12, 9, depth=5, seed=21, full=False)
    ij = inspect_model(1, j)
    for kw in ({}, {"magic": True}, {"deleted": 2}, {"with_attributes": False}, {"objective": "rank:pairwise"}):
        b = synth.xgboost_model_binary(12, 9, depth=5, seed=21, full=False, **kw)
        ib = inspect_model(1, b)
        assert (ib.n_trees, ib.n_features, ib.n_internal_nodes, ib.max_leaves) == (ij.n_trees, ij.n_features, ij.n_internal_nodes, ij.max_leaves)
    b = synth.xgboost_model_binary(12, 9, depth=5, seed=21, full=False, with_attributes=False)
    for cut in list(range(0, 400, 7)) + [len(b) // 2, len(b) - 20, len(b) - 1]:
        with pytest.raises(mb.MrError) as e:
            inspect_model(1, b[:cut])
        assert e.value.status in (2, 5), cut
    for bad, status in ((synth.xgboost_model_binary(3, 4, objective="binary:logistic"), 5),
                        (synth.xgboost_model_binary(3, 4).replace(b"gbtree", b"dart\0\0"), 5)):
        with pytest.raises(mb.MrError) as e:
            inspect_model(1, bad)
        assert e.value.status == status
