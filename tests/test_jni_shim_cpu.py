"""The JNI boundary compiles and behaves: jni_shim.cpp is built with -DWITH_JNI
against the jni.h test double in tests/c/jni_stub/, its exported Java_* symbols must be exactly the @native
list of INTEGRATION.md, and a C++ driver calls it through a fake JNIEnv."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _build(tmp):
    from metarank_b200 import _capi
    _capi.build()
    so = os.path.join(tmp, "libmrgpu_jni.so")
    inc = ["-I", os.path.join(ROOT, "tests", "c", "jni_stub"), "-I", os.path.join(ROOT, "include")]
    lib_dir = os.path.join(ROOT, "metarank_b200")
    subprocess.run([CXX, "-std=c++17", "-O1", "-Wall", "-Werror", "-fPIC", "-shared", "-DWITH_JNI", *inc,
                    os.path.join(ROOT, "metarank_b200", "csrc", "jni_shim.cpp"), "-o", so,
                    "-L", lib_dir, "-lmrgpu", f"-Wl,-rpath,{lib_dir}"], check=True)
    exe = os.path.join(tmp, "jni_driver")
    subprocess.run([CXX, "-std=c++17", "-O1", "-Wall", *inc, os.path.join(ROOT, "tests", "c", "jni_driver.cpp"), "-o", exe,
                    "-L", tmp, "-lmrgpu_jni", "-L", lib_dir, "-lmrgpu", f"-Wl,-rpath,{tmp}", f"-Wl,-rpath,{lib_dir}"], check=True)
    return so, exe


def test_jni_symbols_match_integration_md(tmp_path):
    so, _ = _build(str(tmp_path))
    out = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r"Java_ai_metarank_b200_Native_(\w+)", out)))
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    natives = sorted(set(re.findall(r"@native def (\w+)\(", doc)))
    assert natives, "INTEGRATION.md lists no @native methods"
    assert exported == natives


def test_jni_shim_host_only_behaviour(tmp_path):
    _, exe = _build(str(tmp_path))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")  # the CPU suite must behave the same on a GPU box
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout + r.stderr
    assert "FAIL" not in r.stdout


@pytest.mark.gpu
def test_jni_shim_scores_on_the_gpu(tmp_path):
    _, exe = _build(str(tmp_path))
    r = subprocess.run([exe, "gpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout + r.stderr
