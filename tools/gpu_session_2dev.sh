#!/bin/bash
# 2-GPU box: the GPU tests that need two devices, and the native CLI on two devices against one
tag=${1:-d2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "several_devices or release_after_free or golden_vectors" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
python - > gpurun_out/${tag}_cli.log 2>&1 <<'PY'
import hashlib, json, os, subprocess, sys, time
sys.path.insert(0, os.getcwd())
from pandelos_b200 import build, synth
w = synth.shape("mycoplasma64")
k = synth.calculate_k(w)
w.write_faa("/tmp/in.faa")
gold = json.load(open("tests/golden/digests/mycoplasma64.json"))
for dev in ("1", "2"):
    t = time.time()
    r = subprocess.run([build.CLI_BIN, "-i", "/tmp/in.faa", "-k", str(k), "-o", "/tmp/out%s.net" % dev], capture_output=True, text=True, env=dict(os.environ, PD_DEVICES=dev))
    sha = hashlib.sha256(open("/tmp/out%s.net" % dev, "rb").read()).hexdigest()
    print("devices", dev, "rc", r.returncode, "%.2fs" % (time.time() - t), "net sha ok:", sha == gold["net"]["sha256"], r.stderr[-300:])
PY
tail -3 gpurun_out/${tag}_pytest.log; cat gpurun_out/${tag}_cli.log
