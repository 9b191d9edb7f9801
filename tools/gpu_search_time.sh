#!/bin/bash
python tools/gpu_latency.py
python - <<'PY'
import sys, os, time, tempfile
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from dropin_util import write_inputs, run_dropin
for mode in (["-f","e"], ["-f","d"]):
    with tempfile.TemporaryDirectory() as tmp:
        b, t, _ = write_inputs("dna49", tmp)
        t0 = time.time(); r, d = run_dropin(b, t, tmp, mode); dt = time.time() - t0
        print("drop-in", mode, "wall %.1f s" % dt, "final", d["final"]["likelihood"])
        import subprocess
        ref = os.path.join("oracle", "_ref", "examl-AVX")
        t0 = time.time()
        subprocess.run([os.path.abspath(ref), "-s", b, "-t", t, "-m", "GAMMA", "-n", "REF", "-w", tmp + "/"] + mode, cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT)
        print("reference AVX 1 core", mode, "wall %.1f s" % (time.time() - t0))
PY
