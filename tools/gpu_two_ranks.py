"""Debug helper: the drop-in on 2 ranks / 2 GPUs, full output of both ranks."""
import os, sys, subprocess, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from dropin_util import DROPIN, write_inputs
from golden_util import load_json
first = load_json("dna49p.first.json")
lo, lines = 1, []
for i, part in enumerate(first["partitions"]):
    n = sum(part["wgt"]); lines.append(f"DNA, p{i} = {lo}-{lo + n - 1}"); lo += n
with tempfile.TemporaryDirectory() as tmp:
    binary, tree, _ = write_inputs("dna49p", tmp, lines)
    env = dict(os.environ, EXAML_PROBE_DUMP=os.path.join(tmp, "dump"), EXAML_PROBE_NOTIPS="1", PYTHONPATH=ROOT)
    env.update({k: v for k, v in [a.split("=", 1) for a in sys.argv[1:]]})
    r = subprocess.run([sys.executable, "-m", "examl_b200.launch", "-np", "2", "--", DROPIN, "-s", binary, "-t", tree,
                        "-m", "GAMMA", "-n", "R", "-f", "e", "-w", tmp + "/"], cwd=tmp, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    print("rc", r.returncode)
    for line in r.stdout.splitlines():
        if "back-end" in line or "failed" in line or "rror" in line or "Likelihood" in line or "NCCL" in line:
            print(line)
