"""Run every case of tests/gpu_cases.py in its own process (a trapped kernel poisons its CUDA context; isolating
cases keeps the rest of the run informative).  Writes gpurun_out/probe.json and prints one line per case.

  python tests/gpu_probe.py [substring ...]
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.gpu_cases import CASES  # noqa: E402


def main():
    pats = sys.argv[1:]
    names = [n for n in CASES if not pats or any(p in n for p in pats)]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    results = []
    for n in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, "-m", "tests.gpu_cases", n], cwd=ROOT, capture_output=True, text=True,
                               timeout=240)
            out = r.stdout + r.stderr
            line = [l for l in out.splitlines() if l.startswith("CASE_RESULT ")]
            if line:
                res = json.loads(line[-1][len("CASE_RESULT "):])
            else:
                res = {"case": n, "ok": False, "errs": {}, "tail": out[-1500:]}
        except subprocess.TimeoutExpired as e:
            res = {"case": n, "ok": False, "errs": {}, "tail": "TIMEOUT " + str(e)[-300:]}
        res["sec"] = round(time.time() - t0, 1)
        results.append(res)
        print(("PASS " if res["ok"] else "FAIL ") + n + " " + json.dumps(res.get("errs")) +
              (" :: " + res.get("tail", "")[-600:].replace("\n", " | ") if not res["ok"] else ""), flush=True)
        with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as f:
            json.dump(results, f, indent=1)
    bad = [r["case"] for r in results if not r["ok"]]
    print("probe: %d/%d passed; failed: %s" % (len(results) - len(bad), len(results), bad))


if __name__ == "__main__":
    main()
