"""Runs bench.py on the parity-test configurations (BASELINE.json configs 2 and 5 and a heavy-hex chi=32 sweep is covered by
tools/profile_hh.py) with each kernel family, and prints one summary line per run.   python tools/other_configs.py"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for w in sys.argv[1:] or ["grid32x32_chi8_c128", "cubic16_chi6_c128"]:
    for pm in (0, 2):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", w, "--path", str(pm), "--no-gates",
                              "--no-cpu-baseline", "--steps", "5"], capture_output=True, text=True)
        for l in out.stdout.splitlines():
            if l.startswith("{"):
                j = json.loads(l)
                print(f"{w} path={pm}: {j['ms_per_step']:.3f} ms/sweep, {j['value']:.3e} updates/s, "
                      f"{j['roofline']['achieved']:.2f} TFLOP/s algorithmic, launches {j['gpu_launches']}, e2e {j['e2e']['value']:.3e}", flush=True)
        if out.returncode != 0:
            print(w, pm, "FAILED", out.stderr[-500:])
