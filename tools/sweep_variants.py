#!/usr/bin/env python3
"""Runs bench.py once per tuning build (micro-aes_b200/lib_<name>/, see tools/build_variants.sh) and per
environment setting, and prints one line each: variant, GiB/s, kernel ms, parity flag.

    python tools/sweep_variants.py [--workload ctr128] [--steps 20] name[:ENV=VAL,...] ...
`lib` names the default build."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:]
workload, steps, gib = "ctr128", "20", None
while args and args[0].startswith("--"):
    k = args.pop(0)
    v = args.pop(0)
    if k == "--workload": workload = v
    if k == "--steps": steps = v
    if k == "--gib": gib = v
for spec in args:
    name, _, envs = spec.partition(":")
    env = dict(os.environ)
    if name != "lib":
        env["UAES_LIBDIR"] = os.path.join(ROOT, "micro-aes_b200", f"lib_{name}")
    for kv in filter(None, envs.split(",")):
        k, _, v = kv.partition("=")
        env[k] = v
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", steps, "--warmup", "3", "--no-cpu", "--no-e2e",
           "--no-secondary", "--workload", workload] + (["--gib-per-gpu", gib] if gib else [])
    try:
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    except subprocess.TimeoutExpired:
        print(f"{spec:40s} TIMEOUT", flush=True)
        continue
    try:
        j = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"{spec:40s} {j['value']:9.2f} GiB/s  kernel {j['roofline']['kernel_ms']:8.4f} ms  parity {j.get('parity_spot_check')}"
              f"  split {j.get('queue_split')}", flush=True)
    except Exception as e:
        print(f"{spec:40s} FAILED: {e}; {r.stderr[-300:]}", flush=True)
