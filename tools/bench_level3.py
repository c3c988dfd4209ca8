"""Runs bench.py on the widened rows and collects the JSON lines.
usage: python tools/bench_level3.py [--n 8192] [--routines dsyrk,dtrsm,...] > profiles/bench_r01/level3_n8192.jsonl
(The measurement itself -- value / e2e / roofline / cpu_baseline -- is bench.py's `run_level3`.)"""
import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--routines", default="dsyrk,dtrsm,dtrmm,dsymm,dsyr2k,dpotrf,dgetrf,ssyrk,strsm,ssymm,spotrf,sgetrf")
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    for name in args.routines.split(","):
        p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "%s%d" % (name, args.n), "--steps", str(args.steps),
                            "--warmup", "3"], stdout=subprocess.PIPE, text=True)
        for line in p.stdout.splitlines():
            if line.startswith("{"):
                print(line, flush=True)


if __name__ == "__main__":
    main()
