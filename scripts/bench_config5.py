#!/usr/bin/env python
"""BASELINE.json config 5 (scoring + greedy assembly wall-clock): thin wrapper around ``python bench.py --config5``.
usage: scripts/bench_config5.py [repeats]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
repeats = int(sys.argv[1]) if len(sys.argv) > 1 else 5
sys.exit(subprocess.call([sys.executable, os.path.join(ROOT, "bench.py"), "--config5", "--steps", str(4 * repeats), "--warmup", "3"]))
