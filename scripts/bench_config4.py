#!/usr/bin/env python
"""Config 4 of BASELINE.json (memory-bound path) alone: full-volume patch gather + pixel-wise entropy map.
The measurement itself lives in bench.py (`run_config4`, also part of the default bench line as key `config4`).

  python scripts/bench_config4.py [--slices 180] [--reps 3]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as Bn          # noqa: E402
import nnal_b200            # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--slices', type=int, default=180)
ap.add_argument('--reps', type=int, default=3)
args = ap.parse_args()
padded, stats, _ = Bn.make_workload(10)
eng = nnal_b200.get_engine()
eng.upload(0, padded)
res = Bn.run_config4(eng, padded, stats, Bn.load_peaks(), slices=args.slices, reps=args.reps)
print(json.dumps({'config4': res}))
