"""Runs only bench.py's region_cells section on cuda:0 and prints its JSON (development aid)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=20_000_000)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--no-cpu-baseline", action="store_true")
args = ap.parse_args()
torch.cuda.set_device(0)
print(json.dumps(bench.region_cells_section(args, torch, torch.device("cuda:0"), 0, 1)))
