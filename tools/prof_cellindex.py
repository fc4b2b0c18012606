#!/usr/bin/env python3
"""Small launch sequence for ncu: the S2CellIndex point join of bench.py's cell_index_join section at 20M points
(3 warm-up + 2 timed launches of ci_point_histogram_kernel). Usage: python tools/prof_cellindex.py [n_points]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
args = argparse.Namespace(points=n, steps=2, no_cpu_baseline=True)
print(json.dumps(bench.cell_index_join_section(args, torch, torch.device("cuda:0"), 0, 1)))
