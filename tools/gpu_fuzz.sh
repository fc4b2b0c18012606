#!/bin/bash
python -m pytest tests/test_fuzz_gpu.py -x -q -m gpu 2>&1 | tail -15
