#!/bin/bash
python tools/full_parity_1b.py 2>&1 | tail -3
