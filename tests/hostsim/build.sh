#!/bin/bash
# Builds tests/hostsim/libhostsim.so (test-only host compilation of the device headers).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
g++ -std=c++17 -O2 -ffp-contract=off -fPIC -shared -I"$HERE/../../s2geometry_b200/csrc" "$HERE"/hostsim*.cc -o "$HERE/libhostsim.so"
