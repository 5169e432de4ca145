#!/bin/bash
# compute-sanitizer memcheck + racecheck over the small golden-vector GPU tests (slow: run on demand)
set -e
cd "$(dirname "$0")/.."
for tool in memcheck racecheck; do
  echo "== $tool"
  timeout 280 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_golden.py -m gpu -q -x 2>&1 | tail -4
done
