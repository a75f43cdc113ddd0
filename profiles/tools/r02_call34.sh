#!/bin/bash
# Last call of round 2: the full GPU suite (90 tests with the re-trace test of the drop-in) on the final build.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -12) > gpurun_out/r02_c34_tests.log 2>&1
cat gpurun_out/r02_c34_tests.log
