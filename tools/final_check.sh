#!/bin/bash
# Final verification of the committed tree: GPU tests, smoke, both bench arms with their DEFAULT arguments.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > gpurun_out/fc_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -4 > gpurun_out/fc_smoke.txt
( time python bench.py ) 2>gpurun_out/fc_bench.err | tail -1 > gpurun_out/fc_bench.json
( time python bench.py --impl reference ) 2>gpurun_out/fc_ref.err | tail -1 > gpurun_out/fc_ref.json
cat gpurun_out/fc_pytest.txt gpurun_out/fc_smoke.txt; cut -c1-250 gpurun_out/fc_bench.json; echo; grep real gpurun_out/fc_bench.err; cut -c1-250 gpurun_out/fc_ref.json; echo; grep real gpurun_out/fc_ref.err
