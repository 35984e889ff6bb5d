#!/bin/bash
# round 2, GPU call 26 (1 GPU): the driver's round-end sequence on the final tree -- GPU suite, smoke(), bench.py with its
# defaults, the reference arm with a short run
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -rs > gpurun_out/r2_pytest_gpu_final.log 2>&1
tail -4 gpurun_out/r2_pytest_gpu_final.log | cut -c1-200
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke_final.log 2>&1; tail -3 gpurun_out/r2_smoke_final.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r2_bench_default_final.json 2> gpurun_out/r2_bench_default_final.err; cut -c1-400 gpurun_out/r2_bench_default_final.json; tail -2 gpurun_out/r2_bench_default_final.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_final.json 2> gpurun_out/r2_bench_reference_final.err; cut -c1-300 gpurun_out/r2_bench_reference_final.json
