set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -40
python bench.py --steps 6 --warmup 3 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; tail -5 gpurun_out/r02g_bench.err; cut -c1-3000 gpurun_out/r02g_bench.json
