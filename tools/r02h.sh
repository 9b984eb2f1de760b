set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; tail -3 gpurun_out/r02h_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02h_bench_reference.json 2>/dev/null
