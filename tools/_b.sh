python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r01c_bench.json 2>/dev/null; cat gpurun_out/r01c_bench.json | cut -c1-1500
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01c_bench_reference.json 2>/dev/null
python tools/bench_tc.py > gpurun_out/r01c_tc_bench.jsonl 2>&1; python tools/bench_tc.py --teacher >> gpurun_out/r01c_tc_bench.jsonl 2>&1
python tools/bench_train_step.py --batch 4 --steps 5 --warmup 2 > gpurun_out/r01c_train_step.json 2>&1; tail -1 gpurun_out/r01c_train_step.json | cut -c1-400
