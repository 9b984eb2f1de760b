# end-of-session measurement pass (one gpurun call): tests, smoke, bench lines, launch list, one full capture, tool benches
set -x
(time python -m pytest tests -m gpu -q 2>&1 | tail -3) 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r01e_bench.json 2> gpurun_out/r01e_bench.err; tail -c 300 gpurun_out/r01e_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01e_bench_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 80 --csv --log-file gpurun_out/r01e_launches.csv python bench.py --steps 3 --warmup 3 --host-output 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_strip_splat -s 3 -c 1 -o gpurun_out/r01e_strip -f python bench.py --steps 2 --warmup 3 --host-output 0 > /dev/null 2>&1
python tools/bench_tc.py > gpurun_out/r01e_tc_bench.jsonl 2>/dev/null
python tools/bench_tc.py --teacher >> gpurun_out/r01e_tc_bench.jsonl 2>/dev/null
python tools/bench_tc.py --config2 >> gpurun_out/r01e_tc_bench.jsonl 2>/dev/null
python tools/bench_train_step.py --batch 4 > gpurun_out/r01e_train_step.json 2>/dev/null
ls -la gpurun_out | grep r01e
