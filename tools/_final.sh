# end-of-session measurement pass (one gpurun call): tests, smoke, bench lines, launch list, full captures, tool benches
set -x
R=${R:-r02}
(time python -m pytest tests -m gpu -q 2>&1 | tail -3) 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; tail -c 300 gpurun_out/${R}_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 3 --warmup 3 --inner 4 --host-output 0 --train-steps 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_strip_splat -s 3 -c 1 -o gpurun_out/${R}_strip -f python bench.py --steps 2 --warmup 3 --inner 2 --host-output 0 --train-steps 0 > /dev/null 2>&1
python tools/bench_tc.py > gpurun_out/${R}_tc_bench.jsonl 2>/dev/null
python tools/bench_tc.py --teacher >> gpurun_out/${R}_tc_bench.jsonl 2>/dev/null
python tools/bench_tc.py --config2 >> gpurun_out/${R}_tc_bench.jsonl 2>/dev/null
python tools/bench_infonce.py > gpurun_out/${R}_infonce.jsonl 2>/dev/null
python tools/bench_headconv.py >> gpurun_out/${R}_tc_bench.jsonl 2>/dev/null
python tools/bench_conv_shapes.py > gpurun_out/${R}_conv_shapes.jsonl 2>/dev/null
python tools/bench_wgrad.py > gpurun_out/${R}_wgrad.jsonl 2>/dev/null
python tools/bench_configs.py --config1 --gpu > gpurun_out/${R}_configs.jsonl 2>/dev/null
python tools/bench_configs.py --config4 >> gpurun_out/${R}_configs.jsonl 2>/dev/null
python tools/bench_configs.py --config5 >> gpurun_out/${R}_configs.jsonl 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_gemm_tf32_p -s 6 -c 1 -o gpurun_out/${R}_gemm -f python tools/bench_tc.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_convlstm_tc_p -s 30 -c 1 -o gpurun_out/${R}_convlstm -f python tools/bench_tc.py > /dev/null 2>&1
OESS_INFONCE=tc compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > gpurun_out/${R}_sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/${R}_sanitizer_memcheck.log
OESS_INFONCE=tc compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > gpurun_out/${R}_sanitizer_racecheck.log 2>&1; tail -3 gpurun_out/${R}_sanitizer_racecheck.log
ls -la gpurun_out | grep ${R}_
