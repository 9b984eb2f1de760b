set -x
python tools/bench_configs.py --config1 --gpu 2>/dev/null | tee gpurun_out/r02_configs.jsonl
python tools/bench_configs.py --config4 2>/dev/null | tee -a gpurun_out/r02_configs.jsonl
python tools/bench_configs.py --config5 2>/dev/null | tee -a gpurun_out/r02_configs.jsonl
python bench.py --steps 6 --train-steps 0 --host-output 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps(d['variants'])); print(d['roofline']['path_frac'], d['roofline']['path_frac_convert_only'], d['value'])"
