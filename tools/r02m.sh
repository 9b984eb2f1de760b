timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
python tools/bench_headconv.py
python tools/bench_tc.py 2>/dev/null | sed -n 8,8p | cut -c1-220
python tools/bench_tc.py --teacher 2>/dev/null | sed -n 3,6p | cut -c1-200
python tools/bench_train_step.py 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d.get('ms_per_step_tf32_operands'), d['loss'])"
