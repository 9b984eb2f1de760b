(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3) 2>&1
python tools/bench_headconv.py
python tools/bench_tc.py 2>/dev/null | cut -c1-330
python tools/bench_tc.py --teacher 2>/dev/null | tail -5 | cut -c1-250
python tools/bench_train_step.py 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d.get('ms_per_step_tf32_operands'), d['loss'])"
