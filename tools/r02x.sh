(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3) 2>&1
python tools/bench_tc.py --config2 2>/dev/null | tail -2 | cut -c1-400
python tools/bench_train_step.py 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d.get('ms_per_step_tf32_operands'), d['loss'])"
python tools/profile_train_step.py 2>/dev/null | tail -43 | head -34 | cut -c1-160
