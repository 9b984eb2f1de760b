timeout 900 python -m pytest tests/test_semseg_head.py tests/test_pretrain_step.py tests/test_openess_step.py -m gpu -q -x 2>&1 | tail -3
python tools/bench_tc.py --teacher 2>/dev/null | tail -2 | cut -c1-250
python tools/bench_train_step.py 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d.get('ms_per_step_tf32_operands'), d['loss'])"
python tools/profile_train_step.py --shapes 2>/dev/null | tail -42 | cut -c1-200
