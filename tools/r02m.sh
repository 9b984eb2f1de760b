timeout 900 python -m pytest tests/test_teacher.py tests/test_pretrain_step.py tests/test_drop_in.py tests/test_deeplab.py -m gpu -q -x 2>&1 | tail -8 | cut -c1-200
for v in 1 0 1 0; do OESS_TEACHER_GRAPH=$v python tools/bench_train_step.py 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('teacher graph=$v', d['ms_per_step'], d.get('ms_per_step_tf32_operands'), d['loss'])"; done
python tools/bench_tc.py --teacher 2>/dev/null | sed -n 1,4p | cut -c1-200
