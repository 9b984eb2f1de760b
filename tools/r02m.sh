timeout 600 python -m pytest tests/test_tc_convlstm.py tests/test_models.py tests/test_pretrain_step.py tests/test_openess_step.py tests/test_drop_in.py -m gpu -q -x -s 2>&1 | grep -i "passed\|failed\|latent\|error" | cut -c1-200
timeout 300 python tools/bench_tc.py 2>/dev/null | sed -n 5,8p | cut -c1-260
python tools/bench_train_step.py 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d.get('ms_per_step_tf32_operands'), d['loss'])"
