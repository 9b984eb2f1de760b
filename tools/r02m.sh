timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2 | cut -c1-200
python tools/bench_train_step.py 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('graph', d['ms_per_step'], d.get('ms_per_step_tf32_operands'), d['loss'], d.get('encoder_loop_cuda_graph'))"
