(time timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4) 2>&1
for cfg in "X=1" "OESS_PREFETCH=0"; do echo $cfg; env $cfg python tools/bench_train_step.py 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d.get('ms_per_step_tf32_operands'), d['loss'])"; done
python tools/profile_train_step.py 2>/dev/null | tail -43 | head -34 | cut -c1-160
