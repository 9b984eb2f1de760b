(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3) 2>&1
python tools/bench_tc.py --teacher 2>/dev/null | tail -1 | cut -c1-250
python tools/bench_configs.py 2>/dev/null | cut -c1-600 | head -8
