set -x
python -m pytest tests/test_openess_step.py -m gpu -x -q 2>&1 | grep -v Warning | grep -B30 "^E  " | tail -60
python tools/profile_train_step.py 2>/dev/null | tail -36
