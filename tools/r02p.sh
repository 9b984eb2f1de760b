set -x
python -m pytest tests/test_models.py tests/test_tc_conv.py tests/test_tc_convlstm.py tests/test_pretrain_step.py -m gpu -x -q 2>&1 | grep -v Warning | tail -25
python tools/bench_train_step.py --batch 4 --steps 5 2>/dev/null | tail -1 | cut -c1-400
python tools/bench_tc.py 2>/dev/null | grep -i "e2vid" | cut -c1-300
