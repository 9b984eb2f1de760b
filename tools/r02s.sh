set -x
python -m pytest tests/test_tc_convlstm.py tests/test_tc_conv.py tests/test_models.py -m gpu -x -q -s 2>&1 | grep -v Warning | grep -E "passed|failed|error|Error|assert|max \|latent|E  " | head -30
OESS_E2VID_DTYPE=tf32 python tools/bench_tc.py 2>/dev/null | grep -E "e2vid|convlstm" | cut -c1-220
OESS_E2VID_DTYPE=bf16 python tools/bench_tc.py 2>/dev/null | grep -E "e2vid" | cut -c1-220
python tools/bench_train_step.py --batch 4 --steps 5 2>/dev/null | tail -1 | cut -c1-300
python -m pytest tests/test_pretrain_step.py tests/test_drop_in.py -m gpu -x -q 2>&1 | tail -2
