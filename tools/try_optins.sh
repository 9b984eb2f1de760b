# Opt-in kernel variants built at the end of round 1 WITHOUT GPU time left to validate them.
# Each is the default path with one scheduling change; run the parity suite and the bench with the switch set, adopt if green + faster:
#   gpurun --timeout 600 -- 'bash tools/try_optins.sh'
set -x
for sw in "OESS_RADIX_PRELOAD=1" "OESS_ROWSORT=regs" "OESS_RADIX_PRELOAD=1 OESS_ROWSORT=regs"; do
  env $sw python -m pytest tests/test_voxel_gpu.py tests/test_dsec_slicer.py -m gpu -x -q 2>&1 | tail -1
  env $sw python bench.py --steps 20 --warmup 3 --host-output 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$sw', round(d['value']), round(d['value_single_stream']['value']), d['roofline']['kernel_share_of_step'])"
done
OESS_MHA_EX2=approx python -m pytest tests/test_maskclip.py -m gpu -x -q -s 2>&1 | grep -E "attention|ViT|passed|failed"
OESS_MHA_EX2=approx python tools/bench_tc.py --config2 2>/dev/null | grep maskclip | cut -c1-400
