#!/bin/bash
# GPU bring-up: runs each group of -m gpu tests in its own process (a trapped kernel poisons only its own
# CUDA context) with a wall-clock limit, and collects everything in gpurun_out/bringup.log.
mkdir -p gpurun_out
LOG=gpurun_out/bringup.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv >> $LOG 2>&1
run() {
  echo "=== $*" >> $LOG
  timeout 600 python -m pytest "$@" -q -m gpu -p no:cacheprovider --timeout 300 >> $LOG 2>&1
  echo "=== exit $?" >> $LOG
}
GROUPS_OPS="test_head_conv test_conv_f32_simt test_meta_attention test_ca_scale_residual test_pool_rows_f32 test_bad_arguments test_conv_tc_descriptor test_conv_tc_bias_relu test_conv_tc_pool_rows test_conv_tc_skip test_conv_tc_tail test_conv_tc_pixel_shuffle"
for g in $GROUPS_OPS; do run tests/test_ops_gpu.py -k $g; done
run tests/test_qrcan_gpu.py -k fp32_mode
run tests/test_qrcan_gpu.py -k "bf16_mode"
run tests/test_qrcan_gpu.py -k "full_depth or batch_composition or roundtrip"
grep -E "^===|passed|failed|error" $LOG | tail -60
