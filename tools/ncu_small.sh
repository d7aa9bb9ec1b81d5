#!/bin/bash
# targeted `ncu --set full` captures, one launch per hot kernel (second iteration = warm), small reports
set -u
N="ncu --set full --clock-control none --kernel-name-base demangled"
cap() { # name mode regex skip
  $N -k regex:"$3" -s $4 -c 1 -o gpurun_out/r01_$1 -f python tools/profile_small.py $2 > gpurun_out/ncu_$1.log 2>&1 || echo "ncu $1 failed"
}
python tools/profile_small.py both > gpurun_out/prof_plain.log 2>&1 || { echo plain run failed; exit 1; }
W="${1:-all}"
if [ "$W" = all ] || [ "$W" = conv ]; then
cap conv1_stats   infer 'conv3x3_c64_tc_kernel<\(int\)64, \(int\)5, \(int\)0>' 2
cap conv2_ss      infer 'conv3x3_c64_tc_kernel<\(int\)64, \(int\)6, \(int\)0>' 4
cap dgrad_mask    train 'conv3x3_c64_tc_kernel<\(int\)64, \(int\)7, \(int\)0>' 2
cap wgrad_tc      train 'wgrad_c64_tc_kernel' 24
fi
if [ "$W" = all ] || [ "$W" = simt ]; then
cap wgrad_reduce  train 'wgrad_reduce_kernel' 24
cap bwd_reduce_ca train 'bwd_reduce_ca_kernel' 2
cap form_dr       train 'form_dr_kernel' 2
cap scale_resid   train 'scale_residual_kernel' 2
fi
ls -la gpurun_out/*.ncu-rep
