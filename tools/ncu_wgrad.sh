python tools/profile_small.py train > gpurun_out/prof_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --kernel-name-base demangled -k regex:'wgrad_c64_tc_kernel' -s 24 -c 1 -o gpurun_out/r01_wgrad_tc_v2 -f python tools/profile_small.py train > gpurun_out/ncu_wgrad_tc_v2.log 2>&1
ls -la gpurun_out/r01_wgrad_tc_v2.ncu-rep
