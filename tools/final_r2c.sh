#!/bin/bash
# Round-2 (third part) validation on the GPU box: full GPU test suite, smoke, both bench arms, ncu launch list + full
# captures of the two trunk kernels as launched (conv1 + statistics, conv2 on the hi / 8-bit lo stream) then the
# prologue / tile-loop trace of conv2 (PROBES build under tools/bin).
set -u
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.log 2>&1
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err
if [ -n "${NO_NCU:-}" ]; then tail -4 gpurun_out/pytest_final.log; tail -3 gpurun_out/smoke_final.log; cut -c1-400 gpurun_out/bench_final.json; exit 0; fi
python tools/fwd_once.py 32 32 2 > gpurun_out/fwd_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02c_launches_infer_32x128.csv \
  python tools/fwd_once.py 32 32 2 > gpurun_out/ncu_launches_b.log 2>&1
N="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
python tools/profile_small.py infer > gpurun_out/prof_plain.log 2>&1 && {
$N -k regex:'conv3x3_c64_tc_kernel<\(int\)64, \(int\)5, \(int\)0>' -s 2 -c 1 -o gpurun_out/r02c_conv1_stats -f python tools/profile_small.py infer > gpurun_out/ncu_conv1_stats_b.log 2>&1
$N -k regex:'conv3x3_c64_tc_kernel<\(int\)64, \(int\)9, \(int\)0>' -s 4 -c 1 -o gpurun_out/r02c_conv2_hl8 -f python tools/profile_small.py infer > gpurun_out/ncu_conv2_hl8_b.log 2>&1
}
tail -4 gpurun_out/pytest_final.log; cat gpurun_out/smoke_final.log | tail -3; cat gpurun_out/bench_final.json | cut -c1-400; cat gpurun_out/bench_ref_final.json | cut -c1-200; ls -la gpurun_out/*r02c*
bash tools/run_tile_trace.sh > gpurun_out/r02c_trace_conv2_fx_final.log 2>&1; tail -6 gpurun_out/r02c_trace_conv2_fx_final.log
