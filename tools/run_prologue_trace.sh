# Prologue / pipeline trace of conv2 as the network launches it (PROBES build of the library swapped in on the GPU box only)
set -u
P=super-resolution-meta-attention-networks_b200
cp $P/libdfir_b200.so /tmp/libdfir_ship.so
cp tools/bin/libdfir_probes.so $P/libdfir_b200.so
for fl in 1 0; do
  echo "== FLUSH=$fl"
  FLUSH=$fl DESC=1 timeout 60 python tools/trace_conv.py sshl8fx 2>&1 | grep -v "^ *[3-9][0-9] \|^ *2[89] " | tail -40
done
cp /tmp/libdfir_ship.so $P/libdfir_b200.so
