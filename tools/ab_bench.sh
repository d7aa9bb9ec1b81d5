# Same-box A/B of the Q-RCAN forward: library variants under tools/bin/libdfir_<name>.so against the library in the tree
set -u
P=super-resolution-meta-attention-networks_b200
cp $P/libdfir_b200.so /tmp/libdfir_new.so
for rep in 1 2; do
  for v in ${VARIANTS:-head}; do
    cp tools/bin/libdfir_$v.so $P/libdfir_b200.so; printf "%-12s" "$v:"; python tools/split_bench.py 32 10
  done
  cp /tmp/libdfir_new.so $P/libdfir_b200.so; printf "%-12s" "tree:"; python tools/split_bench.py 32 10
done
