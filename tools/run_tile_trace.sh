# Tile-loop trace of conv2's epilogue group 0 (PROBES build swapped in on the GPU box only)
set -u
P=super-resolution-meta-attention-networks_b200
cp $P/libdfir_b200.so /tmp/libdfir_ship.so
cp tools/bin/libdfir_probes.so $P/libdfir_b200.so
EXTRA_PROBE=262144 FLUSH=0 DESC=1 timeout 60 python tools/trace_conv.py ${1:-sshl8fx} 2>&1 | python -c "
import sys
names=None
for l in sys.stdin:
    t=l.split()
    if t and t[0]=='row': names=t[1:]; print(' '.join('%9s'%n for n in ['row','tfull-wait','ld','issue','landed','upd0','upd1+fence','store','gap']))
    elif t and t[0].isdigit() and names and int(t[9])>0:
        v=dict(zip(names,map(int,t[1:])))
        print(' '.join('%9d'%x for x in [int(t[0]), v['e0_tfull']-v['e0_wait'], v['e0_release']-v['e0_tfull'], v['t_issued']-v['e0_release'], v['t_landed']-v['t_issued'], v['t_updated']-v['t_landed'], v['t2_fenced']-v['t_updated'], v['e0_end']-v['t2_fenced'], 0]))
    elif 'period' in l or 'prologue' in l or l.startswith('update') or l.startswith('  urow'): print(l.rstrip())
"
cp /tmp/libdfir_ship.so $P/libdfir_b200.so
