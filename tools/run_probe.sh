for bits in 262144 786432 1310720 1835008; do
  echo "EXTRA_PROBE=$bits"
  EXTRA_PROBE=$bits DESC=1 timeout 60 python tools/trace_conv.py sshl8stats 2>&1 | python -c "
import sys
L=[l.split() for l in sys.stdin if l.strip() and l.split()[0].isdigit()]
w=[int(r[9]) for r in L if int(r[9])>0]
print('group-0 row period (clk):', [w[i+1]-w[i] for i in range(2,min(10,len(w)-1))])
u=[(int(r[14])-int(r[13])) for r in L[2:10] if int(r[13])>0]
print('update half0 (clk):', u)
"
done
