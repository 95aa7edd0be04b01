for ks in 1 2 4 8; do
  echo "== KSPLIT $ks"
  RA_UMMA_KSPLIT=$ks python tools/dbg_parity.py kitti 256 512 20 2 2>&1 | grep -E "ctrl_out|attn_box|y_out " | cut -c1-60
  RA_UMMA_KSPLIT=$ks python tools/dbg_parity.py kitti 256 512 20 2 2>&1 | grep -E "ctrl_out|attn_box|y_out " | python -c "
import sys
for l in sys.stdin:
  f=l.split(); v=[float(x) for x in f[6:]]; print(f[0], 'max', max(v), 'mean', sum(v)/len(v))"
  RA_UMMA_KSPLIT=$ks python bench.py --steps 10 --no-train-step --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step', d['ms_per_step'])"
done
RA_CONV_FP32=1 python tools/dbg_parity.py kitti 256 512 20 2 2>&1 | grep -E "ctrl_out|attn_box|y_out " | python -c "
import sys
for l in sys.stdin:
  f=l.split(); v=[float(x) for x in f[6:]]; print('fp32conv', f[0], 'max', max(v), 'mean', sum(v)/len(v))"
