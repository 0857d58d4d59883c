for cfg in "2048 512 2048 1.5" "256 64 8192 1.2" "512 128 8192 1.2" "1024 256 8192 1.2" "1024 256 8192 0.8" "2048 512 8192 1.2" "4096 1024 8192 1.2" "2048 128 2048 1.2" "1024 256 32768 1.25"; do
  set -- $cfg
  python bench.py --frame $1 --hop $2 --channels $3 --pitch $4 --steps 300 --warmup 30 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('N=%s hop=%s C=%s pf=%s: %.3e frames/s  %.1f us/launch  %.0f GB/s  frac %.3f  kernel %s  l2res %.3e streams %.3e' % ('$1','$2','$3','$4', d['value'], r['avg_launch_us'], r['achieved'], r['frac'], r['kernel'], d['l2_resident_value'], d['concurrent_streams_value']))"
done
