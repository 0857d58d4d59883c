for cfg in "4096 0.8" "8192 0.8" "32768 1.25" "32768 0.8"; do set -- $cfg
for k in 1 2; do
PVB_KERNEL_1024=$k python bench.py --channels $1 --pitch $2 --steps 600 --warmup 50 --no-cpu-baseline --no-e2e --no-other-configs | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('C=$1 pf=$2', d['roofline']['kernel'], '%.3e'%d['value'], 'frac %.3f'%d['roofline']['frac'], '%.1f us'%d['roofline']['avg_launch_us'], 'streams %.3e'%d['concurrent_streams_value'])"
done; done
