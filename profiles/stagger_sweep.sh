for k in 1 2; do for st in 0 200 400 700 1000 1500; do
PVB_KERNEL_1024=$k PVB_STAGGER_NS=$st python bench.py --steps 1500 --warmup 100 --no-cpu-baseline --no-e2e --no-other-configs | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('kernel',$k,'stagger',$st, d['roofline']['kernel'], '%.3e'%d['value'], '%.3f'%d['roofline']['frac'], '%.1f us'%d['roofline']['avg_launch_us'], 'l2 %.3e'%d['l2_resident_value'], 'streams %.3e'%d['concurrent_streams_value'])"
done; done
