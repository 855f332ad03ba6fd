cd /root/repo
for h in 12 16 24 32; do
python bench.py --no-extra --steps 40 --warmup 5 --e2e-handles $h 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('handles $h: value %.1f e2e %.1f serial %.1f' % (d['value'], d['e2e']['value'], d['e2e']['serial_value']))"
done
