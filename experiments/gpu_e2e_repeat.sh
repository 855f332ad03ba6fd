cd /root/repo
for i in 1 2 3 4; do
python bench.py --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.1f e2e %.1f serial %.1f' % (d['value'], d['e2e']['value'], d['e2e']['serial_value']))"
done
