t0=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-80
timeout 400 python bench.py --steps 30 --no-extras 2>/dev/null | tail -1 | python -c "
import sys, json
j=json.loads(sys.stdin.read())
print('C2', round(j['value']/1e6,1), 'e2e', round(j['e2e']['value']/1e6,1), j['e2e']['parity_ok'], j['parity'], j['roofline']['bound'], round(j['roofline']['frac'],3), j['clocks'])
"
echo "total: $(( $(date +%s) - t0 )) s"
