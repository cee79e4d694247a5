for v in A B C; do
echo "== variant $v (A: old potrf + DMMA tile, B: new potrf + DFMA tile, C: old + old)"
XB200_LIB=$PWD/tools/gpu/variants/libxb200_$v.so timeout 600 python bench.py --no-cpu-baseline --steps 30 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
print({k:v for k,v in d['stage_ms_per_update'].items() if 'chol' in k})"
done
