# usage: run_variants.sh name...   (libraries built by build_variant.sh under _variants/)
for v in "$@"; do
  echo "== variant $v"
  DTFILL_LIB=/root/repo/_variants/libdtfill_$v.so python profiles/exp.py --depths 1,4 --check --kernels --steps 40 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['depth'], d['ms'], d['frac'], 'k2', d['kernel_ms']['k2_chamfer'], 'parity', d.get('parity8'))"
done
