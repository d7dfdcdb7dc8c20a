for v in "$@"; do
  echo "== variant $v"
  DTFILL_LIB=/root/repo/_variants/libdtfill_$v.so python profiles/exp.py --depths 1 --sky 8 --caps 110,130 --check --kernels
  DTFILL_LIB=/root/repo/_variants/libdtfill_$v.so python profiles/exp.py --depths 4 --sky 8 --caps 180
done
