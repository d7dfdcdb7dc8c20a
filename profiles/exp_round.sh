# scheduling experiments for the pipelined step (each line: env | exp.py output)
run() { echo "== $1"; env $1 python profiles/exp.py --depths 4 --steps 100 --check --caps ${2:--1} 2>&1 | tail -n ${3:-1}; }
run "X=0"
run "DTFILL_K1B_SMEM_PAD=100000"
run "DTFILL_K1B_SMEM_PAD=100000 DTFILL_PRIO_MODE=2"
run "DTFILL_K1B_SMEM_PAD=60000"
echo "== stage probe pad"; DTFILL_K1B_SMEM_PAD=100000 python profiles/stage_probe.py 2>&1 | tail -8
