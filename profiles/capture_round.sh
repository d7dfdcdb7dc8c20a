set -x
cd /root/repo
python profiles/run_step.py --pipeline 4 --steps 1 --warmup 3 > gpurun_out/s3_plain_step.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -s 18 -c 6 -o gpurun_out/s3_step python profiles/run_step.py --pipeline 4 --steps 1 --warmup 3 > gpurun_out/s3_ncu_step.log 2>&1
ncu -i gpurun_out/s3_step.ncu-rep --page raw --csv > gpurun_out/s3_step_raw.csv 2>/dev/null
python profiles/run_step.py --pipeline 1 --steps 1 --warmup 3 > gpurun_out/s3_plain_strict.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 18 -c 6 --csv --log-file gpurun_out/s3_strict_launches.csv python profiles/run_step.py --pipeline 1 --steps 1 --warmup 3 > gpurun_out/s3_ncu_strict.log 2>&1
python bench.py --steps 2 --warmup 3 > gpurun_out/s3_bench_small.json 2> gpurun_out/s3_bench_small.err || exit 2
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/s3_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/s3_ncu_bench.log 2>&1
for wl in kitti64 kitti8 nyu; do python profiles/full_parity.py --workload $wl 2>&1 | tail -6; done
