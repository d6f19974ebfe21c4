R=r02; O=gpurun_out
python bench.py --steps 5 --warmup 3 > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1.err; tail -c 200 $O/${R}_bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_kron -s 2 -c 1 -f -o /tmp/${R}_kron_7 python tools/prof_one.py 1024 2048 7 7 auto 3 > $O/${R}_ncu_kron_7.log 2>&1
ncu -i /tmp/${R}_kron_7.ncu-rep --page raw --csv > $O/${R}_kron_7_raw.csv 2>/dev/null
ncu -i /tmp/${R}_kron_7.ncu-rep --page source --csv > $O/${R}_kron_7_src.csv 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"score_|topk_|finalize_" -c 400 --csv --log-file $O/${R}_launches_bench_resnet50.csv python bench.py --steps 2 --warmup 3 --per-site --no-cpu-baseline --no-e2e --no-strong --no-u2netp > $O/${R}_ncu_launches.log 2>&1
echo done
