#!/bin/bash
# Round evidence in one GPU call: tests, bench lines, ncu launch list and one `ncu --set full` capture per kernel.
# Everything lands in gpurun_out/; summaries are made from it on the build host (tools/ncu_summary.py, tools/traffic_from_ncu.py).
R=${1:-r02}
O=gpurun_out
if [ -z "$SKIP_TESTS" ]; then timeout 900 python -m pytest tests -m gpu -x -q > $O/${R}_pytest_gpu.log 2>&1; tail -2 $O/${R}_pytest_gpu.log; fi
python bench.py --steps 5 --warmup 3 > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1.err; tail -c 300 $O/${R}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${R}_bench_reference_arm.json 2> $O/${R}_bench_reference_arm.err
for net in resnet_56 vgg_16_bn googlenet densenet_40; do
  python bench.py --net $net --steps 5 --warmup 3 --no-cpu-baseline --no-u2netp > $O/${R}_bench_$net.json 2> $O/${R}_bench_$net.err
  python bench.py --net $net --steps 5 --warmup 3 --no-cpu-baseline --no-u2netp --graph > $O/${R}_bench_${net}_graph.json 2>> $O/${R}_bench_$net.err
done
# launch list of the bench command (per-launch durations are cold-cache and serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"score_|topk_|finalize_" -c 400 --csv \
  --log-file $O/${R}_launches_bench_resnet50.csv python bench.py --steps 2 --warmup 3 --per-site --no-cpu-baseline --no-e2e --no-strong --no-u2netp > $O/${R}_ncu_launches.log 2>&1
cap() {  # name, kernel regex, prof_one arguments...
  local name=$1 k=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o /tmp/${R}_$name python tools/prof_one.py "$@" > $O/${R}_ncu_$name.log 2>&1
  # (a report is ~20 MB and gpurun brings back 64 MB at most: only the two CSV pages travel)
  ncu -i /tmp/${R}_$name.ncu-rep --page raw --csv > $O/${R}_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/${R}_$name.ncu-rep --page source --csv > $O/${R}_${name}_src.csv 2>/dev/null
}
cap stack_56 score_stack 256 256 56 56 auto 3
cap stack_28 score_stack 256 512 28 28 auto 3
cap stack_14 score_stack 256 1024 14 14 auto 3
cap kron_7 score_kron 256 2048 7 7 auto 3
cap large_320 score_large 12 64 320 320 auto 3
du -sh $O
