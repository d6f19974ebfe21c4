R=r02; O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_stack -s 2 -c 1 -f -o /tmp/${R}_stack_56 python tools/prof_one.py 256 256 56 56 auto 3 > $O/${R}_ncu_stack_56.log 2>&1
ncu -i /tmp/${R}_stack_56.ncu-rep --page raw --csv > $O/${R}_stack_56_raw.csv 2>/dev/null
ncu -i /tmp/${R}_stack_56.ncu-rep --page source --csv > $O/${R}_stack_56_src.csv 2>/dev/null
tail -2 $O/${R}_ncu_stack_56.log
