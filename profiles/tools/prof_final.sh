set -x
python profiles/prof_workload.py > gpurun_out/r02f_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'rollout_kernel|nn_scan_kernel|nn_topk_kernel' -c 7 -f -o gpurun_out/r02f_full python profiles/prof_workload.py > gpurun_out/r02f_ncu.log 2>&1
python profiles/tools/prof_throughput.py > gpurun_out/r02f_thr_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -c 1 -f -o gpurun_out/r02f_throughput python profiles/tools/prof_throughput.py > gpurun_out/r02f_thr_ncu.log 2>&1
python bench.py --steps 1 --warmup 3 --legs "" --skip-cpu > gpurun_out/r02f_bench_short.json 2> gpurun_out/r02f_bench_short.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 1 --warmup 3 --legs "" --skip-cpu > gpurun_out/r02f_launch_ncu.log 2>&1
ls -la gpurun_out/r02f_*
tail -3 gpurun_out/r02f_ncu.log gpurun_out/r02f_thr_ncu.log gpurun_out/r02f_launch_ncu.log
