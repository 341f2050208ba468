# Late round-2 single-GPU bundle (every step under its own timeout). Outputs -> gpurun_out/r2g_*
mkdir -p gpurun_out
O=gpurun_out
timeout -k 10 420 python -m pytest tests -m gpu -q --timeout 150 -p no:cacheprovider > $O/r2g_gpu_tests.log 2>&1; tail -3 $O/r2g_gpu_tests.log
timeout -k 10 300 python bench.py --steps 20 --warmup 3 > $O/r2g_bench_n1_s20.json 2> $O/r2g_bench_n1_s20.err; tail -c 400 $O/r2g_bench_n1_s20.json
NCU="ncu --set full --clock-control none --import-source on"
timeout -k 10 200 $NCU -k regex:stage_b_umma -s 2 -c 1 -f -o $O/r2g_prof_stage_b_tab python tools/run_decode.py c3 fp16 3 > $O/r2g_ncu_b.log 2>&1
LL="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout -k 10 200 $LL -s 9 -c 60 --log-file $O/r2g_launches_bench_c3.csv python bench.py --steps 4 --warmup 3 --no-extra > $O/r2g_bench_under_ncu.log 2>&1
timeout -k 10 200 python bench.py --no-extra > $O/r2g_bench_n1_s100.json 2> $O/r2g_bench_n1_s100.err; tail -c 300 $O/r2g_bench_n1_s100.json
ls -la $O | tail -12
