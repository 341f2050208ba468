set -x
compute-sanitizer --tool memcheck python tools/run_decode.py c1 bf16 1 2>&1 | tail -2 > gpurun_out/sanitizer_memcheck.txt
compute-sanitizer --tool synccheck python tools/run_decode.py c1 bf16 1 2>&1 | tail -2 > gpurun_out/sanitizer_synccheck.txt
cat gpurun_out/sanitizer_memcheck.txt gpurun_out/sanitizer_synccheck.txt
ncu --set full --clock-control none --import-source on -k regex:stage_b_umma -s 2 -c 1 -o gpurun_out/prof_stage_b_r1c -f python tools/run_decode.py c3 bf16 3 > gpurun_out/ncu_b3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stage_a_umma -s 2 -c 1 -o gpurun_out/prof_stage_a_r1c -f python tools/run_decode.py c3 bf16 3 > gpurun_out/ncu_a3.log 2>&1
python bench.py > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err
tail -c 600 gpurun_out/bench_r1_n1.json; tail -2 gpurun_out/bench_r1_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 45 --csv --log-file gpurun_out/launches_r1_bench.csv python bench.py --steps 10 --warmup 3 --no-extra > gpurun_out/bench_under_ncu.log 2>&1
tail -4 gpurun_out/launches_r1_bench.csv | cut -c1-200
