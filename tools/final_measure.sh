# Round-end measurement bundle on ONE GPU (run under gpurun; every risky step under its own timeout). Outputs -> gpurun_out/r2f_*
mkdir -p gpurun_out
O=gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -q --timeout 500 --timeout-method=thread -p no:cacheprovider > $O/r2f_gpu_tests.log 2>&1; tail -3 $O/r2f_gpu_tests.log
timeout -k 10 600 python bench.py --steps 20 --warmup 3 > $O/r2f_bench_n1_s20.json 2> $O/r2f_bench_n1_s20.err; tail -c 300 $O/r2f_bench_n1_s20.json
timeout -k 10 600 python bench.py --no-extra > $O/r2f_bench_n1_s100.json 2> $O/r2f_bench_n1_s100.err
timeout -k 10 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2f_bench_reference_arm.json 2>&1
timeout -k 10 300 python tools/time_wirings.py c3 > $O/r2f_wirings_c3.txt 2>&1
for e in 1 0; do timeout -k 10 100 python tools/run_liif.py c2x4 fp16 5 $e >> $O/r2f_liif.txt 2>&1; done
timeout -k 10 100 python tools/run_liif.py c2x4 fp32 3 1 >> $O/r2f_liif.txt 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout -k 10 300 $NCU -k regex:stage_b_umma -s 2 -c 1 -f -o $O/r2f_prof_stage_b_sel python tools/run_decode.py c3 fp16 3 > $O/r2f_ncu_b.log 2>&1
timeout -k 10 300 $NCU -k regex:stage_a_umma -s 2 -c 1 -f -o $O/r2f_prof_stage_a_p16 python tools/run_decode.py c3 fp16 3 > $O/r2f_ncu_a.log 2>&1
timeout -k 10 300 $NCU -k regex:stage_a_umma -s 4 -c 2 -f -o $O/r2f_prof_stage_a_matrix python tools/run_decode.py c2x4 fp16 1 3 1 > $O/r2f_ncu_am.log 2>&1
timeout -k 10 300 $NCU -k regex:stage_b_umma -s 0 -c 1 -f -o $O/r2f_prof_stage_b_liif python tools/run_liif.py c2x4 fp16 1 1 > $O/r2f_ncu_liif.log 2>&1
LL="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout -k 10 300 $LL -s 9 -c 60 --log-file $O/r2f_launches_bench_c3.csv python bench.py --steps 4 --warmup 3 --no-extra > $O/r2f_bench_under_ncu.log 2>&1
timeout -k 10 300 $LL -c 200 --log-file $O/r2f_launches_initq_c3.csv python tools/run_decode.py c3 fp16 1 3 1 > /dev/null 2>&1
timeout -k 10 300 $LL -c 100 --log-file $O/r2f_launches_mode1_c3.csv python tools/run_decode.py c3 fp16 1 1 0 > /dev/null 2>&1
timeout -k 10 300 $LL -c 40 --log-file $O/r2f_launches_liif_c2x4.csv python tools/run_liif.py c2x4 fp16 1 1 > /dev/null 2>&1
for args in "c1 fp16 1 3 1" "c1 fp16 1 1 0" "c1 fp16 1 2 1" "c1 fp16 1 4 0"; do
  echo "== memcheck run_decode $args" >> $O/r2f_sanitizer.txt
  timeout -k 10 300 compute-sanitizer --tool memcheck python tools/run_decode.py $args 2>&1 | tail -3 >> $O/r2f_sanitizer.txt
done
echo "== memcheck run_liif c1 fp16 1 1" >> $O/r2f_sanitizer.txt
timeout -k 10 300 compute-sanitizer --tool memcheck python tools/run_liif.py c1 fp16 1 1 2>&1 | tail -3 >> $O/r2f_sanitizer.txt
echo "== synccheck run_decode c1 fp16 1 3 1" >> $O/r2f_sanitizer.txt
timeout -k 10 300 compute-sanitizer --tool synccheck python tools/run_decode.py c1 fp16 1 3 1 2>&1 | tail -3 >> $O/r2f_sanitizer.txt
tail -30 $O/r2f_sanitizer.txt
