mkdir -p gpurun_out
O=gpurun_out
timeout -k 10 330 python -m pytest tests -m gpu -q --timeout 150 -p no:cacheprovider > $O/r2h_gpu_tests.log 2>&1; tail -3 $O/r2h_gpu_tests.log
timeout -k 10 150 python bench.py --steps 20 --warmup 3 > $O/r2h_bench_n1_s20.json 2> $O/r2h_bench_n1_s20.err; tail -c 200 $O/r2h_bench_n1_s20.json
timeout -k 10 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
