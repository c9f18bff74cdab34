# Round evidence run: tests, smoke, bench (full), ncu launch list + full capture of the dominant kernel.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r1_tests.txt; cat gpurun_out/r1_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 30 --warmup 5 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; tail -c 2500 gpurun_out/r1_bench.json; tail -3 gpurun_out/r1_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 140 -c 130 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 1 --eager --no-extras > gpurun_out/r1_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemv[34]_kernel -s 70 -c 1 -o gpurun_out/r1_prof_gemv3 python bench.py --steps 2 --warmup 1 --eager --no-extras > gpurun_out/r1_ncu_b.log 2>&1
TMAC_BENCH_PREFETCH=0 ncu --set full --clock-control none -k regex:gemv[34]_kernel -s 70 -c 1 -o gpurun_out/r1_prof_gemv3_nopf python bench.py --steps 2 --warmup 1 --eager --no-extras > gpurun_out/r1_ncu_c.log 2>&1
ls -la gpurun_out | tail -12
