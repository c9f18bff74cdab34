# Round evidence run (one B200): tests, smoke, bench (full), ncu launch list + full captures of the dominant kernels.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r2_tests.txt; cat gpurun_out/r2_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 30 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 600 gpurun_out/r2_bench.json; tail -3 gpurun_out/r2_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --eager --no-extras > gpurun_out/r2_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemv3_kernel -s 70 -c 1 -o gpurun_out/r2_prof_gemv3 python bench.py --steps 2 --warmup 1 --eager --no-extras > gpurun_out/r2_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:seq_kernel -s 9 -c 1 -o gpurun_out/r2_prof_seq python tools/seq_bench.py --reps 2 > gpurun_out/r2_ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:prefill16 -s 1 -c 1 -o gpurun_out/r2_prof_pf16 python tools/pf_one.py 256 1 0 > gpurun_out/r2_ncu_d.log 2>&1
ls -la gpurun_out | grep r2_ | tail -14
