# Late-round evidence run (one B200) for the kernels added after tools/round_run.sh: ncu launch list of the bench step, one
# --set full capture of the resident chain kernel (the bench's headline submission), compute-sanitizer memcheck over the chain
# kernel, the fused integer path and the grouped one-call form.
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --eager --no-extras > gpurun_out/r2_ncu_a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 3 -c 1 -o gpurun_out/r2_prof_chain python tools/seq_bench.py --impl 1 --reps 2 > gpurun_out/r2_ncu_e.log 2>&1
tail -2 gpurun_out/r2_ncu_e.log | cut -c1-200
OUT=gpurun_out/r2_sanitizer_late.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sequence.py -m gpu -x -q -k "integer_path or gemv_grouped or (sequence_chain and grid0 and (w2zp or partial or ragged or w3zp)) or resident_chain" 2>&1 | tail -6 > $OUT
echo "memcheck done" >> $OUT
cat $OUT | cut -c1-200
ls -la gpurun_out | grep -E "r2_prof_chain|r2_launches" 
