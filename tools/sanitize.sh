# compute-sanitizer passes over the GPU suite: memcheck on everything that launches the production kernels (gemv3 incl. bulk
# block loads and the fused LUT, the tcgen05 prefill tiles, the sequence kernel, block-format uploads); racecheck / initcheck on
# a representative subset (they are 20-50x slower).
mkdir -p gpurun_out
OUT=gpurun_out/r2_sanitizer.txt
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sequence.py -m gpu -x -q -k "golden or fused or grouped or general or prefill or ggml_block or gptq or host_call or (sequence_chain and (grid0 or grid3 or grid37)) or long_chain or w2_zp_g128" 2>&1 | tail -6 > $OUT
echo "memcheck done" >> $OUT
compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sequence.py -m gpu -x -q -k "w2_zp_g128 or w4_sym or bitnet_int32 or (fused and w2zp) or (sequence_chain and w2zp and (grid3 or grid7)) or (sequence_chain and w4-grid7)" > gpurun_out/r2_racecheck_full.txt 2>&1
grep -E "hazard|Hazard|passed|failed|SUMMARY" gpurun_out/r2_racecheck_full.txt | sort | uniq -c | sort -rn | head -20 >> $OUT
grep -m2 -B2 -A14 "^E " gpurun_out/r2_racecheck_full.txt | head -40 >> $OUT
compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sequence.py -m gpu -x -q -k "w2_zp_g128 or general or (sequence_chain and w2zp and grid7)" 2>&1 | tail -8 >> $OUT
cat $OUT | cut -c1-220
