# compute-sanitizer passes over the GPU suite (memcheck on everything that launches the production kernels incl. the
# prefill tile, the stream-K experiment and the block-format uploads; racecheck / initcheck on a representative subset).
mkdir -p gpurun_out
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or fused or grouped or general or prefill or stream_k or ggml_block or gptq or host_call" 2>&1 | tail -6 > gpurun_out/r1_sanitizer.txt
echo "memcheck done" >> gpurun_out/r1_sanitizer.txt
compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "w2_zp_g128 or w4_sym or bitnet_int32 or fused or (stream_k and grid7)" > gpurun_out/r1_racecheck_full.txt 2>&1
grep -E "hazard|Hazard|passed|failed|SUMMARY" gpurun_out/r1_racecheck_full.txt | sort | uniq -c | sort -rn | head -20 >> gpurun_out/r1_sanitizer.txt
grep -m2 -B2 -A14 "^E " gpurun_out/r1_racecheck_full.txt | head -40 >> gpurun_out/r1_sanitizer.txt
compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "w2_zp_g128 or general or (stream_k and grid37)" 2>&1 | tail -8 >> gpurun_out/r1_sanitizer.txt
cat gpurun_out/r1_sanitizer.txt | cut -c1-220
