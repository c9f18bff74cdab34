# compute-sanitizer pass over a subset of the GPU parity tests (memcheck + racecheck of the shared-memory LUT / stage / DSMEM code)
mkdir -p gpurun_out
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or fused or grouped or general" 2>&1 | tail -6 > gpurun_out/r1_sanitizer.txt
echo "memcheck rc=$?" >> gpurun_out/r1_sanitizer.txt
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "w2_zp_g128 or w4_sym or bitnet_int32 or fused" 2>&1 | tail -6 >> gpurun_out/r1_sanitizer.txt
echo "racecheck rc=$?" >> gpurun_out/r1_sanitizer.txt
cat gpurun_out/r1_sanitizer.txt
python bench.py --steps 10 --warmup 3 2>gpurun_out/b4.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('e2e', d['e2e']); print(d['tokens_per_s'].get('prefill_seq256_one_tensor_11008x4096_w2'))"
tail -3 gpurun_out/b4.err
