mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
P='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"]), round(d["ms_per_step"],4), round(d["roofline"]["us_per_launch"],2), round(d["roofline"]["frac"],3), round(d["e2e"]["value"]), d["roofline"].get("grouped_launch"))'
for v in "X=0" "TMAC_B200_CS=4 TMAC_B200_WPC=8" "TMAC_B200_CS=4 TMAC_B200_WPC=8 TMAC_B200_LIB=tools/lib_b3.so" "TMAC_B200_CS=8 TMAC_B200_WPC=4 TMAC_B200_LIB=tools/lib_b3.so" ; do echo "== $v"; env $v python bench.py --steps 20 --warmup 3 --no-extras 2>>gpurun_out/b3.err | python -c "$P"; done
tail -3 gpurun_out/b3.err
