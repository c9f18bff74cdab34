mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
P='import json,sys; d=json.loads(sys.stdin.read()); r=d["roofline"]; g=r.get("grouped_launch",{}); print(round(d["value"]), round(d["ms_per_step"],4), "lone", round(r["us_per_launch"],2), round(r["frac"],3), "grouped", round(g.get("us_per_gemv",0),2), round(g.get("frac",0),3), g.get("launch"))'
for v in "X=0" "TMAC_B200_MINB=3" "TMAC_B200_CS=1 TMAC_B200_WPC=4" "TMAC_B200_CS=2 TMAC_B200_WPC=8"; do echo "== $v"; env $v python bench.py --steps 20 --warmup 3 --no-extras 2>>gpurun_out/b3.err | python -c "$P"; done
tail -3 gpurun_out/b3.err
