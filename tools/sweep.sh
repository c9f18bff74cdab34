mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
P='import json,sys; d=json.loads(sys.stdin.read()); r=d["roofline"]; g=r.get("grouped_launch",{}); print("value", round(d["value"]), round(d["ms_per_step"]*1e3/32,2), "us/layer | two-call", round(r["two_call_step"]["GBps"]), "| lone", round(r["us_per_launch"],2), round(r["frac"],3), "| grouped", round(g.get("us_per_gemv",0),2), round(g.get("frac",0),3), "| e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"]*1e3/32,1)); print(d.get("tokens_per_s"))'
python bench.py --steps 20 --warmup 3 2>>gpurun_out/b3.err | python -c "$P"
TMAC_B200_PDL_LATE=0 python bench.py --steps 20 --warmup 3 --no-extras 2>>gpurun_out/b3.err | python -c "$P"
tail -3 gpurun_out/b3.err
