# First GPU call of the next round (one B200): what could not be run at the end of round 1 for lack of GPU minutes.
#   1. GPU suite + smoke (sanity on the fresh box)
#   2. compute-sanitizer over the code added late in round 1 (bulk block loads, stream-K experiment, block-format uploads)
#   3. steady-state timeline + launch-shape sweeps of the lone chain, for the VERDICT's kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 bash tools/sanitize.sh
timeout 300 python tools/pf16_check.py 2>&1 | tail -14 > gpurun_out/r2_pf16.txt; cat gpurun_out/r2_pf16.txt
timeout 200 python tools/sweep_lone.py 2>&1 | tail -22 > gpurun_out/r2_sweep_lone.txt; head -4 gpurun_out/r2_sweep_lone.txt
timeout 200 python tools/sweep_grouped.py 2>&1 | tail -8 > gpurun_out/r2_sweep_grouped.txt; cat gpurun_out/r2_sweep_grouped.txt
