"""Turns the artefacts a tools/round_run.sh run left in gpurun_out/ into the tracked evidence under profiles/ (round 2):
bench line, launch list, summarised ncu metrics of the three dominant kernels (gemv3 launch chain, decode sequence kernel,
fp16 prefill tile), a SASS excerpt proving the Blackwell instructions, DRAM traffic per launch."""
import csv
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct',
        'dram__bytes_read.sum.per_second', 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_active.avg', 'sm__cycles_elapsed.avg'] + \
       ['smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % k for k in
        ('long_scoreboard', 'wait', 'short_scoreboard', 'barrier', 'math_pipe_throttle', 'not_selected', 'dispatch_stall', 'sleeping')]
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))


def summary(rep, note, dst):
    d, u = raw(rep)
    lines = ["kernel: " + d.get('Kernel Name', '?'), note] + ["%-90s %s %s" % (k, d[k], u[k]) for k in KEYS if k in d]
    open(os.path.join(P, dst), "w").write("\n".join(lines) + "\n")
    return d, u


def main():
    shutil.copy(os.path.join(G, "r2_bench.json"), os.path.join(P, "r2_bench.json"))
    shutil.copy(os.path.join(G, "r2_launches.csv"), os.path.join(P, "r2_launches.csv"))
    for extra in ("r2_bench_n2.json", "r2_bench_n8.json", "r2_sanitizer.txt", "r2_sanitizer_late.txt", "r2_pf16.txt"):
        if os.path.exists(os.path.join(G, extra)):
            shutil.copy(os.path.join(G, extra), os.path.join(P, extra))
    d, u = summary(os.path.join(G, "r2_prof_gemv3.ncu-rep"), "(one launch of the chain of bench.py --eager: fused LUT build, cold cache, serialised by ncu: --set full --clock-control none)",
                   "r2_gemv3_ncu_summary.txt")
    traffic = float(d['dram__bytes_read.sum']) * UNIT[u['dram__bytes_read.sum']] + float(d['dram__bytes_write.sum']) * UNIT[u['dram__bytes_write.sum']]
    # profiles/traffic.json is maintained by hand from these captures (see its `source`): the chain launch also prefetches the next tensor
    summary(os.path.join(G, "r2_prof_seq.ncu-rep"), "(one persistent launch = 32 GEMVs 11008x4096 W2, dependent chain, tools/seq_bench.py; --set full --clock-control none)",
            "r2_seq_ncu_summary.txt")
    if os.path.exists(os.path.join(G, "r2_prof_chain.ncu-rep")):
        summary(os.path.join(G, "r2_prof_chain.ncu-rep"), "(one persistent launch = 32 GEMVs 11008x4096 W2, dependent chain, resident chain kernel, tools/seq_bench.py --impl 1; --set full --clock-control none)",
                "r2_chain_ncu_summary.txt")
    summary(os.path.join(G, "r2_prof_pf16.ncu-rep"), "(prefill tile, N = 256 tokens x 11008x4096 W2 g128 zp, tools/pf_one.py 256 1 0; --set full --clock-control none)",
            "r2_prefill16_ncu_summary.txt")
    # SASS excerpt: the Blackwell-only mnemonics of the in-tree library
    so = os.path.join(ROOT, "t-mac_b200", "libtmac_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout.splitlines()
    pats = ("UTCHMMA", "UTCIMMA", "UTCBAR", "LDTM", "UBLKCP", "UBLKPF", "SYNCS", "IDP.4A", "PRMT", "UTCATOM", "ACQBULK", "NANOSLEEP", "ST.E.64.STRONG.SYS", "STG.E.STRONG.SYS", "CCTL", "STAS", "UCGABAR_ARV", "UCGABAR_WAIT")
    out = ["cuobjdump -sass t-mac_b200/libtmac_b200.so: occurrences of Blackwell / hot-path mnemonics, then the first 3 lines of each", ""]
    for pt in pats:
        hits = [l.strip() for l in sass if pt in l]
        out.append("%-22s x %d" % (pt, len(hits)))
    out.append("")
    for pt in pats:
        for l in [l.strip() for l in sass if pt in l][:3]:
            out.append(l[:150])
    open(os.path.join(P, "r2_sass_excerpt.txt"), "w").write("\n".join(out) + "\n")
    print("profiles/ updated; gemv3 traffic per launch = %.0f B" % traffic)


if __name__ == "__main__":
    main()
