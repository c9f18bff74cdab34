"""Turns the artefacts a tools/round_run.sh run left in gpurun_out/ into the tracked evidence under profiles/:
bench line, launch list, raw + summarised ncu metrics of the dominant kernel (lone launch, and the grouped launch when
gpurun_out/r1_prof_grouped.ncu-rep exists), executed opcode mix per block, DRAM traffic per launch."""
import collections
import csv
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct',
        'dram__bytes_read.sum.per_second'] + ['smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % k for k in
        ('long_scoreboard', 'wait', 'short_scoreboard', 'barrier', 'math_pipe_throttle', 'not_selected', 'dispatch_stall')]
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def page(rep, which):
    out = subprocess.run(["ncu", "-i", rep, "--page", which, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def summary(rep, note):
    rows = page(rep, "raw")
    d, u = dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))
    lines = ["kernel: " + d.get('Kernel Name', '?'), note] + ["%-90s %s %s" % (k, d[k], u[k]) for k in KEYS if k in d]
    return rows, d, u, "\n".join(lines) + "\n"


def mix(rep, blocks, title):
    rows = page(rep, "source")
    hi = [i for i, r in enumerate(rows) if "Source" in r][0]
    si, ei = rows[hi].index("Source"), rows[hi].index("Instructions Executed")
    c, tot = collections.Counter(), 0
    for r in rows[hi + 1:]:
        try:
            n = int(r[ei])
        except (ValueError, IndexError):
            continue
        toks = r[si].split()
        c[toks[1] if toks[0].startswith('@') else toks[0]] += n
        tot += n
    lines = [title, "executed warp-instructions per (super-block, chunk) block of 256 lookups per lane (total %d / %d blocks = %.1f):" % (tot, blocks, tot / blocks)]
    return "\n".join(lines + ["  %-22s %8.1f" % (op, n / blocks) for op, n in c.most_common(24)]) + "\n"


def main():
    shutil.copy(os.path.join(G, "r1_bench.json"), os.path.join(P, "r1_bench.json"))
    shutil.copy(os.path.join(G, "r1_launches.csv"), os.path.join(P, "r1_launches.csv"))
    rows, d, u, txt = summary(os.path.join(G, "r1_prof_gemv3_nopf.ncu-rep"),
                              "(one lone launch, no next-tensor L2 prefetch, cold cache, serialised by ncu: --set full --clock-control none)")
    open(os.path.join(P, "r1_gemv3_ncu_summary.txt"), "w").write(txt)
    with open(os.path.join(P, "r1_gemv3_ncu_raw.csv"), "w", newline="") as f:
        csv.writer(f).writerows(rows)
    traffic = float(d['dram__bytes_read.sum']) * UNIT[u['dram__bytes_read.sum']] + float(d['dram__bytes_write.sum']) * UNIT[u['dram__bytes_write.sum']]
    json.dump({"gemv_kernel_dram_bytes_per_launch": traffic, "source": "ncu --set full, r1_gemv3_ncu_raw.csv (lone launch, no next-tensor prefetch)",
               "algorithmic_bytes_per_launch": 12741632}, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    text = ""
    grouped = os.path.join(G, "r1_prof_grouped.ncu-rep")
    if os.path.exists(grouped):
        _, _, _, t2 = summary(grouped, "(ONE grouped launch = 32 GEMVs of the bench workload, serialised by ncu: --set full --clock-control none; 64-register variant)")
        open(os.path.join(P, "r1_gemv3_grouped_ncu_summary.txt"), "w").write(t2)
        text += mix(grouped, 32 * 2752, "gemv3_kernel<2,sym,8,4,minb4> grouped launch (32 GEMVs), ncu --page source:") + "\n"
    text += mix(os.path.join(G, "r1_prof_gemv3.ncu-rep"), 2752,
                "gemv3_kernel<2,sym,8,4,minb3,fused> lone launch (4 CTAs x 8 warps per super-block, LUT built in the kernel), ncu --page source:")
    text += "\nALU pipe: PRMT, LOP3, SHF, IADD3, ISETP, VIADD, LEA, MOV, SEL ...; FMA pipe: IDP (DP4A), IMAD*, FFMA/FADD/FMUL.\n"
    open(os.path.join(P, "r1_gemv3_opcode_mix.txt"), "w").write(text)
    print(txt)


if __name__ == "__main__":
    main()
