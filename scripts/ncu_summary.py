#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into the handful of numbers we track.  usage: ncu_summary.py file.ncu-rep [points]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; points = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    g = lambda k: float(d[k][1].replace(",", "")) if k in d and d[k][1] not in ("", "no data") else float("nan")
    print("kernel:", d.get("Kernel Name", ("", "?"))[1][:100])
    cyc = g("sm__cycles_elapsed.avg")
    print(f"  duration {g('gpu__time_duration.sum'):.4f} {d['gpu__time_duration.sum'][0]}  cycles {cyc:.0f}  regs {g('launch__registers_per_thread'):.0f}  grid {g('launch__grid_size'):.0f} block {g('launch__block_size'):.0f}")
    print(f"  warps active % {g('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f}  issue active % {g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f}  fp64 pipe % {g('sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active'):.1f}  alu % {g('sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active'):.1f} lsu % {g('sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active'):.1f} xu % {g('sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active'):.1f} fma % {g('sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active'):.1f}")
    inst = g("smsp__inst_executed.sum")
    fp64 = sum(g(f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed") for op in ("dadd", "dmul", "dfma")) * cyc
    print(f"  warp inst {inst:.4g}  thread FP64 (dadd+dmul+dfma) {fp64:.4g}  shared ld {g('smsp__sass_inst_executed_op_shared_ld.sum'):.4g} st {g('smsp__sass_inst_executed_op_shared_st.sum'):.4g} branch {g('smsp__inst_executed_op_branch.sum'):.4g}")
    if points:
        print(f"  per point: thread inst {inst*g('smsp__thread_inst_executed_per_inst_executed.ratio')/points:.1f}  FP64 ops {fp64/points:.1f}  LDS {g('smsp__sass_inst_executed_op_shared_ld.sum')*32/points:.1f} STS {g('smsp__sass_inst_executed_op_shared_st.sum')*32/points:.1f}")
    print(f"  dram read {d['dram__bytes_read.sum']}  write {d['dram__bytes_write.sum']}  smem wavefronts {g('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'):.4g} bank conflicts {g('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'):.4g}")
    st = []
    for h in hdr:
        if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            try: st.append((float(d[h][1]), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
            except ValueError: pass
    print("  stalls/issue:", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
