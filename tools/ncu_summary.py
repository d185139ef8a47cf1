"""Summarise an .ncu-rep (first kernel): key raw metrics, and source lines ranked by instructions / stall samples."""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 18
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, val = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_static', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed']
for i, h in enumerate(hdr):
    if h in want or ('issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h):
        try:
            if 'stalled' in h and float(val[i].replace(',', '')) < 0.15:
                continue
        except ValueError:
            pass
        print("%s | %s | %s" % (h, units[i], val[i]))
if top > 0:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[2]
    iL, iSamp, iInst = hdr.index('Line No'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    out, ti, ts = [], 0, 0
    for r in rows[3:]:
        if len(r) <= iInst or r[iL] == '':
            continue
        try:
            inst, samp = int(r[iInst]), int(r[iSamp])
        except ValueError:
            continue
        ti += inst; ts += samp; out.append((inst, samp, r[iL], r[1][:100]))
    print("# source lines by stall samples (total inst %d, samples %d)" % (ti, ts))
    for o in sorted(out, key=lambda x: -x[1])[:top]:
        print("%5.1f%% inst %5.1f%% samp  L%s  %s" % (100 * o[0] / max(ti, 1), 100 * o[1] / max(ts, 1), o[2], o[3]))
    print("# source lines by instructions")
    for o in sorted(out, reverse=True)[:top]:
        print("%5.1f%% inst %5.1f%% samp  L%s  %s" % (100 * o[0] / max(ti, 1), 100 * o[1] / max(ts, 1), o[2], o[3]))
