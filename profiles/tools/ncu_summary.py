"""Summarise an .ncu-rep (raw page) into a small metric x kernel CSV for profiles/."""
import csv, subprocess, sys
rep, out = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread',
 'lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed',
 'launch__grid_size','launch__block_size','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','l1tex__m_xbar2l1tex_read_bytes.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_tensor.sum', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio']
ki = hdr.index('Kernel Name')
names = [r[ki].split('(')[0].replace('void bg::', '') for r in data]
lines = [['metric', 'unit'] + names]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        lines.append([w, units[i]] + [r[i] for r in data])
txt = '\n'.join(','.join(l) for l in lines)
print(txt)
if out:
    open(out, 'w').write(txt + '\n')
