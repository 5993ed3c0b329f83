"""Print the headline metrics + SASS block histogram of an .ncu-rep (scratch helper for profiles/*.md)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
want += [c for c in h if c.startswith('smsp__average_warps_issue_stalled') and c.endswith('per_issue_active.ratio')]
for r in rows[2:]:
    for w in want:
        if w in h:
            v = r[h.index(w)]
            try:
                if float(v) == 0: continue
            except ValueError:
                pass
            print(w, v, rows[1][h.index(w)])
if len(sys.argv) > 2:
    sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(sass)))
    h = rows[1]; ie = h.index('Instructions Executed'); s = h.index('# Samples')
    data = [(r[1].strip(), int(r[ie]), int(r[s])) for r in rows[2:]]
    tot = sum(d[1] for d in data); ts = sum(d[2] for d in data)
    nsub = float(sys.argv[2])
    print('total instr', tot, 'per subgraph', tot / nsub, 'samples', ts)
    i = 0
    while i < len(data):
        j = i
        while j + 1 < len(data) and abs(data[j + 1][1] - data[i][1]) <= 0.03 * max(data[i][1], 1): j += 1
        bi = sum(d[1] for d in data[i:j + 1]); bs = sum(d[2] for d in data[i:j + 1])
        if bi > 0.01 * tot or bs > 0.01 * ts:
            print(f"[{i}-{j}] n={j - i + 1} exec/subg {data[i][1] / nsub:.1f} instr share {bi / tot:.3f} sample share {bs / ts:.3f}")
        i = j + 1
    if len(sys.argv) > 3:
        a, b = map(int, sys.argv[3].split('-'))
        for d in data[a:b + 1]: print(f"{d[0][:100]:100s} {d[1] / nsub:8.1f} {d[2]}")
