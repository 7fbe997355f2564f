"""profiles/r2_ncu_igemm250_full_summary.txt, dominant_kernel.json and the AFTER half of the source-page summary from the
latest gpurun_out/r2_igemm250.ncu-rep / r2_igemm250_raw.csv (tools/ncu_round2.sh)"""
import csv
import json
import subprocess

rows = list(csv.reader(open('gpurun_out/r2_igemm250_raw.csv')))
hdr, units, vals = rows[0], rows[1], rows[2]
old = open('profiles/r2_ncu_igemm250_full_summary.txt').read().splitlines()
want = [l.split()[0] for l in old[4:] if l.strip()]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
out = old[:3] + ['%-106s %s' % ('Kernel Name', d['Kernel Name'][1])]
for w in want:
    if w in d:
        out.append('%-90s %-16s %s' % (w, d[w][0], d[w][1]))
open('profiles/r2_ncu_igemm250_full_summary.txt', 'w').write('\n'.join(out) + '\n')
g = lambda k: float(d[k][1])
p = 'profiles/dominant_kernel.json'
j = json.load(open(p))
r = j['round2_batch250']
rd, wr = int(g('dram__bytes_read.sum') * 1e6), int(g('dram__bytes_write.sum') * 1e6)
r.update(dram_bytes_read=rd, dram_bytes_write=wr, dram_bytes_per_launch=rd + wr,
         tensor_pipe_active_pct_of_elapsed=round(g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'), 2),
         tensor_pipe_active_pct_of_active=round(g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'), 2),
         duration_us_under_ncu=g('gpu__time_duration.sum'), l2_to_sm_bytes=int(g('l1tex__m_xbar2l1tex_read_bytes.sum') * 1e6))
j['dram_bytes_per_launch_b250'] = rd + wr
json.dump(j, open(p, 'w'), indent=1)
src = subprocess.run(['ncu', '-i', 'gpurun_out/r2_igemm250.ncu-rep', '--page', 'source', '--csv'], capture_output=True, text=True).stdout
open('/tmp/src_after.csv', 'w').write(src)
after = subprocess.run(['python', 'tools/ncu_source_summary.py', '/tmp/src_after.csv'], capture_output=True, text=True).stdout
cur = open('profiles/r2_ncu_igemm250_source_summary.txt').read()
head = cur[:cur.index('## AFTER')]
open('profiles/r2_ncu_igemm250_source_summary.txt', 'w').write(head + '## AFTER: the same launch with the final round-2 kernel (plain epilogue variant)\n' + after)
print(rd, wr, r['tensor_pipe_active_pct_of_elapsed'], r['duration_us_under_ncu'])
