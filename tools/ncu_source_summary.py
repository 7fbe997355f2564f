"""Summarise the source (SASS) page of an `ncu --set full --import-source on` capture of igemm_kernel:
warp-stall samples by kernel region (producer / MMA issuer / epilogue warps are contiguous address ranges), the stall
reasons per region and the opcode mix of the epilogue's per-chunk body.

    ncu -i gpurun_out/r2_igemm250.ncu-rep --page source --csv > /tmp/src.csv
    python tools/ncu_source_summary.py /tmp/src.csv > profiles/r2_ncu_igemm250_source_summary.txt
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
S, E = ix['# Samples'], ix['Instructions Executed']
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[S]) for r in data)
ex = [int(r[E]) for r in data]
# the epilogue body executes once per (epilogue warp, tile, box, chunk): the largest execution count shared by > 100 instructions
cnt = collections.Counter(e for e in ex if e > 1000)
chunk_exec = max(c for c in cnt if cnt[c] > 100)
first = next(i for i, e in enumerate(ex) if e == chunk_exec)
last = len(ex) - 1 - next(i for i, e in enumerate(reversed(ex)) if e == chunk_exec)
regions = [(0, first, 'before the epilogue body (setup, TMA producer, MMA issuer, barrier waits)'),
           (first, last + 1, 'epilogue body (8 warps; executed %d times = warps x tiles x boxes x chunks)' % chunk_exec),
           (last + 1, len(data), 'after the epilogue body (store hand-off, teardown, out-of-line paths)')]
print('# %s' % rows[0][1][:120])
print('# %d SASS instructions, %d warp-stall samples' % (len(data), tot))
for lo, hi, name in regions:
    n = sum(int(r[S]) for r in data[lo:hi])
    t = collections.Counter()
    for r in data[lo:hi]:
        for h in stalls:
            t[h[6:]] += int(r[ix[h]])
    live = sum(1 for r in data[lo:hi] if int(r[E]) > 0)
    print('\n%s\n  instructions %d (%d executed), samples %d (%.1f %%)' % (name, hi - lo, live, n, 100.0 * n / tot))
    print('  stall reasons: ' + ', '.join('%s %d' % kv for kv in t.most_common(7)))
ops = collections.Counter()
for r in data[first:last + 1]:
    if int(r[E]) == chunk_exec:
        tok = r[1].split()
        ops[tok[1] if tok[0].startswith('@') else tok[0]] += 1
print('\nopcode mix of one pass of the epilogue body (32 pixels of one channel per thread): %d instructions' % sum(ops.values()))
print('  ' + ', '.join('%s %d' % kv for kv in ops.most_common(16)))
top = sorted(range(len(data)), key=lambda i: -int(data[i][S]))[:12]
print('\nmost-sampled instructions:')
for i in sorted(top):
    r = data[i]
    st = sorted(((int(r[ix[h]]), h[6:]) for h in stalls), reverse=True)[:2]
    print('  #%-5d samples %-4s executed %-7s %-46s %s' % (i, r[S], r[E], r[1].strip()[:46], ', '.join('%s %d' % (h, v) for v, h in st if v)))
