"""Instruction mix of one ncu capture (source page): executed warp instructions per opcode, top lines.
    python tools/sass_mix.py file.ncu-rep [top]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if 'Source' in r and 'Address' in r)
H = rows[h]; si = H.index('Source'); ei = H.index('Instructions Executed'); sm = H.index('# Samples')
mix = collections.Counter(); samp = collections.Counter(); total = 0
for r in rows[h + 1:]:
    if len(r) <= ei: continue
    try: n = int(r[ei])
    except ValueError: continue
    src = r[si].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith('@') and len(toks) > 1 else (toks[0] if toks else '?')
    op = op.split('.')[0]
    mix[op] += n; total += n
    try: samp[op] += int(r[sm])
    except ValueError: pass
print(f"total warp instructions {total}")
st = sum(samp.values()) or 1
for op, n in mix.most_common(top):
    print(f"  {op:12s} {n:12d} {100.0 * n / total:5.1f} %   samples {100.0 * samp[op] / st:5.1f} %")
