"""Attribute ncu SASS-level samples / instruction counts to CUDA source lines.
usage: python scripts/ncu_lines.py <rep.ncu-rep> <cubin> <mangled kernel substring> [top]"""
import collections, csv, io, re, subprocess, sys
rep, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
# locate the function section
line_of = {}
cur = None; infunc = False
for ln in dis.splitlines():
    if ln.strip().startswith('.section'):
        infunc = kern in ln
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m and infunc:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
    m2 = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m2 and infunc:
        line_of[int(m2.group(1), 16)] = cur
src = list(csv.reader(io.StringIO(subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout)))
hdr = src[1]; data = src[2:]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
tot_s = tot_i = 0
stall_keys = [k for k in hdr if k.startswith('stall_') and '(' not in k]
addr0 = None
for r in data:
    try:
        addr = int(r[ix['Address']], 16)
    except ValueError:
        continue
    if addr0 is None:
        addr0 = addr
    key = line_of.get(addr - addr0) or ('?', 0)
    s = float(r[ix['# Samples']] or 0); i = float(r[ix['Instructions Executed']] or 0)
    a = agg[key]; a[0] += s; a[1] += i
    for k in stall_keys:
        a[2][k] += float(r[ix[k]] or 0)
    tot_s += s; tot_i += i
print(f'total samples {tot_s:.0f}, warp instructions {tot_i:.4g}, mapped lines {len(agg)}')
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ', '.join(f'{k[6:]} {v / max(a[0], 1) * 100:.0f}%' for k, v in a[2].most_common(3))
    print(f'{a[0] / tot_s * 100:5.1f}% samples {a[1] / tot_i * 100:5.1f}% inst  {key[0]}:{key[1]}   [{st}]')
