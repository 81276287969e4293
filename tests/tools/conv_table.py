"""Offline: pair the conv launches of an ncu launch list (tools/time_unet.py run) with the FFHQ UNet's conv shapes."""
import csv, sys, os
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))]
from oracle import unet_ref
path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches_unet.csv'
N = int(sys.argv[2]) if len(sys.argv) > 2 else 32
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]; data = rows[hi + 1:]
ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
names = [(r[ki].split('(')[0].replace('kdip::', '').replace('void ', ''), float(r[vi].replace(',', '')) / 1e3) for r in data if len(r) > vi]
idx = [i for i, (n, _) in enumerate(names) if n.startswith('im2col')]
seq = names[idx[-2]:]
import collections
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for n, t in seq:
    a = agg[n.split('<')[0]]; a[0] += 1; a[1] += t; a[2] = max(a[2], t)
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:28s} n={a[0]:4d} total={a[1]/1e3:8.3f} ms {100*a[1]/tot:5.1f}% max={a[2]:8.1f} us")
print('total', tot / 1e3)
convs = [t for n, t in seq if n.startswith('conv_gemm')]
cfg = unet_ref.ffhq_config(); plan = unet_ref.block_plan(cfg)
fw = []; res = 256
for b in plan:
    if b['kind'] == 'conv_in':
        fw.append((f"in 3->{b['cout']}@{res} (i2c K=64)", 2 * N * res * res * b['cout'] * 64))
    elif b['kind'] == 'res':
        if b['updown'] == 'down': res //= 2
        elif b['updown'] == 'up': res *= 2
        ci, co = b['cin'], b['cout']
        fw.append((f"c1 {ci}->{co}@{res}", 2 * N * res * res * co * ci * 9))
        k2 = co * 9 + (ci if ci != co else 0)
        fw.append((f"c2 {co}->{co}@{res}" + (f"+skip{ci}" if ci != co else f"+res{b['updown'] or ''}"), 2 * N * res * res * co * k2))
    else:
        c = b['cin']; fw.append((f"qkv {c}@{res}", 2 * N * res * res * 3 * c * c)); fw.append((f"proj {c}@{res}", 2 * N * res * res * c * c))
fw.append(("head 128->6(16)@256", 2 * N * 256 * 256 * 6 * 128 * 9))
nf = len(fw); tot = 0
for (d, f), t in zip(fw, convs[:nf]):
    print(f"{d:34s} {t:8.1f} us  {f/t/1e6:7.1f} TF/s"); tot += t
print('fwd conv total ms', tot / 1e3, 'TF/s', sum(f for _, f in fw) / tot / 1e6)
bw = convs[nf:]
print('bwd convs:', len(bw), 'total ms', sum(bw) / 1e3)
print([round(t) for t in bw])
