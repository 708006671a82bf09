import csv, sys
for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    H = rows[hdr]; ki = H.index('Kernel Name'); vi = H.index('Metric Value')
    seq = [(r[ki][:50], float(r[vi].replace(',', ''))) for r in rows[hdr + 1:] if len(r) > vi]
    idx = [i for i, (k, _) in enumerate(seq) if 'final' in k]
    start = idx[-2] + 1
    last = seq[start:idx[-1] + 1]
    rounds = [round(v / 1e3) for k, v in last if 'ba_round' in k]
    print(f, "rounds us", rounds, "sum", sum(rounds), "| all kernels us", round(sum(v for k, v in last) / 1e3))
