"""Summarise an `ncu --page source --csv --print-source cuda,sass` export: samples per CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
def num(v):
    try: return int(float(v))
    except ValueError: return 0
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == 'File Path':
        path = rows[i][1]; j = i + 1
        while j < len(rows) and '# Samples' not in rows[j]: j += 1
        hdr = rows[j]; k = j + 1; data = []
        while k < len(rows) and not (rows[k] and rows[k][0] == 'File Path'):
            if len(rows[k]) == len(hdr): data.append(rows[k])
            k += 1
        ix = {}
        for n, h in enumerate(hdr): ix.setdefault(h, n)
        stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
        tot = sum(num(r[ix['# Samples']]) for r in data)
        print("==", path, "lines", len(data), "samples", tot)
        agg = {s: sum(num(r[ix[s]]) for r in data) for s in stalls}
        print("   ", ", ".join("%s %.1f%%" % (s[6:], 100.0 * v / max(tot, 1)) for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
        for r in sorted(data, key=lambda r: -num(r[ix['# Samples']]))[:topn]:
            st = sorted(stalls, key=lambda s: -num(r[ix[s]]))[:2]
            print("%6d L%-4s %-86s | %s | inst %s" % (num(r[ix['# Samples']]), r[0], r[1].strip()[:86],
                  " ".join("%s=%s" % (s[6:], r[ix[s]]) for s in st), r[ix['Instructions Executed']]))
        i = k
    else:
        i += 1
