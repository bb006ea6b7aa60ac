"""Where a kernel's warp-stall samples fall: `ncu -i X.ncu-rep --page source --csv > X.csv`, then
python profiles/sass_regions.py X.csv [instructions per bucket].  Prints, per bucket of consecutive SASS
instructions, the share of samples and of executed instructions, the top stall reasons and the opcode mix."""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    hdr, data = rows[1], rows[2:]
    isamp, iinst, isrc = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Source')
    stall = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(float(r[isamp] or 0) for r in data)
    totinst = sum(float(r[iinst] or 0) for r in data)
    print(rows[0][1])
    print('samples %d, warp instructions %d, SASS lines %d' % (tot, totinst, len(data)))
    allst = {h: sum(float(r[i] or 0) for r in data) for i, h in stall}
    print('stalls overall: ' + ' '.join('%s %.1f%%' % (h[6:], v / tot * 100) for h, v in sorted(allst.items(), key=lambda x: -x[1])[:8]))
    for k in range(0, len(data), chunk):
        seg = data[k:k + chunk]
        s = sum(float(r[isamp] or 0) for r in seg)
        ins = sum(float(r[iinst] or 0) for r in seg)
        if s < tot * 0.003:
            continue
        st = sorted(((h, sum(float(r[i] or 0) for r in seg)) for i, h in stall), key=lambda x: -x[1])[:4]
        ops = {}
        for r in seg:
            w = r[isrc].split()
            op = w[1] if w[0].startswith('@') else w[0]
            ops[op] = ops.get(op, 0) + 1
        top = sorted(ops.items(), key=lambda x: -x[1])[:5]
        print('%5d-%5d samples %5.1f%% inst %5.1f%%  %s | %s' % (
            k, k + chunk, s / tot * 100, ins / totinst * 100,
            ' '.join('%s:%.0f' % (a[6:], b / max(s, 1) * 100) for a, b in st), ' '.join('%s x%d' % t for t in top)))


if __name__ == '__main__':
    main()
