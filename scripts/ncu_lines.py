"""Per-source-line stall samples of one kernel from an .ncu-rep captured with --import-source on (cuda,sass source page)."""
import csv, collections, subprocess, sys

def main(path, top=45):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    cur, agg, hdr = None, collections.OrderedDict(), None
    num = lambda s: int(s) if s.strip().lstrip('-').isdigit() else 0
    for r in rows:
        if len(r) == 2 and r[0] == 'File Path':
            cur = r[1].split('/')[-1]; continue
        if len(r) > 5 and r[0] == 'Line No':
            hdr = r; continue
        if hdr is None or len(r) < 7 or not r[0].isdigit():
            continue
        col = lambda name: num(r[hdr.index(name)])
        agg[(cur, int(r[0]))] = (num(r[6]), num(r[7]), r[1].strip()[:100], col('stall_long_sb'), col('stall_barrier'), col('stall_membar'),
                                 col('stall_lg'), col('stall_math'), col('stall_short_sb'), col('stall_mio'))
    tot = sum(v[0] for v in agg.values())
    print('total samples', tot)
    print('file line samples share | long_sb barrier membar lg math short_sb mio | source')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print('%-16s %4d %6d %5.1f%% | %5d %5d %5d %5d %5d %5d %5d | %s' % (k[0], k[1], v[0], 100 * v[0] / max(tot, 1), *v[3:], v[2]))

if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
