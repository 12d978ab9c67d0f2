"""Render the TR lines printed by a -DV2W_TRACE build (tools/trace_fused.py) as one timeline per kernel."""
import sys

def main(path):
    lines = [l.split() for l in open(path) if l.startswith('TR ')]
    kern, cur, seen = [], {'mma': [], 'epi': []}, {'mma': -1, 'epi': -1}
    for _, name, q, tag, t in lines:
        q = int(q)
        if q <= seen[name] and name == 'mma' and cur['mma']:
            kern.append(cur); cur = {'mma': [], 'epi': []}; seen = {'mma': -1, 'epi': -1}
        seen[name] = q
        cur[name].append((int(tag), int(t)))
    kern.append(cur)
    for ki, k in enumerate(kern):
        print('=== kernel', ki)
        ev = [('M', tag, t) for tag, t in k['mma']] + [('E', tag, t) for tag, t in k['epi']]
        ev.sort(key=lambda x: x[2])
        t0 = ev[0][2] if ev else 0
        for who, tag, t in ev:
            tile, sub = tag // 100, tag % 100
            if who == 'M':
                print(f"{t - t0:8d} M tile{tile} seg{sub // 2} {'issued' if sub % 2 else 'start'}")
            elif sub >= 50:
                print(f"{t - t0:8d}      E tile{tile} E2 {['begin', 't2full', 'end'][sub - 50]}")
            else:
                print(f"{t - t0:8d}      E tile{tile} E1_{sub // 10} {['begin', 't1full', 'a1empty', 'end'][sub % 10]}")

if __name__ == '__main__':
    main(sys.argv[1])
