"""Developer tool: side-by-side table of tools/timeline.py outputs.  usage: tl_table.py dir name [name ...]"""
import sys
d, names = sys.argv[1], sys.argv[2:]
T = {}
for n in names:
    t = {}
    for l in open('%s/timeline_%s.txt' % (d, n)):
        p = l.split()
        if len(p) == 4 and p[0] not in ('kernel', 'step'):
            t[p[0]] = (float(p[1]), float(p[2]))
    T[n] = t
ks = sorted(T[names[0]], key=lambda k: T[names[0]][k][0])
print('%-12s' % 'kernel' + ''.join('%14s' % n[:13] for n in names))
for k in ks:
    print('%-12s' % k + ''.join(('%7.1f%7.1f' % T[n][k]) if k in T[n] else ' ' * 14 for n in names))
print('%-12s' % 'span' + ''.join('%14.1f' % max(v[1] for v in T[n].values()) for n in names))
