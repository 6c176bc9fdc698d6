"""Print one step of a tools/timeline.py recording: start, end, duration, stream, kernel.   python tools/timeline_view.py <json> [min_us] [t_lo t_hi]"""
import json, re, sys
d = json.load(open(sys.argv[1]))
min_us = float(sys.argv[2]) if len(sys.argv) > 2 else 15.0
starts = [i for i, e in enumerate(d) if 'k_pack_weights_batch' in e[0]]
s, e = starts[1], starts[2]
t0 = d[s][2]
lo = float(sys.argv[3]) if len(sys.argv) > 3 else -1
hi = float(sys.argv[4]) if len(sys.argv) > 4 else 1e18
def short(n):
    n = n.replace('void ', '').replace('l3::', '')
    m = re.match(r'([\w]+)(<[^>]*>)?', n)
    return (m.group(1) + (m.group(2) or '')).replace('__nv_bfloat16', 'bf')[:40]
for ev in d[s:e]:
    if ev[3] < min_us or not (lo <= ev[2] - t0 <= hi): continue
    print("%8.1f %8.1f %7.1f s%-3d %s" % (ev[2] - t0, ev[2] - t0 + ev[3], ev[3], ev[1], short(ev[0])))
print("step span us:", d[e][2] - t0)
