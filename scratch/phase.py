import sys, re, collections
ev = []
for l in open(sys.argv[1]):
    if l.startswith("PH "):
        _, h, lab, t0, t1 = l.split()
        ev.append((h, lab, float(t0), float(t1)))
calls = [e for e in ev if e[1] == "call"]
T0 = min(e[2] for e in calls); T1 = max(e[3] for e in calls)
# restrict to the last 60 % of the run (timed steps)
lo = T0 + float(sys.argv[2]) * (T1 - T0) if len(sys.argv) > 2 else T0 + 0.4 * (T1 - T0)
gpu = sorted((max(e[2], lo), e[3]) for e in ev if e[1] in ("scan", "select", "chain", "dp", "fetch") and e[3] > lo)
cov = 0.0; cur0, cur1 = None, None
for a, b in gpu:
    if cur1 is None or a > cur1:
        if cur1 is not None: cov += cur1 - cur0
        cur0, cur1 = a, b
    else: cur1 = max(cur1, b)
cov += cur1 - cur0
print("window", T1 - lo, "s; union of GPU-pending phases", cov, "=", cov / (T1 - lo))
tot = collections.defaultdict(float)
for h, lab, a, b in ev:
    if b > lo: tot[lab] += b - max(a, lo)
print({k: round(v, 3) for k, v in tot.items()})
