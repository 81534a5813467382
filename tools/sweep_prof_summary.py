import sys, collections
tag = sys.argv[2] if len(sys.argv) > 2 else 'PF'   # PF = whole launch, PW = first task of every CTA only
acc = collections.defaultdict(lambda: [0]*13); n = collections.Counter()
perwarp = collections.defaultdict(lambda: [0]*13)
for line in open(sys.argv[1]):
    if not line.startswith(tag + ' '): continue
    p = line.split(); sec, cta, warp = int(p[1]), int(p[2]), int(p[3]); v = list(map(int, p[4:17]))
    for i in range(13): acc[sec][i] += v[i]; perwarp[(sec, warp)][i] += v[i]
    n[sec] += 1
for sec in sorted(acc):
    t = acc[sec]; tot = t[0] + t[4]
    print(f"pass {sec+1}: warps*launches {n[sec]}  in-task {100*t[0]/tot:.1f}%  between-tasks {100*t[4]/tot:.1f}% | of in-task: wait-left {100*t[1]/t[0]:.1f}%  wait-right {100*t[2]/t[0]:.1f}%  wait-global {100*t[3]/t[0]:.1f}%  cycles/step {t[0]/max(t[5],1):.0f}")
    for w in range(8):
        u = perwarp[(sec, w)]
        print(f"   warp {w}: in-task {u[0]/1e6:8.1f}M  left {100*u[1]/max(u[0],1):5.1f}%  right {100*u[2]/max(u[0],1):5.1f}%  global {100*u[3]/max(u[0],1):5.1f}%  between {u[4]/1e6:8.1f}M cyc/step {u[0]/max(u[5],1):.0f}  phaseA/step {u[6]/max(u[5],1):.0f}  phaseB/step {u[7]/max(u[5],1):.0f}  blocking {100*u[8]/max(u[5],1):.1f}%  stage-miss {100*u[9]/max(u[5],1):.1f}%  A1 {u[10]/max(u[5],1):.0f} A3 {u[11]/max(u[5],1):.0f} A5 {u[12]/max(u[5],1):.0f}")
