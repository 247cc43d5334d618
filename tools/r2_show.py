import json, sys
cur = None
for l in open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/r2/quick.log'):
    l = l.rstrip()
    if l.startswith('=='):
        cur = l[3:]
    elif l.startswith('{'):
        d = json.loads(l)
        print(f"{cur:55s} x_rt {d.get('x_realtime')}  ms {d.get('kernel_ms')}")
    else:
        print('   ', l[:300])
