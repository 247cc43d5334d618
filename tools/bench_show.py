"""one-line summary of a bench.py JSON line (developer convenience)"""
import json
import sys

for path in sys.argv[1:]:
    try:
        j = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:   # noqa: BLE001
        print(path, "unreadable:", e)
        continue
    r, e = j.get("roofline") or {}, j.get("e2e") or {}
    print(f"{path}: value={j.get('value'):.1f} {j.get('unit')} ms={j.get('ms_per_step'):.2f} n_gpus={j.get('n_gpus')} "
          f"x_rt={(j.get('config') or {}).get('x_realtime')} frac={r.get('frac')} e2e={e.get('value')} "
          f"locked={j.get('locked_channels')} parity={j.get('parity_max_rel')} kernel_ms={r.get('kernel_ms_per_launch')}")
