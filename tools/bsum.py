"""Prints the figures of a bench.py JSON line that matter when comparing two builds."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as ex:                                   # noqa: BLE001
        print(path, "unreadable:", ex)
        continue
    st = d.get("stage_ms_per_step", {})
    print(f"{path}: value {d['value'] / 1e6:.3f} M/s ({d['ms_per_step'] * 1e3:.1f} us/step), one at a time {d.get('latency', {}).get('ms_per_step', 0) * 1e3:.1f} us,"
          f" e2e {d.get('e2e', {}).get('value', 0) / 1e6:.3f} M/s, stages us " + " ".join(f"{k}={v * 1e3:.1f}" for k, v in st.items())
          + f", roofline frac {d.get('roofline', {}).get('frac', 0):.3f}")
