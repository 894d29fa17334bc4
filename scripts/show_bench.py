"""Readable summary of a bench.py JSON line (primary + `also` entries)."""
import json, sys
b = json.load(open(sys.argv[1]))
def brief(d):
    if "error" in d:
        print(f"  {d.get('workload')}: ERROR {d['error']}"); return
    r = d.get("roofline") or {}
    print(f"== {d.get('workload') or 'primary'}: {d['value']:.4g} {d['unit']}  ms/step {d['ms_per_step']:.3f}  steps {d['steps']}  prec {d.get('precision')}  "
          f"launches {d.get('gpu_launches')}  clocks {d['clocks']['sm_mhz']}/{d['clocks']['sm_mhz_min']} {d['clocks']['reasons']} ({d['clocks']['samples']} samples)")
    print(f"   roofline: {r.get('bound')} achieved {r.get('achieved')} peak {r.get('peak')} frac {r.get('frac')} executed_frac {r.get('executed_frac')}")
    if r.get("sustained"):
        q = r["sustained"]; print(f"   sustained: {q['value']:.4g} frac {q['frac']:.3f} over {q['seconds']:.2f}s clocks {q['clocks']['sm_mhz']} {q['clocks']['reasons']}")
    if r.get("hbm"): print(f"   hbm {r['hbm']['frac']:.3f} tensor {r['tensor']['frac']:.3f}")
    print(f"   parity: {d.get('parity')}")
    if d.get("gather"):
        g = d["gather"]; print(f"   gather: {g['how'][:60]} | with {g['ms_per_step_with_gather']:.3f} ms, none {g['ms_per_step_no_gather']:.3f} ms, nccl-after {g.get('nccl_after_kernel', {}).get('ms_per_step')} | recv {g['receive_gbs_per_rank']:.0f} GB/s/rank")
    e = d.get("e2e")
    if e:
        print(f"   e2e: {e['value']:.4g} {e['unit']} ({e['api'][:70]})")
        for k, v in e.items():
            if isinstance(v, dict): print(f"      {k}: {v['value']:.4g}")
    if d.get("abs_features"): print(f"   abs: {d['abs_features']['value']:.4g}")
    if d.get("frames_per_sec"): print(f"   frames/s: {d['frames_per_sec']:.1f}")
    if d.get("cpu_baseline"): print(f"   cpu: {d['cpu_baseline'].get('value')} ({d['cpu_baseline'].get('cores')} cores)")
brief(b)
for o in b.get("also", []): brief(o)
