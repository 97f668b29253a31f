import sys,json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception:
        print(l.rstrip()); continue
    print(r["layer"], {k:(round(v,1) if k.endswith("_us") else round(v,4)) for k,v in r.items() if k.endswith("_us") or k.endswith("relerr") or k.endswith("frac")})
