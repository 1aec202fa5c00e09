import sys, json
for ln in sys.stdin:
    try: r = json.loads(ln)
    except Exception:
        print(ln.rstrip()); continue
    print(r["case"], "fast ms", r["fast"]["kernel_ms"], "Mrays/s", r["fast"]["Mrays_s"], "gt1", r["fast"]["gt1"], "neq", r["fast"]["neq"], "max", r["fast"]["maxdiff"], "| strict ms", r["strict"]["kernel_ms"], "neq", r["strict"]["neq"])
