import csv, json, sys
tag = sys.argv[1]
for f in ("d1", "d3"):
    print(f, open(f"gpurun_out/{tag}_{f}.log").read().strip().splitlines()[-1])
print(open(f"gpurun_out/{tag}_tests.log").read().strip().splitlines()[-1])
try:
    d = json.loads(open(f"gpurun_out/{tag}_bench.json").read().strip().splitlines()[-1])
    print("bench %.3e agent-steps/s  ms/step %.4f  frac %.3f  step_ms %.4f  e2e %.3e  launches %d" % (
        d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["gpu_launches"]))
    print(d.get("episode_stats"))
except Exception as e:
    print("bench ERR", e, open(f"gpurun_out/{tag}_bench.err").read()[-1500:])
try:
    rows = list(csv.reader(open(f"gpurun_out/{tag}_launches.csv")))
    i = [k for k, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[i]
    cur = {}
    for r in rows[i + 1:]:
        d = dict(zip(hdr, r))
        cur.setdefault((d["ID"], d["Kernel Name"][:28]), {})[d["Metric Name"].split(".")[0]] = d["Metric Value"]
    for k, v in list(cur.items())[:6]:
        print(k, v)
except Exception as e:
    print("launches ERR", e)
