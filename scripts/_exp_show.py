import json, subprocess
print(subprocess.run("python scripts/ffn_timeline_print.py gpurun_out/ffn_timeline.txt 2 | awk '$1<100 || $1>=507' | tr '\\n' ';'; python scripts/ffn_timeline_print.py gpurun_out/ffn_timeline.txt 1 | awk '$1<10' | tr '\\n' ';'", shell=True, capture_output=True, text=True, cwd="/root/repo").stdout)
for w in ("random", "copy"):
    d = json.loads(open(f"/root/repo/gpurun_out/e_{w}.json").read().strip().splitlines()[-1])
    print(w, round(d["value"], 1), round(d["one_batch_in_flight"]["value"], 1), round(d["roofline"]["frac"], 3), d["roofline"]["in_situ"]["avg_added_us"], d["roofline"]["in_situ"]["frac"])
    if w == "random":
        for k, v in d.get("kernels_in_situ", {}).items(): print("   ", k, v)
