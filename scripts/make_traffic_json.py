"""profiles/roofline_traffic.json <- ncu DRAM byte counters (dram__bytes_read.sum + dram__bytes_write.sum):
  gemm   : mean bytes per gemm_epi_kernel launch over one training micro-batch (the bench's dominant kernel)
  search : bytes of one ip_topk_kernel launch of all 190 k queries against N candidate rows, keyed by N
           (profiles/dram_search_r02_N<rows>.csv: the 1 / 2 / 4 / 8-GPU shard sizes of the 5.6 M pool)"""
import csv, json, os, re, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def per_kernel(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: collections.defaultdict(float))
    cnt = collections.Counter()
    for r in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        m = r["Metric Name"]
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"].lower()
        if "byte" in u:
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}[u]
        agg[name][m] += v
        if m == "dram__bytes_read.sum":
            cnt[name] += 1
    return agg, cnt


out = {}
tr = os.path.join(ROOT, "profiles", "dram_train_step_r01.csv")
if os.path.exists(tr):
    agg, cnt = per_kernel(tr)
    b = sum(v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"] for k, v in agg.items() if "gemm_epi_kernel" in k)
    n = sum(c for k, c in cnt.items() if "gemm_epi_kernel" in k)
    out["gemm"] = b / n
    out["gemm_launches_profiled"] = n
out["search"] = {}
for f in sorted(os.listdir(os.path.join(ROOT, "profiles"))):
    m = re.fullmatch(r"dram_search_r02_N(\d+)\.csv", f)
    if not m:
        continue
    agg, cnt = per_kernel(os.path.join(ROOT, "profiles", f))
    for k, v in agg.items():
        if "ip_topk" in k:
            out["search"][m.group(1)] = (v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]) / max(1, cnt[k])
json.dump(out, open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w"), indent=1)
print(out)
