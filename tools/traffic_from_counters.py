"""Fold ncu counter csv files (tools/gpu_final*.sh: ...cN_fp32_counters.csv) into profiles/traffic.json.
usage: python tools/traffic_from_counters.py cN=path.csv ..."""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tpath = os.path.join(ROOT, "profiles", "traffic.json")
t = json.load(open(tpath))
for arg in sys.argv[1:]:
    name, path = arg.split("=")
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    m = {r[12]: float(r[14]) for r in rows[1:]}
    g = lambda k: m[k]
    t[name] = int(g("dram__bytes_read.sum") + g("dram__bytes_write.sum"))
    t[name + "_executed_flop_per_launch"] = int(g("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum") + g("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum")
                                                + 2 * g("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum"))
    t[name + "_thread_instructions_per_launch"] = int(g("smsp__thread_inst_executed.sum"))
    t[name + "_launch_ms_under_ncu"] = g("gpu__time_duration.sum") / 1e6
    print(name, {k: v for k, v in t.items() if k.startswith(name) and k != name}, t[name])
json.dump(t, open(tpath, "w"), indent=1)
