#!/usr/bin/env python3
"""Regenerate profiles/traffic.json and profiles/trace_efficiency.json (the two tables bench.py quotes next to its live numbers)
from ONE `ncu --set full` capture of a steady-state frame.
usage: tools/profile_tables.py <frame.ncu-rep> <capture id, e.g. r3p> [--sass-steps N]

traffic.json           dram__bytes_read.sum + dram__bytes_write.sum per launch, keyed like bench.py's kernels[]
trace_efficiency.json  issue-slot / lane figures of the DDA engine and the other step-dominating kernels; for the branch-free
                       engine also the ALIVE lanes per warp-step, from the predicated-on thread count of its predicated FADDs
                       (`--page source`): a finished lane still executes the block, so threads/instruction overstates it."""
import csv, json, os, subprocess, sys

rep, cap = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def raw_rows():
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1, 'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9, 's': 1, 'usecond': 1e-6, 'msecond': 1e-3,
             'nsecond': 1e-9, 'second': 1}
    res = []
    for r in rows[2:]:
        def g(k, sc=False):
            try:
                v = float(r[ix[k]].replace(',', ''))
            except Exception:
                return float('nan')
            return v * scale.get(units[ix[k]], 1) if sc else v
        res.append({"name": r[ix['Kernel Name']].split('(')[0].replace('void ', '').replace('vpt::', ''),
                    "us": g('gpu__time_duration.sum', True) * 1e6,
                    "dram": g('dram__bytes_read.sum', True) + g('dram__bytes_write.sum', True),
                    "issue": g('smsp__issue_active.avg.pct_of_peak_sustained_active'),
                    "thr": g('smsp__thread_inst_executed_per_inst_executed.ratio'),
                    "alu": g('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'),
                    "inst": g('smsp__inst_executed.sum'),
                    "dram_pct": g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
                    "regs": g('launch__registers_per_thread')})
    return res


def alive_lanes(kernel_regex):
    """alive lanes per warp-step of every captured launch of the branch-free DDA engine."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv", "--kernel-name", "regex:" + kernel_regex],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
    res, seen = [], set()
    for hi in starts:
        hdr = rows[hi]
        ix = {h: i for i, h in enumerate(hdr)}
        data = []
        for r in rows[hi + 1:]:
            if len(r) != len(hdr) or r[0] == 'Address':
                break
            data.append(r)
        fadd = [r for r in data if 'FADD' in r[ix['Source']] and r[ix['Source']].strip().startswith('@')]
        inst = sum(int(r[ix['Instructions Executed']] or 0) for r in fadd) / 3.0
        on = sum(int(r[ix['Predicated-On Thread Instructions Executed']] or 0) for r in fadd)
        key = (len(data), inst, on)
        if inst == 0 or key in seen:  # ncu prints every launch twice in some versions
            continue
        seen.add(key)
        res.append({"warp_steps": inst, "alive_lanes_per_warp_step": round(on / inst, 2)})
    return res


K = raw_rows()
def pick(pred): return [k for k in K if pred(k["name"])]
def first(pred):
    p = pick(pred)
    return p[0] if p else None

traffic = {"_capture": "%s (profiles/%s_ncu_full_summary.txt)" % (cap, cap),
           "_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch, from the %s `ncu --set full` capture of one steady-state frame, 1920x1080, 4 spp" % cap}
names = [("prep_firefly_sky", lambda n: n.startswith("prepKernel")), ("temporal", lambda n: n.startswith("temporalKernel")),
         ("history_clamp", lambda n: n.startswith("historyClamp")), ("atrous_smem", lambda n: n.startswith("atrousFirst")),
         ("atrous", lambda n: "atrousTileKernel<2" in n or n.startswith("atrousKernel")), ("atrous_last", lambda n: "atrousTileKernel<8" in n),
         ("gen", lambda n: n.startswith("genKernel")), ("shade1", lambda n: n.startswith("shade1")), ("shade2", lambda n: n.startswith("shade2")),
         ("shade3", lambda n: n.startswith("shade3")), ("shade4", lambda n: n.startswith("shade4")), ("shade5", lambda n: n.startswith("shade5"))]
for key, pred in names:
    k = first(pred)
    if k:
        traffic[key] = int(k["dram"])
closest = pick(lambda n: n.startswith("ddaKernel<1, 1") or n.startswith("ddaFlatKernel<1"))
anyhit = pick(lambda n: (n.startswith("ddaKernel") or n.startswith("ddaFlatKernel")) and not (n.startswith("ddaKernel<1, 1") or n.startswith("ddaFlatKernel<1")))
if closest: traffic["dda_closest"] = int(closest[0]["dram"])
if anyhit: traffic["dda_any"] = int(anyhit[0]["dram"])
trace = [k for k in K if any(k["name"].startswith(p) for p in ("genKernel", "ddaKernel", "ddaFlatKernel", "shade", "accumulateKernel"))]
den = [k for k in K if not any(k["name"].startswith(p) for p in ("genKernel", "ddaKernel", "ddaFlatKernel", "shade", "accumulateKernel"))]
traffic["trace_total_per_frame"] = int(sum(k["dram"] for k in trace))
traffic["denoiser_total_per_frame"] = int(sum(k["dram"] for k in den))
json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)

alive = alive_lanes("ddaFlatKernel")
eff = {"_comment": "ncu --set full, %s capture (profiles/%s_ncu_full_summary.txt). Peak = 4 warp-instructions/clk/SM; the alu pipe takes one warp "
                   "instruction per 2 clk per SM sub-partition. alive_lanes_per_warp_step: lanes whose ray is still walking (a finished lane of the "
                   "branch-free engine executes the block predicated off, so threads/instruction overstates it)" % (cap, cap),
       "dda_closest": None if not closest else {"us": round(closest[0]["us"], 1), "issue_slot_utilisation_pct": round(closest[0]["issue"], 1),
                                                "alu_pipe_pct": round(closest[0]["alu"], 1), "threads_per_instruction": round(closest[0]["thr"], 1),
                                                "registers": int(closest[0]["regs"]), "warp_instructions": int(closest[0]["inst"])},
       "dda_any": {"us": [round(k["us"], 1) for k in anyhit], "issue_slot_utilisation_pct": [round(k["issue"], 1) for k in anyhit],
                   "alu_pipe_pct": [round(k["alu"], 1) for k in anyhit], "threads_per_instruction": [round(k["thr"], 1) for k in anyhit],
                   "warp_instructions": [int(k["inst"]) for k in anyhit], "registers": int(anyhit[0]["regs"]) if anyhit else None,
                   "alive_lanes_per_warp_step": [a["alive_lanes_per_warp_step"] for a in alive],
                   "sass_instructions_per_step": 16}}
for key, pred in (("shade1", lambda n: n.startswith("shade1")), ("shade3", lambda n: n.startswith("shade3")), ("temporal", lambda n: n.startswith("temporalKernel"))):
    k = first(pred)
    if k:
        eff[key] = {"us": round(k["us"], 1), "dram_pct_of_peak": round(k["dram_pct"], 1), "issue_slot_utilisation_pct": round(k["issue"], 1),
                    "warp_instructions": int(k["inst"])}
at = pick(lambda n: n.startswith("atrousTileKernel"))
if at:
    eff["atrous_tile"] = {"us": [round(k["us"], 1) for k in at], "issue_slot_utilisation_pct": [round(k["issue"], 1) for k in at],
                          "warp_instructions": [int(k["inst"]) for k in at]}
json.dump(eff, open(os.path.join(ROOT, "profiles", "trace_efficiency.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))
print(json.dumps(eff, indent=1))
