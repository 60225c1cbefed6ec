"""Turns two ncu launch lists (--metrics smsp__inst_executed.sum,smsp__inst_executed_pipe_alu.sum,dram__bytes_read.sum,
dram__bytes_write.sum,gpu__time_duration.sum --csv) of tools/ncu_step_target.py into profiles/r2_instr_counts.json,
the per-step constants bench.py's roofline uses.  usage: ncu_counts.py join.csv exhaustive.csv bound.csv out.json"""
import csv
import json
import sys


def load(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 12 and r[0] != "ID"]
    per = {}
    for r in rows:
        lid, kernel, metric, val = r[0], r[4], r[12], float(r[14].replace(",", ""))
        unit = r[13]
        if unit == "Mbyte": val *= 1e6
        elif unit == "Kbyte": val *= 1e3
        elif unit == "Gbyte": val *= 1e9
        elif unit == "us": val *= 1e3
        elif unit == "ms": val *= 1e6
        elif unit == "s": val *= 1e9
        per.setdefault(lid, {"kernel": kernel})[metric] = val
    return list(per.values())


def summarise(launches, main_prefix):
    inst = sum(l.get("smsp__inst_executed.sum", 0) for l in launches)
    alu = sum(l.get("smsp__inst_executed_pipe_alu.sum", 0) for l in launches)
    main = [l for l in launches if main_prefix in l["kernel"]]
    dram = sum(l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0) for l in main) / max(len(main), 1)
    t_all = sum(l.get("gpu__time_duration.sum", 0) for l in launches)
    t_main = sum(l.get("gpu__time_duration.sum", 0) for l in main)
    return inst, alu / max(inst, 1), dram, len(launches), len(main), t_main / max(t_all, 1)


j = load(sys.argv[1]); x = load(sys.argv[2]); d = load(sys.argv[3])
ji, ja, jd, jn, jm, js = summarise(j, "spr_join_score_kernel")
xi, xa, xd, xn, xm, xs = summarise(x, "spr_score_lattice_kernel")
di, da, dd, dn, dm, ds = summarise(d, "spr_bound_lattice_kernel")
out = {"source": "profiles/r2_step_join_launches.csv, profiles/r2_step_exhaustive_launches.csv, profiles/r2_step_bound_launches.csv "
                 "(ncu launch lists of tools/ncu_step_target.py, one config-2 search step each)",
       "join_c2_winst_per_step": ji, "join_c2_alu_share": ja, "join_c2_dram_bytes_per_launch": jd,
       "join_c2_launches": jn, "join_c2_main_kernel_launches": jm, "join_c2_main_kernel_time_share": js,
       "exhaustive_c2_winst_per_step": xi, "exhaustive_c2_alu_share": xa, "exhaustive_c2_dram_bytes_per_launch": xd,
       "exhaustive_c2_launches": xn, "exhaustive_c2_main_kernel_launches": xm, "exhaustive_c2_main_kernel_time_share": xs,
       "search_c2_winst_per_step": di, "search_c2_alu_share": da, "search_c2_bound_dram_bytes_per_launch": dd,
       "search_c2_launches": dn, "search_c2_bound_kernel_launches": dm, "search_c2_bound_kernel_time_share": ds}
json.dump(out, open(sys.argv[4], "w"), indent=1)
print(json.dumps(out, indent=1))
