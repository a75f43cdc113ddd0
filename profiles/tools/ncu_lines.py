"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by source line.

usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > x.csv
       python profiles/tools/ncu_lines.py x.csv [top_n]
Prints executed warp instructions and stall samples per CUDA source line (file:line).
"""
import csv
import sys
from collections import defaultdict


def main(path, top=40):
    inst = defaultdict(int)
    samp = defaultdict(int)
    reasons = defaultdict(lambda: defaultdict(int))
    rcols = {}
    text = {}
    cur_file = None
    cur_line = None
    cols = None
    for row in csv.reader(open(path, newline="")):
        if not row:
            continue
        if row[0] == "File Path":
            cur_file = row[1].split("/")[-1]
            continue
        if row[0] == "Function Name":
            continue
        if row[0] == "Line No":
            cols = {name: i for i, name in enumerate(row)}
            i_inst = row.index("Instructions Executed")
            i_samp = row.index("# Samples")
            rcols = {name: i for i, name in enumerate(row) if name.startswith("stall_") and "Not Issued" not in name}
            continue
        if cols is None:
            continue
        if row[0] != "":
            cur_line = (cur_file, int(row[0]))
            text[cur_line] = row[1].strip()
            continue
        try:
            inst[cur_line] += int(row[i_inst])
            samp[cur_line] += int(row[i_samp])
            for name, i in rcols.items():
                v = int(row[i] or 0)
                if v:
                    reasons[name][cur_line] += v
        except (ValueError, IndexError):
            pass
    tot_i = sum(inst.values())
    tot_s = sum(samp.values())
    print(f"total warp instructions {tot_i}  samples {tot_s}")
    byfile_i = defaultdict(int)
    byfile_s = defaultdict(int)
    for k, v in inst.items():
        byfile_i[k[0]] += v
        byfile_s[k[0]] += samp[k]
    for f in sorted(byfile_i, key=byfile_i.get, reverse=True):
        print(f"  {f:24s} inst {byfile_i[f]:>12d} ({100*byfile_i[f]/max(tot_i,1):5.1f}%)  samples {byfile_s[f]:>7d} ({100*byfile_s[f]/max(tot_s,1):5.1f}%)")
    print("--- top lines by instructions")
    for k in sorted(inst, key=inst.get, reverse=True)[:top]:
        print(f"{k[0]}:{k[1]:<5d} inst {inst[k]:>10d} ({100*inst[k]/tot_i:4.1f}%) samp {samp[k]:>5d}  {text.get(k,'')[:90]}")
    print("--- top lines by samples")
    for k in sorted(samp, key=samp.get, reverse=True)[:top]:
        print(f"{k[0]}:{k[1]:<5d} samp {samp[k]:>6d} ({100*samp[k]/max(tot_s,1):4.1f}%) inst {inst[k]:>10d}  {text.get(k,'')[:90]}")


    if reasons:
        print("--- stall reasons (top lines each)")
        for name in sorted(reasons, key=lambda n: sum(reasons[n].values()), reverse=True)[:6]:
            d = reasons[name]
            print(f"{name}: {sum(d.values())}")
            for k in sorted(d, key=d.get, reverse=True)[:10]:
                print(f"    {k[0]}:{k[1]:<5d} {d[k]:>6d}  {text.get(k,'')[:90]}")
    return
    main.__dict__["reasons"] = reasons


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
