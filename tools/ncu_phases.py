"""Summarise an `ncu --page source --csv` dump by phase: consecutive SASS
instructions are grouped into segments delimited by barriers, and each
segment's stall samples are totalled.  Scratch tool for reading profiles here
(no GPU needed):  ncu -i X.ncu-rep --page source --csv > s.csv; python tools/ncu_phases.py s.csv"""
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    seg, segs = {"ops": {}, "samples": 0, "stalls": {}, "first": None}, []
    total = 0
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        sass = r[col["Source"]].strip()
        toks = sass.split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "")
        base = op.split(".")[0]
        if base in ("DFMA", "DADD", "DMUL"):
            base = "FP64"
        n = int(r[col["# Samples"]] or 0)
        total += n
        if seg["first"] is None:
            seg["first"] = r[col["Address"]][-5:]
        seg["ops"][base] = seg["ops"].get(base, 0) + 1
        seg["samples"] += n
        for s in stall_cols:
            v = int(r[col[s]] or 0)
            if v:
                seg["stalls"][s[6:]] = seg["stalls"].get(s[6:], 0) + v
        if base == "BAR" or base == "EXIT":
            seg["end"] = sass[:40]
            segs.append(seg)
            seg = {"ops": {}, "samples": 0, "stalls": {}, "first": None}
    segs.append(seg)
    print("total samples", total)
    for s in segs:
        if s["samples"] < total * 0.002:
            continue
        ops = " ".join("%s:%d" % kv for kv in sorted(s["ops"].items(), key=lambda kv: -kv[1])[:5])
        st = " ".join("%s:%.1f%%" % (k, 100.0 * v / total) for k, v in
                      sorted(s["stalls"].items(), key=lambda kv: -kv[1])[:6])
        print("%s %5.1f%% | %s | %s | -> %s" % (s["first"], 100.0 * s["samples"] / total, ops, st,
                                              s.get("end", "")))


if __name__ == "__main__":
    main(sys.argv[1])
