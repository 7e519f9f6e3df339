"""Summarise an ncu report's source page per SASS basic block: share of issued instructions, average active threads, stall samples.
Usage: python scripts/ncu_blocks.py report.ncu-rep [kernel-regex] [min-share-%]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    kre = sys.argv[2] if len(sys.argv) > 2 else "k_trace"
    min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, out, k = None, [], 0
    for r in rows:
        if r and r[0] == "Kernel Name":
            k += 1
            if k == 1:
                print(r[1][:100])
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if k == 1 and hdr and len(r) == len(hdr):
            out.append(dict(zip(hdr, r)))
    tot_inst = sum(int(o["Instructions Executed"]) for o in out)
    tot_thr = sum(int(o["Thread Instructions Executed"]) for o in out)
    tot_s = max(1, sum(int(o["# Samples"]) for o in out))
    print("SASS lines %d, warp-instructions %d, thread-instructions %d, avg active threads %.2f, samples %d" %
          (len(out), tot_inst, tot_thr, tot_thr/max(1, tot_inst), tot_s))
    stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    tot_st = {c: sum(int(o[c]) for o in out) for c in stall_cols}
    print("stall mix: " + ", ".join("%s %.1f%%" % (c[6:], 100.0*v/tot_s) for c, v in sorted(tot_st.items(), key=lambda x: -x[1])[:8]))
    blocks, cur = [], None
    for i, o in enumerate(out):
        ie = int(o["Instructions Executed"])
        if cur is None or ie != cur["ie"]:
            cur = {"start": i, "ie": ie, "n": 0, "thr": 0, "s": 0, "ops": {}, "st": {}}
            blocks.append(cur)
        cur["n"] += 1
        cur["thr"] += int(o["Thread Instructions Executed"])
        cur["s"] += int(o["# Samples"])
        for c in stall_cols:
            cur["st"][c] = cur["st"].get(c, 0) + int(o[c])
        toks = o["Source"].strip().split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = op.split(".")[0]
        cur["ops"][op] = cur["ops"].get(op, 0) + 1
    for b in blocks:
        share = b["ie"]*b["n"]/max(1, tot_inst)*100
        if share < min_share and b["s"]/tot_s*100 < min_share:
            continue
        ops = sorted(b["ops"].items(), key=lambda x: -x[1])[:6]
        st = sorted(b["st"].items(), key=lambda x: -x[1])[:3]
        print("@%4d n=%3d exec=%9d share=%5.1f%% thr=%5.1f samp=%4.1f%% %s | %s" %
              (b["start"], b["n"], b["ie"], share, b["thr"]/max(1, b["ie"]*b["n"]), b["s"]/tot_s*100,
               " ".join("%s:%d" % x for x in ops), " ".join("%s:%d" % (c[6:], v) for c, v in st)))


if __name__ == "__main__":
    main()
