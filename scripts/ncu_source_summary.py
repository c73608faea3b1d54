"""Summarise an `ncu --page source --print-source cuda,sass --csv` export: instructions and stall samples per source line."""
import csv
import sys


def summarise(path, top=25):
    rows = list(csv.reader(open(path)))
    hdr = rows[2]
    i_line, i_src, i_inst, i_samp = 0, 1, hdr.index("Instructions Executed"), hdr.index("# Samples")
    out = []
    for r in rows[3:]:
        if r[i_line] and r[i_line].isdigit():
            try:
                out.append((int(r[i_inst]), int(r[i_samp]), int(r[i_line]), r[i_src].strip()))
            except ValueError:
                pass
    tot_i = sum(o[0] for o in out) or 1
    tot_s = sum(o[1] for o in out) or 1
    print(rows[1][1])
    print(f"total warp-instructions {tot_i}, samples {tot_s}")
    for inst, samp, line, src in sorted(out, reverse=True)[:top]:
        print(f"{100*inst/tot_i:5.1f}% inst {100*samp/tot_s:5.1f}% samp  L{line:<4} {src[:110]}")


if __name__ == "__main__":
    summarise(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
