"""Top source lines of an `ncu --page source --print-source cuda --csv` export, per kernel.
usage: ncu_source_top.py <csv> [top_n]"""
import csv
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path, errors="replace")))
    kernel = "?"
    hdr = None
    acc = []

    def flush():
        if not acc or hdr is None:
            return
        cols = {h: i for i, h in enumerate(hdr)}
        samp = next((cols[h] for h in hdr if h.startswith("Warp Stall Sampling (All")), None)
        inst = next((cols[h] for h in hdr if h.startswith("Instructions Executed")), None)
        src = cols.get("Source", 1)
        line = cols.get("#", 0)
        def num(r, i):
            try:
                return float(r[i].replace(",", "")) if i is not None and i < len(r) else 0.0
            except ValueError:
                return 0.0
        tot_s = sum(num(r, samp) for r in acc) or 1.0
        tot_i = sum(num(r, inst) for r in acc) or 1.0
        print(f"== {kernel}  (samples {tot_s:.0f}, warp-instructions {tot_i:.0f})")
        for r in sorted(acc, key=lambda r: -num(r, samp))[:top]:
            print(f"  L{r[line]:>5} samp {100 * num(r, samp) / tot_s:5.1f}%  inst {100 * num(r, inst) / tot_i:5.1f}%  | {r[src].strip()[:110]}")

    for r in rows:
        if not r:
            continue
        if len(r) == 1 or (r[0].startswith("Kernel") or "Kernel Name" in r[0]):
            flush(); acc = []; hdr = None
            kernel = ",".join(r)[:120]
            continue
        if r[0] == "#" or (hdr is None and "Source" in r):
            flush(); acc = []
            hdr = r
            continue
        if hdr is not None and len(r) >= len(hdr) - 2:
            acc.append(r)
    flush()


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
