"""Per-source-line instruction / stall-sample shares of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).
usage: ncu_hotlines.py report.ncu-rep kernel-regex [top] [function-substring]
(the function substring picks one template instantiation when the regex matches several launches)"""
import csv, subprocess, sys
def load(path, kregex, fsub=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kregex],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname = None; hdr = None; lines = []; func_ok = True
    for r in rows:
        if not r: continue
        if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
        if r[0] == "Function Name": func_ok = (fsub is None) or (fsub in r[1]); continue
        if r[0] == "Line No": hdr = r; continue
        if hdr and r[0].isdigit() and func_ok:
            ie = hdr.index("Instructions Executed"); ss = hdr.index("# Samples")
            try: lines.append((int(r[ie]), int(r[ss]), fname, int(r[0]), r[1]))
            except ValueError: pass
    return lines
def main(path, kregex, top=40, fsub=None):
    lines = load(path, kregex, fsub)
    tot = sum(l[0] for l in lines) or 1; tots = sum(l[1] for l in lines) or 1
    print(f"total warp-instructions {tot}, samples {tots}")
    for ins, smp, f, ln, src in sorted(lines, reverse=True)[:top]:
        print(f"{100*ins/tot:5.1f}% inst {100*smp/tots:5.1f}% smp  {f}:{ln:<4d} {src.strip()[:100]}")
if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40, sys.argv[4] if len(sys.argv) > 4 else None)
