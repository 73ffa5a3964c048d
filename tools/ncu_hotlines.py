"""Per-source-line instruction / stall-sample shares of one kernel from an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys
def main(path, kregex, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kregex],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname = None; hdr = None; lines = []
    for r in rows:
        if not r: continue
        if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
        if r[0] == "Function Name": continue
        if r[0] == "Line No": hdr = r; continue
        if hdr and r[0].isdigit():
            ie = hdr.index("Instructions Executed"); ss = hdr.index("# Samples")
            try: lines.append((int(r[ie]), int(r[ss]), fname, int(r[0]), r[1]))
            except ValueError: pass
    tot = sum(l[0] for l in lines) or 1; tots = sum(l[1] for l in lines) or 1
    print(f"total warp-instructions {tot}, samples {tots}")
    for ins, smp, f, ln, src in sorted(lines, reverse=True)[:top]:
        print(f"{100*ins/tot:5.1f}% inst {100*smp/tots:5.1f}% smp  {f}:{ln:<4d} {src.strip()[:100]}")
if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
