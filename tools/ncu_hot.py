"""Top stall locations of one kernel from an ncu report's source page (SASS level):
    python tools/ncu_hot.py report.ncu-rep regex:kernel [N]"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    skip = sys.argv[4] if len(sys.argv) > 4 else "0"
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", kern, "-s", skip, "-c", "1"],
                         capture_output=True, text=True).stdout
    lines = out.splitlines()
    print(lines[0][:160])
    rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    data = []
    for r in rows[1:]:
        try:
            s = int(r[col["# Samples"]])
        except Exception:
            continue
        data.append((s, r))
    total = sum(s for s, _ in data)
    print("total samples", total)
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for s, r in sorted(data, key=lambda t: -t[0])[:n]:
        stalls = sorted(((int(r[col[h]] or 0), h) for h in stall_cols), reverse=True)[:2]
        print(f"{100.0 * s / max(total, 1):5.1f}%  {r[col['Address']][-5:]}  {r[col['Source']][:90]:90s}  "
              + ", ".join(f"{h[6:]}={v}" for v, h in stalls if v))


if __name__ == "__main__":
    main()
