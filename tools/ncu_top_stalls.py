"""Top-N SASS instructions by warp-stall samples from an `ncu --set full --import-source on` report (source page).

    python tools/ncu_top_stalls.py gpurun_out/prof_decoder_X.ncu-rep [N] [launch-index]"""
import csv
import io
import subprocess
import sys


def main():
    path = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for line in out.splitlines():
        if line.startswith('"Kernel Name"'):
            cur = {"name": next(csv.reader([line]))[1], "lines": []}
            blocks.append(cur)
        elif cur is not None:
            cur["lines"].append(line)
    idx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    b = blocks[idx]
    rows = list(csv.DictReader(io.StringIO("\n".join(b["lines"]))))
    key = "Warp Stall Sampling (All Samples)"
    tot = sum(int(r[key] or 0) for r in rows)
    print(f"# {path}: {b['name']}")
    print(f"# {len(rows)} SASS instructions, {tot} warp-stall samples; top {n} instructions by samples")
    rows.sort(key=lambda r: -int(r[key] or 0))
    for r in rows[:n]:
        stalls = {k[len("stall_"):]: int(v) for k, v in r.items() if k.startswith("stall_") and v and int(v) > 0}
        top = ", ".join(f"{k}={v}" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:3])
        print(f"{100.0 * int(r[key]) / max(tot, 1):5.1f}%  {r['Source'].strip():60s} {top}")


if __name__ == "__main__":
    main()
