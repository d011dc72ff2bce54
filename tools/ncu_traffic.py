"""Extracts dram bytes (read + write) per launch from `ncu --set full` reports into profiles/ncu_traffic.json.

    python tools/ncu_traffic.py c3:bf16 speller.steps=gpurun_out/prof_decoder_X.ncu-rep listener.L1.input_gemm=...:LAUNCH_INDEX
bench.py reads the file to fill `roofline.traffic`."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def dram_bytes(path, index=0):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, r = rows[0], rows[1], rows[2 + index]
    tot = 0.0
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(key)
        tot += float(r[i].replace(",", "")) * UNIT[units[i]]
    return tot


def main():
    key = sys.argv[1]
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    d = json.load(open(p)) if os.path.exists(p) else {}
    e = d.setdefault(key, {})
    for a in sys.argv[2:]:
        name, path = a.split("=")
        idx = 0
        if ":" in path:
            path, idx = path.rsplit(":", 1)
        e[name] = dram_bytes(path, int(idx))
        print(name, e[name])
    json.dump(d, open(p, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
