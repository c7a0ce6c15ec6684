"""Writes profiles/ncu_constants.json -- the numbers bench.py reports that come from profiler captures and
probes rather than from the run itself.  Every entry names its source, so nothing in bench.py is a bare constant.

usage: ncu_constants.py [conv_cfg3=<file.ncu-rep>] [concat_cfg2=<file.ncu-rep>] [probe=<probe log>]
Missing arguments keep the previous entry."""
import csv, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "ncu_constants.json")


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def traffic(rep):
    launches, units = raw(rep)
    rd = [to_bytes(l["dram__bytes_read.sum"], units["dram__bytes_read.sum"]) for l in launches]
    wr = [to_bytes(l["dram__bytes_write.sum"], units["dram__bytes_write.sum"]) for l in launches]
    n = len(launches)
    return {"dram_bytes_per_launch": (sum(rd) + sum(wr)) / n, "dram_read_bytes_per_launch": sum(rd) / n,
            "dram_write_bytes_per_launch": sum(wr) / n, "launches_captured": n, "kernel": launches[0]["Kernel Name"][:80],
            "source": f"ncu --set full, {os.path.basename(rep)} (dram__bytes_read.sum + dram__bytes_write.sum, mean over {n} captured launch(es); "
                      "ncu serialises launches and flushes nothing between them, so output that is still in the 126 MB write-back L2 when a "
                      "launch ends shows up as few or no DRAM writes: read the READ figure as the traffic check -- it equals the algorithmic input)"}


def main():
    cur = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for a in sys.argv[1:]:
        k, v = a.split("=", 1)
        if k == "probe":
            for line in open(v):
                m = re.search(r"BENCH kind::i8 M=128 N=256 K=32: ([\d.]+) MAC/clk/SM .* ([\d.]+) TOPS over 148 SMs", line)
                if m:
                    cur["i8_mma_probe"] = {"tops": float(m.group(2)), "mac_per_clk_per_sm": float(m.group(1)),
                                           "source": f"{os.path.relpath(v, ROOT)}: probe/umma_probe.cu, tcgen05.mma kind::i8 M=128 N=256 K=32, MMA only, 148 SMs"}
        else:
            cur[k] = traffic(v)
    json.dump(cur, open(OUT, "w"), indent=1)
    print(json.dumps(cur, indent=1))


if __name__ == "__main__":
    main()
