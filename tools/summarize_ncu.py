"""Summarise an `ncu --set full` report (.ncu-rep) per launch: duration, tensor / XU / issue utilisation, L2 and
DRAM traffic.  Usage: python tools/summarize_ncu.py report.ncu-rep  (needs `ncu` on PATH to read the report)."""
import csv
import io
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "us", 1e-3), ("launch__grid_size", "grid", 1),
        ("launch__registers_per_thread", "regs", 1),
        ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "tc%", 1),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%", 1),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%", 1),
        ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue%", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%", 1),
        ("dram__bytes_read.sum", "dramR_MB", 1), ("dram__bytes_write.sum", "dramW_MB", 1),
        ("lts__t_sectors_srcunit_tex_op_read.sum", "l2rd_MB", 32e-6)]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {n: hdr.index(n) for n, _, _ in COLS if n in hdr}
    print(f"{'kernel':44s} " + " ".join(f"{lab:>9s}" for n, lab, _ in COLS if n in idx))
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        name = name.replace("void pfd::", "").replace("pfd::", "").split("(")[0]
        vals = []
        for n, lab, sc in COLS:
            if n not in idx:
                continue
            v = r[idx[n]].replace(",", "")
            u = units[idx[n]]
            try:
                f = float(v) * sc
                if lab == "us" and u in ("ns", "nsecond"):
                    f = float(v) * 1e-3
                elif lab == "us" and u in ("us", "usecond"):
                    f = float(v)
                if lab.startswith("dram") and u in ("Kbyte",):
                    f = float(v) * 1e-3
                elif lab.startswith("dram") and u in ("byte",):
                    f = float(v) * 1e-6
                vals.append(f"{f:9.2f}")
            except ValueError:
                vals.append(f"{v[:9]:>9s}")
        print(f"{name[:44]:44s} " + " ".join(vals))


if __name__ == "__main__":
    main(sys.argv[1])
