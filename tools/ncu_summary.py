"""Condense `ncu -i X.ncu-rep --page raw --csv` into one line of key metrics per kernel launch.
    ncu -i rep.ncu-rep --page raw --csv | python tools/ncu_summary.py [--csv out.csv]"""
import csv
import sys

KEYS = [
    ("us", "gpu__time_duration.sum"),
    ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("xu%", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("alu%", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
    ("fma%", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("dram_rd_MB", "dram__bytes_read.sum"),
    ("dram_wr_MB", "dram__bytes_write.sum"),
    ("dram%", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("local_st", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum"),
    ("smem_st_conf", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum"),
    ("regs", "launch__registers_per_thread"),
    ("warps_act%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rows = list(csv.reader(sys.stdin))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        d = {"kernel": name[:70]}
        for k, m in KEYS:
            if m in idx and r[idx[m]] not in ("", "n/a"):
                v = float(r[idx[m]].replace(",", ""))
                u = units[idx[m]]
                if k.endswith("_MB"):
                    v = v / 1e6 * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                if k == "us":
                    v = v * {"ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1)
                d[k] = round(v, 2)
        st = sorted(((float(r[i]), h[len(STALL):].replace("_per_issue_active.ratio", "")) for h, i in idx.items()
                     if h.startswith(STALL) and h.endswith("per_issue_active.ratio") and r[i] not in ("", "n/a")), reverse=True)
        d["stalls"] = " ".join(f"{n}={v:.2f}" for v, n in st[:5])
        out.append(d)
    cols = ["kernel"] + [k for k, _ in KEYS] + ["stalls"]
    w = csv.writer(sys.stdout)
    w.writerow(cols)
    for d in out:
        w.writerow([d.get(c, "") for c in cols])


if __name__ == "__main__":
    main()
