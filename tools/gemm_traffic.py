"""Per-launch DRAM traffic of the backbone's gather-GEMM launches from an ncu CSV
(`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum -k regex:gather_gemm_tc`
over tools/profile_step.py): the first N launches of a forward step are the backbone convolutions."""
import csv
import json
import sys


def main():
    path, n_conv, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    lines = [l for l in open(path) if not l.startswith("==")]
    per = {}
    order = []
    for row in csv.DictReader(lines):
        i = int(row["ID"])
        if i not in per:
            per[i] = {}
            order.append(i)
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(unit, 1)
        per[i][row["Metric Name"]] = v * scale
    ids = order[:n_conv]
    rd = sum(per[i]["dram__bytes_read.sum"] for i in ids)
    wr = sum(per[i]["dram__bytes_write.sum"] for i in ids)
    l2 = sum(per[i].get("lts__t_bytes.sum", 0.0) for i in ids)
    t = sum(per[i]["gpu__time_duration.sum"] for i in ids)
    res = {"kernel": "gather_gemm_tc_kernel", "launches": len(ids), "dram_bytes_per_launch": (rd + wr) / len(ids),
           "dram_read_bytes": rd, "dram_write_bytes": wr, "l2_bytes": l2, "time_us_sum": t,
           "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum "
                     "-k regex:gather_gemm_tc over one scannet_b8 step (tools/profile_step.py); first %d launches = the "
                     "backbone convolutions" % len(ids)}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
