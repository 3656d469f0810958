"""Turn an `ncu --set full` report of tools/syrk_lab.py once into profiles/r02_syrk_traffic.json + a metrics table.

    python tools/ncu_summarize.py gpurun_out/r2_prof_syrk_final.ncu-rep gpurun_out/r2_syrk_once_order.json
"""
import csv
import io
import json
import subprocess
import sys

rep, order_path = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
order = json.load(open(order_path))
assert len(order) == len(data), (len(order), len(data))
want = [("gpu__time_duration.sum", "duration"), ("sm__cycles_elapsed.avg.per_second", "SM clock"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % of elapsed"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> SM read"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "LSU shared wavefronts"),
        ("launch__registers_per_thread", "registers / thread"), ("launch__grid_size", "grid"), ("launch__cluster_size", "cluster"),
        ("smsp__cycles_active.avg", "SMSP active cycles")]


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def to_bytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)


shapes, lines = {}, []
lines.append("| K | tokens | launch | " + " | ".join(n for _, n in want) + " |")
lines.append("|---|---|---|" + "---|" * len(want))
seen = {}
for (K, T), r in zip(order, data):
    seen[(K, T)] = seen.get((K, T), 0) + 1
    cells = []
    for m, _ in want:
        cells.append((r[ix[m]] + " " + units[ix[m]]) if m in ix else "n/a")
    lines.append(f"| {K} | {T} | {seen[(K, T)]} | " + " | ".join(cells) + " |")
    if seen[(K, T)] == 2:
        rd = to_bytes(num(r[ix["dram__bytes_read.sum"]]), units[ix["dram__bytes_read.sum"]])
        wr = to_bytes(num(r[ix["dram__bytes_write.sum"]]), units[ix["dram__bytes_write.sum"]])
        nb = (K + 255) // 256
        shapes[f"K{K}_T{T}"] = {"dram_read_bytes": rd, "dram_write_bytes": wr,
                               "algorithmic_bytes": 2.0 * K * T + 2 * 4.0 * (nb * (nb + 1) // 2) * 65536,
                               "duration_ns": to_bytes(num(r[ix["gpu__time_duration.sum"]]), "byte") * {"ns": 1, "us": 1e3, "ms": 1e6}.get(units[ix["gpu__time_duration.sum"]], 1),
                               "tensor_pipe_active_pct": num(r[ix["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]])}
json.dump({"kernel": "gq::syrk_tc2_kernel<128,3>", "source": rep, "how": "ncu --set full --clock-control none, second launch of each shape",
           "launch_shapes": shapes}, open("profiles/r02_syrk_traffic.json", "w"), indent=1)
open("profiles/r02_syrk_tc2_ncu_full.md", "w").write(
    "# Round 2: ncu --set full of the Hessian kernel (gq::syrk_tc2_kernel<128,3>, tcgen05 cta_group::2) at every launch shape of the bench\n\n"
    "`tools/syrk_lab.py once` on one B200 (`--clock-control none`); two launches per shape.  Algorithmic bytes per launch = X slab "
    "(2 K tokens) + read-modify-write of the 256 x 256 fp32 tiles touching the lower triangle.\n\n" + "\n".join(lines) + "\n")
print("\n".join(lines))
print(json.dumps(shapes, indent=1))
