"""Summarise an .ncu-rep (raw page + source page): key metrics, stall mix, top stalled instructions.
usage: python tools/ncu_summary.py REPORT.ncu-rep [> profiles/xxx.txt]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit", "sm__cycles_elapsed.avg.per_second", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum",
        "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "smsp__average_warps_issue_stalled", "smsp__pcsamp_sample_buffer", "l1tex__throughput", "sm__throughput.avg.pct"]
for k in range(2, len(rows)):
    vals = rows[k]
    print("=" * 100)
    for h, u, v in zip(hdr, units, vals):
        if any(h.startswith(x) or x == h for x in KEYS):
            if v not in ("0", "", "0.000000"):
                print(f"{h:95s} {u:12s} {v}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot, byop, st, data = 0, collections.Counter(), collections.Counter(), []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[ix["# Samples"]])
    except ValueError:
        continue
    srcl = r[ix["Source"]]
    tok = srcl.split()
    op = tok[1] if tok and tok[0].startswith("@") and len(tok) > 1 else (tok[0] if tok else "")
    byop[op.split(".")[0]] += n
    tot += n
    for s in stalls:
        try:
            st[s] += int(r[ix[s]])
        except ValueError:
            pass
    data.append((n, srcl, r))
print("=" * 100)
print("warp-state samples:", tot)
print("by opcode:", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in byop.most_common(12)))
print("by stall :", ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in st.most_common(10)))
print("top instructions:")
for n, srcl, r in sorted(data, key=lambda x: -x[0])[:14]:
    ss = {s[6:]: int(r[ix[s]]) for s in stalls if r[ix[s]] not in ("", "0")}
    print(f"  {100 * n / tot:5.1f}%  {srcl[:64]:64s} {dict(sorted(ss.items(), key=lambda x: -x[1])[:2])}")
