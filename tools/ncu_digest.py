#!/usr/bin/env python
"""Text digest of an `ncu --set full --import-source on` capture of nmpc_ipm_kernel (runs on the CPU box):
   python tools/ncu_digest.py gpurun_out/prof.ncu-rep mpc-implementation_b200/csrc/inst_15_3.o > profiles/<name>.txt
Launch / SM metrics, pipe utilisation (FP64 and tensor), DRAM bytes, stall reasons per issue, and -- joined with
nvdisasm of the instantiation's object -- executed instructions and stall samples per device function, opcode mix."""
import bisect, collections, csv, os, re, shutil, subprocess, sys, tempfile

rep, obj = sys.argv[1], sys.argv[2]
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h, units, v = raw[0], raw[1], raw[2]
col = {n: i for i, n in enumerate(h)}
get = lambda n: (v[col[n]], units[col[n]]) if n in col else ("n/a", "")
print(f"# {os.path.basename(rep)}: {v[col['Kernel Name']] if 'Kernel Name' in col else ''}  grid {get('launch__grid_size')[0]} x block {get('launch__block_size')[0]}")
WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__icc_request_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
print("\n## launch / SM metrics")
for n in WANT:
    a, u = get(n)
    print(f"{n:78s} {a} {u}")
for n in ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "sm__ops_path_tensor_src_fp64.sum",
          "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active"):
    a, u = get(n)
    print(f"{n:78s} {a} {u}")
print("\n## stall reasons (warps stalled per issue-active cycle)")
st = [(n.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), float(v[col[n]])) for n in h if "issue_stalled_" in n and n.endswith("per_issue_active.ratio") and v[col[n]] not in ("", "n/a")]
for n, x in sorted(st, key=lambda t: -t[1])[:10]:
    print(f"  {n:28s} {x:.3f}")

# ---- per-function breakdown
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", cub], cwd=d, capture_output=True, text=True).stdout
labels, cur, insec = [], None, False
for ln in dis.splitlines():
    m = re.match(r"^\s*\.section\s+(\S+)", ln)
    if m:
        insec = m.group(1).startswith(".text") and "nmpc_ipm_kernel" in m.group(1); continue
    if not insec:
        continue
    m = re.match(r"^([$_a-zA-Z][^ \t]*):$", ln)
    if m and not m.group(1).startswith(".L"):
        cur = m.group(1); continue
    m = re.match(r"^\s+/\*([0-9a-f]+)\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
    if m and cur:
        labels.append((int(m.group(1), 16), cur, m.group(3)))
shutil.rmtree(d, ignore_errors=True)
offs = [o for o, _, _ in labels]
src = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout.splitlines()))
hi = next(i for i, r in enumerate(src) if "Address" in r and "Source" in r)
sh = src[hi]; sc = {n: i for i, n in enumerate(sh)}
data = [r for r in src[hi + 1:] if len(r) == len(sh)]
addr = lambda a: int(a, 16) if a.startswith("0x") else int(a)
base = min(addr(r[sc["Address"]]) for r in data)
agg = collections.defaultdict(collections.Counter); ops = collections.Counter()
stall_cols = [c for c in sh if c.startswith("stall_") and "Not Issued" not in c]
for r in data:
    i = max(bisect.bisect_right(offs, addr(r[sc["Address"]]) - base) - 1, 0)
    fn = labels[i][1].split("$")[-1]; n_ex = int(float(r[sc["Instructions Executed"]] or 0))
    g = agg[fn]; g["inst"] += n_ex; g["samples"] += int(float(r[sc["# Samples"]] or 0))
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", r[sc["Source"]])
    if m:
        ops[m.group(2)] += n_ex
    for c in stall_cols:
        g[c] += int(float(r[sc[c]] or 0))
names = subprocess.run(["c++filt"], input="\n".join(agg), capture_output=True, text=True).stdout.splitlines()
ti = sum(g["inst"] for g in agg.values()); ts = sum(g["samples"] for g in agg.values())
print(f"\n## per device function: executed warp instructions ({ti:.3e}) and stall samples ({ts})")
for (fn, g), nm in sorted(zip(agg.items(), names), key=lambda t: -t[0][1]["samples"])[:40]:
    if g["samples"] < 0.0002 * ts and not any(k in nm for k in ("ph_load", "ph_output")):
        continue
    nm = re.sub(r"nmpc::Lay<\d+, \d+>", "L", nm).replace("nmpc::SolveArgs const&", "A"); nm = re.sub(r"_INTERNAL_\w+::", "", nm)
    top = sorted(((c[6:], g[c]) for c in stall_cols), key=lambda t: -t[1])[:3]
    print(f"  {100 * g['inst'] / ti:5.1f}% inst {100 * g['samples'] / ts:5.1f}% samples  {nm[:60]:60s} " + " ".join(f"{c} {100 * x / max(1, g['samples']):.0f}%" for c, x in top))
print("\n## opcode mix (executed warp instructions)")
tot = sum(ops.values())
print("  " + "  ".join(f"{o} {100 * c / tot:.1f}%" for o, c in ops.most_common(16)))
f64 = sum(c for o, c in ops.items() if o in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print(f"  FP64 arithmetic opcodes (DFMA DMUL DADD DSETP DMNMX): {100 * f64 / tot:.1f}% ; tensor opcodes (HMMA/DMMA/UTCMMA...): "
      f"{sum(c for o, c in ops.items() if 'MMA' in o)}")
