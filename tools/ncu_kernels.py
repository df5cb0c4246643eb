#!/usr/bin/env python
"""Summarise an `ncu --set full` report of the hot-path kernels (tools/gpu_profile.sh).

    python tools/ncu_kernels.py gpurun_out/<tag>_prof.ncu-rep profiles/ncu_<tag>_kernels [points]

writes <out>.json (read by bench.py: executed FP64 flop and DRAM bytes per launch and per point, FP64
pipe activity) and <out>.txt (the raw metrics a reviewer wants to see: time, occupancy, stalls)."""
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__cycles_active.avg", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmma_pred_on.sum", "smsp__inst_executed_pipe_fp64.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def source_hash():
    """Hash of the CUDA sources the capture was made from: the working tree, or -- when the tree has moved on
    since the gpurun call -- the git revision named in CHOMP_PROFILE_REV."""
    h = hashlib.sha256()
    rev = os.environ.get("CHOMP_PROFILE_REV")
    if rev:
        names = subprocess.run(["git", "-C", ROOT, "ls-tree", "--name-only", rev, "chomp_b200/csrc/"], capture_output=True,
                               text=True).stdout.split()
        for f in sorted(os.path.basename(n) for n in names):
            if f.endswith((".cu", ".cuh")):
                h.update(subprocess.run(["git", "-C", ROOT, "show", "%s:chomp_b200/csrc/%s" % (rev, f)], capture_output=True).stdout)
        return h.hexdigest()[:16]
    d = os.path.join(ROOT, "chomp_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def kernel_short_name(name):
    """chomp::wtheta_kernel<(bool)0>(...) / void chomp::mass_tables_kernel<false>(...) -> the bare kernel name."""
    base = name.split("(")[0] if not name.startswith("void ") else name[5:].split("(")[0]
    base = base.split("<")[0]
    return base.split("::")[-1].strip()


def to_float(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def dmma_warp_instructions(rep):
    """Executed DMMA warp instructions per kernel from the SASS source page (the
    smsp__sass_thread_inst_executed_op_dmma counter is not collected on this part)."""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    out, cur, col_src, col_n = {}, None, None, None
    for r in csv.reader(txt.splitlines()):
        if not r:
            continue
        if r[0] == "Kernel Name":
            cur = kernel_short_name(r[1])
            if cur in out:          # the page lists every kernel twice (two views of the same SASS): first one only
                cur = None
            else:
                out[cur] = 0.0
            continue
        if "Instructions Executed" in r and "Source" in r:
            col_src, col_n = r.index("Source"), r.index("Instructions Executed")
            continue
        if cur is None or col_src is None or len(r) <= col_n:
            continue
        ops = r[col_src].split()
        if ops and (ops[0].startswith("DMMA") or (ops[0].startswith("@") and len(ops) > 1 and ops[1].startswith("DMMA"))):
            try:
                out[cur] += float(r[col_n])
            except ValueError:
                pass
    return out


def main():
    rep, out = sys.argv[1], sys.argv[2]
    points = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
             "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}
    kernels, lines = {}, []
    dmma_warp = dmma_warp_instructions(rep)
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        short = kernel_short_name(name)
        lines.append("--- " + name[:100])
        m = {}
        for i, h in enumerate(hdr):
            stall = "issue_stalled" in h and "per_issue_active" in h
            if h in KEEP or stall:
                v = to_float(r[i])
                if v is None:
                    continue
                if stall and v < 0.15:
                    continue
                lines.append("  %-95s %s %s" % (h, r[i], units[i]))
                m[h] = v*scale.get(units[i], 1.0) if (h.startswith("dram__bytes") or h == "gpu__time_duration.sum") else v
        dfma = m.get("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", 0.0)
        dmul = m.get("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", 0.0)
        dadd = m.get("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", 0.0)
        dmma = m.get("smsp__sass_thread_inst_executed_op_dmma_pred_on.sum", 0.0) or 32.0*dmma_warp.get(short, 0.0)
        # DMMA m8n8k4: 8*8*4 FMA per warp instruction = 16 flop per thread instruction
        flop = 2.0*dfma + dmul + dadd + 16.0*dmma
        t = m.get("gpu__time_duration.sum", 0.0)
        kernels[short] = {
            "ncu_ms": 1e3*t, "grid": m.get("launch__grid_size"), "registers": m.get("launch__registers_per_thread"),
            "smem_dynamic_kb": m.get("launch__shared_mem_per_block_dynamic"),
            "dfma": dfma, "dmul": dmul, "dadd": dadd, "dmma": dmma,
            "executed_fp64_flop": flop, "executed_fp64_flop_per_point": flop/points,
            "executed_tflops_under_ncu": flop/t/1e12 if t else None,
            "fp64_pipe_pct": m.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            "dmma_pipe_pct": m.get("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active"),
            "lsu_wavefront_pct": m.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
            "warps_active_pct": m.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "issue_active_pct": m.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "threads_per_inst": m.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "dram_read_bytes": m.get("dram__bytes_read.sum"), "dram_write_bytes": m.get("dram__bytes_write.sum"),
            "dram_bytes_per_point": ((m.get("dram__bytes_read.sum") or 0) + (m.get("dram__bytes_write.sum") or 0))/points,
        }
    doc = {"report": os.path.basename(rep), "points": points, "source_hash": source_hash(),
           "how": "ncu --set full --clock-control none + smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd,dmma}_pred_on.sum; "
                  "flop = 2 dfma + dmul + dadd + 16 dmma (thread instructions)",
           "kernels": kernels}
    json.dump(doc, open(out + ".json", "w"), indent=1)
    open(out + ".txt", "w").write("\n".join(lines) + "\n")
    tot = sum(k["ncu_ms"] for k in kernels.values())
    for k, v in kernels.items():
        print("%-22s %6.3f ms (%4.1f%%)  fp64 pipe %5.1f%%  dmma %4.1f%%  lsu %4.1f%%  exec %6.2f TF/s  flop/pt %.3g  dram %6.1f MB  thr/inst %.1f  warps %.0f%%" % (
            k, v["ncu_ms"], 100*v["ncu_ms"]/tot, v["fp64_pipe_pct"] or 0, v.get("dmma_pipe_pct") or 0, v.get("lsu_wavefront_pct") or 0,
            v["executed_tflops_under_ncu"] or 0,
            v["executed_fp64_flop_per_point"], ((v["dram_read_bytes"] or 0) + (v["dram_write_bytes"] or 0))/1e6,
            v["threads_per_inst"] or 0, v["warps_active_pct"] or 0))


if __name__ == "__main__":
    main()
