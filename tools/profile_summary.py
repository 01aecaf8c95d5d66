#!/usr/bin/env python
"""Turns the raw ncu CSVs written by tools/profile_all.sh into the tracked summaries under profiles/:

  r02_pair_kernel_ncu_full_summary.csv   one row per launch of tools/prof_shapes.py with the counters DESIGN.md cites
  r02_qkv_lora_traffic.json              DRAM bytes of ONE launch of the roofline kernel (read by bench.py)
  r02_launch_list_one_step_summary.csv   kernels of one routed step grouped by name: launches, total / avg us, share
"""
import csv
import json
import sys
from collections import OrderedDict
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))

COLS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]


def to_bytes(val: str, unit: str) -> float:
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    return float(val.replace(",", "")) * mult


def shapes_summary(src: Path, dst: Path):
    rows = list(csv.reader(open(src / "r02_shapes_raw.csv")))
    hdr, units, data = rows[0], rows[1], rows[2:]
    names = []
    try:
        text = (Path(__file__).resolve().parent / "prof_shapes.py").read_text()
        start = text.index("NAMES = [")
        names = eval(text[start + 8: text.index("]", start) + 1])      # noqa: S307 (our own file)
    except Exception:
        pass
    # launches of one round in program order: k1v2 / ln / k2 kernels (ln_lora_u, U pass and K2 add extra launches)
    labels = ["plain 96000x768->768", "fc1 96000x768->3072", "fc1+GELU (8 epilogue warps)",
              "out_proj head-major x + residual (8 epilogue warps)", "LayerNorm + U fused (one pass over h)",
              "q|k|v + routed LoRA r16: dense 256-wide tiles + extra K block, U ready (north-star kernel)",
              "fc2 96000x3072->768 + residual", "plain LayerNorm", "two-launch form: U pass",
              "two-launch form: dense launch", "q|k|v base only (LID pass)", "K2 pool (bf16 states)", "K2 head"]
    ik = hdr.index("Kernel Name")
    idx = [hdr.index(c) for c in COLS if c in hdr]
    out = [["shape", "kernel"] + [hdr[i] for i in idx], ["", ""] + [units[i] for i in idx]]
    traffic = None
    off = int(__import__("os").environ.get("PROFILE_LABEL_OFFSET", "0"))   # capture window not aligned to a round
    for n, r in enumerate(data):
        label = labels[(n + off) % len(labels)] if n < len(labels) else f"launch {n}"
        out.append([label, r[ik][:70]] + [r[i] for i in idx])
        if "north-star" in label:
            ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            traffic = to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw])
    with open(dst / "r02_pair_kernel_ncu_full_summary.csv", "w", newline="") as f:
        csv.writer(f).writerows(out)
    if traffic:
        (dst / "r02_qkv_lora_traffic.json").write_text(json.dumps(
            {"bytes": traffic, "source": "profiles/r02_pair_kernel_ncu_full_summary.csv (ncu --set full, one launch of "
                                         "k1v2<256,AUG> at M=96000, 768->2304, r=16; tools/profile_all.sh)"}) + "\n")
    return traffic


def launch_list_summary(src: Path, dst: Path):
    rows = [r for r in csv.reader(open(src / "r02_launch_list_one_step.csv")) if len(r) > 10]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        us = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1e-3)
        name = r[ik]
        name = name.split("(")[0] if not name.startswith("void") else name[5:].split("(")[0]
        a = agg.setdefault(name[:100], [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    fam = {"pair kernel k1v2": 0.0, "LayerNorm (plain + fused LN+U)": 0.0, "K2 router": 0.0, "own attention (fa_fwd)": 0.0,
           "cuDNN SDPA": 0.0, "torch / other": 0.0}
    with open(dst / "r02_launch_list_one_step_summary.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["# ncu --nvtx --nvtx-include timed/ --metrics gpu__time_duration.sum --clock-control none: every kernel of ONE "
                    "routed step (bench.py --steps 1 --warmup 3 --extras none); cold-cache serialised per-launch times, shares "
                    "are what is meaningful"])
        w.writerow(["kernel", "launches", "total_us", "avg_us", "share_pct"])
        for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([name, n, f"{us:.1f}", f"{us / n:.1f}", f"{100 * us / total:.2f}"])
            key = ("pair kernel k1v2" if "k1v2" in name or "skinny" in name or "k1_qv" in name else
                   "LayerNorm (plain + fused LN+U)" if "ln_" in name else "K2 router" if "k2_" in name else
                   "own attention (fa_fwd)" if "fa_fwd" in name or "decode_attn" in name else
                   "cuDNN SDPA" if "cudnn" in name or "sdpa" in name.lower() or "fmha" in name.lower() else "torch / other")
            fam[key] += us
        w.writerow(["# family shares (%)"] + [f"{k}={100 * v / total:.1f}" for k, v in fam.items()] + [f"total_us={total:.0f}"])


def main():
    src, dst = Path(sys.argv[1]), Path(sys.argv[2])
    t = shapes_summary(src, dst)
    launch_list_summary(src, dst)
    print("traffic bytes of the roofline kernel:", t)


if __name__ == "__main__":
    main()
