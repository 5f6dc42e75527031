"""Turn the outputs of tools/gpu_final.sh (gpurun_out/final_*) into the tracked files under profiles/ (round 2)."""
from __future__ import annotations

import collections
import csv
import json
import re
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT, PROF = ROOT / "gpurun_out", ROOT / "profiles"


def copy(src: str, dst: str) -> None:
    if (OUT / src).exists():
        shutil.copyfile(OUT / src, PROF / dst)


def launch_table(path: Path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = rows[0]
    kn, mv, mn, idc = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name"), h.index("ID")
    per = collections.defaultdict(dict)
    for r in rows[1:]:
        per[(r[idc], r[kn])][r[mn]] = float(r[mv].replace(",", ""))
    tot, cnt, dram = collections.Counter(), collections.Counter(), collections.Counter()
    for (_, k), m in per.items():
        k = re.sub(r"\(.*", "", k).replace("void ", "").replace("tl::", "")
        tot[k] += m.get("gpu__time_duration.sum", 0.0)
        cnt[k] += 1
        dram[k] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    return tot, cnt, dram


def launches_md() -> None:
    lines = ["# Launch lists of the final tree (ncu `gpu__time_duration.sum`, `--clock-control none`)\n",
             "Serialised and cold-cache: use the SHARES; the step times quoted in DESIGN.md come from CUDA events.\n"]
    traffic = {}
    for name, title in [("final_launches_decode_b1.csv", "decode, batch 1, context 128 (one step; `tools/launch_list.py --mode decode --batch 1 --context 128`)"),
                        ("final_launches_decode_b64.csv", "decode, 64 slots x 1024-token contexts (one step)"),
                        ("final_launches_chunk128.csv", "one 128-token chunked-prefill step at context 512")]:
        path = OUT / name
        if not path.exists():
            continue
        tot, cnt, dram = launch_table(path)
        s = sum(tot.values())
        lines.append(f"\n## {title}\n\n`profiles/r02_{name[6:]}`: {sum(cnt.values())} launches, {s / 1e3:.0f} us serialised\n")
        lines.append("| kernel | launches | total us | share | avg us | DRAM MB / launch |\n|---|---|---|---|---|---|")
        for k, v in tot.most_common(12):
            d = f"{dram[k] / cnt[k] / 1e6:.2f}" if dram[k] else "-"
            lines.append(f"| `{k[:70]}` | {cnt[k]} | {v / 1e3:.1f} | {100 * v / s:.1f} % | {v / cnt[k] / 1e3:.2f} | {d} |")
        for k in tot:
            if "stream5" in k and dram[k] and cnt[k] >= 100 and "b1" in name:
                traffic["w4a16_stream5_kernel"] = {"dram_bytes_per_launch": round(dram[k] / cnt[k]), "launches": cnt[k], "source": f"profiles/r02_{name[6:]}"}
        copy(name, "r02_" + name[6:])
    (PROF / "r02_launches_summary.md").write_text("\n".join(lines) + "\n")
    if traffic:
        (PROF / "traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")


METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
TITLES = ["tcgen05 flash prefill, L = S = 4096 (causal, 32/8 heads)", "same kernel as split-KV decode attention, B = 64, S = 8192",
          "swap-AB GEMM M = 64, 2560 -> 19456 (gate|up)", "swap-AB GEMM M = 64, 2560 -> 151936 (tied head)",
          "streaming matvec M = 1, 2560 -> 19456 (gate|up)", "streaming matvec M = 1, 2560 -> 151936 (tied head)",
          "CTA-pair GEMM M = 4096, 2560 -> 19456 (gate|up)"]


def ncu_md() -> None:
    rep = OUT / "final_kernels.ncu-rep"
    if not rep.exists():
        return
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    out = ["# `ncu --set full --clock-control none` of the round-2 kernels (final tree)\n",
           "One launch each after three warm-up rounds (`tools/ncu_round2.py`); durations under ncu are cold-cache, the timed numbers are in DESIGN.md.\n"]
    stall_cols = [c for c in h if c.startswith("smsp__pcsamp_warps_issue_stalled_") and not c.endswith("_not_issued")]
    for i, r in enumerate(rows[2:]):
        out.append(f"\n### {TITLES[i] if i < len(TITLES) else 'kernel ' + str(i)}\n| metric | value |\n|---|---|")
        out.append(f"| `Kernel Name` | {r[h.index('Kernel Name')][:90]} |")
        for k in ("Grid Size", "Block Size"):
            out.append(f"| `{k}` | {r[h.index(k)]} |")
        for m in METRICS:
            if m in h:
                out.append(f"| `{m}` | {r[h.index(m)]} {units[h.index(m)]} |")
        st = sorted(((float(r[h.index(c)] or 0), c.replace("smsp__pcsamp_warps_issue_stalled_", "")) for c in stall_cols), reverse=True)
        total = sum(v for v, _ in st) or 1.0
        out.append("| warp stall samples | " + ", ".join(f"{n} {100 * v / total:.0f} %" for v, n in st[:6]) + " |")
    # SASS evidence from the shipped library
    lib = ROOT / "tiny-llm_b200" / "extensions_b200" / "tiny_llm_ext_b200" / "libtiny_llm_b200.so"
    sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
    out.append("\n## SASS evidence (`cuobjdump -sass libtiny_llm_b200.so`, mnemonic counts per kernel)\n\n| kernel | UTCHMMA (tcgen05.mma) | of which `.2CTA` | UTCBAR (commit) | LDTM / STTM (tcgen05.ld / st) | UTMALDG (TMA) |\n|---|---|---|---|---|---|")
    cur, counts = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur:
            for mn in ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG"):
                if mn in line:
                    counts[cur][mn] += 1
            if "UTCHMMA.2CTA" in line:
                counts[cur]["2CTA"] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    for name, c in zip(demangle, counts.values()):
        if c["UTCHMMA"] or c["UTMALDG"]:
            short = name.split("(")[0].replace("void ", "").replace("tl::", "")[:80]
            out.append(f"| `{short}` | {c['UTCHMMA']} | {c['2CTA']} | {c['UTCBAR']} | {c['LDTM']} / {c['STTM']} | {c['UTMALDG']} |")
    (PROF / "r02_ncu_kernels.md").write_text("\n".join(out) + "\n")


def main() -> None:
    PROF.mkdir(exist_ok=True)
    for w in ("decode", "prefill", "serve", "serve8k", "reference"):
        copy(f"final_bench_{w}.json", f"r02_bench_{w}.json")
    copy("final_kbench.json", "r02_kbench_projections_final.json")
    copy("final_kbench_att.json", "r02_kbench_decode_attention_final.json")
    copy("final_smi.txt", "r02_final_smi.txt")
    if (OUT / "final_gemm_bench.txt").exists():
        last = (OUT / "final_gemm_bench.txt").read_text().strip().splitlines()[-1]
        try:
            json.loads(last)
            (PROF / "r02_gemm_bench_default.json").write_text(last + "\n")
        except ValueError:
            pass
    launches_md()
    ncu_md()
    print("profiles updated:", sorted(p.name for p in PROF.glob("r02_*")))


if __name__ == "__main__":
    sys.exit(main())
