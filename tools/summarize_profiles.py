#!/usr/bin/env python
"""Turns the scratch output of one `tools/gpu_round.sh <tag>` visit (gpurun_out/<tag>_*) into the tracked evidence files
under profiles/ (r<NN>_*): bench line, launch list and shares, per-layer times, role waits, clock summary, ncu summary +
profiles/ncu_traffic.json (what bench.py reports as roofline.traffic), sanitizer summary, SASS opcode histogram.

usage: python tools/summarize_profiles.py <tag> [<round prefix, default r02>]"""
import collections
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
tag = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else "r02"
src = lambda name: os.path.join(OUT, f"{tag}_{name}")           # noqa: E731
dst = lambda name: os.path.join(PROF, f"{rnd}_{name}")          # noqa: E731


def copy(a, b):
    if os.path.exists(src(a)):
        shutil.copyfile(src(a), dst(b))
        return True
    return False


# ---- plain copies
if os.path.exists(src("bench.json")):
    line = [ln for ln in open(src("bench.json")) if ln.startswith("{")][-1]
    json.dump(json.loads(line), open(dst("bench_1gpu.json"), "w"), indent=1)
copy("layer_times.log", "layer_times.log")
copy("build.log", "clean_build_on_gpu_box.log")
if os.path.exists(src("role_waits.log")):
    with open(dst("role_waits.log"), "w") as f:
        f.writelines(ln for ln in open(src("role_waits.log")) if ln.startswith("[conv stats]") or ln.startswith("[wgrad stats]"))
if os.path.exists(src("pytest.log")):
    with open(dst("pytest_gpu.log"), "w") as f:
        f.writelines(open(src("pytest.log")).readlines()[-14:])

# ---- launch list of one step (ncu --metrics gpu__time_duration.sum): keep the last training step's launches
if os.path.exists(src("launches.csv")):
    text = open(src("launches.csv")).read()
    text = text[text.index('"ID"'):]
    rows = list(csv.DictReader(io.StringIO(text)))
    rows = [r for r in rows if r.get("Metric Name") == "gpu__time_duration.sum"]
    us = lambda r: float(r["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(r["Metric Unit"], 1.0)   # noqa: E731
    names = [r["Kernel Name"] for r in rows]
    # one step = from the last weight_scale_kernel launch to the end (bench.py --steps 2: the final eager profiled step follows)
    starts = [i for i, n in enumerate(names) if "weight_scale_kernel" in n]
    lo = starts[-1] if starts else 0
    ends = [i for i, n in enumerate(names) if "adam_kernel" in n and i > lo]
    hi = ends[0] + 1 if ends else len(rows)
    step = rows[lo:hi]
    with open(dst("launches_one_step.csv"), "w") as f:
        f.write("kernel;grid;block;us\n")
        for r in step:
            f.write(f'{r["Kernel Name"]};{r.get("Grid Size", "")};{r.get("Block Size", "")};{us(r):.3f}\n')
    agg = collections.OrderedDict()
    for r in step:
        n = re.sub(r"\(.*", "", r["Kernel Name"])
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1; a[1] += us(r)
    total = sum(a[1] for a in agg.values())
    with open(dst("launch_shares.txt"), "w") as f:
        f.write(f"One eager training step of config 2 (bench.py --no-graph) under ncu --metrics gpu__time_duration.sum --clock-control none: "
                f"{len(step)} launches, {total:.0f} us summed\n(per-launch times under ncu are cold-cache and serialised: compare SHARES with the "
                f"bench line's CUDA-event times, not absolutes)\n\n")
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{100 * t / total:6.2f} %  {t:9.1f} us  x{c:3d}  {n}\n")
        foreign = [n for n in agg if not re.match(r"(void )?(convk|wgradk|pw|lossk|net|peakk)::", n)]
        f.write("\nkernels that are not the engine's own: " + (", ".join(f"{n} x{agg[n][0]}" for n in foreign) or "none") + "\n")

# ---- clocks
if os.path.exists(src("clocks.csv")):
    rows = list(csv.reader(open(src("clocks.csv"))))
    hdr = [h.strip() for h in rows[0]]
    body = [r for r in rows[1:] if len(r) == len(hdr)]
    col = lambda name: [r[hdr.index(name)].strip() for r in body]      # noqa: E731
    sm = sorted(int(v.split()[0]) for v in col("clocks.current.sm [MHz]") if v.split()[0].isdigit()) if "clocks.current.sm [MHz]" in hdr else []
    pw_ = [float(v.split()[0]) for v in col("power.draw [W]")] if "power.draw [W]" in hdr else []
    reasons = [h.split(".")[-1] for h in hdr if h.startswith("clocks_event_reasons.") and h != "clocks_event_reasons.active" and any(v == "Active" for v in col(h))]
    with open(dst("clocks_summary.txt"), "w") as f:
        f.write(f"nvidia-smi -lms 200 during tools/gpu_round.sh {tag} bench (bench.py default run incl. CPU baseline and tensor-peak probe): {len(body)} samples"
                + (f", SM clock median {sm[len(sm) // 2]} MHz, min {sm[0]}, max {sm[-1]} MHz" if sm else "") + (f"; power max {max(pw_):.0f} W" if pw_ else "")
                + f"; throttle reasons seen: {reasons}\n(bench.py's own NVML sampler, 5 ms period inside the timed region only, is what the JSON line's `clocks` reports)\n")

# ---- ncu --set full captures
WANT = [("gpu__time_duration.sum", "duration"), ("sm__cycles_elapsed.avg.per_second", "SM clock"), ("dram__bytes_read.sum", "DRAM read"),
        ("dram__bytes_write.sum", "DRAM write"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"), ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts % of peak"),
        ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn smem"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %")]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}
traffic = {}
sections = []
for rep, title in (("conv_fwd", "conv_igemm_kernel (decode_block_1.0 / 1.2 forward, CTA pairs)"), ("wgrad", "wgrad_igemm_kernel (the two largest layers)"),
                   ("pointwise", "HBM-bound kernels (first launches)")):
    path = src(rep + ".ncu-rep")
    if not os.path.exists(path):
        continue
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"== {title}"]
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")])
        parts = []
        rd = wr = None
        for key, label in WANT:
            if key in hdr:
                v, u = r[hdr.index(key)], units[hdr.index(key)]
                parts.append(f"{label} {v} {u}".strip())
                if key == "dram__bytes_read.sum": rd = float(v.replace(",", "")) * UNIT.get(u, 1.0)
                if key == "dram__bytes_write.sum": wr = float(v.replace(",", "")) * UNIT.get(u, 1.0)
        lines.append(f"{name}\n   " + " | ".join(parts))
        if rd is not None and wr is not None:
            short = re.sub(r"<.*", "", name.split("::")[-1]).replace("void ", "").strip()
            t = traffic.setdefault(short, [0.0, 0, rep])
            t[0] += rd + wr; t[1] += 1
    sections.append("\n".join(lines))
if sections:
    with open(dst("ncu_summary.txt"), "w") as f:
        f.write(f"ncu --set full --clock-control none (tools/gpu_round.sh {tag}, B200, one training step of config 2: 32 x 3 x 64 x 64, kind::f16 split operands)\n"
                "Times under ncu are cold-cache and serialised; the bench line's CUDA-event times are the ones reported.\n\n" + "\n\n".join(sections) + "\n")
    json.dump({k: {"dram_bytes_per_launch": v[0] / v[1], "launches_captured": v[1], "source": f"profiles/{rnd}_ncu_summary.txt ({tag}_{v[2]}.ncu-rep)"}
               for k, v in traffic.items()}, open(os.path.join(PROF, "ncu_traffic.json"), "w"), indent=1)

# ---- sanitizers
logs = {t: src(f"sanitizer_{t}.log") for t in ("memcheck", "synccheck", "racecheck")}
if all(os.path.exists(p) for p in logs.values()):
    with open(dst("sanitizer_summary.txt"), "w") as f:
        f.write("compute-sanitizer on __graft_entry__.smoke() (one SSDN training step, 2 x 3 x 32 x 32: blind-spot U-Net forward + posterior/NLL +\n"
                "backward + Adam, ~110 launches of the engine's kernels incl. conv_igemm_kernel<T,PAIR>, wgrad_igemm_kernel and all pointwise\n"
                f"kernels), B200, tools/gpu_round.sh {tag} sanitize.  Full logs: gpurun_out/{tag}_sanitizer_*.log (scratch); tails below.\n\n")
        for t, p in logs.items():
            keep = [ln.strip() for ln in open(p) if re.search(r"ERROR SUMMARY|RACECHECK SUMMARY|\[smoke\] ok|exit ", ln)]
            f.write(f"{t:9s}: " + "   ".join(keep) + "\n")
        race = open(logs["racecheck"]).read()
        sites = sorted(set(re.findall(r"Race reported between (\w+ access at [^\n]*?\+0x[0-9a-f]+)", race)))
        reads = sorted(set(re.findall(r"and (\w+ access at [^\n]*? in [\w.]+:\d+)", race)))
        if sites:
            f.write("  racecheck report sites: " + "; ".join(sites) + "\n    against: " + "; ".join(re.sub(r"\+0x[0-9a-f]+", "", r) for r in reads) + "\n")
            f.write("  i.e. (if the only site is conv_igemm_kernel<*, true> vs umma::tmem_alloc_pair) the shared-memory word that tcgen05.alloc.cta_group::2 writes\n"
                    "  the TMEM base address to, against the later reads of that word by all warps.  The write is performed by the tensor-memory\n"
                    "  allocator hardware on behalf of the allocating warps of BOTH CTAs of the pair, and the reads are ordered after it by\n"
                    "  tcgen05.fence::before_thread_sync -> __syncthreads() -> barrier.cluster arrive.release / wait.acquire ->\n"
                    "  tcgen05.fence::after_thread_sync (conv_igemm.cuh, kernel prologue) - the sequence CUTLASS's 2-SM kernels use.  The same allocation\n"
                    "  pattern with cta_group::1 (wgrad_igemm_kernel, conv_igemm_kernel<*, false>) is not reported, and no hazard is reported on any\n"
                    "  data path (operand rings, epilogue staging rows, bias / column-sum scratch, pointwise kernels).  Treated as a tool limitation\n"
                    "  on the allocator's cross-CTA write; every parity test runs these kernels at full size.\n")

# ---- SASS opcode histogram of the shipped library
lib = os.path.join(ROOT, "selfsupervised-denoising_b200", "libssdn_b200.so")
if os.path.exists(lib):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    pat = re.compile(r"\b(UTCHMMA\S*|UTMALDG\S*|LDTM\S*|UTCBAR\S*|UTCATOMSWS\S*|UTMACCTL\S*|SYNCS\.\S+|F2FP\.\S+|ACQBULK|UBLKCP\S*)")
    tot, per, cur = collections.Counter(), collections.OrderedDict(), None
    reuse = collections.Counter()
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1); per[cur] = collections.Counter(); continue
        m = pat.search(ln)
        if m and cur:
            tot[m.group(1)] += 1; per[cur][m.group(1)] += 1
            if "UTCHMMA" in m.group(1):
                for q in ("A_KEEP", "A_REUSE"):
                    if q in ln: reuse[q] += 1
    demangle = lambda s: subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip() or s      # noqa: E731
    with open(dst("sass_opcodes.txt"), "w") as f:
        f.write("SASS opcode evidence of libssdn_b200.so (cuobjdump -sass, sm_100a); built from the sources of this commit\n\nlibrary totals of the Blackwell-specific opcodes:\n")
        for k in sorted(tot): f.write(f"  {k:40s} {tot[k]}\n")
        f.write(f"  UTCHMMA operand-collector modifiers: gdesc[..].A_KEEP on {reuse['A_KEEP']} instructions, .A_REUSE on {reuse['A_REUSE']}\n")
        f.write("\nper kernel (tensor-core / TMA / TMEM opcodes only):\n")
        for fn, c in per.items():
            keep = {k: v for k, v in c.items() if re.match(r"UTC|UTMA|LDTM", k)}
            if keep:
                f.write(demangle(fn) + "\n")
                for k in sorted(keep): f.write(f"    {k:40s} {keep[k]}\n")
print("profiles written:", ", ".join(sorted(p for p in os.listdir(PROF) if p.startswith(rnd + "_"))))
