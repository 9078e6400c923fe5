"""Collect the logs a gpurun call left in gpurun_out/ into the committed,
judge-readable files under profiles/ (round 1).  Usage: python tools/make_profiles.py"""

import collections
import csv
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GP = os.path.join(ROOT, "gpurun_out")


def last_json(path):
    if not os.path.exists(path):
        return None
    for line in reversed(open(path).read().strip().splitlines()):
        line = line.strip()
        if line.startswith("{"):
            return json.loads(line)
    return None


def bench_md():
    lines = ["# Round 1 — bench.py on B200 (driver contract lines)", ""]
    for title, name in (("N = 1, `python bench.py` (defaults: 40 steps, 5 warm-up)", "bench_full.log"),
                        ("N = 1, `python bench.py --impl reference --steps 3 --warmup 1`", "bench_ref.log"),
                        ("N = 2, `torchrun --nproc-per-node 2 bench.py --gpus 2 --steps 20 --warmup 5`",
                         "bench_2gpu.log"),
                        ("N = 4, `torchrun --nproc-per-node 4 bench.py --gpus 4 --steps 20 --warmup 5`",
                         "bench_4gpu.log"),
                        ("N = 8, `torchrun --nproc-per-node 8 bench.py --gpus 8 --steps 20 --warmup 5`",
                         "bench_8gpu.log")):
        d = last_json(os.path.join(GP, name))
        if d is None:
            continue
        lines += ["## " + title, ""]
        if d.get("impl") == "reference":
            lines += ["* CPU oracle port, %d cores: **%.1f M channel-samples/s**" %
                      (d["cpu_baseline"]["cores"], d["value"] / 1e6), ""]
        else:
            lines += ["* value (HBM-resident): **%.2f G channel-samples/s**, %.3f ms per step of "
                      "%d x %d samples" % (d["value"] / 1e9, d["ms_per_step"],
                                           d["config"]["rows_per_gpu"], d["config"]["chunk"]),
                      "* e2e (pinned host chunks, H2D inside): **%.2f G channel-samples/s** "
                      "(%.1f GB/s of float64 over PCIe per GPU)" %
                      (d["e2e"]["value"] / 1e9, d["e2e"]["value"] * 8 / 1e9 / d["n_gpus"]),
                      "* our kernels launched in the timed region: %d; clocks %s" %
                      (d["gpu_launches"], json.dumps(d["clocks"]))]
            nar = (d.get("e2e") or {}).get("narrow_inputs")
            if nar:
                lines += ["* e2e with narrower samples (widened on the device; extra, not the "
                          "contract's e2e): " + ", ".join(
                              "%s **%.2f G/s** (%.1f GB/s over PCIe)" %
                              (k, v["value"] / 1e9,
                               v["value"] / (d["config"]["rows_per_gpu"] * d["config"]["chunk"])
                               * v["h2d_bytes_per_step"] / 1e9 / d["n_gpus"])
                              for k, v in nar.items())]
            f32 = d.get("float32_compute")
            if f32:
                lines += ["* the same pipeline with the opt-in float32 arithmetic (extra): **%.2f G "
                          "channel-samples/s**, %.3f ms per step, PSD within %.1e of the float64 "
                          "result (per channel, of its largest bin)" %
                          (f32["value"] / 1e9, f32["ms_per_step"], f32["max_rel_diff_vs_float64"])]
            if d.get("host_enqueue"):
                lines += ["* host enqueue pace inside the timed region: %s" %
                          json.dumps(d["host_enqueue"])]
            if d.get("per_rank"):
                lines += ["* per rank: own ms/step %s; sum of kernel times per step %s" %
                          (d["per_rank"]["ms_per_step"], d["per_rank"]["kernel_ms_per_step"])]
            if "cpu_baseline" in d:
                one = d["cpu_baseline"].get("single_core")
                lines += ["* cpu_baseline (oracle port, %d cores): %.1f M channel-samples/s%s" %
                          (d["cpu_baseline"]["cores"], d["cpu_baseline"]["value"] / 1e6,
                           "; one core (the reference as shipped is single-threaded): %.2f M" %
                           (one["value"] / 1e6) if one else "")]
            r = d.get("roofline")
            if r:
                lines += ["* roofline (dominant kernel `%s`, %.0f %% of the step): %.0f GB/s "
                          "algorithmic = **%.1f %%** of %.0f GB/s measured HBM; DRAM traffic per "
                          "launch (ncu) %.3f GB vs %.3f GB algorithmic" %
                          (r["kernel"], 100 * r["share_of_step"], r["achieved"], 100 * r["frac"],
                           r["peak"], (r["traffic"] or 0) / 1e9,
                           (r.get("alg_bytes_per_launch") or 0) / 1e9)]
            lines += ["", "| kernel | launches | ms / step | algorithmic GB/s |", "|---|---|---|---|"]
            for k, v in d["kernels"].items():
                lines.append("| %s | %d | %.3f | %.0f |" % (k, v["launches"],
                                                             v["ms_total"] / d["steps"], v["alg_GBps"]))
            if "named_kernels" in d:
                lines += ["", "north_star kernels alone (256 x 1e6 float64, CUDA events):", "",
                          "| kernel | ms | G ch-samples/s | % of measured HBM roofline |", "|---|---|---|---|"]
                for k, v in d["named_kernels"].items():
                    lines.append("| %s | %.3f | %.1f | %.1f %% |" %
                                 (k, v["ms"], v["channel_samples_per_s"] / 1e9, 100 * v["frac"]))
        lines += ["", "```json", json.dumps(d), "```", ""]
    open(os.path.join(OUT, "r01_bench.md"), "w").write("\n".join(lines))


def kbench_md():
    path = os.path.join(GP, "kbench.log")
    if not os.path.exists(path):
        return
    body = open(path).read().strip()
    text = ("# Round 1 — per-kernel device timings (`python tools/kernel_bench.py`, B200)\n\n"
            "CUDA events, median of 5 after 2 warm-up launches, inputs 256 rows x 1e6 float64\n"
            "(2 GB, >> 126 MB L2) unless the line says otherwise.  Last column: algorithmic bytes\n"
            "(SURVEY.md 8d) / time as a fraction of the measured HBM copy bandwidth.\n\n```\n"
            + body + "\n```\n")
    open(os.path.join(OUT, "r01_kernel_bench.md"), "w").write(text)


def launches_md():
    path = os.path.join(GP, "r01_launches.csv")
    if not os.path.exists(path):
        return
    raw = [l for l in open(path) if l.startswith('"')]
    open(os.path.join(OUT, "r01_launches.csv"), "w").write("".join(raw))
    rows = list(csv.reader(raw))
    idx = {h: i for i, h in enumerate(rows[0])}
    tot, cnt = {}, collections.Counter()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "")[:64]
        tot[name] = tot.get(name, 0.0) + float(r[idx["Metric Value"]])
        cnt[name] += 1
    total = sum(tot.values())
    lines = ["# Round 1 — ncu launch list of the bench command", "",
             "    ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \\",
             "        --log-file gpurun_out/r01_launches.csv \\",
             "        python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --no-named", "",
             "(after the same command exited 0 without ncu).  Per-launch times under ncu are cold",
             "and serialised: compare shares.  Raw list: `r01_launches.csv` (%d launches;"
             % (len(rows) - 1),
             "the run covers 3 warm-up + 4 timed + 3 cool-down chunks and the one-off input",
             "generation by torch's `normal_` kernel).", "",
             "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        lines.append("| `%s` | %d | %.3f | %.1f %% |" % (k, cnt[k], v / 1e6, 100 * v / total))
    ours = sum(v for k, v in tot.items() if k.startswith("osz::"))
    lines += ["", "Kernels of `libosz_b200.so` account for %.1f %% of the GPU time; the rest is the"
              % (100 * ours / total),
              "synthetic input generation and small torch copies/fills of the staging rings."]
    open(os.path.join(OUT, "r01_launches.md"), "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    bench_md()
    kbench_md()
    launches_md()
    print(sorted(os.listdir(OUT)))
