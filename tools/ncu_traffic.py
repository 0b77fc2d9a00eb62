"""profiles/traffic.json from `ncu --set full` captures of the pipeline kernels (one launch = 1M Franka rows):
executed FP32 / FP64 flops, thread instructions and DRAM bytes per row, read by bench.py for the roofline block.

    python tools/ncu_traffic.py gpurun_out/prof_fk_cull_kernel.ncu-rep gpurun_out/prof_mid_kernel.ncu-rep ... [--rows 1000000]
"""
import csv, json, subprocess, sys
from pathlib import Path

reps = [a for a in sys.argv[1:] if a.endswith(".ncu-rep")]
rows = int(sys.argv[sys.argv.index("--rows") + 1]) if "--rows" in sys.argv else 1_000_000
out = {"source": "ncu --set full --clock-control none, one launch per kernel of `bench.py` (1M Franka rows, scene_with_obstacles); "
                 + ", ".join(Path(r).name for r in reps),
       "algorithmic_bytes_per_launch": 37.0 * rows, "rows_per_launch": rows, "kernels": {}}
for rep in reps:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(txt.splitlines()))
    h, v = r[0], r[2]
    g = lambda name: float(v[h.index(name)].replace(",", "")) if name in h else 0.0
    name = v[h.index("Kernel Name")].split("(")[0].split("::")[-1]
    unit = lambda name: r[1][h.index(name)] if name in h else ""
    def bytes_of(name):
        x, u = g(name), unit(name)
        return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    # op counts come as chip-wide rates per elapsed SM cycle: x elapsed cycles = thread-instructions of the launch
    cyc = g("smsp__cycles_elapsed.avg")
    op = lambda k: g(f"smsp__sass_thread_inst_executed_op_{k}_pred_on.sum.per_cycle_elapsed") * cyc
    fadd, fmul, ffma = op("fadd"), op("fmul"), op("ffma")
    dadd, dmul, dfma = op("dadd"), op("dmul"), op("dfma")
    lanes = g("smsp__thread_inst_executed_per_inst_executed.ratio")
    out["kernels"][name] = {
        "fp32_flop_per_row": (fadd + fmul + 2 * ffma) / rows,
        "fp64_flop_per_row": (dadd + dmul + 2 * dfma) / rows,
        "thread_inst_per_row": g("smsp__inst_executed.sum") * lanes / rows,
        "warp_inst_per_row": g("smsp__inst_executed.sum") / rows,
        "dram_bytes_per_launch": bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum"),
        "duration_us_under_ncu": g("gpu__time_duration.sum"),
        "registers": g("launch__registers_per_thread"),
        "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "lanes_per_inst": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
    }
Path("profiles/traffic.json").write_text(json.dumps(out, indent=1))
print(json.dumps(out, indent=1))
