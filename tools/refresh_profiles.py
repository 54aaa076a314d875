"""Turn the captures of tools/profile_round.sh (gpurun_out/) into the tracked summaries under profiles/
and print the numbers the docs quote.  usage: python tools/refresh_profiles.py"""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")


def run(script, arg):
    return subprocess.run([sys.executable, os.path.join(ROOT, "tools", script), arg], capture_output=True, text=True).stdout


def dram_per_launch(path):
    vals = {}
    for line in open(path):
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            if line.startswith(key):
                unit = re.search(r"\[(\w+)\]", line).group(1)
                scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}[unit]
                vals[key] = [float(x) * scale for x in line.split(":", 1)[1].split("|")]
    rd, wr = vals["dram__bytes_read.sum"], vals["dram__bytes_write.sum"]
    return int(sum(a + b for a, b in zip(rd, wr)) / len(rd))


def main():
    for mode, name in (("fp32", "default"), ("tf32", "tf32"), ("simt", "simt")):
        with open(os.path.join(PROF, f"r1_launches_{name}_mode.txt"), "w") as f:
            f.write(run("summarize_launches.py", os.path.join("gpurun_out", f"launches_r1_{mode}.csv")))
    for k in ("x3_gather", "x3_wgrad", "umma_gather", "umma_wgrad", "smallc"):
        with open(os.path.join(PROF, f"r1_ncu_{k}.txt"), "w") as f:
            f.write(run("ncu_summary.py", os.path.join("gpurun_out", f"prof_{k}_r1.ncu-rep")))
    tpath = os.path.join(PROF, "ncu_traffic.json")
    t = json.load(open(tpath))
    xg, xw = dram_per_launch(os.path.join(PROF, "r1_ncu_x3_gather.txt")), dram_per_launch(os.path.join(PROF, "r1_ncu_x3_wgrad.txt"))
    ug, uw = dram_per_launch(os.path.join(PROF, "r1_ncu_umma_gather.txt")), dram_per_launch(os.path.join(PROF, "r1_ncu_umma_wgrad.txt"))
    for m in ("fp32", "tf32x3"):
        t[m + ":conv_fprop"] = xg
        t[m + ":conv_dgrad"] = xg
        t[m + ":conv_wgrad"] = xw
    t["tf32:conv_fprop"] = ug
    t["tf32:conv_dgrad"] = ug
    t["tf32:conv_wgrad"] = uw
    json.dump(t, open(tpath, "w"), indent=1)
    d = json.loads(open(os.path.join(OUT, "bench_r1_final.json")).read().strip().splitlines()[-1])
    for k in ("", "simt_mode", "tf32_mode"):
        o = d[k] if k else d
        r = o["roofline"]
        print(k or "default", o["value"], "e2e", o["e2e"]["value"], "ms", o.get("ms_per_step"), "roofline", r["achieved"], "/",
              r["peak"], "=", round(r["frac"], 3), {c: (v["ms_per_step"], v["tflops"]) for c, v in r["per_class"].items()})
    print("cpu", d["cpu_baseline"]["value"], "mnist", d["mnist_cnn"]["images_per_sec"])
    print("sweep C=K=256", d["conv2d_sweep"]["rows"][-1])


if __name__ == "__main__":
    os.chdir(ROOT)
    main()
