"""Turn the captures of tools/profile_round2.sh (gpurun_out/) into the tracked summaries under profiles/ and
update profiles/ncu_traffic.json.  usage: python tools/refresh_profiles_r2.py"""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")


def run(script, *args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "tools", script)] + list(args), capture_output=True, text=True).stdout


def dram_per_launch(text):
    vals = {}
    for line in text.splitlines():
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            if line.startswith(key + " ["):
                unit = re.search(r"\[(\w+)\]", line).group(1)
                scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}[unit]
                vals[key] = [float(x) * scale for x in line.split(":", 1)[1].split("|")]
    rd, wr = vals["dram__bytes_read.sum"], vals["dram__bytes_write.sum"]
    return int(sum(a + b for a, b in zip(rd, wr)) / len(rd))


def main():
    os.chdir(ROOT)
    for mode, name in (("fp32", "default"), ("tf32", "tf32"), ("simt", "simt")):
        keys = "gather_f16,wgrad_f16,stage_f16,bn_reduce,smallc,pool_" if mode == "fp32" else "umma_conv,conv_gather,conv_wgrad"
        with open(os.path.join(PROF, f"r2_launches_{name}_mode.txt"), "w") as f:
            f.write(run("step_launches.py", os.path.join("gpurun_out", f"launches_r2_{mode}.csv"), "--list", keys))
    texts = {}
    for k in ("f16_gather", "f16_wgrad"):
        texts[k] = run("ncu_summary.py", os.path.join("gpurun_out", f"prof_{k}_r2.ncu-rep"))
        with open(os.path.join(PROF, f"r2_ncu_{k}.txt"), "w") as f:
            f.write(texts[k])
    for k in ("membound", "smallc"):   # summarised on the GPU box (the captures exceed gpurun's 64 MiB return limit)
        src = os.path.join(OUT, f"r2_ncu_{k}.txt")
        if os.path.exists(src):
            with open(os.path.join(PROF, f"r2_ncu_{k}.txt"), "w") as f:
                f.write(open(src).read())
    tpath = os.path.join(PROF, "ncu_traffic.json")
    t = json.load(open(tpath))
    g, w = dram_per_launch(texts["f16_gather"]), dram_per_launch(texts["f16_wgrad"])
    for m in ("fp32", "f16x3"):
        t[m + ":conv_fprop"] = g
        t[m + ":conv_dgrad"] = g
        t[m + ":conv_wgrad"] = w
    t["_note_r2"] = ("fp32 / f16x3 entries: FP16x3 kernels, profiles/r2_ncu_f16_*.txt (the staged operand copies are read, "
                     "so a launch's DRAM bytes are about those of the FP32 tensors); tf32x3 / fp32_tf32x3 keep round 1's captures")
    t["fp32_tf32x3:conv_fprop"] = t["tf32x3:conv_fprop"]
    t["fp32_tf32x3:conv_dgrad"] = t["tf32x3:conv_dgrad"]
    t["fp32_tf32x3:conv_wgrad"] = t["tf32x3:conv_wgrad"]
    json.dump(t, open(tpath, "w"), indent=1)
    print("traffic per launch: gather", g, "wgrad", w)


if __name__ == "__main__":
    main()
