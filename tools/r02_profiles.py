"""Turns the files tools/r02_capture.sh left under gpurun_out/ into the round-2 files of profiles/ (run here, after gpurun
merged them back): launch tables, ncu summaries with SASS hot spots, the SASS census, the bench lines, and the
per-launch DRAM bytes / instruction counts bench.py reads from profiles/traffic.json."""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
G, P, T = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles"), os.path.join(ROOT, "tools")


def run(*a, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, *a], capture_output=True, text=True, env=e).stdout


def write(name, head, body):
    open(os.path.join(P, name), "w").write(head.rstrip("\n") + "\n" + body)
    print("wrote", name, len(body.splitlines()), "lines")


for wl, lim in (("1080p_b32", 48), ("vga_b128", 43), ("single_1080p", 60)):
    src = os.path.join(G, f"r02_launches_{wl}.csv")
    write(f"r02_launches_{wl}.txt", "", run(os.path.join(T, "launch_table.py"), src, str(lim)))
    if wl != "single_1080p":
        shutil.copy(src, os.path.join(P, f"r02_launches_{wl}.csv"))
write("r02_ncu_k_descriptor.txt",
      "# ncu --set full --import-source on, k_descriptor, 1080p workload, 32 images per launch (tools/r02_capture.sh; tools/ncu_summary.py, tools/ncu_source.py)",
      run(os.path.join(T, "ncu_summary.py"), os.path.join(G, "r02_desc_raw.csv")) +
      run(os.path.join(T, "ncu_source.py"), os.path.join(G, "r02_desc_src.csv"), "30"))
write("r02_ncu_k_orient.txt", "# ncu --set full --import-source on, k_orient, 1080p workload, 32 images per launch",
      run(os.path.join(T, "ncu_summary.py"), os.path.join(G, "r02_orient_raw.csv")) +
      run(os.path.join(T, "ncu_source.py"), os.path.join(G, "r02_orient_src.csv"), "16"))
write("r02_ncu_pyramid_octave0.txt",
      "# ncu --set full: k_upsample2x, k_blur_march<0..5> and k_extrema_tma of octaves 0 and 1, 1080p workload, 32 images per launch",
      run(os.path.join(T, "ncu_summary.py"), os.path.join(G, "r02_pyr_raw.csv")))
write("r02_ncu_k_match_nn.txt", "# ncu --set full: k_match_nn, 8648 x 8648 descriptors (bench.py --workload match)",
      run(os.path.join(T, "ncu_kernel.py"), os.path.join(G, "r02_match_raw.csv"), "k_match_nn"))
write("r02_sass_census.txt", "", run(os.path.join(T, "sass_census.py")))
for f in ("r02_fine_1080p.txt", "r02_fine_vga.txt", "r02_fine_4k.txt"):
    shutil.copy(os.path.join(G, f), os.path.join(P, f))
for f in ("r02_bench_1080p.json", "r02_bench_reference_1080p.json"):
    lines = [l for l in open(os.path.join(G, f)) if l.startswith("{")]
    open(os.path.join(P, f), "w").write(lines[-1])

# per-launch DRAM bytes / instruction counts of the captured kernels (32 images per launch)
tp = os.path.join(P, "traffic.json")
tj = json.load(open(tp))


def rows_of(path):
    r = list(csv.reader(open(path)))
    return r[0], r[2:]


hdr, rows = rows_of(os.path.join(G, "r02_desc_raw.csv"))
r = rows[0]
val = lambda h: float(r[hdr.index(h)].replace(",", ""))
unit = lambda h: list(csv.reader(open(os.path.join(G, "r02_desc_raw.csv"))))[1][hdr.index(h)]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
dram = val("dram__bytes_read.sum") * scale[unit("dram__bytes_read.sum")] + val("dram__bytes_write.sum") * scale[unit("dram__bytes_write.sum")]
bench = json.loads(open(os.path.join(P, "r02_bench_1080p.json")).read())
kp_img = bench["keypoints_per_image"]
tj["k_descriptor_1080p"].update({
    "dram_bytes_per_image": int(dram / 32), "source": "profiles/r02_ncu_k_descriptor.txt",
    "warp_instructions_per_keypoint": round(val("smsp__inst_executed.sum") / (32 * kp_img), 1),
    "warp_instructions_note": "smsp__inst_executed.sum of the captured launch / (32 images x keypoints per image)"})
uh = list(csv.reader(open(os.path.join(G, "r02_pyr_raw.csv"))))
hdr, units, rows = uh[0], uh[1], uh[2:]
kn = hdr.index("Kernel Name")


def dram_of(row):
    tot = 0.0
    for h in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(row[hdr.index(h)].replace(",", "")) * scale[units[hdr.index(h)]]
    return tot


b5 = [r for r in rows if "k_blur_march<5" in r[kn] or "k_blur_march<(int)5" in r[kn]]
ex = [r for r in rows if "k_extrema_tma" in r[kn]]
if b5:
    tj["k_blur_march5_1080p_r02"] = {"dram_bytes_per_image": int(dram_of(b5[0]) / 32), "algorithmic_bytes_per_image": 66355200,
                                     "source": "profiles/r02_ncu_pyramid_octave0.txt"}
if ex:
    tj["k_extrema_tma_1080p_octave0_r02"] = {"dram_bytes_per_image": int(dram_of(ex[0]) / 32), "algorithmic_bytes_per_image": 199065600,
                                             "source": "profiles/r02_ncu_pyramid_octave0.txt"}
json.dump(tj, open(tp, "w"), indent=1)
print("traffic.json:", json.dumps(tj["k_descriptor_1080p"]))
