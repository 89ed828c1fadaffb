#!/usr/bin/env python
"""profiles/r02_traffic.json from the ncu metrics pass of one step (scripts/gpu_profile_r02.sh, part A): per stage, DRAM
bytes and warp instructions per unit of work, which bench.py scales to the launch it measures.
usage: make_traffic.py launches.csv <raw_rows> <target_rows> <nn_queries> [last_step_only=1]"""
import collections
import csv
import json
import sys

path = sys.argv[1]
raw_rows, target_rows, nn_queries = float(sys.argv[2]), float(sys.argv[3]), float(sys.argv[4])
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    per.setdefault(int(r[idi]), {"name": r[ki].split("(")[0].replace("void ", "").split("<")[0]})[r[mi]] = v
launches = list(per.values())
# the LAST step only: everything after the last k_synth_* launch (input generation precedes the warm-up steps too, so
# take the launches after the final k_icp_init)
last_init = max(i for i, l in enumerate(launches) if l["name"] == "k_icp_init")
first = max(i for i, l in enumerate(launches[:last_init]) if l["name"].startswith("k_vox_clear"))
step = launches[first:]
stage_of = lambda n: ("voxel" if n.startswith(("k_vox", "k_voxel", "k_seg_sort", "k_sort", "k_scan", "k_mark", "k_copy_words", "k_widen"))
                      else "index_build" if n in ("k_bbox", "k_morton", "k_tree_bounds", "k_gather_leaves", "k_boxes_up")
                      else "normals" if n in ("k_self_knn", "k_knn_redo", "k_normals_from_graph", "k_knn", "k_grid_params", "k_grid_build", "k_tree_attach")
                      else "icp_loop" if n.startswith("k_icp") else "other")
agg = collections.defaultdict(lambda: collections.defaultdict(float))
kern = collections.defaultdict(lambda: collections.defaultdict(float))
seen_index = False
for l in step:
    st = stage_of(l["name"])
    # the sorts after the first k_bbox belong to the index build
    if l["name"] == "k_bbox":
        seen_index = True
    if st == "voxel" and seen_index and l["name"].startswith(("k_seg_sort", "k_sort", "k_scan", "k_copy_words")):
        st = "index_build"
    for m in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum"):
        agg[st][m] += l.get(m, 0.0)
        kern[l["name"]][m] += l.get(m, 0.0)
    kern[l["name"]]["n"] += 1
units = {"voxel": ("raw rows", raw_rows), "index_build": ("indexed rows", target_rows), "normals": ("target rows", target_rows / 2),
         "icp_loop": ("nearest-neighbour queries", nn_queries)}
out = {}
for st, (uname, u) in units.items():
    a = agg[st]
    out[st] = {"unit": uname, "units_in_capture": u, "dram_bytes_per_unit": (a["dram__bytes_read.sum"] + a["dram__bytes_write.sum"]) / u,
               "warp_inst_per_unit": a["smsp__inst_executed.sum"] / u, "kernel_time_us_under_ncu": a["gpu__time_duration.sum"] / 1e3,
               "source": "profiles/r02_ncu_launches_c5_1024pairs.csv (ncu metrics pass, cold caches, serialised launches)"}
json.dump(out, open("profiles/r02_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
print("per kernel (one step):")
for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    print(f"{k:26s} n={int(v['n']):4d} time {v['gpu__time_duration.sum'] / 1e3:10.1f} us  dram {(v['dram__bytes_read.sum'] + v['dram__bytes_write.sum']) / 1e6:9.1f} MB  warp-inst {v['smsp__inst_executed.sum'] / 1e6:9.1f} M")
