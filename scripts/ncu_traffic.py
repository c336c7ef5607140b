"""profiles/ncu_traffic.json: measured DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum of one
`ncu --set full` capture inside the bench command) for the kernel classes bench.py reports, read back by bench.py for
`roofline.traffic`.  usage: python scripts/ncu_traffic.py  (after scripts/gpu_final.sh merged gpurun_out/)"""
import csv, json, pathlib, subprocess
ROOT = pathlib.Path(__file__).resolve().parents[1]
CLASSES = {"transform_bwd_axis0": "prof_xbwd", "pencil_solve": "prof_solve", "pointwise": "prof_pointwise",
           "pencil_matvec": "prof_matvec", "transform_bwd_axis2": "prof_zbwd", "transform_fwd_axis2": "prof_zfwd"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
out = {"workload": "rb3d 256^3 fp64 RK222, 1 GPU", "how": "ncu --set full --clock-control none, one launch inside bench.py", "classes": {}}
for cls, rep in CLASSES.items():
    path = ROOT / "gpurun_out" / f"{rep}.ncu-rep"
    if not path.exists():
        continue
    rows = list(csv.reader(subprocess.run(["ncu", "-i", str(path), "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    def get(name):
        i = hdr.index(name)
        return float(vals[i]) * UNIT[units[i]]
    out["classes"][cls] = dict(kernel=vals[hdr.index("Kernel Name")], dram_bytes_per_launch=get("dram__bytes_read.sum") + get("dram__bytes_write.sum"),
                               dram_read=get("dram__bytes_read.sum"), dram_write=get("dram__bytes_write.sum"), capture=f"profiles/r01_ncu_final_{rep[5:]}.txt")
json.dump(out, open(ROOT / "profiles" / "ncu_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
