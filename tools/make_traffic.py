#!/usr/bin/env python
"""profiles/r02_traffic.json from full Nsight Compute captures: DRAM bytes (read + written) per image and
launch of every captured kernel -- the `traffic` figure bench.py puts beside the algorithmic bytes.

    python tools/make_traffic.py IMAGES_PER_LAUNCH profiles/r02_traffic.json gpurun_out/r02z_*.ncu-rep

The captures come from `GG_SUBBATCH=1 ncu --set full ... python tools/kernel_times.py` (config B, every
launch covers the whole batch of IMAGES_PER_LAUNCH images); the text summary of each report
(tools/ncu_summary.py) is committed under profiles/ and named as the source."""
import csv
import json
import os
import subprocess
import sys


def main(images, out_path, reports):
    kernels = {}
    for rep in reports:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        i_name = hdr.index("Kernel Name")
        i_r, i_w = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows[2:]:
            name = r[i_name].split("(")[0].split("<")[0].replace("void ", "").replace("gg::", "").strip()
            b = float(r[i_r].replace(",", "")) * scale[units[i_r]] + float(r[i_w].replace(",", "")) * scale[units[i_w]]
            src = "r02_ncu_" + os.path.basename(rep).split("_", 1)[1].replace(".ncu-rep", ".txt")
            kernels.setdefault(name, {"bytes_per_image_per_launch": round(b / images, 1), "source": src})
    json.dump({"note": f"DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch from the committed "
                       f"`ncu --set full` captures of tools/kernel_times.py ({images} images per launch, 320x480, "
                       f"GG_SUBBATCH=1), divided by {images}: bytes per image per launch",
               "kernels": kernels}, open(out_path, "w"), indent=1)
    for k, v in kernels.items():
        print(f"{k:28s} {v['bytes_per_image_per_launch'] / 1e6:8.3f} MB/image  ({v['source']})")


if __name__ == "__main__":
    main(int(sys.argv[1]), sys.argv[2], sys.argv[3:])
