"""Turn one `ncu --set full` capture of tools/frame_probe.py (fast mode, headline frame) into profiles/ncu_traffic.json:
DRAM bytes (read + written) per frame for every stage of the fast path, keyed by the digest of the kernel sources so
that bench.py only quotes it for the kernels that were profiled.

    python tools/ncu_traffic.py gpurun_out/frame.ncu-rep profiles/r2x_ncu_full_summary.txt
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                    # noqa: E402

STAGE_OF = (("sample_window", "percentile_sample"), ("window_count", "percentile_count"),
            ("hist_percentile", "percentile"), ("decimate", "decimate"), ("coarse_", "coarse"),
            ("interp_argmax", "interp_argmax"), ("band_project", "band"))


def main(rep, summary_name):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}

    def nbytes(r, key):
        v = float(r[idx[key]].replace(",", ""))
        unit = units[idx[key]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]

    per_stage, frames = {}, 0
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        for pat, stage in STAGE_OF:
            if pat in name:
                per_stage[stage] = per_stage.get(stage, 0.0) + nbytes(r, "dram__bytes_read.sum") + nbytes(r, "dram__bytes_write.sum")
                frames += stage == "decimate"
                break
    frames = max(frames, 1)
    out = {"capture": os.path.basename(summary_name), "frames_in_capture": frames,
           "source_digest": bench.kernel_source_digest(),
           "stage_bytes": {k: v / frames for k, v in sorted(per_stage.items())},
           "what": "dram__bytes_read.sum + dram__bytes_write.sum per frame, ncu --set full, 2048x2048x64 fast mode"}
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
