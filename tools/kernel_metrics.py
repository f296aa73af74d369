"""Join an `ncu --csv` metric dump of tools/forward_once.py with that run's launch labels and write the per-kernel-class table
bench.py reads for `roofline.traffic` (profiles/<tag>_kernel_metrics.json): per class the mean DRAM bytes per launch
(dram__bytes_read.sum + dram__bytes_write.sum), mean duration, tensor-pipe %, plus the sha256 of the library that was profiled
(bench.py drops the traffic figure when the library has changed since).
    python tools/kernel_metrics.py gpurun_out/r02_fwd_metrics.csv gpurun_out/r02_labels.json profiles/r02_kernel_metrics.json"""
import csv, hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def lib_hash():
    """Identity of the profiled kernels: hash of the kernel SOURCES (two nvcc builds of the same sources are not byte-identical)."""
    sys.path.insert(0, ROOT)
    from uformer_b200 import _lib
    return _lib.csrc_hash()


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(unit, 1)


def main(csv_path, labels_path, out_path):
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 5]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    iid, ikn, imn, imu, imv = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value")
    launches = {}
    for r in rows[hdr + 1:]:
        d = launches.setdefault(int(r[iid]), {"name": r[ikn]})
        d[r[imn]] = (r[imv], r[imu])
    ordered = [launches[k] for k in sorted(launches)]
    OURS = ("input_proj_kernel", "input_proj_tc_kernel", "output_proj_kernel", "output_proj_tc_kernel", "wmsa_kernel", "wmsa16_kernel", "wmsa_tma_kernel", "leff_fused_kernel", "ares_kernel", "leff2_kernel", "down_kernel",
            "charbonnier", "adamw_kernel")
    is_ours = lambda n: n.replace("void ", "").replace("lw::", "").startswith(OURS)          # noqa: E731
    ours = [d for d in ordered if is_ours(d["name"])]
    labels = json.load(open(labels_path))
    # charbonnier / adamw never run in a forward; labels and our kernels are both in launch order
    assert len(ours) == len(labels), (len(ours), len(labels))
    agg = {}
    for lab, d in zip(labels, ours):
        a = agg.setdefault(lab, dict(n=0, dram=0.0, us=0.0, tensor=0.0, kernel=d["name"][:90]))
        a["n"] += 1
        a["dram"] += to_bytes(*d["dram__bytes_read.sum"]) + to_bytes(*d["dram__bytes_write.sum"])
        a["us"] += to_us(*d["gpu__time_duration.sum"])
        a["tensor"] += float(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"][0])
    out = dict(source=f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active... "
                      f"--clock-control none over one batch-32 forward ({os.path.basename(csv_path)}); cold-cache, serialised launches",
               csrc_sha256_16=lib_hash(), others_aten=[d["name"][:80] for d in ordered if not is_ours(d["name"])],
               dram_bytes_per_launch={k: round(v["dram"] / v["n"]) for k, v in agg.items()},
               us_per_launch={k: round(v["us"] / v["n"], 2) for k, v in agg.items()},
               tensor_pipe_pct={k: round(v["tensor"] / v["n"], 2) for k, v in agg.items()},
               launches={k: v["n"] for k, v in agg.items()}, kernel={k: v["kernel"] for k, v in agg.items()})
    json.dump(out, open(out_path, "w"), indent=1)
    tot = sum(v["us"] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        print(f"{k:24s} n={v['n']:2d} {v['us']/v['n']:8.1f} us/launch  share {v['us']/tot*100:5.1f}%  dram {v['dram']/v['n']/1e6:8.1f} MB  tensor {v['tensor']/v['n']:5.1f}%")


if __name__ == "__main__":
    main(*sys.argv[1:4])
