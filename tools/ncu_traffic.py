"""profiles/ncu_traffic.json (stage -> DRAM bytes per step, read + write, from one `ncu --set full` capture of
tools/ncu_step.py) out of a tools/ncu_summary.py summary.   python tools/ncu_traffic.py SUMMARY.json OUT.json"""
import json
import sys

STAGE = [("project_fwd", ("project_fwd_",)), ("project_bwd", ("project_bwd_kernel",)),
         ("blend_fwd", ("blend_fwd_kernel", "acc_fwd_kernel")), ("blend_bwd", ("blend_bwd_kernel", "acc_bwd_kernel")),
         ("adam", ("adam_kernel",)),
         ("bin_sort", ("depth_keys", "DeviceRadixSort", "DeviceScan", "inverse_perm", "write_total", "emit_keys", "bin_edges",
                       "class_count", "class_compact", "sched_kernel"))]


def to_bytes(s):
    v, u = s.split()[0], s.split()[1].lower()
    return float(v) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]


def main(src, dst):
    d = json.load(open(src))
    out = {k: 0.0 for k, _ in STAGE}
    for k in d["kernels"]:
        for stage, pats in STAGE:
            if any(p in k["kernel"] for p in pats):
                out[stage] += to_bytes(k["dram__bytes_read.sum"]) + to_bytes(k["dram__bytes_write.sum"])
                break
    out = {k: int(v) for k, v in out.items()}
    out["_source"] = src
    json.dump(out, open(dst, "w"), indent=1)
    print(out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
