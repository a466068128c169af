"""One process, many variants: bench.run_engine with the host fast paths / scheduling heuristics switched off one at a
time (Python setters + the C library's per-call environment switches).  Prints one line per variant."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from languagegroundedsemseg_b200 import minkowski as E

ENVS = ["LGS_TC_NO_BALANCE", "LGS_BN_SCALAR", "LGS_WGRAD_NO_BALANCE", "LGS_TC_TS1"]
VARIANTS = [
    ("all_on", {}, {}),
    ("all_off", {"fuse": 0, "overlap": 0, "batch": 0}, {"LGS_TC_NO_BALANCE": "1", "LGS_BN_SCALAR": "1", "LGS_WGRAD_NO_BALANCE": "1"}),
    ("no_fuse_conv_bn", {"fuse": 0}, {}),
    ("no_overlap", {"overlap": 0}, {}),
    ("no_batch_prep", {"batch": 0}, {}),
    ("no_tile_balance", {}, {"LGS_TC_NO_BALANCE": "1"}),
    ("bn_scalar", {}, {"LGS_BN_SCALAR": "1"}),
    ("no_wgrad_balance", {}, {"LGS_WGRAD_NO_BALANCE": "1"}),
    ("ts1", {}, {"LGS_TC_TS1": "1"}),
    ("all_on_again", {}, {}),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    args = argparse.Namespace(gpus=1, steps=20, warmup=5, impl="engine", algo="tc", dtype="f32", cpu_sample_voxels=0,
                              model=bench.MODEL, voxels=bench.TARGET_VOXELS, voxel_size=0.02, no_cpu_baseline=True,
                              no_prefetch=False, profile_run=False, step_program=False)
    for name, flags, env in VARIANTS:
        if a.only and name not in a.only.split(","):
            continue
        for k in ENVS:
            os.environ.pop(k, None)
        os.environ.update(env)
        E.set_conv_bn_fusion(flags.get("fuse", 1)), E.set_wgrad_overlap(flags.get("overlap", 1))
        E.set_batched_weight_prep(flags.get("batch", 1))
        try:
            r = bench.run_engine(args, 0, 1, 0)
            rf = r["roofline"]
            print(f"ABLATE {name:18s} ms/step {r['ms_per_step']:7.3f}  e2e {r['e2e']['ms_per_step']:7.3f}  loss {r['loss']:.5f}  "
                  f"dominant {rf['avg_launch_ms']:.4f} ms ({rf['frac']:.3f})  conv share {rf['all_conv_fwd_dgrad']['share_of_step']:.3f} "
                  f"wgrad share {rf['all_wgrad']['share_of_step']:.3f} launches {r['gpu_launches']}", flush=True)
            if name == "all_on":
                print("BENCHLINE " + json.dumps(r), flush=True)
        except Exception as e:  # keep going: one broken variant must not hide the others
            print(f"ABLATE {name:18s} FAILED: {type(e).__name__}: {e}", flush=True)


if __name__ == "__main__":
    main()
