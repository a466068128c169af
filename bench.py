#!/usr/bin/env python
"""bench.py — active voxels/sec of Res16UNet34C forward+backward (BASELINE.json metric, config 2) on N x B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl engine|reference] [--algo tc|simt] [--dtype f32|bf16]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N

A step = one pass of the hot path over one batch: SparseTensor construction (cuckoo coordinate-map build),
lazy build of the 4 strided maps + 9 kernel maps, Res16UNet34C forward, CrossEntropy(ignore -1), backward,
SGD step.  One synthetic ScanNet-shaped scene (~150 K voxels @ 2 cm) per GPU; the only collective is DDP's NCCL
gradient all-reduce (weak scaling).
  value  = voxels of all ranks / step time, inputs resident in HBM.
  e2e    = same step through the reference-facing API with pinned HOST inputs (H2D inside) and loss.item() (D2H).
  roofline = conv fwd/dgrad kernel: SURVEY.md §8(d) per-offset HBM-gather bytes / CUDA-event time of its launches.
  cpu_baseline = the CPU oracle (restatement of MinkowskiEngine's CPU algorithm) on this box's host cores.
--impl reference times that CPU restatement alone (MinkowskiEngine itself is not installable here).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "active voxels/sec Res16UNet34C fwd+bwd"
UNIT = "voxels/s"
MODEL = "Res16UNet34C"
TARGET_VOXELS = 150_000


def workload_string(model, n_vox, voxel_size, voxels_target, config=2):
    tag = {2: " (BASELINE configs[1])", 3: " + CLIP text-anchor CE loss, 200 x 512 anchors, learned projection (BASELINE configs[2])",
           4: " (BASELINE configs[3]: bf16)", 5: " + CLIP text-anchor CE loss (BASELINE configs[4])"}.get(config, "")
    if config == 2 and not (model == MODEL and voxels_target == TARGET_VOXELS):
        tag = ""
    return f"{model} fwd+bwd+SGD, 1 synthetic ScanNet-shaped scene/GPU, {n_vox} voxels @{voxel_size * 100:g}cm, 200 classes" + tag


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1400.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled while the timed regions run: NVML in-process every 100 ms (nvidia-smi -lms 200 as
    the fallback); rows = [host time, sm MHz, max sm MHz, power, hw_slowdown, hw_thermal, sw_thermal, sw_power_cap]."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.window = None          # (t0, t1) host times of the timed region: only samples inside it are reported

    def _nvml_loop(self, nv, h):
        names = {0x8: 3, 0x40: 4, 0x20: 5, 0x4: 6}          # HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap -> row column
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        while not self._stop.is_set():
            try:
                row = [time.time(), str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(mx), "", "Not Active", "Not Active", "Not Active", "Not Active"]
                mask = get_reasons(h)
                for bit, col in names.items():
                    if mask & bit:
                        row[col + 1] = "Active"
                self.rows.append(row)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        # in-process NVML (nvidia_ml_py) first: one clock + one reasons query per 100 ms from a thread.  The nvidia-smi child
        # process it replaces stalled the CUDA launch path for 50-80 ms in one timed region out of three (one step of 87 ms
        # among 11 ms steps, profiles/r2_bench_stalls.txt); nvidia-smi stays as the fallback
        try:
            if self.index < 0:
                raise RuntimeError('disabled')
            import pynvml as nv
            nv.nvmlInit()
            h = None
            try:        # the CUDA device's own identity (NVML indices ignore CUDA_VISIBLE_DEVICES): UUID, then PCI bus id
                props = torch.cuda.get_device_properties(self.index)
                uuid = str(getattr(props, "uuid", "") or "")
                if uuid:
                    h = nv.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
            except Exception:
                h = None
            if h is None:
                try:
                    props = torch.cuda.get_device_properties(self.index)
                    bus = f"{props.pci_domain_id:08x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
                    h = nv.nvmlDeviceGetHandleByPciBusId(bus)
                except Exception:
                    h = None
            if h is None:
                h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self._stop = threading.Event()
            self.t = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.t.start()
            self.nvml = True
            return self
        except Exception:
            self.nvml = False
        try:
            if self.index < 0:
                raise RuntimeError('disabled')
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            import atexit
            atexit.register(self._kill)          # never leave the poller behind, whatever happens to the run
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [x.strip() for x in line.split(",")])

    def wait_first_sample(self, timeout=15.0):
        """block until the poller has printed its first row, i.e. its NVML start-up (1-3 s on an 8-GPU box) is over"""
        t0 = time.time()
        if getattr(self, "nvml", False):
            while not self.rows and time.time() - t0 < timeout:
                time.sleep(0.02)
            return
        while self.proc is not None and self.proc.poll() is None and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.05)

    def _kill(self):
        if self.proc and self.proc.poll() is None:
            self.proc.terminate()

    def __exit__(self, *a):
        if getattr(self, "nvml", False):
            self._stop.set()
            self.t.join(timeout=2)
            return
        if self.proc and self.proc.poll() is None:
            time.sleep(0.25)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self, window=None):
        window = window or self.window
        rows = [r[1:] for r in self.rows if window is None or window[0] <= r[0] <= window[1] + 0.25]
        if not rows:                # region shorter than one polling period: the nearest samples
            rows = [r[1:] for r in self.rows[-2:]]
        rows = [r for r in rows if len(r) >= 7]
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_scene(seed, target=TARGET_VOXELS, voxel_size=0.02):
    from languagegroundedsemseg_b200 import scenes
    return scenes.synthetic_voxel_scene(seed=seed, target_voxels=target, voxel_size=voxel_size)


def build_net(engine, device, dtype, model=None):
    from languagegroundedsemseg_b200 import nets
    torch.manual_seed(42)
    net = nets.build_model(model or MODEL, 3, 200, nets.DefaultConfig(), engine=engine).to(device).train()
    # stock torch SGD with the reference's hyper-parameters (lib/solvers.py:58-63); fused=True is torch's own single-
    # kernel implementation of the same update (CUDA only)
    kw = {"fused": True} if str(device).startswith("cuda") else {}
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, dampening=0.1, weight_decay=1e-4, **kw)
    return net, opt


def _aten_criterion(logits, labels):
    return torch.nn.functional.cross_entropy(logits.float(), labels, ignore_index=-1)


def train_step(ST, net, opt, coords, feats, labels, reducer=None, st=None, criterion=_aten_criterion, program=None, native=None, clip=None):
    if st is None:
        st = ST(feats, coords)                               # pl_BaselineTrainer.py:300
    if native is not None:
        # default driver: the whole forward + loss + backward is ONE C-ABI call (languagegroundedsemseg_b200/program.py);
        # gradients are written into the flat gradient buffer, the all-reduce buckets go out from inside run()
        loss = native.run(st, labels)
        opt.step()
        return loss
    if program is not None:
        # opt-in (--step-program): forward + loss + backward as one explicit program over the same entry points
        # (languagegroundedsemseg_b200/step.py), no autograd graph / module dispatch
        if reducer is not None:
            reducer.zero_grad()
        else:
            opt.zero_grad(set_to_none=True)
        with torch.no_grad():
            loss = program.run(st, labels, ignore_index=-1)
        if reducer is not None:
            reducer()
        opt.step()
        return loss
    if clip is not None:
        # pl_RepresentationTrainer.py:183-216: features (+ projected anchors) -> text-anchor loss
        crit, anchors = clip
        if hasattr(net, "projection_layer"):
            feat, anc = net(st, anchors)
        else:
            feat, anc = net(st), anchors
        loss = crit(feat.F, labels, anc)[0]
    else:
        out, _ = net(st)                                     # res16unet.py:196
        loss = criterion(out.F, labels)                      # :350  CrossEntropyLoss(ignore_index)
    if reducer is not None:
        reducer.zero_grad()                                  # gradients are views of the reducer's flat buffer: one memset
    else:
        opt.zero_grad(set_to_none=True)
    loss.backward()
    if reducer is not None:
        reducer()                                            # N > 1: the only collective (flat NCCL gradient all-reduce)
    opt.step()
    return loss


# ------------------------------------------------------------------------------------------------------------
# work model: per conv launch with P pairs, K offsets, element size s
# ------------------------------------------------------------------------------------------------------------
def launch_work(meta, pair_cache):
    kind, K, c_in, c_out, n_in, n_out, km, dtype = meta
    s = 2 if dtype == torch.bfloat16 else 4
    if km is None:
        P = n_out
        idx = 0
    else:
        if id(km) not in pair_cache:
            pair_cache[id(km)] = int(km.counts.sum().item())
        P = pair_cache[id(km)]
        idx = 8 * P
    flops = 2.0 * P * c_in * c_out
    if kind == "wgrad":
        byts = P * (c_in + c_out) * s + idx + K * c_in * c_out * 4
    else:
        byts = P * (c_in + c_out) * s + idx + K * c_in * c_out * s
    return flops, byts, P


def roofline_report(prof, ms_per_step, algo, dtype, PROF_STEPS, pair_counts=None):
    """`roofline` entry of the JSON line from the per-launch records [(meta, (start event, end event))] of PROF_STEPS
    profiled steps (meta = kind, K, c_in, c_out, n_in, n_out, kernel map, feature dtype)."""
    # Launches are grouped by (kernel kind, K, c_in, c_out, n_out); the group with the largest summed duration is "the
    # dominant kernel".  achieved = SURVEY §8(d) gather-model bytes of ONE launch / its average CUDA-event duration.
    hbm_gbs, tf_peak, peak_src = load_peaks()
    pair_cache = {} if pair_counts is None else pair_counts
    groups, agg = {}, {}
    for meta, (a, b) in prof:
        fl, by, _ = launch_work(meta, pair_cache)
        dt_ms = a.elapsed_time(b)
        kname = "wgrad" if meta[0] == "wgrad" else "conv"
        d = agg.setdefault(kname, [0.0, 0.0, 0.0, 0])
        d[0] += dt_ms; d[1] += by; d[2] += fl; d[3] += 1
        g = groups.setdefault((kname,) + tuple(meta[1:6]), [0.0, by, fl, 0, meta[7]])
        g[0] += dt_ms; g[3] += 1
    if os.environ.get("LGS_BENCH_LAYERS"):
        for key, g in sorted(groups.items(), key=lambda kv: -kv[1][0])[:40]:
            print(f"LAYER {key[0]:6s} K={key[1]:2d} {key[2]:4d}->{key[3]:4d} n_out={key[5]:7d} launches/step={g[3] // PROF_STEPS:3d} "
                  f"ms/step={g[0] / PROF_STEPS:7.3f}", file=sys.stderr)
    step_flops = sum(v[2] for v in agg.values()) / PROF_STEPS
    step_bytes = sum(v[1] for v in agg.values()) / PROF_STEPS
    conv_ms, conv_bytes, conv_flops, conv_n = agg.get("conv", [1e-9, 0, 0, 1])
    wg = agg.get("wgrad", [0, 0, 0, 0])
    dom_key, dom = max(((k, g) for k, g in groups.items() if k[0] == "conv"), key=lambda kg: kg[1][0])
    dom_ms = dom[0] / dom[3]
    achieved = dom[1] / (dom_ms * 1e-3) / 1e9
    # DRAM traffic of that launch from the committed `ncu --set full` capture (profiles/ncu_traffic.json), if present
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    shape_tag = f"conv K={dom_key[1]} {dom_key[2]}->{dom_key[3]} n_out={dom_key[5]} {algo} {dtype}"
    traffic_src = tensor_pct = None
    if os.path.exists(tpath):
        t = json.load(open(tpath)).get(f"K={dom_key[1]} {dom_key[2]}->{dom_key[3]} {algo} {dtype}")
        if t:
            traffic, traffic_src = t["bytes"], t["source"]      # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch
            tensor_pct = t.get("tensor_pipe_pct")
    kname = {"bx3": ("conv_nb_kernel (neighbourhood cache in shared memory -> bf16x3 tcgen05 GEMM in TS form, fwd and dgrad): "
                     if dom_key[1] == 27 else
                     "conv_bx3_kernel (output-stationary gather -> bf16x3 tcgen05 GEMM in TS form, fwd and dgrad): "),
             "simt": "conv_simt_kernel: "}.get(algo, "conv_tc2_kernel (output-stationary gather -> tcgen05 GEMM, fwd and dgrad): ")
    # tensor work the kernel issues for this launch: dense over all K offsets of every 128-row tile, x3 for bf16x3
    issued = 2.0 * dom_key[1] * dom_key[5] * dom_key[2] * dom_key[3] * (3 if algo in ("bx3", "tc") else 1)
    roofline = {"bound": "hbm", "kernel": kname + shape_tag,
                "achieved": round(achieved, 1), "peak": hbm_gbs, "unit": "GB/s", "frac": round(achieved / hbm_gbs, 4),
                "traffic": traffic, "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                "traffic_source": traffic_src, "peak_source": peak_src,
                "tensor_pipe_active_pct": tensor_pct,   # sm__pipe_tensor_cycles_active of the same ncu capture: what the SM actually did
                "tensor_issued_tflops": round(issued / (dom_ms * 1e-3) / 1e12, 1), "tensor_peak_tflops": tf_peak,
                "note": "achieved/peak is the SURVEY 8(d) per-offset gather model (bytes an ME-style gather -> GEMM -> scatter would move) over "
                        "the measured HBM copy bandwidth; it exceeds 1 where the kernel does not move those bytes at all: it is "
                        "output-stationary and serves the 27-offset gather from a shared-memory cache of each supertile's unique rows "
                        "(see traffic: DRAM bytes are 20x below the model).  What binds it is the tensor pipe (bf16x3 = 3 MMAs per product, "
                        "dense over the 27 offsets: tensor_issued_tflops against tensor_peak_tflops, the measured cuBLAS bf16 rate) and "
                        "the shared-memory data pipe; the step-level fraction of the model floor is step_model.frac_of_floor",
                "algorithmic_bytes_per_launch": int(dom[1]), "avg_launch_ms": round(dom_ms, 4), "launches_per_step": dom[3] // PROF_STEPS,
                "share_of_step": round((dom[0] / PROF_STEPS) / ms_per_step, 3),
                "tflops": round(dom[2] / (dom_ms * 1e-3) / 1e12, 2),
                "timed_over": f"{PROF_STEPS} extra steps right after the timed region, CUDA events around every launch",
                "all_conv_fwd_dgrad": {"share_of_step": round((conv_ms / PROF_STEPS) / ms_per_step, 3), "launches_per_step": conv_n // PROF_STEPS,
                                       "achieved_gbs": round(conv_bytes / max(conv_ms, 1e-9) / 1e6, 1),
                                       "tflops": round(conv_flops / max(conv_ms, 1e-9) / 1e9, 2)},
                "all_wgrad": {"share_of_step": round((wg[0] / PROF_STEPS) / ms_per_step, 3),
                              "achieved_gbs": round(wg[1] / max(wg[0], 1e-9) / 1e6, 1), "tflops": round(wg[2] / max(wg[0], 1e-9) / 1e9, 2)},
                "step_model": {"gflop": round(step_flops / 1e9, 1), "gbyte": round(step_bytes / 1e9, 2),
                               "hbm_floor_ms": round(step_bytes / hbm_gbs / 1e6, 3),
                               "frac_of_floor": round((step_bytes / hbm_gbs / 1e6) / ms_per_step, 3)}}
    return roofline


def run_engine(args, rank, world, local_rank):
    from languagegroundedsemseg_b200 import _lib, minkowski as E
    from languagegroundedsemseg_b200.csrc import build as _build
    if rank == 0:
        _build.build()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    _lib.load()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if os.environ.get("LGS_MAIN_PRIORITY", "0") != "0":
        # experiment: training stream = a high-priority stream, so its (critical-path) kernels get SMs ahead of the wgrad side stream
        torch.cuda.set_stream(torch.cuda.Stream(dev, priority=-1))
    # ONE nvidia-smi poller (rank 0 only: concurrent pollers slow the driver), started now — seconds before the first timed
    # region — because its NVML start-up stalls CUDA calls for a few hundred ms; it then polls every 200 ms through both
    # timed regions and each region reports the samples that fall inside it.
    clk = ClockSampler(local_rank if (rank == 0 and os.environ.get("LGS_BENCH_NO_CLOCKS", "0") == "0") else -1)   # knob: diagnose sampler-induced stalls
    clk.__enter__()
    E.set_conv_algo(args.algo)
    fdtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32

    from languagegroundedsemseg_b200.ddp import shard_scenes
    coords_np, feats_np, labels_np = make_scene(seed=shard_scenes(world, world, rank)[0], target=args.voxels,
                                                voxel_size=args.voxel_size)   # one scene per rank
    if args.permute_rows:
        # random voxel order (what RandomDropout / a raw PLY order gives) instead of the generator's surface-by-surface order
        perm = np.random.RandomState(1234 + rank).permutation(coords_np.shape[0])
        coords_np, feats_np, labels_np = coords_np[perm].copy(), feats_np[perm].copy(), labels_np[perm].copy()
    n_vox = coords_np.shape[0]
    # host (pinned) copies for the e2e leg; device-resident copies for `value`
    h_coords = torch.from_numpy(coords_np).pin_memory()
    h_feats = torch.from_numpy(feats_np).pin_memory()
    h_labels = torch.from_numpy(labels_np).pin_memory()
    d_coords, d_feats, d_labels = h_coords.to(dev), h_feats.to(dev).to(fdtype), h_labels.to(dev)

    from languagegroundedsemseg_b200 import losses as lgs_losses
    net, opt = build_net(None, dev, fdtype, args.model)
    if args.dtype == "bf16":
        # bf16 features; parameters stay fp32 (master weights), BN in fp32 statistics via autocast-free mixed dtype
        pass
    from languagegroundedsemseg_b200 import ddp
    model = net
    # same seed on every rank => same init; the native driver launches the buckets itself, the facade path through hooks
    reducer = ddp.GradAllReducer(net.parameters(), overlap=args.driver != "native") if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # Every step stages the NEXT step's batch (copies + coordinate/kernel maps) on a side stream while it runs
    # (languagegroundedsemseg_b200/prefetch.py); the map build is still done once per step, inside the timed region.
    from languagegroundedsemseg_b200.prefetch import SparseBatchPrefetcher
    pf = (SparseBatchPrefetcher(dev, fdtype, threaded=os.environ.get("LGS_STAGE_THREAD", "0") != "0",
                                high_priority=os.environ.get("LGS_STAGE_PRIORITY", "0") != "0")
          if not args.no_prefetch else None)
    tickets = {}
    if pf is not None and os.environ.get("LGS_BENCH_NO_PRERESERVE", "0") == "0":
        # The staging stream's allocations (coordinate maps, tables, plans of the NEXT batch) are released only after the
        # training stream's recorded use of them: when that event is still pending the caching allocator finds no cached
        # block and falls through to cudaMalloc, which waits for the GPU — the 50-100 ms `stage` call behind the occasional
        # long step (profiles/r2_bench_stage_stall.txt).  One large block cached in the staging stream's pool up front
        # gives the allocator something to split instead.
        # Both pools (large blocks and the <= 1 MB small-block pool) of both streams that allocate during a step.
        for strm in (pf.stream, torch.cuda.current_stream(dev)):
            with torch.cuda.stream(strm):
                spare = [torch.empty(1 << 30, dtype=torch.uint8, device=dev)] + \
                        [torch.empty(512 << 10, dtype=torch.uint8, device=dev) for _ in range(256)]
            del spare

    # the engine's fused softmax cross-entropy (lgs_seg_ce: one pass over the logits) unless LGS_ATEN_CE=1
    crit = _aten_criterion if os.environ.get("LGS_ATEN_CE") else (lambda x, y: lgs_losses.cross_entropy(x, y, ignore_index=-1))

    program = native = None
    if args.step_program:
        from languagegroundedsemseg_b200.step import StepProgram
        program = StepProgram(model)
    clip = None
    if args.clip:
        net.representation_only(True)                       # models/clip_models.py:106-109
        torch.manual_seed(1)
        anchors = torch.nn.functional.normalize(torch.randn(200, 512, device=dev), dim=1)     # CLIP text features stand-in
        clip = (lgs_losses.ContrastiveLanguageCELoss(num_labels=200, ignore_label=-1), anchors)
    if not args.step_program and args.driver == "native" and args.dtype == "f32" and args.algo == "bx3":
        from languagegroundedsemseg_b200.program import NativeStep
        if reducer is None:
            reducer = ddp.GradAllReducer(net.parameters(), overlap=False)
        head = None
        if clip is not None:
            crit_c, anc_c = clip
            proj = getattr(net, "projection_layer", None)

            def head(feats, labels):
                a = proj(anc_c.unsqueeze(-1)).squeeze() if proj is not None else anc_c
                return crit_c(feats, labels, a)[0]
        native = NativeStep(model, ignore_index=-1, reducer=reducer, head=head)
    native_step = native

    prof_mode = {"on": False}     # per-launch CUDA events need the facade's hooks: the profiled steps run module by module

    phases = []              # host ms per step: (get, stage next batch, step + optimiser) — diagnostics of host-side stalls

    def staged_step(key, src):
        flush.fill_(0.0)
        native = None if prof_mode["on"] else native_step
        if pf is None:
            c, f, lab = (t.to(dev, non_blocking=True) for t in src)
            return train_step(E.SparseTensor, model, opt, c, f.to(fdtype), lab, reducer, criterion=crit, program=program, native=native, clip=clip)
        if key not in tickets:
            tickets[key] = pf.stage(*src)
        t0 = time.perf_counter()
        st, lab = pf.get(tickets[key])
        t1 = time.perf_counter()
        tickets[key] = pf.stage(*src)                        # next step's batch, overlapped with this step
        t2 = time.perf_counter()
        out = train_step(None, model, opt, None, None, lab, reducer, st=st, criterion=crit, program=program, native=native, clip=clip)
        phases.append((1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (time.perf_counter() - t2)))   # get, stage next, run + optimiser
        return out

    def resident_step():
        return staged_step("resident", (d_coords, d_feats, d_labels))

    # e2e: the loss of every step is copied device->host into pinned memory inside the timed region; its VALUE is consumed
    # one step later (after that copy's event), so the training stream is never drained by a blocking .item()
    loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_evt = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_state = {"i": 0, "last": None}

    def e2e_step():
        loss = staged_step("e2e", (h_coords, h_feats, h_labels))
        i = e2e_state["i"]
        loss_host[i & 1].copy_(loss.detach().float(), non_blocking=True)
        loss_evt[i & 1].record()
        if i > 0:
            loss_evt[(i - 1) & 1].synchronize()
            e2e_state["last"] = float(loss_host[(i - 1) & 1])
        e2e_state["i"] = i + 1

    def e2e_drain():
        i = e2e_state["i"]
        if i > 0:
            loss_evt[(i - 1) & 1].synchronize()
            e2e_state["last"] = float(loss_host[(i - 1) & 1])

    # W untimed warm-up steps, never fewer than 10: the caching allocator's per-stream pools (training stream + staging
    # stream, blocks handed over with record_stream) take several steps to stop growing, and a cudaMalloc inside the
    # timed region costs milliseconds (seen as a 19 ms outlier in one of four runs with 5 warm-up steps)
    warm_done = args.warmup if args.profile_run else max(args.warmup, 10)   # --profile-run: launch lists under ncu
    clk.wait_first_sample()             # the poller's start-up must not overlap a timed region
    for _ in range(warm_done):
        resident_step()
    barrier()
    if not args.profile_run:
        # settle: further UNTIMED batches of 5 steps — at least 8, until four consecutive batches agree within 3 % (max over
        # ranks), at most 60.  (The single 30-100 ms step seen in about one first timed region in ten — never in the second
        # one, which starts ~45 steps into the process — points at warm-up that is still going on: allocator pools of the
        # training / staging / side streams whose reuse depends on event timing.)  A fresh box pages the image / driver in for the first seconds (one run in three showed a 0.6 s host stall
        # inside the first timed region after 10 warm-up steps: 40 ms/step instead of 11); the total goes into "warmup_done" ("warmup" echoes the requested count).
        prev, stable = None, 0
        for it in range(60):
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(5):
                resident_step()
            s1.record()
            torch.cuda.synchronize()
            warm_done += 5
            cur = torch.tensor([s0.elapsed_time(s1)], device=dev)
            if world > 1:
                import torch.distributed as dist
                dist.all_reduce(cur, op=dist.ReduceOp.MAX)
            cur = float(cur.item())
            stable = stable + 1 if (prev is not None and abs(cur - prev) <= 0.03 * min(cur, prev)) else 0
            prev = cur
            if it >= 7 and stable >= 3:     # at least 40 further steps, the last four batches within 3 % of their neighbours
                break
        barrier()

    # ---- timed region: `value` --------------------------------------------------------------------------
    # the cyclic garbage collector stays off inside the timed regions (collected right before): a generation-2 pass over the
    # process's objects is one candidate for the single 50-100 ms host stall seen in about one value leg in fifteen
    import gc
    gc.collect()
    gc.disable()
    l0 = _lib.launch_count()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]      # per-step diagnostics, created outside the region
    for m in marks:
        m.record()                                                                 # cudaEventCreate happens at the first record
    torch.cuda.synchronize()
    t_w0 = time.time()
    host_t = [time.perf_counter()]
    ev0.record()
    del phases[:]
    segs_before = torch.cuda.memory_stats(dev).get("segment.all.allocated", 0)       # cudaMalloc calls so far (caching allocator)
    for i in range(args.steps):
        loss = resident_step()
        marks[i].record()
        host_t.append(time.perf_counter())
    value_phases = list(phases)
    segs_after = torch.cuda.memory_stats(dev).get("segment.all.allocated", 0)
    ev1.record()
    barrier()
    clocks_value = clk.summary((t_w0, time.time()))
    ms = ev0.elapsed_time(ev1)
    per_step = [a.elapsed_time(b) for a, b in zip([ev0] + marks[:-1], marks)]      # diagnostics only: the value is ms / steps
    host_ms = [1e3 * (b - a) for a, b in zip(host_t[:-1], host_t[1:])]             # host time of each step's issue (no sync)
    i_max = max(range(len(per_step)), key=per_step.__getitem__)
    launches = _lib.launch_count() - l0
    loss_val = float(loss.item())
    if args.profile_run:
        clk.__exit__()
        if rank == 0:
            emit({"profile_run": True, "ms_per_step": ms / args.steps, "gpu_launches": int(launches)})
        return None

    # ---- e2e leg (right after the value leg, GPU warm; its own >= 3 warm-up steps; clocks sampled too) -----------
    for _ in range(max(3, args.warmup)):
        e2e_step()
    e2e_drain()
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t_e0 = time.time()
    t0.record()
    for _ in range(args.steps):
        e2e_step()
    e2e_drain()                     # the last step's loss is read inside the timed region too
    t1.record()
    barrier()
    clocks_e2e = clk.summary((t_e0, time.time()))
    gc.enable()
    clk.__exit__()
    ms_e2e = t0.elapsed_time(t1)

    # ---- per-launch CUDA-event timing of the engine's conv kernels (same process, same data, right after the timed
    # region; kept out of it because ~380 event records per step starve the launch queue and double the step time)
    PROF_STEPS = 2
    prof_mode["on"] = True
    resident_step()                      # the facade's lazily built state (weight-operand cache) outside the profile
    E.profile_begin()
    for _ in range(PROF_STEPS):
        resident_step()
    torch.cuda.synchronize()
    prof = E.profile_end() or []
    prof_mode["on"] = False

    # ---- kernel-map build time (coordinate map + 4 strided maps + 9 kernel maps of one scene) -------------
    def build_maps():
        st = E.SparseTensor(d_feats[:, :1], d_coords)
        m, k = st.coordinate_manager, st.coordinate_map_key
        for lvl in range(5):
            m.kernel_map(k, k, [3, 3, 3], [1, 1, 1])
            if lvl < 4:
                k2 = m.stride(k, 2)
                m.kernel_map(k, k2, [2, 2, 2], [1, 1, 1])
                k = k2
        return m
    build_maps()
    torch.cuda.synchronize()
    m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m0.record()
    for _ in range(5):
        build_maps()
    m1.record()
    torch.cuda.synchronize()
    kmap_ms = m0.elapsed_time(m1) / 5

    # ---- reduce over ranks ---------------------------------------------------------------------------------
    tot_vox = n_vox
    per_rank = None
    if world > 1:
        import torch.distributed as dist
        mine = torch.tensor([n_vox, ms / args.steps, ms_e2e / args.steps], device=dev, dtype=torch.float64)
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)                    # each rank's scene size and own step time: slowest-rank gating made visible
        per_rank = {"voxels": [int(e[0].item()) for e in every], "ms_per_step": [round(e[1].item(), 3) for e in every],
                    "e2e_ms_per_step": [round(e[2].item(), 3) for e in every]}
        t = torch.tensor([ms, ms_e2e, kmap_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, kmap_ms = t.tolist()
        v = torch.tensor([n_vox], device=dev, dtype=torch.float64)
        dist.all_reduce(v)
        tot_vox = int(v.item())
    if rank != 0:
        return None

    # ---- roofline of the dominant kernel (roofline_report above; a failure there must not lose the measured line) ------
    try:
        roofline = roofline_report(prof, ms / args.steps, args.algo, args.dtype, PROF_STEPS)
    except Exception as e:  # noqa: BLE001
        roofline = {"error": f"{type(e).__name__}: {e}"}

    value = tot_vox * args.steps / (ms * 1e-3)
    e2e_value = tot_vox * args.steps / (ms_e2e * 1e-3)
    h2d = h_coords.numel() * 4 + h_feats.numel() * 4 + h_labels.numel() * 8
    res = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "warmup_done": warm_done, "ms_per_step": round(ms / args.steps, 3),
        "step_ms": {"median": round(statistics.median(per_step), 3), "min": round(min(per_step), 3), "max": round(max(per_step), 3),
                    "argmax": i_max, "host_issue_ms_median": round(statistics.median(host_ms), 3),
                    "host_issue_ms_around_max": [round(h, 2) for h in host_ms[max(0, i_max - 2): i_max + 2]],
                    "host_phases_ms_at_max(get,stage_next,run+opt)": [round(x, 2) for x in value_phases[i_max]] if i_max < len(value_phases) else None,
                    "cudaMalloc_segments_during_region": int(segs_after - segs_before),
                    "host_phases_ms_median": [round(statistics.median(p[j] for p in value_phases), 2) for j in range(3)] if value_phases else None},
        **({"per_rank": per_rank} if per_rank else {}), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if args.dtype == "f32" else "bf16", "data": "synthetic",
        "config": {"workload": workload_string(args.model, n_vox, args.voxel_size, args.voxels, args.config), "voxels_per_gpu": n_vox, "algo": args.algo,
                   "binding": _lib.binding() + " (Python -> C ABI)",
                   "driver": ("StepProgram (explicit program, no autograd)" if args.step_program else
                              "native step driver (one lgs_program_run per step, languagegroundedsemseg_b200/program.py)" if native is not None
                              else "MinkowskiEngine facade + autograd"),
                   "math": {"bx3": "tcgen05 bf16x3 error-compensated products (2^-16) fwd/dgrad, TF32 wgrad, fp32 accumulate in TMEM",
                            "tc": "tcgen05 3xTF32 products (2^-21) fwd/dgrad, TF32 wgrad, fp32 accumulate in TMEM",
                            "tf32": "tcgen05 single-pass TF32, fp32 accumulate", "simt": "fp32 FMA"}[args.algo]
                   if args.dtype == "f32" else "tcgen05 bf16 products, fp32 accumulate in TMEM",
                   "l2": "256 MB buffer written between steps (L2 flush); per-step activations >> 126 MB L2",
                   **({"row_order": "random permutation of the scene's voxels"} if args.permute_rows else {}),
                   "parallelism": f"dp{world}" + (" (one scene per rank, one flat NCCL gradient all-reduce per step)" if world > 1 else ""),
                   "step": "coordinate+kernel maps" + ("" if args.no_prefetch else " (staged on a side stream during the previous step)")
                           + (", fwd, CLIP text-anchor CE loss (fused tcgen05 kernel), bwd, SGD" if args.clip else ", fwd, CE loss (fused lgs_seg_ce), bwd, SGD")},
        "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": round(ms_e2e / args.steps, 3), "last_loss": e2e_state["last"], "clocks": clocks_e2e,
                "how": "pinned host coords/feats/labels -> H2D every step (staged on a side stream one step ahead), "
                       "SparseTensor + fwd + loss + bwd + SGD through the facade, loss -> pinned host every step "
                       "(value consumed one step later)"},
        "gpu_launches": int(launches),
        "kernel_map_build_ms": round(kmap_ms, 3),
        "roofline": roofline,
        "clocks": clocks_value,
        "loss": round(loss_val, 5),
    }
    return res


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (port of MinkowskiEngine's CPU algorithm) on the host cores
# ------------------------------------------------------------------------------------------------------------
def cpu_arm(steps, warmup, sample_voxels, seed=0):
    from oracle import me_cpu
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    coords, feats, labels = make_scene(seed=seed, target=sample_voxels)
    net, opt = build_net(me_cpu, "cpu", torch.float32)
    c, f, lab = torch.from_numpy(coords), torch.from_numpy(feats), torch.from_numpy(labels)
    for _ in range(warmup):
        train_step(me_cpu.SparseTensor, net, opt, c, f, lab)
    t0 = time.perf_counter()
    for _ in range(steps):
        train_step(me_cpu.SparseTensor, net, opt, c, f, lab)
    dt = time.perf_counter() - t0
    n = coords.shape[0]
    return {"value": round(n * steps / dt, 1), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{MODEL} fwd+bwd+SGD on a {n}-voxel scene from the same generator, {steps} step(s) after "
                      f"{warmup} warm-up; ME-CPU-algorithm restatement (torch index_select->mm->index_add_, "
                      f"{cores} threads)",
            "seconds": round(dt, 2)}, n, dt


def cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


_JSON_FD = None


def claim_stdout():
    """stdout carries exactly ONE line, the JSON result: keep a private duplicate of fd 1 for it and point fd 1 at stderr,
    so that nothing else in the process (NCCL's version banner, library warnings written with printf) lands on stdout"""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--algo", default="bx3", choices=["bx3", "tc", "tf32", "simt"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--cpu-sample-voxels", type=int, default=TARGET_VOXELS,
                    help="scene size of the CPU arm (default: the full BASELINE configs[1] scene, ~4 s per step on 16 cores)")
    ap.add_argument("--model", default=MODEL, help="topology (default: the BASELINE metric's Res16UNet34C)")
    ap.add_argument("--voxels", type=int, default=TARGET_VOXELS)
    ap.add_argument("--voxel-size", type=float, default=0.02)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-prefetch", action="store_true", help="build the coordinate manager on the training stream")
    ap.add_argument("--permute-rows", action="store_true", help="random voxel order instead of the scene generator's surface order")
    ap.add_argument("--driver", default="native", choices=["native", "facade"],
                    help="native: forward + loss + backward as one lgs_program_run per step (default; fp32 / bx3); facade: the "
                         "reference's module-by-module path through the MinkowskiEngine facade and autograd")
    ap.add_argument("--step-program", action="store_true",
                    help="forward + loss + backward through languagegroundedsemseg_b200.step.StepProgram (explicit program over "
                         "the same C-ABI calls, no autograd graph) instead of the module-by-module facade")
    ap.add_argument("--profile-run", action="store_true",
                    help="for ncu launch lists: allow fewer warm-up steps and skip the e2e / map-build / roofline legs")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="BASELINE.json configs[] (1-based): 2 Res16UNet34C fwd+bwd ~150K voxels (default, the metric's config); "
                         "3 Res16UNet34CR_Proj + CLIP text-anchor loss (200 x 512 anchors, learned projection); 4 config 2 in bf16 "
                         "(launch with torchrun for DDP); 5 Res16UNet34D @1cm ~600K voxels + CLIP loss")
    args = ap.parse_args()
    args.clip = False
    if args.config == 3:
        args.model, args.clip = "Res16UNet34CR_Proj", True
    elif args.config == 4:
        args.dtype, args.driver = "bf16", "facade"
    elif args.config == 5:
        args.model, args.clip, args.voxels, args.voxel_size = "Res16UNet34D", True, 600_000, 0.01
    args.warmup = max(args.warmup, 3) if (args.impl == "engine" and not args.profile_run) else args.warmup

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    claim_stdout()

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = max(1, min(args.steps, 2))
        cb, n, dt = cpu_arm(steps, min(args.warmup, 1), args.cpu_sample_voxels)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": round(dt / steps * 1e3, 1),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                # same workload as the engine arm: the full scene of BASELINE configs[1], fewer steps (cpu_baseline.sample)
                "config": {"workload": workload_string(MODEL, make_scene(0, TARGET_VOXELS)[0].shape[0], 0.02, TARGET_VOXELS),
                           "sample_voxels": n,
                           "implementation": "CPU restatement of MinkowskiEngine 0.5.4's CPU algorithm (oracle/me_cpu.py; the "
                                             "reference's own ME is not installable here), all host threads",
                           "cpu": cpu_model_name()},
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    if world > 1:
        from languagegroundedsemseg_b200 import ddp
        ddp.init_process_group("nccl")
    res = run_engine(args, rank, world, local_rank)
    if rank == 0 and res is not None:
        if world == 1 and not args.no_cpu_baseline:
            cb, _, _ = cpu_arm(1, 0, args.cpu_sample_voxels)
            cb["cpu"] = cpu_model_name()
            res["cpu_baseline"] = cb
        emit(res)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
