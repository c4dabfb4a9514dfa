#!/usr/bin/env python
"""Headline benchmark: images/sec of the CAM -> SVC -> PAR hot path (BASELINE.json: configs[1],
"ViT-B/16 CAM+SVC+PAR, synthetic VOC 512x512 batch=16, 1xB200"), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                         # the reference's CPU path (oracle port)

A step = one pass of the whole path (CLIP-surgery ViT-B/16 forward -> patch x text CAM -> SVC -> PAR 20 it ->
pseudo labels) over one batch of 16 synthetic 512^2 images per GPU (weak scaling: the batch shards by image,
SURVEY.md §8e).  `value` is timed with the inputs resident in HBM; `e2e` runs the same call with HOST (pinned)
inputs and a device->host read of the labels inside the timed region.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec CAM+SVC+PAR @512^2 batch"
WORKLOAD = "ViT-B/16 CAM+SVC+PAR, synthetic VOC 512x512 batch=16, 1xB200"
SIZE, BATCH, NUM_FG, T_BANK = 512, 16, 20, 45
N_PRESENT = None   # classes per image drawn from the empirical VOC distribution (mean 1.55, max 6)
PAR_ITERS, DIL = 20, (1, 2, 4, 8, 12, 24)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"   # /opt/skills/guides/B200_PROFILING.md


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (profiling recipe's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def synthetic_batch(seed):
    from excel_b200 import synth
    return (synth.images(BATCH, SIZE, seed=seed), synth.class_labels(BATCH, NUM_FG, seed=seed + 100, n_fixed=N_PRESENT))


def cpu_port_step(W, text, imgs, cls):
    """One pass of the reference's CPU path (oracle/port.py) over `imgs`."""
    from oracle import port
    return port.hot_path(W, text, imgs, cls, NUM_FG, num_iter=PAR_ITERS, use_cv2=True)


def time_cpu(n_images, reps=1, warm=0):
    """images/sec of the oracle port on the host cores (all threads), on n_images of the bench workload."""
    from excel_b200 import synth
    torch.set_num_threads(os.cpu_count())
    W = synth.random_visual_weights(seed=0)
    text = synth.text_bank(T_BANK, 512, seed=1)
    imgs, cls = synthetic_batch(10)
    imgs, cls = imgs[:n_images], cls[:n_images]
    with torch.no_grad():
        for _ in range(warm):
            cpu_port_step(W, text, imgs[:1], cls[:1])
        t0 = time.perf_counter()
        for _ in range(reps):
            cpu_port_step(W, text, imgs, cls)
        dt = (time.perf_counter() - t0) / reps
    return n_images / dt, dt


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (the oracle port; the reference is pure
    Python and cannot travel to the GPU box), all host threads, a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = 2
    W_, K_ = max(args.warmup, 0), max(args.steps, 1)
    K_ = min(K_, 3)  # ~5 s per image per step on 8 cores: keep the run within minutes
    ips, dt = time_cpu(n, reps=K_, warm=1 if W_ else 0)
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus, "steps": K_,
            "warmup": 1 if W_ else 0, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": f"{n} of the {BATCH} images per step"},
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{n} images x {K_} steps of the bench workload, torch {torch.__version__} CPU"},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_cuda(args):
    import torch.distributed as dist
    from excel_b200 import _lib, synth, evaluate
    from excel_b200.encoder import SurgeryViT
    from excel_b200.pipeline import ExCELHotPath, HostPipeline
    from excel_b200.par import par_refine_planes, par_affinity

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    arch = _lib.lib().excel_device_arch(local)
    if arch != 100 and not os.environ.get("EXCEL_ALLOW_ANY_ARCH"):
        raise RuntimeError(f"bench: expected sm_100 (B200), found sm_{arch}")

    hbm_peak, tf_peak, tf_sust, peak_kind = peaks()
    # --graph: the encoder's launches are captured once and replayed as a CUDA graph (its outputs are consumed by the SVC/PAR
    # stages of the same step before the next forward overwrites them).  Measured: no gain on this pool -- the step runs into
    # the board power cap, not into launch gaps -- so the default launches kernel by kernel.
    enc = SurgeryViT(synth.random_visual_weights(seed=0), device=dev, graph=args.graph)
    hp = ExCELHotPath(enc, synth.text_bank(T_BANK, 512, seed=1), NUM_FG)
    # 3 rotating input batches (151 MB > the 126 MB L2) + ~1.4 GB of per-step intermediates: no L2 carry-over
    host = [synthetic_batch(10 + 3 * rank + i) for i in range(3)]
    host = [(i.pin_memory(), c.pin_memory()) for i, c in host]
    devb = [(i.to(dev), c.to(dev)) for i, c in host]
    gt = [torch.randint(0, NUM_FG + 1, (BATCH, SIZE // 32, SIZE // 32), device=dev).repeat_interleave(32, 1).repeat_interleave(32, 2)
          for _ in range(3)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(i):
        # images resident in HBM; the image-level labels are host logic input (which planes exist) and stay on the host
        return hp(devb[i % 3][0], host[i % 3][1])

    # e2e: pinned host images -> H2D -> hot path -> D2H of the int64 labels into pinned memory, through the public
    # HostPipeline API (copies of neighbouring batches overlap the kernels; every byte moves inside the timed region)
    pipe = HostPipeline(hp)

    def step_e2e(i):
        if pipe.staged is None:
            pipe.stage(*host[i % 3])
        return pipe.submit(stage_next=host[(i + 1) % 3])

    def timed(fn, steps, warmup, sample_clocks=False, finish=None):
        for i in range(warmup):
            fn(i)
        barrier()
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
        l0 = _lib.lib().excel_launch_count() + enc.replayed_launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            out = fn(i)
        if finish is not None:
            out = finish()          # e.g. wait for the last step's D2H copy: it belongs to the timed region
        e1.record()
        barrier()
        launches = _lib.lib().excel_launch_count() + enc.replayed_launches - l0
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), launches, (sampler.stop() if sampler else None), out

    # clocks / throttle reasons are sampled on rank 0's GPU only (one nvidia-smi poller per job, not per rank)
    ms, launches, clocks, labels = timed(step_resident, args.steps, args.warmup, sample_clocks=(rank == 0))
    value = world * BATCH * args.steps / (ms / 1e3)
    ms_e, _, _, labels_e = timed(step_e2e, args.steps, max(args.warmup, 3), finish=pipe.flush)
    assert labels_e is not None and labels_e.shape == (BATCH, SIZE, SIZE) and not labels_e.is_cuda
    e2e = world * BATCH * args.steps / (ms_e / 1e3)

    # the path's single collective: confusion histogram of the last step's labels, summed over ranks
    hist = evaluate.confusion_hist(gt[(args.steps - 1) % 3], labels, NUM_FG + 1)
    if (labels_e.to(dev) != labels).any().item():   # both arms ended on batch (steps-1) % 3
        raise RuntimeError("bench: e2e labels differ from the resident-input labels of the same batch")
    evaluate.all_reduce_hist(hist)

    # ---- roofline of the kernel the metric names: the PAR propagation step (HBM-bound), on the plane counts of the
    # bench workload (images sorted by plane count, one launch per count class -- what refine_batch does)
    from excel_b200.affutils import _segments
    imgs = devb[0][0]

    def t_par(planes, off, segs, iters):
        f = (lambda: par_refine_planes(imgs, planes, off, max(m for _, _, m in segs), DIL, iters, group=0, segments=segs)) if iters \
            else (lambda: [par_affinity(imgs[b0:b1], (SIZE, SIZE), DIL) for b0, b1, _ in segs])
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            f()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / 5

    def par_rate(counts):
        counts = sorted(counts)
        offs = [0]
        for c in counts:
            offs.append(offs[-1] + c)
        planes = torch.softmax(torch.randn(offs[-1], SIZE, SIZE, device=dev), 0)
        off = torch.tensor(offs, dtype=torch.int32, device=dev)
        segs = _segments(counts)
        per_step_ms = (t_par(planes, off, segs, PAR_ITERS) - t_par(planes, off, segs, 0)) / PAR_ITERS
        alg = sum(4.0 * SIZE * SIZE * (48 + 2 * c) for c in counts)          # 4*(K + 2C) B/pixel/step
        return alg / (per_step_ms * 1e-3) / 1e9, per_step_ms, len(segs)
    counts = [int(c.sum().item()) + 1 for c in host[0][1]]
    achieved, per_step_ms, nseg = par_rate(counts)
    by_planes = {str(c): round(par_rate([c] * BATCH)[0], 1) for c in (2, 3, 4)}
    traffic = None   # DRAM bytes of the same step from the committed ncu capture (profiles/), per step like `achieved`
    tp = os.path.join(ROOT, "profiles", "r01g_par_traffic.json")
    if os.path.exists(tp) and sorted(counts) == [2] * 8 + [3] * 7 + [4]:
        traffic = json.load(open(tp))["dram_bytes_per_step"]
    roofline = {"kernel": "par_iterate_kernel", "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "peak_kind": peak_kind,
                "algorithmic_bytes": sum(4.0 * SIZE * SIZE * (48 + 2 * c) for c in counts),
                "note": f"one propagation step over the {BATCH} images of a bench batch (planes/image {sorted(counts)}): "
                        f"{per_step_ms*1e3:.1f} us in {nseg} launches; uniform-C batches GB/s: {by_planes}; "
                        "traffic = ncu dram bytes of the same three launches (profiles/r01g_par_traffic.json)"}

    line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "images_per_gpu_per_step": BATCH,
                       "classes_per_image": "empirical VOC distribution (mean 1.55, max 6), seeded",
                       "par_iters": PAR_ITERS, "text_bank_rows": T_BANK, "weights": "seeded random-init ViT-B/16", "encoder_cuda_graph": args.graph,
                       "l2": "3 rotating input batches (151 MB > L2) + >1 GB of per-step intermediates"},
            "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": BATCH * 3 * SIZE * SIZE * 4 + BATCH * NUM_FG * 4,
                    "d2h_bytes_per_step": BATCH * SIZE * SIZE * 8, "ms_per_step": ms_e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "hist_pixels_all_ranks": int(hist.sum().item())}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            ips, dt = time_cpu(2, reps=1, warm=0)                  # ~10-20 s of CPU work
            line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"2 of the {BATCH} images of one step, oracle/port.py, torch {torch.__version__} CPU"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", action="store_true", help="replay the encoder as a CUDA graph instead of launching kernel by kernel")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun (the driver launches torchrun itself)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_cuda(args)


if __name__ == "__main__":
    main()
