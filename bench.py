#!/usr/bin/env python
"""Headline benchmark: images/sec of the CAM -> SVC -> PAR hot path (BASELINE.json: configs[1],
"ViT-B/16 CAM+SVC+PAR, synthetic VOC 512x512 batch=16, 1xB200"), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path, batched public API
    python bench.py --surface dropin [...]                         # the reference's per-image loop through install()
    python bench.py --config cfg3|cfg4|cfg5 [...]                  # the other BASELINE.json configs, same JSON contract
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


def vit_flops(n_tok, n_patch, patch, D, L, n_sur, E):
    """fp32-equivalent FLOPs of one image through the surgery ViT."""
    return (2.0 * n_patch * 3 * patch * patch * D + (L - n_sur) * (24.0 * n_tok * D * D + 4.0 * n_tok * n_tok * D)
            + n_sur * (26.0 * n_tok * D * D + 12.0 * n_tok * n_tok * D) + 2.0 * n_tok * D * E)


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
    """images/sec of the oracle port on the host cores (all threads), on n_images of the bench workload.
    Also returns the outputs of the last pass (the parity record checks the GPU step against them)."""
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
            out = cpu_port_step(W, text, imgs, cls)
        dt = (time.perf_counter() - t0) / reps
    return n_images / dt, dt, out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (the oracle port; the reference is pure
    Python and cannot travel to the GPU box), all host threads, a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = 2
    W_, K_ = max(args.warmup, 0), max(args.steps, 1)
    K_ = min(K_, 3)  # ~5 s per image per step on 8 cores: keep the run within minutes
    ips, dt, _ = time_cpu(n, reps=K_, warm=1 if W_ else 0)
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus, "steps": K_,
            "warmup": 1 if W_ else 0, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": f"{n} of the {BATCH} images per step"},
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{n} images x {K_} steps of the bench workload, torch {torch.__version__} CPU"},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


class Job:
    """Process-per-GPU plumbing shared by every bench mode: rank / device / NCCL, barrier, max-over-ranks timing."""

    def __init__(self):
        import torch.distributed as dist
        from excel_b200 import _lib
        self.dist, self.lib = dist, _lib.lib()
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        arch = self.lib.excel_device_arch(self.local)
        if arch != 100 and not os.environ.get("EXCEL_ALLOW_ANY_ARCH"):
            raise RuntimeError(f"bench: expected sm_100 (B200), found sm_{arch}")
        self.extra_launches = lambda: 0   # kernels replayed from CUDA graphs (the library counter sees host launches only)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps, warmup, sample_clocks=False, finish=None):
        """W untimed warm-up steps, then exactly `steps` steps between barrier + synchronize, CUDA events, max over ranks."""
        for i in range(warmup):
            fn(i)
        self.barrier()
        sampler = ClockSampler(self.local) if sample_clocks else None
        if sampler:
            sampler.start()
        l0 = self.lib.excel_launch_count() + self.extra_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for i in range(steps):
            out = fn(i)
        if finish is not None:
            out = finish()          # e.g. wait for the last step's D2H copy: it belongs to the timed region
        e1.record()
        self.barrier()
        launches = self.lib.excel_launch_count() + self.extra_launches() - l0
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return ms.item(), launches, (sampler.stop() if sampler else None), out

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def event_ms(fn, warm=3, rep=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(rep):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / rep


def par_roofline(dev, imgs, counts, size, hbm_peak, peak_kind, traffic_file=None):
    """Roofline of the kernel the metric names: the PAR propagation step (HBM-bound) on the plane counts `counts`
    (images sorted by plane count, one launch per count class on forked streams -- what refine_batch does).  Measured
    live: (t(20 steps) - t(affinity only)) / 20 against the algorithmic bytes 4*H*W*(48 + 2C) per image and step."""
    from excel_b200.affutils import _segments
    from excel_b200.par import par_affinity, par_refine_planes
    B = len(counts)

    def t_par(planes, off, segs, iters):
        f = (lambda: par_refine_planes(imgs[:B], planes, off, max(m for _, _, m in segs), DIL, iters, group=0, segments=segs)) if iters \
            else (lambda: [par_affinity(imgs[b0:b1], (size, size), DIL) for b0, b1, _ in segs])
        return event_ms(f)

    def par_rate(cnt):
        cnt = sorted(cnt)
        offs = [0]
        for c in cnt:
            offs.append(offs[-1] + c)
        planes = torch.softmax(torch.randn(offs[-1], size, size, device=dev), 0)
        off = torch.tensor(offs, dtype=torch.int32, device=dev)
        segs = _segments(cnt)
        per_step_ms = (t_par(planes, off, segs, PAR_ITERS) - t_par(planes, off, segs, 0)) / PAR_ITERS
        alg = sum(4.0 * size * size * (48 + 2 * c) for c in cnt)          # 4*(K + 2C) B/pixel/step
        return alg / (per_step_ms * 1e-3) / 1e9, per_step_ms, len(segs)
    achieved, per_step_ms, nseg = par_rate(counts)
    by_planes = {str(c): round(par_rate([c] * B)[0], 1) for c in (2, 3, 4)}
    traffic = None   # DRAM bytes of the same step from the committed ncu capture (profiles/), per step like `achieved`
    if traffic_file and os.path.exists(traffic_file):
        t = json.load(open(traffic_file))
        if sorted(counts) == t.get("planes_per_image", [2] * 8 + [3] * 7 + [4]):
            traffic = t["dram_bytes_per_step"]
    return {"kernel": "par_iterate_kernel", "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
            "frac": achieved / hbm_peak, "traffic": traffic, "peak_kind": peak_kind,
            "algorithmic_bytes": sum(4.0 * size * size * (48 + 2 * c) for c in counts),
            "note": f"one propagation step over {B} images (planes/image {sorted(counts)}): "
                    f"{per_step_ms*1e3:.1f} us in {nseg} launches; uniform-C batches GB/s: {by_planes}; "
                    "traffic = ncu dram bytes of the same launches (profiles/)"}


def parity_record(hp, dev, imgs_dev, cls_host, ref):
    """GPU step vs the oracle outputs `ref` of the cpu_baseline leg (same images), outside every timed region.
    Gate: oracle/parity.py (oracle top-2 margin 1e-5)."""
    from excel_b200 import affutils
    from oracle.parity import MARGIN, label_parity
    n = len(ref["labels"])
    attr, attn, _ = hp.cams(imgs_dev)
    lab, planes, off, _ = affutils.refine_batch(attr, attn, cls_host, imgs_dev, hp.par, return_cams=True)
    lab_iso, planes_iso, off_iso, _ = affutils.refine_batch(ref["attr_maps_raw"].to(dev), ref["attn_weights"].to(dev), cls_host[:n],
                                                           imgs_dev[:n], hp.par, return_cams=True)
    off, off_iso = off.cpu().tolist(), off_iso.cpu().tolist()
    rec = {"images": n, "pixels": n * SIZE * SIZE, "margin": MARGIN,
           "cam_max_abs": (attr[:n].cpu() - ref["attr_maps_raw"]).abs().max().item(),
           "attn_max_abs": (attn[:, :n].cpu() - ref["attn_weights"]).abs().max().item()}
    e2e = dict(plane_max_abs=0.0, label_mismatch_px=0, hard_mismatch_px=0, hard_mismatch_px_strict=0)
    iso = dict(plane_max_abs=0.0, label_mismatch_px=0, hard_mismatch_px=0, hard_mismatch_px_strict=0)
    for b in range(n):
        for d, L, P, O in ((e2e, lab, planes, off), (iso, lab_iso, planes_iso, off_iso)):
            err = (P[O[b]:O[b + 1]].cpu() - ref["cams"][b]).abs().max().item()
            hard, total = label_parity(ref["refined"][b], ref["labels"][b][0], L[b].cpu(), plane_err=err)
            strict, _ = label_parity(ref["refined"][b], ref["labels"][b][0], L[b].cpu())
            d["plane_max_abs"] = max(d["plane_max_abs"], err)
            d["label_mismatch_px"] += total
            d["hard_mismatch_px"] += hard
            d["hard_mismatch_px_strict"] += strict
    rec.update(e2e)
    rec["tail_on_oracle_cams"] = iso
    from oracle.parity import svc_self_noise
    rec["oracle_fp32_vs_fp64_plane_max_abs"] = max(svc_self_noise(ref["attr_maps_raw"][b], ref["attn_weights"][:, b], cls_host[b], (SIZE, SIZE))
                                                   for b in range(n))
    rec["note"] = ("label_mismatch_px: GPU vs oracle labels; hard = oracle top-2 margin > 1e-5 + what the measured PAR-input "
                   "difference (plane_max_abs) can move an output (oracle/parity.py); _strict = margin 1e-5 alone; "
                   "tail_on_oracle_cams = GPU SVC+PAR+argmax fed with the oracle's CAMs and attention; "
                   "oracle_fp32_vs_fp64_plane_max_abs = the reference's OWN fp32 rounding noise on these PAR input planes (random-init "
                   "weights give nearly uniform attention, the per-class min-max of utils/affutils.py:69-78 amplifies it): the floor for plane_max_abs")
    return rec


def run_cuda(args):
    from excel_b200 import synth, evaluate
    from excel_b200.encoder import SurgeryViT
    from excel_b200.pipeline import ExCELHotPath, HostPipeline

    job = Job()
    dev, rank, world = job.dev, job.rank, job.world
    hbm_peak, tf_peak, tf_sust, peak_kind = peaks()
    # --graph: the encoder's launches are captured once and replayed as a CUDA graph (its outputs are consumed by the SVC/PAR
    # stages of the same step before the next forward overwrites them).
    enc = SurgeryViT(synth.random_visual_weights(seed=0), device=dev, graph=args.graph)
    job.extra_launches = lambda: enc.replayed_launches
    hp = ExCELHotPath(enc, synth.text_bank(T_BANK, 512, seed=1), NUM_FG)
    # 3 rotating input batches (151 MB > the 126 MB L2) + ~1.4 GB of per-step intermediates: no L2 carry-over
    host = [synthetic_batch(10 + 3 * rank + i) for i in range(3)]
    host = [(i.pin_memory(), c.pin_memory()) for i, c in host]
    devb = [(i.to(dev), c.to(dev)) for i, c in host]
    gt = [torch.randint(0, NUM_FG + 1, (BATCH, SIZE // 32, SIZE // 32), device=dev).repeat_interleave(32, 1).repeat_interleave(32, 2)
          for _ in range(3)]

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

    # clocks / throttle reasons are sampled on rank 0's GPU only (one nvidia-smi poller per job, not per rank)
    ms, launches, clocks, labels = job.timed(step_resident, args.steps, args.warmup, sample_clocks=(rank == 0))
    value = world * BATCH * args.steps / (ms / 1e3)
    ms_e, _, _, labels_e = job.timed(step_e2e, args.steps, max(args.warmup, 3), finish=pipe.flush)
    assert labels_e is not None and labels_e.shape == (BATCH, SIZE, SIZE) and not labels_e.is_cuda
    e2e = world * BATCH * args.steps / (ms_e / 1e3)

    # the path's single collective: confusion histogram of the last step's labels, summed over ranks
    hist = evaluate.confusion_hist(gt[(args.steps - 1) % 3], labels, NUM_FG + 1)
    if (labels_e.to(dev) != labels).any().item():   # both arms ended on batch (steps-1) % 3
        raise RuntimeError("bench: e2e labels differ from the resident-input labels of the same batch")
    evaluate.all_reduce_hist(hist)

    counts = [int(c.sum().item()) + 1 for c in host[0][1]]
    roofline = par_roofline(dev, devb[0][0], counts, SIZE, hbm_peak, peak_kind, os.path.join(ROOT, "profiles", "par_traffic.json"))
    # the 75 % of the step that is tensor-bound: encoder forward alone, fp32-equivalent FLOPs x 3 split-fp16 MMA passes
    enc_ms = event_ms(lambda: enc(devb[0][0]), warm=2, rep=5)
    fl = BATCH * vit_flops(1025, 1024, 16, 768, 12, 5, 512)
    enc_roof = {"kernel": "surgery ViT-B/16 forward (gemm_tc + attn_tc + attn_pv)", "bound": "tensor", "unit": "TFLOP/s",
                "achieved": 3 * fl / (enc_ms * 1e-3) / 1e12, "peak": tf_sust, "frac": 3 * fl / (enc_ms * 1e-3) / 1e12 / tf_sust,
                "peak_kind": peak_kind + " (sustained bf16/fp16 dense)", "ms": enc_ms, "fp32_equiv_tflops": fl / (enc_ms * 1e-3) / 1e12,
                "note": "tensor work = 3 fp16 MMA passes per fp32-quality product (hi*hi + hi*lo + lo*hi); SURVEY.md §8d FLOP count"}

    line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "surface": "batched (excel_b200.pipeline.ExCELHotPath / HostPipeline)",
                       "images_per_gpu_per_step": BATCH,
                       "classes_per_image": "empirical VOC distribution (mean 1.55, max 6), seeded",
                       "par_iters": PAR_ITERS, "text_bank_rows": T_BANK, "weights": "seeded random-init ViT-B/16", "encoder_cuda_graph": args.graph,
                       "l2": "3 rotating input batches (151 MB > L2) + >1 GB of per-step intermediates"},
            "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": BATCH * 3 * SIZE * SIZE * 4 + BATCH * NUM_FG * 4,
                    "d2h_bytes_per_step": BATCH * SIZE * SIZE * 8, "ms_per_step": ms_e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_encoder": enc_roof,
            "hist_pixels_all_ranks": int(hist.sum().item())}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            ips, dt, ref = time_cpu(2, reps=1, warm=0)                  # ~10-20 s of CPU work
            line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"2 of the {BATCH} images of one step, oracle/port.py, torch {torch.__version__} CPU"}
            line["parity"] = parity_record(hp, dev, devb[0][0], host[0][1], ref)   # rank 0's batch 0 is seed 10: the CPU leg's images
        print(json.dumps(line), flush=True)
    job.close()


# ---------------------------------------------------------------------------------------------------------------------
# --surface dropin: the reference's own per-image loop (tools/infer_lam.py:70-94, engine/validatation_engine.py:18-38)
# through the install()-patched symbols, on the stand-in module tree tests/dropin_tree (the reference is not on the box).

def run_dropin(args):
    from excel_b200 import synth, install as inst
    job = Job()
    dev, rank, world = job.dev, job.rank, job.world
    sys.path.insert(0, os.path.join(ROOT, "tests", "dropin_tree"))
    originals = inst.install(graph=args.graph)
    # --- from here on the code reads like the reference script: module paths, names and call order are the reference's
    from model.model_excel import ExCEL_model
    from utils.affutils import refine_cams_with_aff, refine_cams_with_bkg_weclip
    from utils.PAR import PAR
    model = ExCEL_model(synth.random_visual_weights(seed=0), synth.text_bank(T_BANK, 512, seed=1).t().contiguous(), NUM_FG + 1)
    model.to(dev).eval()
    par = PAR(num_iter=PAR_ITERS, dilations=list(DIL)).to(dev)
    results = {}
    for size in (SIZE, 320):
        imgs_h, cls_h = synthetic_batch(10 + 3 * rank)
        if size != SIZE:
            imgs_h = torch.nn.functional.interpolate(imgs_h, size=[size, size], mode="bilinear", align_corners=False)
        imgs_h, cls_h = imgs_h.pin_memory(), cls_h.pin_memory()

        def step(_i):
            out = None
            with torch.no_grad():
                for k in range(BATCH):                                           # batch_size = 1 data loader
                    inputs = imgs_h[k:k + 1].to(dev, non_blocking=True)          # :75
                    cls_labels = cls_h[k:k + 1].to(dev, non_blocking=True)       # :77
                    _, ex_feats, attr_maps_raw, attn_weights, attn_pred = model(inputs)          # :79
                    for i, attr_map in enumerate(attr_maps_raw):                 # :88
                        refined, cls_lst = refine_cams_with_aff(attr_map, attn_weights[:, i, ...], cls_labels[i],
                                                                size=inputs.shape[2:], seg_attn=None, caa_thre=0.79)   # :93
                        labels, normed = refine_cams_with_bkg_weclip(refined, inputs[i], cls_lst, par, inputs.shape[-2:])  # :94
                    out = labels.cpu().numpy()                                   # :113 (blocking D2H per image)
            return out
        ms, launches, clocks, last = job.timed(step, args.steps, args.warmup, sample_clocks=(rank == 0 and size == SIZE))
        results[size] = dict(value=world * BATCH * args.steps / (ms / 1e3), ms_per_step=ms / args.steps, launches=int(launches),
                             clocks=clocks)
        assert last.shape == (1, size, size)
    # f4 (decoder-side inference) timed on the batched call the seg-eval scripts make: model(inputs) for 16 images of 512^2
    from excel_b200 import decoder
    from excel_b200.encoder import generate_clip_fts
    imgs16 = synthetic_batch(10 + 3 * rank)[0].to(dev)
    with torch.no_grad():
        ms_model = event_ms(lambda: model(imgs16), warm=2, rep=3)
        _, _, feats16 = generate_clip_fts(imgs16, model.encoder)
        Lf, Bf, Nf, Df = feats16.shape
        ms_head = event_ms(lambda: decoder.segformer_head_tokens(model.decoder_fts_fuse, feats16.reshape(Lf, Bf * Nf, Df)), warm=2, rep=3)
        fts = torch.randn(Bf, 256, SIZE // 16, SIZE // 16, device=dev)
        ms_pred = event_ms(lambda: decoder.attn_pred(fts), warm=2, rep=3)
    f4 = {"model_forward_ms": ms_model, "images": BATCH, "size": SIZE, "segformer_head_ms": ms_head, "attn_pred_ms": ms_pred,
          "segformer_head_tflops_fp32_equiv": 2.0 * Bf * Nf * (12 * (Df * 256 + 256 * 256) + 12 * 256 * 256) / (ms_head * 1e-3) / 1e12,
          "note": "ExCEL_model.forward at inference through install(): encoder + CAM + SegFormerHead (4 launches: split, 2 batched GEMMs, "
                  "fuse GEMM) + attn_pred; the stand-in decoder of tests/dropin_tree is a 1x1 conv"}
    inst.uninstall(originals)
    r = results[SIZE]
    line = {"metric": METRIC, "value": r["value"], "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "surface": "dropin: the reference's per-image loop (tools/infer_lam.py:70-94) through "
                       "excel_b200.install() on tests/dropin_tree -- batch 1, model(inputs) incl. the decoder head, "
                       "refine_cams_with_aff, refine_cams_with_bkg_weclip, labels.cpu() per image",
                       "images_per_gpu_per_step": BATCH, "par_iters": PAR_ITERS, "encoder_cuda_graph": args.graph},
            "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": BATCH * (3 * SIZE * SIZE * 4 + NUM_FG * 4),
                    "d2h_bytes_per_step": BATCH * SIZE * SIZE * 8, "ms_per_step": r["ms_per_step"],
                    "note": "this surface is end to end by construction: host images in, host labels out, per image"},
            "gpu_launches": r["launches"], "clocks": r["clocks"],
            "dropin_320": {"value": results[320]["value"], "unit": "images/s", "ms_per_16_images": results[320]["ms_per_step"]},
            "f4": f4}
    if rank == 0:
        print(json.dumps(line), flush=True)
    job.close()


# ---------------------------------------------------------------------------------------------------------------------
# the other BASELINE.json configs (same JSON contract; the headline stays cfg2)

def run_cfg3(args):
    """configs[2]: ViT-B/16 CAM+SVC+PAR, synthetic COCO 448^2, batch 64 sharded over the ranks (strong scaling)."""
    from excel_b200 import synth
    from excel_b200.encoder import SurgeryViT
    from excel_b200.pipeline import ExCELHotPath, HostPipeline
    job = Job()
    dev, rank, world = job.dev, job.rank, job.world
    G, S, K, T = 64, 448, 80, 103
    from excel_b200.evaluate import shard_slice
    mine = shard_slice(G, rank, world)          # rank r owns a contiguous 64/W slice of each global batch
    per = mine.stop - mine.start
    hp = ExCELHotPath(SurgeryViT(synth.random_visual_weights(seed=0), device=dev), synth.text_bank(T, 512, seed=1), K)
    host = []
    for i in range(3):
        im = synth.images(G, S, seed=20 + i)[mine].contiguous().pin_memory()
        cl = synth.class_labels(G, K, seed=120 + i, n_fixed=None, dataset="ms_coco")[mine].contiguous()
        host.append((im, cl))
    devb = [(i.to(dev), c) for i, c in host]
    ms, launches, clocks, _ = job.timed(lambda i: hp(devb[i % 3][0], devb[i % 3][1]), args.steps, args.warmup, sample_clocks=(rank == 0))
    pipe = HostPipeline(hp)

    def step_e2e(i):
        if pipe.staged is None:
            pipe.stage(*host[i % 3])
        return pipe.submit(stage_next=host[(i + 1) % 3])
    ms_e, _, _, lab = job.timed(step_e2e, args.steps, max(args.warmup, 3), finish=pipe.flush)
    assert lab.shape == (per, S, S)
    hbm_peak, _, _, peak_kind = peaks()
    counts = [int(c.sum().item()) + 1 for c in host[0][1]]
    line = {"metric": "images/sec CAM+SVC+PAR @448^2 batch", "value": G * args.steps / (ms / 1e3), "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ViT-B/16 CAM+SVC+PAR, synthetic COCO 448x448 batch=64, sharded 8xB200", "global_batch": G,
                       "images_per_gpu_per_step": per, "text_bank_rows": T, "classes": K,
                       "classes_per_image": "empirical COCO distribution (mean 2.84, max 18), seeded", "par_iters": PAR_ITERS,
                       "l2": "3 rotating input batches + >1 GB of per-step intermediates"},
            "e2e": {"value": G * args.steps / (ms_e / 1e3), "unit": "images/s", "h2d_bytes_per_step": per * 3 * S * S * 4,
                    "d2h_bytes_per_step": per * S * S * 8, "ms_per_step": ms_e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": par_roofline(dev, devb[0][0], [min(c, 4) for c in counts], S, hbm_peak, peak_kind)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    job.close()


def run_cfg4(args):
    """configs[3]: ViT-L/14@336 dense attention + 103-row text bank, batch 8 (encoder + CAM only: SURVEY.md §8d caveat)."""
    from excel_b200 import synth
    from excel_b200.clip import clip_feature_surgery
    from excel_b200.encoder import SurgeryViT, generate_clip_fts
    job = Job()
    dev, rank, world = job.dev, job.rank, job.world
    B, S = 8, 336
    enc = SurgeryViT(synth.random_visual_weights(layers=24, width=1024, patch=14, grid0=24, embed=768, seed=4), device=dev)
    text = synth.text_bank(103, 768, seed=5).to(dev)
    host = [synth.images(B, S, seed=30 + 3 * rank + i).pin_memory() for i in range(3)]
    devb = [h.to(dev) for h in host]

    def step(i, src=devb):
        tok, attn, feats = generate_clip_fts(src[i % 3], enc)
        return clip_feature_surgery(tok, text)[:, 1:, :80]
    ms, launches, clocks, _ = job.timed(step, args.steps, args.warmup, sample_clocks=(rank == 0))
    out_h = torch.empty((B, 576, 80), dtype=torch.float32).pin_memory()

    def step_e2e(i):
        out_h.copy_(step(0, [host[i % 3].to(dev, non_blocking=True)] * 3), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out_h
    ms_e, _, _, _ = job.timed(step_e2e, args.steps, max(args.warmup, 3))
    _, _, tf_sust, peak_kind = peaks()
    fl = B * vit_flops(577, 576, 14, 1024, 24, 5, 768)
    ach = 3 * fl / (ms / args.steps * 1e-3) / 1e12
    line = {"metric": "images/sec ViT-L/14@336 dense forward + CAM", "value": world * B * args.steps / (ms / 1e3), "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ViT-L/14@336 dense attention + 81-class text bank, batch=8, 1xB200 (tensor-core path stress)",
                       "images_per_gpu_per_step": B, "text_bank_rows": 103, "l2": "3 rotating input batches + ~2 GB of outputs per step"},
            "e2e": {"value": world * B * args.steps / (ms_e / 1e3), "unit": "images/s", "h2d_bytes_per_step": B * 3 * S * S * 4,
                    "d2h_bytes_per_step": B * 576 * 80 * 4, "ms_per_step": ms_e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"kernel": "surgery ViT-L/14 forward (gemm_tc + attn_tc + attn_pv)", "bound": "tensor", "achieved": ach,
                         "peak": tf_sust, "unit": "TFLOP/s", "frac": ach / tf_sust, "traffic": None,
                         "peak_kind": peak_kind + " (sustained bf16/fp16 dense)", "fp32_equiv_tflops": ach / 3,
                         "note": "3 fp16 MMA passes per fp32-quality product; 402 GFLOP/image fp32-equivalent (SURVEY.md §8d); includes the CAM kernels"}}
    if rank == 0:
        print(json.dumps(line), flush=True)
    job.close()


def run_cfg5(args):
    """configs[4]: PAR iteration sweep 1-50 @1024^2, batch 4, 4 planes (HBM-roofline stress)."""
    from excel_b200 import synth
    from excel_b200.par import par_affinity, par_refine_planes
    job = Job()
    dev, rank, world = job.dev, job.rank, job.world
    B, S, C = 4, 1024, 4
    hbm_peak, _, _, peak_kind = peaks()
    imgs = [synth.images(B, S, seed=40 + 3 * rank + i).to(dev) for i in range(2)]
    planes = [torch.softmax(torch.randn(B * C, S, S, device=dev), 0).contiguous() for _ in range(2)]
    off = torch.arange(0, (B + 1) * C, C, dtype=torch.int32, device=dev)
    sweep = {}
    for it in (1, 2, 5, 10, 50, 20):       # 20 last: its timed() call is the line's headline (clocks sampled there)
        fn = lambda i, it=it: par_refine_planes(imgs[i % 2], planes[i % 2], off, C, DIL, it)
        ms, launches, clocks, _ = job.timed(fn, args.steps, args.warmup, sample_clocks=(rank == 0 and it == 20))
        alg = 4.0 * S * S * ((3 + 48) + it * (48 + 2 * C)) * B
        sweep[str(it)] = {"ms": ms / args.steps, "GBps": alg / (ms / args.steps * 1e-3) / 1e9, "frac": alg / (ms / args.steps * 1e-3) / 1e9 / hbm_peak}
    t_aff = event_ms(lambda: par_affinity(imgs[0], (S, S), DIL))
    step_ms = (sweep["50"]["ms"] - sweep["10"]["ms"]) / 40
    alg_step = 4.0 * S * S * (48 + 2 * C) * B
    host_m = planes[0].cpu().pin_memory()
    host_i = imgs[0].cpu().pin_memory()
    out_h = torch.empty_like(host_m)

    def step_e2e(i):
        out_h.copy_(par_refine_planes(host_i.to(dev, non_blocking=True), host_m.to(dev, non_blocking=True), off, C, DIL, 20), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    ms_e, _, _, _ = job.timed(step_e2e, args.steps, max(args.warmup, 3))
    line = {"metric": "PAR HBM GB/s @1024^2 batch 4 (20 iterations incl. affinity set-up, algorithmic bytes)", "value": world * sweep["20"]["GBps"],
            "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sweep["20"]["ms"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "PAR iteration sweep 1-50 iters @1024x1024, batch=4, 1xB200 (HBM-roofline stress)", "planes_per_image": C,
                       "l2": "805 MB affinity stream per step + 2 rotating inputs (>> L2)"},
            "e2e": {"value": world * 4.0 * S * S * ((3 + 48) + 20 * (48 + 2 * C)) * B / (ms_e / args.steps * 1e-3) / 1e9, "unit": "GB/s",
                    "h2d_bytes_per_step": B * (3 + C) * S * S * 4, "d2h_bytes_per_step": B * C * S * S * 4, "ms_per_step": ms_e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "sweep": sweep,
            "roofline": {"kernel": "par_iterate_kernel<4>", "bound": "hbm", "achieved": alg_step / (step_ms * 1e-3) / 1e9, "peak": hbm_peak,
                         "unit": "GB/s", "frac": alg_step / (step_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None, "peak_kind": peak_kind,
                         "algorithmic_bytes": alg_step, "note": f"one propagation step, (t(50) - t(10)) / 40 = {step_ms*1e3:.1f} us; "
                         f"affinity set-up {t_aff*1e3:.0f} us = {4.0*S*S*51*B/(t_aff*1e-3)/1e9:.0f} GB/s"}}
    if rank == 0:
        print(json.dumps(line), flush=True)
    job.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--config", default="cfg2", choices=["cfg2", "cfg3", "cfg4", "cfg5"],
                    help="BASELINE.json configs[1..4]; cfg2 (default) is the headline the metric is quoted on")
    ap.add_argument("--surface", default="batched", choices=["batched", "dropin"],
                    help="cfg2 only: the batched public API (default) or the reference's per-image loop through install()")
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
    if args.config == "cfg2":
        return run_dropin(args) if args.surface == "dropin" else run_cuda(args)
    {"cfg3": run_cfg3, "cfg4": run_cfg4, "cfg5": run_cfg5}[args.config](args)


if __name__ == "__main__":
    main()
