#!/usr/bin/env python
"""Benchmark of the hot path: `VisionEmbedder::embed_images` (reference src/vision.rs:102-117) on
ViT-SO400M-16-SigLIP2-384, batch 1024 per GPU, preprocessing included (BASELINE.json configs[2]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

One process per GPU (torchrun for N > 1, ranks read from the environment), an independent engine replica per
rank, no collective on the data path (weak scaling: every rank embeds its own 1024 synthetic images per step).

A step = one call of the C-ABI embed entry point over the whole per-rank batch.
  value ...... images/s with the uint8 batch already resident in HBM (clipb200_vision_embed_rgb8_device), timed
               with CUDA events on the engine's compute stream, max over ranks.
  e2e ........ the same metric through the host-buffer entry point (clipb200_vision_embed_rgb8): pinned host
               uint8 in, host fp32 embeddings out, H2D/D2H inside the timed region.
  roofline ... the tcgen05 GEMM kernel class: algorithmic 2*M*N*K of all GEMM launches / their summed CUDA-event
               durations, against the measured bf16 peak in MEASURED_PEAKS.json.
  cpu_baseline the CPU oracle port (oracle/reference_forward.py, torch fp32, all host threads) on a bounded sample.
`--impl reference` times that CPU port as the reference arm (the reference's ort CPU EP cannot be built here:
no Rust toolchain and no onnxruntime in the image).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

WORKLOADS = {
    # name: (export config, tower, per-GPU batch, GFLOP per unit, unit, description)
    "so400m_vision": ("so400m_siglip2_384", "vision", 1024, 518.94, "images/s",
                      "ViT-SO400M-16-SigLIP2-384 vision embedding, batch 1024 per GPU, uint8 384x384 inputs, "
                      "GPU preprocessing included"),
    "gopt_vision": ("gopt_siglip2_384", "vision", 512, 1392.98, "images/s",
                    "ViT-gopt-16-SigLIP2-384 vision embedding, batch 512 per GPU"),
    "dfn5b_text": ("dfn5b_h14_378", "text", 8192, 47.09, "texts/s",
                   "DFN5B-CLIP-ViT-H-14-378 text encoder, batch 8192 per GPU, context 77"),
    "vit_b32_vision": ("vit_b32", "vision", 1024, 8.82, "images/s", "ViT-B/32 vision embedding, batch 1024 per GPU"),
    "mobileclip2_vision": ("mobileclip2_s2", "vision", 256, 15.67, "images/s",
                           "MobileCLIP2-S2 (FastViT-MCi2) vision embedding, batch 256 per GPU, uint8 256x256 inputs"),
    "mobileclip2_s3_vision": ("mobileclip2_s3", "vision", 256, 0.0, "images/s",
                              "MobileCLIP2-S3 (FastViT-MCi3, 5 stages) vision embedding, batch 256 per GPU"),
    "mobileclip2_s4_vision": ("mobileclip2_s4", "vision", 256, 0.0, "images/s",
                              "MobileCLIP2-S4 (FastViT-MCi4, 5 stages) vision embedding, batch 256 per GPU"),
    "mobileclip2_text": ("mobileclip2_s2", "text", 256, 5.96, "texts/s", "MobileCLIP2-S2 text encoder, batch 256, context 77"),
    "small_vision": ("small_siglip", "vision", 256, 0.0, "images/s", "small SigLIP-shaped test tower"),
    "so400m_photos": ("so400m_siglip2_384", "vision", 56, 518.94, "images/s",
                      "ViT-SO400M-16-SigLIP2-384 embed_images on photo-sized inputs (GPU resize included)"),
    # a crop-mode model (resize the shorter side, centre crop): only the rows / columns inside the crop cross PCIe
    "mobileclip2_photos": ("mobileclip2_s2", "vision", 56, 0.0, "images/s",
                           "MobileCLIP2-S2 embed_images on photo-sized inputs (GPU resize + centre crop included)"),
}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def model_dir_for(config: str, towers, rank: int, world: int, barrier) -> str:
    """Rank 0 writes the synthetic model directory once (seeded), everyone else waits for the marker."""
    import export_synthetic as ex

    base = os.environ.get("CLIPB200_MODEL_CACHE", os.path.join(tempfile.gettempdir(), "clipb200_models"))
    path = os.path.join(base, f"{config}_{'_'.join(towers)}_s0")
    marker = os.path.join(path, ".complete")
    if rank == 0 and not os.path.exists(marker):
        ex.write_model_dir(ex.CONFIGS[config], path, seed=0, towers=towers)
        open(marker, "w").close()
    barrier()
    while not os.path.exists(marker):
        time.sleep(0.2)
    return path


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"burst": float(d["bf16_tflops"]), "sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                "hbm": float(d["hbm_gbs"]), "source": "measured"}
    return {"burst": 1590.0, "sustained": 1400.0, "hbm": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.path = os.path.join(tempfile.gettempdir(), f"clipb200_clocks_{os.getpid()}.csv")
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(gpu_index)], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1]))
                    mx.append(float(parts[2]))
                except ValueError:
                    continue
                for name, val in zip(names, parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            load = [s for s in sm if s >= 0.5 * max(sm)] or sm
            out["sm_mhz"] = statistics.median(load)
            out["sm_max_mhz"] = max(mx) if mx else None
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def cpu_reference_rate(model_dir: str, tower: str, seconds: float, sample_items: int, seed: int):
    """Times the CPU oracle port on a bounded sample with all host threads.  Returns (units/s, cores, sample)."""
    import numpy as np
    import torch

    from oracle import reference_forward as R

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    with open(os.path.join(model_dir, "open_clip_config.json")) as f:
        cfg = json.load(f)
    rng = np.random.default_rng(seed)
    if tower == "vision":
        t = R.Tower(os.path.join(model_dir, "visual.onnx"))
        size = int(cfg["model_cfg"]["vision_cfg"]["image_size"])
        pc = cfg["preprocess_cfg"]
        imgs = rng.integers(0, 256, size=(sample_items, size, size, 3), dtype=np.uint8)

        def run():
            return R.vision_forward(t, R.preprocess_batch(list(imgs), size, pc["mean"], pc["std"]))
    else:
        t = R.Tower(os.path.join(model_dir, "text.onnx"))
        ctx = int(cfg["model_cfg"]["text_cfg"]["context_length"])
        ids = rng.integers(1, 40000, size=(sample_items, ctx), dtype=np.int64)
        ids[:, -1] = 49407

        def run():
            return R.text_forward(t, ids)
    run()  # warm-up (thread pool, allocator)
    n, t0 = 0, time.perf_counter()
    while True:
        run()
        n += 1
        el = time.perf_counter() - t0
        if el >= seconds or n >= 8:
            break
    return n * sample_items / el, cores, f"{sample_items} items x {n} passes in {el:.1f} s after 1 warm-up pass"


def run_reference_arm(args, wl, rank, world):
    config, tower, batch, gflop, unit, desc = wl
    if rank != 0:
        return
    mdir = model_dir_for(config, (tower,), 0, 1, lambda: None)
    sample = 4 if tower == "vision" else 32
    rates = []
    cores = os.cpu_count() or 1
    note = ""
    for i in range(args.warmup + args.steps):
        r, cores, note = cpu_reference_rate(mdir, tower, seconds=6.0, sample_items=sample, seed=100 + i)
        if i >= args.warmup:
            rates.append(r)
    value = statistics.mean(rates)
    line = {
        "impl": "reference", "metric": "SigLIP2-SO400M-384 images/sec" if args.workload == "so400m_vision" else f"{args.workload} {unit}",
        "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sample / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "sample_per_step": sample,
                   "note": "reference's ort CPU EP is not buildable here (no cargo / onnxruntime); this is the CPU "
                           "oracle port (torch fp32) on all host threads"},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": note},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# Sizes of the reference's example photos (assets/img/*.jpg, width x height): the inputs `embed_images(&[DynamicImage])`
# sees in its README / integration test.  Content is synthetic; only the sizes matter for resize + PCIe cost.
PHOTO_SIZES = [(4608, 3456), (2592, 1944), (5312, 2988), (4160, 2336), (4608, 3456), (1944, 2592), (2592, 1456)]


def pinned_u8(lib, nbytes):
    import numpy as np

    p = lib.clipb200_host_alloc(nbytes)
    if not p:
        raise RuntimeError("pinned allocation failed")
    return p, np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,))


def run_pool(args):
    """`--pool`: ONE host process drives N GPUs through the in-process pool (clipb200_pool_*), the deployment shape of a
    Rust drop-in (one `VisionEmbedder` serving a caller's batch).  End to end only: pinned host uint8 in, host fp32
    out, every replica running its own H2D / compute / D2H pipeline.  Not the driver's scaling line (that stays
    torchrun); results go to profiles/r02_pool_*gpu.json."""
    import numpy as np

    import clip_embedder_rs_b200 as cb
    from clip_embedder_rs_b200 import _native

    lib = _native.lib
    config, tower, batch, gflop, unit, desc = WORKLOADS[args.workload]
    if args.batch > 0:
        batch = args.batch
    n = args.gpus
    mdir = model_dir_for(config, (tower,), 0, 1, lambda: None)
    t0 = time.perf_counter()
    build = cb.VisionEmbedder if tower == "vision" else cb.TextEmbedder
    emb = build.from_local_dir(mdir).micro_batch(args.micro_batch).devices(list(range(n))).build()
    load_s = time.perf_counter() - t0
    sess = emb.session
    E = sess.embed_dim
    total = batch * n  # weak scaling: 1024 images per GPU, one caller-side batch of N*1024
    rng = np.random.default_rng(4)
    if tower == "vision":
        size = sess.image_size
        item = size * size * 3
        h_in, host = pinned_u8(lib, total * item)
        block = rng.integers(0, 256, size=min(len(host), 1 << 28), dtype=np.uint8)
        for o in range(0, len(host), len(block)):
            host[o:o + len(block)] = block[:len(host) - o]
    else:
        ctx = sess.context_length
        item = ctx * 8
        h_in, raw = pinned_u8(lib, total * item)
        host = raw.view(np.int64).reshape(total, ctx)
        host[:] = 0
        lens = rng.integers(4, ctx - 1, size=total)
        host[:, 0] = 49406
        for i in range(total):
            host[i, 1:lens[i]] = rng.integers(1, 49000, size=lens[i] - 1)
            host[i, lens[i]] = 49407
    h_out, out_raw = pinned_u8(lib, total * E * 4)

    def step():
        if tower == "vision":
            sess.run_rgb8(h_in, total, size, size, emb._pp, h_out)
        else:
            sess.run_ids(h_in, None, total, ctx, h_out)

    for _ in range(max(args.warmup, 1)):
        step()
    l0 = sess.launch_count
    sampler = ClockSampler(0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    el = time.perf_counter() - t0
    clocks = sampler.stop()
    out = out_raw.view(np.float32).reshape(total, E)
    norm_err = float(np.abs(np.linalg.norm(out[:: max(1, total // 64)], axis=1) - 1.0).max())
    value = total * args.steps / el
    line = {"mode": "pool", "metric": f"{args.workload} {unit} (one process, in-process pool)", "value": value, "unit": unit,
            "n_gpus": n, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
            "higher_is_better": True, "scaling": "weak", "dtype": "bf16", "data": "synthetic",
            "config": {"workload": desc, "per_gpu_batch": batch, "global_batch": total,
                       "parallelism": f"one host process, {n} replicas, one host thread per replica, rows split "
                                      "contiguously, no collective",
                       "host_cores": os.cpu_count(), "pool_load_s": load_s},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": total * item, "d2h_bytes_per_step": total * E * 4},
            "h2d_gb_per_s": total * item * args.steps / el / 1e9,
            "per_gpu": value / n, "gpu_launches": sess.launch_count - l0, "clocks": clocks, "out_norm_err": norm_err}
    print(json.dumps(line), flush=True)
    sess.close()


def run_photos(args):
    """`--workload so400m_photos`: `embed_images(&[DynamicImage])` on photo-sized inputs (the sizes of the reference's
    assets/img JPEGs, synthetic content, pageable host memory like a decoded `DynamicImage`): staging, H2D, GPU resize
    (vision.rs:164-198), normalise and tower, all inside the timed region.  Reports images/s, the PCIe floor of the
    same bytes, and the CPU oracle's resize rate next to the reference README's 10-20 ms per image."""
    import numpy as np

    import clip_embedder_rs_b200 as cb
    from clip_embedder_rs_b200 import _native

    lib = _native.lib
    cfg_name, _, _, _, _, label = WORKLOADS[args.workload]
    mdir = model_dir_for(cfg_name, ("vision",), 0, 1, lambda: None)
    emb = cb.VisionEmbedder.from_local_dir(mdir).micro_batch(args.micro_batch).profile(True).build()
    sess = emb.session
    pc = emb.config.preprocess_cfg
    size = emb.config.model_cfg.vision_cfg.image_size
    rng = np.random.default_rng(4)
    base = {}
    for (w, h) in set(PHOTO_SIZES):  # one random texture per size, shifted per image (generation is not what is timed)
        base[(w, h)] = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    imgs = []
    for i in range(args.photos):
        w, h = PHOTO_SIZES[i % len(PHOTO_SIZES)]
        imgs.append(np.ascontiguousarray(np.roll(base[(w, h)], i * 17, axis=1)))
    src_bytes = sum(a.nbytes for a in imgs)
    for _ in range(max(1, min(args.warmup, 2))):
        emb.embed_images(imgs)
    sess.profile(reset=True)
    l0 = sess.launch_count
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = emb.embed_images(imgs)
    el = time.perf_counter() - t0
    prof = sess.profile(reset=True)
    value = len(imgs) * args.steps / el
    # PCIe floor: the same bytes from pinned memory, one plain H2D copy
    h_p, host = pinned_u8(lib, 1 << 28)
    d_p = lib.clipb200_device_alloc(0, 1 << 28)
    lib.clipb200_memcpy_h2d(0, d_p, h_p, 1 << 28)
    t1 = time.perf_counter()
    for _ in range(4):
        lib.clipb200_memcpy_h2d(0, d_p, h_p, 1 << 28)
    pcie = 4 * (1 << 28) / (time.perf_counter() - t1)
    # staging floor: one host thread copying the same photos pageable -> pinned
    t1 = time.perf_counter()
    n_st = 0
    for a in imgs[:8]:
        m = min(a.nbytes, 1 << 28)
        host[:m] = a.reshape(-1)[:m]
        n_st += m
    memcpy_1t = n_st / (time.perf_counter() - t1)
    # CPU baseline: the oracle's restatement of resize_with_fast_image_resize on the same photos
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import resize as RZ

        t1 = time.perf_counter()
        k = 0
        for a in imgs[:2]:
            RZ.resize_rgb8(a, size, pc.interpolation, pc.resize_mode)
            k += 1
        cpu_el = time.perf_counter() - t1
        cpu = {"value": k / cpu_el, "unit": "images/s (resize only)", "cores": 1, "kind": "port",
               "sample": f"{k} photos, oracle/resize.py (numpy) in {cpu_el:.1f} s; the reference README quotes 10-20 ms "
                         "per image for its preprocessing on the author's CPU"}
    ms = prof["ms"]
    line = {"metric": f"{label.split(' embed_images')[0]} images/sec from photo-sized inputs (resize on the GPU)", "value": value,
            "unit": "images/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{label.split(' embed_images')[0]} embed_images on {len(imgs)} photo-sized RGB8 images per step "
                                   f"(sizes of the reference's assets/img: {sorted(set(PHOTO_SIZES))}), pageable host memory",
                       "mean_photo_mb": src_bytes / len(imgs) / 1e6, "resize_mode": pc.resize_mode,
                       "interpolation": pc.interpolation, "image_size": size},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": src_bytes, "d2h_bytes_per_step": len(imgs) * out.shape[1] * 4},
            "h2d_gb_per_s": src_bytes * args.steps / el / 1e9,
            "pcie_floor": {"measured_h2d_gb_per_s": pcie / 1e9, "images_per_s": pcie / (src_bytes / len(imgs))},
            "staging_floor_one_thread": {"memcpy_gb_per_s": memcpy_1t / 1e9, "images_per_s": memcpy_1t / (src_bytes / len(imgs))},
            "kernel_ms_per_step": {k: v / args.steps for k, v in ms.items()},
            "gpu_launches": sess.launch_count - l0, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    sess.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="so400m_vision", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override")
    ap.add_argument("--micro-batch", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-text", action="store_true", help="skip the secondary text-encoder measurement")
    ap.add_argument("--no-extras", action="store_true", help="skip the strong-scaling and MobileCLIP2 extra keys")
    ap.add_argument("--pool", action="store_true",
                    help="ONE process feeding --gpus N devices through the in-process pool (clipb200_pool_*), no torchrun")
    ap.add_argument("--photos", type=int, default=56, help="photos per step of --workload so400m_photos / mobileclip2_photos")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    wl = WORKLOADS[args.workload]
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)

    if args.impl == "reference":
        run_reference_arm(args, wl, rank, world)
        return
    if args.workload in ("so400m_photos", "mobileclip2_photos"):
        run_photos(args)
        return
    if args.pool:
        run_pool(args)
        return

    import numpy as np
    import torch

    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    import clip_embedder_rs_b200 as cb
    from clip_embedder_rs_b200 import _native

    lib = _native.lib
    pk = peaks()

    def measure(workload_name: str, steps: int, warmup: int, with_e2e: bool, batch_override: int = 0):
        config, tower, batch, gflop, unit, desc = WORKLOADS[workload_name]
        if args.batch > 0 and workload_name == args.workload:
            batch = args.batch
        if batch_override > 0:
            batch = batch_override
        mdir = model_dir_for(config, (tower,), rank, world, barrier)
        dev = local_rank
        if tower == "vision":
            emb = cb.VisionEmbedder.from_local_dir(mdir).device(dev).micro_batch(args.micro_batch).profile(True).build()
            size = emb.session.image_size
            item_bytes = size * size * 3
        else:
            emb = cb.TextEmbedder.from_local_dir(mdir).device(dev).micro_batch(args.micro_batch).profile(True).build()
            ctx = emb.session.context_length
            item_bytes = ctx * 8
        sess = emb.session
        h = sess.handle
        E = sess.embed_dim
        in_bytes, out_bytes = batch * item_bytes, batch * E * 4
        # pinned host input / output (the caller's buffers for the e2e arm)
        h_in = lib.clipb200_host_alloc(in_bytes)
        h_out = lib.clipb200_host_alloc(out_bytes)
        if not h_in or not h_out:
            raise RuntimeError("pinned allocation failed")
        rng = np.random.default_rng(4 + rank)
        if tower == "vision":
            host = np.ctypeslib.as_array(C.cast(h_in, C.POINTER(C.c_uint8)), shape=(in_bytes,))
            chunk = 1 << 26
            for o in range(0, in_bytes, chunk):
                host[o:o + chunk] = rng.integers(0, 256, size=min(chunk, in_bytes - o), dtype=np.uint8)
        else:
            host = np.ctypeslib.as_array(C.cast(h_in, C.POINTER(C.c_int64)), shape=(batch, ctx))
            host[:] = 0
            lens = rng.integers(4, ctx - 1, size=batch)
            for i in range(batch):
                host[i, 0] = 49406
                host[i, 1:lens[i]] = rng.integers(1, 49000, size=lens[i] - 1)
                host[i, lens[i]] = 49407
        d_in = lib.clipb200_device_alloc(dev, in_bytes)
        d_out = lib.clipb200_device_alloc(dev, out_bytes)
        if not d_in or not d_out:
            raise RuntimeError("device allocation failed")
        sess.check(lib.clipb200_memcpy_h2d(dev, d_in, h_in, in_bytes))

        def step_device():
            if tower == "vision":
                sess.check(lib.clipb200_vision_embed_rgb8_device(h, d_in, batch, emb._pp, d_out))
            else:
                sess.check(lib.clipb200_text_embed_device(h, d_in, batch, ctx, d_out))

        def step_host():
            if tower == "vision":
                sess.check(lib.clipb200_vision_embed_rgb8(h, h_in, batch, size, size, emb._pp, h_out))
            else:
                sess.check(lib.clipb200_text_embed(h, h_in, None, batch, ctx, h_out))

        # ---------------- device-resident timing (value) ----------------
        for _ in range(warmup):
            step_device()
        sess.synchronize()
        sess.profile(reset=True)
        launches0 = sess.launch_count
        barrier()
        sampler = ClockSampler(dev) if rank == 0 else None
        sess.check(lib.clipb200_engine_record_event(h, 0))
        for _ in range(steps):
            step_device()
        sess.check(lib.clipb200_engine_record_event(h, 1))
        ms = C.c_double()
        sess.check(lib.clipb200_engine_elapsed_ms(h, 0, 1, C.byref(ms)))
        sess.synchronize()
        barrier()
        clocks = sampler.stop() if sampler else None
        prof = sess.profile(reset=True)
        launches = sess.launch_count - launches0
        dev_ms = max_over_ranks(ms.value)
        res = {"unit": unit, "desc": desc, "batch": batch, "ms_per_step": dev_ms / steps, "steps": steps,
               "value": world * batch * steps / (dev_ms * 1e-3), "launches": launches, "prof": prof,
               "clocks": clocks, "gflop": gflop, "h2d": in_bytes, "d2h": out_bytes,
               "weight_bytes": sess.weight_bytes}
        # ---------------- end-to-end timing (host buffers) ----------------
        if with_e2e:
            for _ in range(max(1, min(warmup, 2))):
                step_host()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                step_host()
            el = time.perf_counter() - t0
            el = max_over_ranks(el)
            res["e2e_value"] = world * batch * steps / el
            out = np.ctypeslib.as_array(C.cast(h_out, C.POINTER(C.c_float)), shape=(batch, E))
            res["out_norm_err"] = float(np.abs(np.linalg.norm(out[:64], axis=1) - 1.0).max())
        lib.clipb200_device_free(dev, d_in)
        lib.clipb200_device_free(dev, d_out)
        lib.clipb200_host_free(h_in)
        lib.clipb200_host_free(h_out)
        sess.close()
        return res, mdir

    res, mdir = measure(args.workload, args.steps, args.warmup, with_e2e=True)
    config, tower, _, gflop, unit, desc = wl

    gemm_ms = res["prof"]["ms"]["gemm"]
    gemm_tflops = res["prof"]["gemm_flops"] / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    total_kernel_ms = sum(v for k, v in res["prof"]["ms"].items() if k not in ("h2d", "d2h"))
    # DRAM bytes per GEMM launch from the committed ncu --set full capture of one SO400M layer at this micro-batch
    # (profiles/r01d_ncu_summary.md); the algorithmic bytes of the same launches are stated beside it.
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "r01d_gemm_traffic.json")
    if args.workload == "so400m_vision" and args.micro_batch in (0, 256) and os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic = tj["mean_dram_bytes_per_gemm_launch"]
        traffic_note = ("mean dram__bytes_read+write per layer-GEMM launch (qkv, proj, fc1, fc2 at M=147456), "
                        "profiles/r01d_gemm_traffic.json; algorithmic bytes of the same launches: 1.83e9")
    roofline = {
        "bound": "tensor", "kernel": "gemm_bf16_tcgen05_kernel (all GEMM launches of the timed steps)",
        "achieved": gemm_tflops, "peak": pk["sustained"], "unit": "TFLOP/s",
        "frac": gemm_tflops / pk["sustained"] if pk["sustained"] else None,
        "frac_of_burst_peak": gemm_tflops / pk["burst"] if pk["burst"] else None,
        "peak_source": f"{pk['source']} (MEASURED_PEAKS.json bf16_tflops_sustained; burst {pk['burst']})",
        "traffic": traffic, "traffic_source": traffic_note,
        "gemm_share_of_kernel_time": gemm_ms / total_kernel_ms if total_kernel_ms > 0 else None,
        "kernel_ms_per_step": {k: v / args.steps for k, v in res["prof"]["ms"].items()},
        "launches_per_step": {k: v / args.steps for k, v in res["prof"]["launches"].items()},
    }
    whole = None
    if gflop > 0:
        tf = res["value"] / world * gflop * 1e9 / 1e12  # per GPU
        whole = {"tflops_per_gpu": tf, "frac_of_sustained_peak": tf / pk["sustained"],
                 "frac_of_burst_peak": tf / pk["burst"], "gflop_per_unit": gflop}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r, cores, note = cpu_reference_rate(mdir, tower, seconds=15.0,
                                                sample_items=4 if tower == "vision" else 32, seed=7)
            cpu_baseline = {"value": r, "unit": unit, "cores": cores, "kind": "port", "sample": note}
        except Exception as e:  # pragma: no cover
            cpu_baseline = {"value": None, "unit": unit, "cores": os.cpu_count(), "kind": "port",
                            "sample": f"failed: {e}"}

    def class_rooflines(r):
        """GEMM class in TFLOP/s and depthwise-conv class in GB/s from the live per-class CUDA-event times."""
        ms, out = r["prof"]["ms"], {}
        if ms["gemm"] > 0:
            tf = r["prof"]["gemm_flops"] / (ms["gemm"] * 1e-3) / 1e12
            out["gemm"] = {"bound": "tensor", "achieved": tf, "peak": pk["sustained"], "unit": "TFLOP/s",
                           "frac": tf / pk["sustained"], "ms_per_step": ms["gemm"] / r["steps"]}
        if ms.get("dwconv", 0) > 0:
            gbs = r["prof"]["conv_bytes"] / (ms["dwconv"] * 1e-3) / 1e9
            out["dwconv"] = {"bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"],
                             "ms_per_step": ms["dwconv"] / r["steps"],
                             "note": "algorithmic bytes (input + output once) of every depthwise conv launch"}
        out["kernel_ms_per_step"] = {k: v / r["steps"] for k, v in ms.items()}
        return out

    # BASELINE config 4 (DFN5B text, batch 8192 per GPU) at every N
    text_extra = None
    if args.workload == "so400m_vision" and not args.no_text:
        try:
            tres, _ = measure("dfn5b_text", max(2, args.steps // 2), 2, with_e2e=True)
            tf = tres["value"] / world * tres["gflop"] * 1e9 / 1e12
            text_extra = {"workload": tres["desc"], "value": tres["value"], "unit": tres["unit"], "n_gpus": world,
                          "e2e": tres.get("e2e_value"), "ms_per_step": tres["ms_per_step"],
                          "tflops_per_gpu": tf, "frac_of_sustained_peak": tf / pk["sustained"]}
        except Exception as e:  # pragma: no cover
            text_extra = {"error": str(e)}

    # BASELINE config 3 read as strong scaling: a global batch of 1024 split over the N GPUs
    strong_extra = None
    if args.workload == "so400m_vision" and world > 1 and not args.no_extras:
        try:
            per_gpu = max(1, 1024 // world)
            sres, _ = measure("so400m_vision", args.steps, max(1, args.warmup), with_e2e=True, batch_override=per_gpu)
            strong_extra = {"global_batch": per_gpu * world, "per_gpu_batch": per_gpu, "value": sres["value"],
                            "unit": sres["unit"], "e2e": sres.get("e2e_value"), "ms_per_step": sres["ms_per_step"],
                            "vs_weak_same_run": sres["value"] / res["value"],
                            "note": "fixed total work: every GPU embeds 1024/N images per step (half a micro-batch at "
                                    "N=8); vs_weak_same_run = this throughput / the weak-scaling value of this run"}
        except Exception as e:  # pragma: no cover
            strong_extra = {"error": str(e)}

    # BASELINE config 2: MobileCLIP2-S2 vision + text, batch 256, one GPU
    mobile_extra = None
    if args.workload == "so400m_vision" and world == 1 and not args.no_extras:
        try:
            mv, _ = measure("mobileclip2_vision", max(3, args.steps), 3, with_e2e=True)
            mt, _ = measure("mobileclip2_text", max(3, args.steps), 3, with_e2e=True)
            mobile_extra = {
                "vision": {"workload": mv["desc"], "value": mv["value"], "unit": mv["unit"], "e2e": mv.get("e2e_value"),
                           "ms_per_step": mv["ms_per_step"], "gpu_launches": mv["launches"],
                           "tflops": mv["value"] * mv["gflop"] * 1e9 / 1e12, "roofline": class_rooflines(mv)},
                "text": {"workload": mt["desc"], "value": mt["value"], "unit": mt["unit"], "e2e": mt.get("e2e_value"),
                         "ms_per_step": mt["ms_per_step"], "gpu_launches": mt["launches"],
                         "tflops": mt["value"] * mt["gflop"] * 1e9 / 1e12, "roofline": class_rooflines(mt)},
            }
        except Exception as e:  # pragma: no cover
            mobile_extra = {"error": str(e)}

    if rank == 0:
        line = {
            "metric": "SigLIP2-SO400M-384 images/sec" if args.workload == "so400m_vision" else f"{args.workload} {unit}",
            "value": res["value"], "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": desc, "per_gpu_batch": res["batch"], "global_batch": res["batch"] * world,
                       "parallelism": f"replicas x{world}, batch row-sharded, no collective",
                       "weights": "random-init, reference ONNX layout (tools/export_synthetic.py), bf16 in HBM",
                       "cache": "inputs_larger_than_l2 (453 MB uint8 batch + >1 GB activations per micro-batch)",
                       "weight_bytes": res["weight_bytes"]},
            "e2e": {"value": res.get("e2e_value"), "unit": unit, "h2d_bytes_per_step": res["h2d"],
                    "d2h_bytes_per_step": res["d2h"]},
            "gpu_launches": res["launches"],
            "clocks": res["clocks"],
            "roofline": roofline,
            "whole_step": whole,
            "cpu_baseline": cpu_baseline,
            "text": text_extra,
            "strong": strong_extra,
            "mobileclip2": mobile_extra,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
