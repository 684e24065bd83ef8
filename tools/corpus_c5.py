"""BASELINE.json configs[4] end to end: "ViT-gopt-16-SigLIP2-384 image-search corpus embedding (100k synthetic images)
sharded over 8 x B200", then one `rank_images`-style query over the whole corpus (src/clip.rs:136-170).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        tools/corpus_c5.py --images 100000

Every rank owns a contiguous index range (sharding.shard_range), regenerates its images from the counter-based
generator (sharding.counter_images: image i is a pure function of (seed, i), nothing is read from disk), embeds them
with its own engine replica and keeps the rows; there is no collective on the data path.  Rank 0 then gathers the
[N, 1536] matrix, loads it into an HBM-resident corpus (clipb200_corpus_*) and ranks it against a query embedding.
Image generation runs in worker threads ahead of the GPU; the JSON line reports the wall time of the whole job and
the time spent inside the embedding calls alone.
"""
from __future__ import annotations

import argparse
import json
import os
import queue
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=100000)
    ap.add_argument("--config", default="gopt_siglip2_384")
    ap.add_argument("--chunk", type=int, default=256)
    ap.add_argument("--workers", type=int, default=3)
    ap.add_argument("--model-root", default="/tmp/clipb200_c5")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    import clip_embedder_rs_b200 as cb
    import export_synthetic as ex
    from clip_embedder_rs_b200 import corpus as corpus_mod
    from clip_embedder_rs_b200 import sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    spec = ex.CONFIGS[args.config]
    mdir = os.path.join(args.model_root, args.config)
    if rank == 0 and not os.path.exists(os.path.join(mdir, "visual.onnx.data")):
        ex.write_model_dir(spec, mdir, seed=0, towers=("vision",))
    if world > 1:
        dist.barrier()
    vis = cb.VisionEmbedder.from_local_dir(mdir).device(local).build()
    size, dim = spec.vision.image_size, spec.embed_dim
    start, stop = sharding.shard_range(args.images, rank, world)

    # producer threads: regenerate images ahead of the GPU (numpy releases the GIL inside the integer kernels)
    chunks = [(s, min(s + args.chunk, stop)) for s in range(start, stop, args.chunk)]
    todo: "queue.Queue" = queue.Queue()
    slots = threading.Semaphore(2 * args.workers)  # bounds the chunks generated ahead of the GPU (113 MB each)
    for i, c in enumerate(chunks):
        todo.put((i, c))
    results = {}
    cond = threading.Condition()

    def produce():
        while True:
            slots.acquire()
            try:
                i, (s, e) = todo.get_nowait()
            except queue.Empty:
                slots.release()
                return
            imgs = sharding.counter_images(s, e, size, seed=9)
            with cond:
                results[i] = imgs
                cond.notify_all()

    vis.embed_images(sharding.counter_images(0, 8, size, seed=1))  # warm-up (allocations, first launches)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    threads = [threading.Thread(target=produce, daemon=True) for _ in range(args.workers)]
    for t in threads:
        t.start()
    rows = np.empty((stop - start, dim), dtype=np.float32)
    embed_s = 0.0
    for i, (s, e) in enumerate(chunks):
        with cond:
            while i not in results:
                cond.wait()
            imgs = results.pop(i)
        slots.release()
        t1 = time.perf_counter()
        rows[s - start:e - start] = vis.embed_images(imgs)
        embed_s += time.perf_counter() - t1
    local_s = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    wall_s = time.perf_counter() - t0

    stats = torch.tensor([local_s, embed_s, float(stop - start)], dtype=torch.float64, device=f"cuda:{local}")
    all_stats = [torch.zeros_like(stats) for _ in range(world)] if world > 1 else [stats]
    if world > 1:
        dist.all_gather(all_stats, stats)
    whole = sharding.gather_rows(rows, args.images, dim, rank, world, dst=0, device=f"cuda:{local}" if world > 1 else None)
    if rank == 0:
        whole = np.ascontiguousarray(whole[:args.images] if world == 1 else whole, dtype=np.float32)
        norms = np.linalg.norm(whole, axis=1)
        t2 = time.perf_counter()
        corp = corpus_mod.EmbeddingCorpus(dim, capacity=args.images, device=local)
        corp.append(whole)
        query = whole[12345 % args.images]  # a corpus row as the query: it must rank itself first
        # softmax across the corpus (clip.rs:144-163 for softmax models): with the model's own sigmoid activation and a
        # logit scale of 112 every near-duplicate of a random image saturates at 1.0 and the ranking is all ties
        probs = corp.probabilities(query, spec.logit_scale, spec.logit_bias, False)
        idx = np.argsort(-probs, kind="stable")  # stable descending sort, clip.rs:167
        order = [(int(i), float(probs[i])) for i in idx[:5]]
        rank_s = time.perf_counter() - t2
        per_rank = [[round(float(v), 3) for v in s.tolist()] for s in all_stats]
        print(json.dumps({
            "workload": f"{spec.name} corpus embedding, {args.images} counter-based {size}x{size} images, {world} GPU(s)",
            "images": args.images, "n_gpus": world, "wall_s": round(wall_s, 3),
            "images_per_s_wall": round(args.images / wall_s, 1),
            "images_per_s_embed_only": round(sum(s[2].item() for s in all_stats) / max(s[1].item() for s in all_stats), 1),
            "per_rank_[local_s, embed_s, images]": per_rank,
            "unit_norm_max_err": float(np.abs(norms - 1.0).max()),
            "corpus_upload_and_rank_s": round(rank_s, 4), "top1_is_query": int(order[0][0]) == 12345 % args.images,
            "top5": [[int(i), float(p)] for i, p in order[:5]]}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
