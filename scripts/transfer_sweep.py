"""BASELINE.json configs[4]: ADiL transfer sweep over six random-init models (ResNet / DenseNet / VGG / GoogLeNet /
MobileNet / Inception), 16 384 images, 200 atoms, 8 x B200 -- the caller side of the hot path (performance.py:183-232).

    torchrun --nproc-per-node 8 scripts/transfer_sweep.py [--images-per-gpu 2048] [--atoms 200] [--epochs 1]

Each rank owns 2048 synthetic images.  (1) The dictionary is learned on the source model with the image-sharded fit
(ADIL.begin_fit(distributed=True) / fit_batch_resident: K = 200 runs the tcgen05 kernels -- synthesis in one pass, the
backward as two column windows).  (2) Every rank crafts adversarial images for its shard with the learned dictionary
(ADIL.__call__, supervised) and evaluates them on all six models through performance.get_transfer_performance; the
counters are all-reduced.  Writes gpurun_out/transfer_sweep_r02_w<world>.json on rank 0.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images-per-gpu", type=int, default=2048)
    ap.add_argument("--atoms", type=int, default=200)
    ap.add_argument("--batch", type=int, default=100)
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--source", default="resnet18")
    ap.add_argument("--steps-inference", type=int, default=10)
    ap.add_argument("--eval-images-per-gpu", type=int, default=512)
    args = ap.parse_args()
    from dl_attack_on_imagenet_b200 import ADIL, build_classifier, ops, performance as perf
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    EPS = 8.0 / 255.0
    N, K, B = args.images_per_gpu, args.atoms, args.batch
    names = ["resnet18", "densenet121", "vgg11", "googlenet", "mobilenet_v2", "inception_v3"]
    models = {n: build_classifier(n, seed=0, device=dev) for n in names}
    for mdl in models.values():
        for p in mdl.parameters():
            p.requires_grad_(False)
    g = torch.Generator().manual_seed(1 + rank)
    x = torch.rand(N, 3, 224, 224, generator=g)
    ADIL.verbose = False
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    os.chdir(os.path.join(ROOT, "gpurun_out"))
    atk = ADIL(models[args.source], eps=EPS, n_atoms=K, batch_size=B, model_name="sweep_%d" % rank, step_size=0.01,
               steps_inference=args.steps_inference)
    torch.manual_seed(1234)
    st = atk.begin_fit(N, (3, 224, 224), distributed=(world > 1))
    atk.set_resident_images(x)
    perm = torch.randperm(N, generator=torch.Generator().manual_seed(7 + rank))
    batches = [perm[i:i + B].contiguous() for i in range(0, N - B + 1, B)]
    for b in batches[:3]:                               # warm-up: graph capture, cuDNN autotuning, label cache
        atk.fit_batch_resident(b)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    with ops.kernel_timer() as kt:
        for _ in range(args.epochs):
            for b in batches:
                loss, fooled = atk.fit_batch_resident(b)
    if st.shard is not None:
        st.shard.wait()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    fit_s = time.perf_counter() - t0
    fit_rate = world * len(batches) * B * args.epochs / fit_s
    # ---- transfer evaluation: craft once per batch, evaluate on all six models ----------------------------------
    ne = min(args.eval_images_per_gpu, N)
    with torch.no_grad():
        y = torch.cat([models[args.source](x[i:i + 256].to(dev)).argmax(-1).cpu() for i in range(0, ne, 256)])
    data = [(x[i:i + B], y[i:i + B]) for i in range(0, ne, B)]
    t0 = time.perf_counter()
    res = perf.get_transfer_performance({"ADIL": [atk]}, models, data)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sweep_s = time.perf_counter() - t0
    if rank == 0:
        out = {"world": world, "images_per_gpu": N, "atoms": K, "batch": B, "source_model": args.source,
               "fit": {"images_per_s": fit_rate, "seconds": fit_s, "epochs": args.epochs,
                       "kernels": {k: {"calls": v["calls"], "us_mean": 1e3 * v["ms_mean"]} for k, v in kt.summary().items()},
                       "tc_supported": int(__import__("dl_attack_on_imagenet_b200")._lib.lib().adil_tc_supported(B, 150528, K)),
                       "dict_step": type(st.shard).__name__ if st.shard is not None else "fused single-GPU step"},
               "sweep": {"evaluated_images": world * ne, "seconds": sweep_s, "images_per_s": world * ne / sweep_s,
                         "steps_inference": args.steps_inference, "models": names},
               "transfer_performance": res["ADIL"]}
        with open(os.path.join(ROOT, "gpurun_out", "transfer_sweep_r02_w%d.json" % world), "w") as f:
            json.dump(out, f, indent=1)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
