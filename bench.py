#!/usr/bin/env python
"""bench.py — ViT-B/16 224x224 bf16 training throughput (images/sec) on N B200s, data-parallel.

  python bench.py --gpus N --steps K --warmup W            # our arm (hand-written sm_100a kernels)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path (oracle port) on host cores

One JSON line on stdout (rank 0).  A step = zero_grad + forward + cross-entropy + backward (+ gradient all-reduce)
+ Adam, on synthetic ImageNet-shaped data with random-init weights (BASELINE.json config 2).
  value : whole-job images/sec with the batch already resident in HBM
  e2e   : the same step through the public Trainer API with HOST (pinned) images/labels copied every step and the
          loss read back every step
  roofline : the tcgen05 GEMM kernel (all launches of one step, CUDA-event timed) against the measured bf16 peak
  cpu_baseline : the oracle (fp32 PyTorch restatement of the reference path) on the host cores, bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(image_size=224, patch_size=16, num_layers=12, num_heads=12, hidden_dim=768, mlp_dim=3072, num_classes=1000)
TRAIN_GFLOP_PER_IMAGE = 105.383  # SURVEY.md §8d / BASELINE.md §3 (3 x forward, GEMM + attention contractions only)


def fwd_gemm_flops_per_image(cfg):
    S = (cfg["image_size"] // cfg["patch_size"]) ** 2 + 1
    D, Fd, L, C, p = cfg["hidden_dim"], cfg["mlp_dim"], cfg["num_layers"], cfg["num_classes"], cfg["patch_size"]
    patch = 2 * (S - 1) * 3 * p * p * D
    layer = 2 * S * D * 3 * D + 2 * S * D * D + 4 * S * D * Fd
    return patch + L * layer + 2 * D * C


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def oracle_cpu_step_rate(batch, steps, warmup, threads=None):
    """The reference's CPU path (oracle port): fp32 zero_grad + forward + cross-entropy + backward + Adam of ViT-B/16
    (the loop body of vanilla_vit.py:235-239); images/sec."""
    import torch
    from oracle import vit_oracle as O
    # explicit thread count: torchrun exports OMP_NUM_THREADS=1 to its workers, which would time the CPU path on ONE core at N >= 2
    torch.set_num_threads(threads or len(os.sched_getaffinity(0)))
    sd = O.seeded_state_dict(O.vit_param_shapes(**CFG), 0)
    sd = {k: v.requires_grad_(True) for k, v in sd.items()}
    images = O.seeded_images(batch, CFG["image_size"], 1)
    labels = O.seeded_labels(batch, CFG["num_classes"], 2)
    kw = dict(patch_size=CFG["patch_size"], num_layers=CFG["num_layers"], num_heads=CFG["num_heads"])
    opt = torch.optim.Adam(list(sd.values()), lr=1e-4)      # vanilla_vit.py:221
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        loss = torch.nn.functional.cross_entropy(O.vit_forward(sd, images, **kw), labels)
        loss.backward()
        opt.step()
        t1 = time.perf_counter()
        if i >= warmup:
            times.append(t1 - t0)
    dt = sum(times) / len(times)
    return batch / dt, dt, torch.get_num_threads()


def gpu_library_step_rate(dev, batch=128, steps=5, warmup=2):
    """Same-box library bar: the reference's module arithmetic (the oracle's PyTorch operators: cuDNN conv, cuBLAS
    GEMMs, ATen SDPA / LayerNorm / GELU) run EAGERLY on this GPU under torch.autocast(bfloat16) with torch.optim.Adam — what a user
    of the unmodified reference gets by moving it to a B200.  Bounded: a few steps of a smaller batch; images/sec."""
    import torch
    from oracle import vit_oracle as O
    sd = {k: v.to(dev).requires_grad_(True) for k, v in O.seeded_state_dict(O.vit_param_shapes(**CFG), 0).items()}
    images = torch.randn(batch, 3, CFG["image_size"], CFG["image_size"], device=dev)
    labels = torch.randint(0, CFG["num_classes"], (batch,), device=dev)
    kw = dict(patch_size=CFG["patch_size"], num_layers=CFG["num_layers"], num_heads=CFG["num_heads"])
    opt = torch.optim.Adam(list(sd.values()), lr=1e-4)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(warmup + steps):
        if i == warmup:
            e0.record()
        opt.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = O.vit_forward(sd, images, **kw)
        torch.nn.functional.cross_entropy(logits.float(), labels).backward()
        opt.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del sd, opt, images
    torch.cuda.empty_cache()
    return batch / ms * 1e3, ms


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    batch = 8
    steps = max(1, min(args.steps, 5))
    warmup = 1
    ips, dt, cores = oracle_cpu_step_rate(batch, steps, warmup)
    sample = f"{steps} timed fwd+CE+bwd+Adam steps of batch {batch} (fp32, PyTorch CPU ops the reference dispatches to), {warmup} warm-up"
    line = {"impl": "reference", "metric": "ViT-B/16 train images/sec", "value": ips, "unit": "images/sec", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ViT-B/16 224x224 training step, synthetic ImageNet-shaped batch (BASELINE.json configs[1])",
                       "batch_per_step": batch, "where": "host CPU cores"},
            "cpu_baseline": {"value": ips, "unit": "images/sec", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": ips, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-roofline", action="store_true", help="skip the instrumented eager steps (used under ncu for the launch list)")
    ap.add_argument("--no-extras", action="store_true", help="skip the same-GPU library baseline and the cfg3/4/5 measurements")
    ap.add_argument("--config", default="vit_b16_train", choices=["vit_b16_train", "deit_s_distill", "vit_l_infer", "detr_enc"],
                    help="BASELINE.json configs[1] (default, the headline) or configs[2] / [3] / [4] as a stand-alone line")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.config != "vit_b16_train":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import configs_bench
        return configs_bench.bench_line(args)

    # stdout carries exactly ONE line (the JSON): everything else a library may print there at the C level — NCCL's version banner
    # under NCCL_DEBUG=VERSION / INFO, which the driver may set to check the communicator — is sent to stderr by pointing fd 1 at fd 2
    # for the lifetime of the run; the line itself is written to the saved descriptor.  NCCL_DEBUG is left as the caller set it.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    import faulthandler
    faulthandler.dump_traceback_later(240, exit=True)   # a hung rank prints where it is stuck instead of burning the time limit
    import torch
    import torch.distributed as dist
    from vitb200 import ops
    from vitb200.trainer import Trainer
    from vitb200.vit import ViT

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    reducer = None
    if world > 1:
        # stdout carries the single JSON line; NCCL's own log (the driver may ask for NCCL_DEBUG=INFO/VERSION to check the
        # communicator's rank count) is left on and routed to stderr
        if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
            os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
        dist.init_process_group(backend="nccl", device_id=dev)
        from vitb200.dp import GradReducer
        reducer = GradReducer(bucket_bytes=int(os.environ.get("VITB200_DP_BUCKET_MB", "48")) << 20)
    W = max(3, args.warmup)
    K = args.steps
    B = args.batch

    torch.manual_seed(0)
    model = ViT(CFG["image_size"], CFG["patch_size"], CFG["num_layers"], CFG["num_heads"], CFG["hidden_dim"], CFG["mlp_dim"], 0.0, 0.0,
                CFG["num_classes"])
    with torch.no_grad():  # random-init weights of the architecture; un-zero the head so every gradient is live
        model.heads.head.weight.normal_(std=0.02)
        model.class_token.normal_(std=0.02)
    model = model.to(dev)
    model.train()
    trainer = Trainer(model, lr=1e-4, reducer=reducer, use_cuda_graph=not args.no_graph)

    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    host_images = torch.randn(B, 3, CFG["image_size"], CFG["image_size"], generator=g).pin_memory()
    host_labels = torch.randint(0, CFG["num_classes"], (B,), generator=g).pin_memory()
    images = host_images.to(dev)
    labels = host_labels.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing (value) ----------------
    for _ in range(W):
        trainer.step(images, labels)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        loss = trainer.step(images, labels)
    e1.record()
    barrier()
    launches = trainer.launches_per_step
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    final_loss = loss.item()
    value = world * B * K / (ms / 1e3)

    # ---------------- end-to-end timing (host buffers in, loss out, every step) ----------------
    e2e = None
    if not args.no_e2e:
        copy_stream = torch.cuda.Stream(device=dev)
        bufs = [(torch.empty_like(images), torch.empty_like(labels)) for _ in range(2)]
        loss_host = torch.zeros(1).pin_memory()
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]

        e2e_dbg = os.environ.get("VITB200_BENCH_E2E_DEBUG", "")   # diagnostics only: "noh2d" / "nod2h" drop one leg (the line is then not an e2e number)

        def prefetch(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[i & 1])
                if e2e_dbg != "noh2d":
                    bufs[i & 1][0].copy_(host_images, non_blocking=True)
                    bufs[i & 1][1].copy_(host_labels, non_blocking=True)
                ready[i & 1].record(copy_stream)

        def e2e_loop(n):
            for c in consumed:
                c.record(torch.cuda.current_stream())
            prefetch(0)
            for i in range(n):
                if i + 1 < n:
                    prefetch(i + 1)
                torch.cuda.current_stream().wait_event(ready[i & 1])
                l = trainer.step(bufs[i & 1][0], bufs[i & 1][1])
                consumed[i & 1].record(torch.cuda.current_stream())
                if e2e_dbg != "nod2h":
                    loss_host.copy_(l, non_blocking=True)

        e2e_loop(3)
        barrier()
        e0.record()
        e2e_loop(K)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B * K / (t.item() / 1e3), "unit": "images/sec",
               "h2d_bytes_per_step": host_images.numel() * 4 + host_labels.numel() * 8, "d2h_bytes_per_step": 4}
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)

    # ---------------- roofline of the dominant kernel (tcgen05 GEMM), CUDA events around every launch of one step ----------
    roofline = None
    if not args.no_roofline:  # every rank runs the instrumented step (it contains the gradient all-reduce); rank 0 reports
        peak, peak_sus, hbm, src = measured_peaks()
        rec = []
        orig = ops.gemm

        def timed_gemm(A, Bm, C, **kw):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = orig(A, Bm, C, **kw)
            b.record()
            Mr, Kr = (A.shape[-2], A.shape[-1]) if kw.get("a_major", 0) == 0 else (A.shape[-1], A.shape[-2])
            Nr = Bm.shape[-2] if kw.get("b_major", 0) == 0 else Bm.shape[-1]
            nb = A.shape[0] if A.dim() == 3 else 1
            rec.append((a, b, 2.0 * Mr * Nr * Kr * nb))
            return out

        ops.gemm = timed_gemm
        was_graph = trainer.use_cuda_graph
        trainer.use_cuda_graph = False      # eager steps so that every GEMM launch can be bracketed by events
        # back-to-back steps: the first ones bring the GPU back to the power-capped clocks of the timed region (a single step after an
        # idle gap would run at boost clocks and flatter the kernel), the last ROOF_STEPS are measured
        ROOF_WARM, ROOF_STEPS = 4, 4
        for i in range(ROOF_WARM + ROOF_STEPS):
            if i == ROOF_WARM:
                rec.clear()
            trainer.step(images, labels)
        torch.cuda.synchronize()
        trainer.use_cuda_graph = was_graph
        ops.gemm = orig
        gemm_ms = sum(a.elapsed_time(b) for a, b, _ in rec) / ROOF_STEPS
        gemm_flops = sum(f for _, _, f in rec) / ROOF_STEPS
        n_gemm = len(rec) // ROOF_STEPS
        achieved = gemm_flops / (gemm_ms / 1e3) / 1e12
        # DRAM traffic per launch from the committed ncu --set full capture of the six representative ViT-B launches
        # (profiles/r1b_gemm_ncu.json; each shape occurs 12x per step, the wgrad shape stands for the 4 wgrad GEMMs per layer)
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r1b_gemm_ncu.json")) as fh:
                prof = json.load(fh)
            traffic = sum(p["dram_bytes"] for p in prof) / len(prof)
        except Exception:
            pass
        roofline = {"bound": "tensor", "kernel": "vb::gemm_kernel (tcgen05, all launches of one step)", "achieved": achieved,
                    "peak": peak_sus, "unit": "TFLOP/s", "frac": achieved / peak_sus, "traffic": traffic,
                    "traffic_note": "mean dram read+write bytes per launch over the 6 profiled ViT-B/16 GEMM shapes (ncu --set full, "
                                    "profiles/r1b_gemm_ncu_summary.txt); algorithmic bytes of the same launches are within 0.83-1.01x",
                    "peak_kind": f"bf16_tflops_sustained ({src}); burst peak {peak}", "frac_of_burst": achieved / peak,
                    # the whole training step (all kernels, model FLOPs of SURVEY.md §8d) against both measured peaks — the figure
                    # north_star's ">= 60 % of tensor peak" refers to
                    "step_tflops": world * B * K / (ms / 1e3) / world * TRAIN_GFLOP_PER_IMAGE / 1e3,
                    "step_frac_of_burst": B * K / (ms / 1e3) * TRAIN_GFLOP_PER_IMAGE / 1e3 / peak,
                    "step_frac_of_sustained": B * K / (ms / 1e3) * TRAIN_GFLOP_PER_IMAGE / 1e3 / peak_sus,
                    "launches_per_step": n_gemm, "gemm_ms_per_step": gemm_ms, "gemm_share_of_step": gemm_ms / (ms / K),
                    "algorithmic_gflop_per_launch_avg": gemm_flops / n_gemm / 1e9,
                    "how": f"CUDA events around every GEMM launch of {ROOF_STEPS} consecutive eager steps after {ROOF_WARM} warm ones (steady-state clocks)"}

    # ---------------- CPU baseline (oracle port) on the host cores, bounded sample ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ips, dt, cores = oracle_cpu_step_rate(8, 2, 1)
        cpu = {"value": ips, "unit": "images/sec", "cores": cores, "kind": "port",
               "sample": "2 timed fwd+CE+bwd+Adam steps of batch 8 after 1 warm-up, fp32 oracle (PyTorch CPU ops the reference dispatches to)"}

    # ---------------- same-box library bar + the other BASELINE.json configs (bounded; 1-GPU runs only) ----------------
    lib = None
    others = None
    if rank == 0 and world == 1 and not args.no_extras:
        del trainer
        torch.cuda.empty_cache()
        try:
            ips, lms = gpu_library_step_rate(dev)
            lib = {"value": ips, "unit": "images/sec", "ms_per_step": lms, "kind": "oracle restatement of the reference modules, eager "
                   "PyTorch (cuDNN/cuBLAS/ATen) under torch.autocast(bfloat16) + torch.optim.Adam on this GPU", "sample": "5 timed steps of batch 128"}
        except Exception as e:  # the headline must not depend on it
            lib = {"unavailable": f"{type(e).__name__}: {e}"}
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import configs_bench
            others = configs_bench.summary(dev, peak)
        except Exception as e:
            others = {"unavailable": f"{type(e).__name__}: {e}"}

    if rank == 0:
        step_tflops = value / world * TRAIN_GFLOP_PER_IMAGE / 1e3
        peak, peak_sus, hbm, src = measured_peaks()
        line = {"metric": "ViT-B/16 train images/sec", "value": value, "unit": "images/sec", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": "ViT-B/16 224x224 training step (fwd + CE + bwd + Adam), synthetic ImageNet-shaped batch, "
                                       "random-init weights (BASELINE.json configs[1])",
                           "per_gpu_batch": B, "global_batch": B * world, "parallelism": f"dp{world}",
                           "l2_policy": "inputs+activations (>10 GB/step) far exceed the 126 MB L2; no explicit flush",
                           "residual_stream": "fp32", "operands": "bf16, fp32 accumulate",
                           "cuda_graph": not args.no_graph},
                "per_gpu_tflops": step_tflops, "mfu_vs_burst_peak": step_tflops / peak, "mfu_vs_sustained_peak": step_tflops / peak_sus,
                "final_loss": final_loss, "gpu_launches": launches, "clocks": sampler.summary(), "e2e": e2e, "roofline": roofline,
                "cpu_baseline": cpu, "gpu_library_baseline": lib, "other_configs": others}
        emit(line)
    faulthandler.cancel_dump_traceback_later()
    sys.stdout.flush()
    if world > 1:
        # the result is out; nothing below may turn a finished measurement into a hung job (communicator teardown with captured
        # collectives has been seen to block): leave hard after 20 s
        threading.Timer(20.0, lambda: os._exit(0)).start()
        try:
            trainer.close()
        except NameError:
            pass
        if os.environ.get("VITB200_BENCH_DEBUG"):
            print(f"[rank {rank}] before destroy_process_group", file=sys.stderr, flush=True)
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        if os.environ.get("VITB200_BENCH_DEBUG"):
            print(f"[rank {rank}] after destroy_process_group", file=sys.stderr, flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
