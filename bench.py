"""Benchmark of the ACL-GAN convolutional training step (BASELINE.json metric: training images/s at 256x256).

    python bench.py --gpus N --steps K --warmup W            # this repo's B200 path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the UNMODIFIED reference's CPU path (oracle/_ref) on host cores

One "step" = one step-pair (dis_update + gen_update on the same batch, what the reference's train.py:71-74 executes on
even iterations) on synthetic uniform[-1,1] 256x256 RGB batches, batch 8 per GPU (BASELINE.json configs[1]; weak scaling).
Prints ONE JSON line (rank 0).  `value` times K step-pairs with the inputs already resident in HBM; `e2e` times the same
K step-pairs through the public trainer API with pinned-host inputs copied H2D and the two total losses read back D2H
inside the timed region.  Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
"""
import argparse
import copy
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "acl-gan_b200")
for p in (PKG, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

F_ALG_TFLOP_PER_IMG = {4: 2.623, 3: 2.616}      # necessary conv+linear FLOPs per image per step-pair @256^2


def executed_tflop_per_image(cfg, size, f_alg, subpixel=True):
    """conv FLOPs the kernels EXECUTE per image per step-pair: F_alg (SURVEY.md 8d, 25 taps per output pixel in the two
    up-blocks) minus what the sub-pixel form of nearest-2x-upsample + 5x5 saves (3x3 main convolution with the 4 phases folded
    into N = 9 taps per output pixel, plus the exact 25-tap ring strips: 2 rows x 2 sides x 2W and 2 columns x 2 sides x (2H-4)
    outputs per image).  Decodes per step-pair: 8 forward (3 in dis_update, 5 in gen_update) + 5 x (dgrad + wgrad)."""
    if not subpixel:
        return f_alg
    dim, n_down = cfg["gen"]["dim"], cfg["gen"]["n_downsample"]
    c, h = dim * 2 ** n_down, size // 2 ** n_down
    saved = 0.0
    for _ in range(n_down):
        full = 25.0 * c * (c // 2) * (2 * h) * (2 * h)
        sub = 9.0 * c * 4 * (c // 2) * h * h + 25.0 * c * (c // 2) * (4 * 2 * h + 4 * (2 * h - 4))
        saved += full - sub
        c, h = c // 2, h * 2
    return f_alg - (8 + 2 * 5) * 2.0 * saved / 1e12


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="male2female.yaml")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32x3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-bar", action="store_true",
                    help="skip the extra leg that times the unmodified reference as eager PyTorch/cuDNN on this GPU")
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--profile-step", action="store_true",
                    help="after warm-up run ONE step-pair between cudaProfilerStart/Stop and exit (for ncu "
                         "--profile-from-start off launch lists); prints no bench line")
    return ap.parse_args()


def load_cfg(name):
    import yaml
    with open(os.path.join(PKG, "configs", name)) as f:
        return yaml.safe_load(f)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(burst=d["bf16_tflops"], sustained=d["bf16_tflops_sustained"], hbm=d["hbm_gbs"], src="measured")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 9]
        if not rows:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if r[col].lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=float(rows[0][2]), reasons=sorted(reasons), samples=len(rows),
                    power_w_max=max(float(r[3]) for r in rows))


def _host_threads():
    # intra-op threads: all host cores up to 32 (beyond that torch's CPU convolutions on these small per-image
    # shapes slow down from oversubscription: 128 threads measured 34x slower than 8)
    return min(os.cpu_count() or 1, 32)


def _reference_available():
    try:
        import ref_shim
        return ref_shim.available()
    except Exception:
        return False


def cpu_baseline(cfg, size, steps=1, warmup=0, budget_s=200.0):
    """the reference's CPU path on the host cores, bs=1 step-pairs at `size`^2 (a bounded sample of the bs=8 workload).
    kind "reference": the UNMODIFIED reference modules (oracle/_ref, verbatim copy made by oracle/make_ref.py) under the
    CPU shim of oracle/ref_shim.py, driven through its own aclgan_Trainer.dis_update / gen_update; kind "port": the
    oracle restatement (oracle/aclgan_oracle.py), used only when the reference copy is absent."""
    import torch
    cores = _host_threads()
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(1234)
    x_a = torch.rand(1, 3, size, size, generator=g) * 2 - 1
    x_b = torch.rand(1, 3, size, size, generator=g) * 2 - 1
    if _reference_available():
        import contextlib
        import ref_shim
        _, trainer_mod, _ = ref_shim.import_reference()
        shim, kind = ref_shim.cpu_shim, "reference"
        with shim():
            torch.manual_seed(0)
            tr = trainer_mod.aclgan_Trainer(copy.deepcopy(cfg))

        def pair():
            with shim():
                tr.dis_update(x_a, x_b, cfg)
                tr.gen_update(x_a, x_b, cfg)
        what = "the UNMODIFIED reference (oracle/_ref: trainer.py dis_update + gen_update under the CPU shim)"
    else:
        import aclgan_oracle as O
        torch.manual_seed(0)
        ot = O.OracleTrainer(copy.deepcopy(cfg))
        kind = "port"

        def pair():
            ot.dis_update(x_a, x_b)
            ot.gen_update(x_a, x_b)
        what = "the oracle port (oracle/aclgan_oracle.py)"
    done_warm = 0
    t_first = None
    for _ in range(warmup):
        t0 = time.time()
        pair()
        t_first = time.time() - t0 if t_first is None else t_first
        done_warm += 1
    if t_first is not None and steps * t_first > budget_s:      # keep the whole run within a few minutes
        steps = max(1, int(budget_s / t_first))
    t0 = time.time()
    for _ in range(steps):
        pair()
    dt = (time.time() - t0) / steps
    return dict(value=1.0 / dt, unit="images/s", cores=cores, kind=kind, steps=steps, warmup=done_warm,
                sample="%d step-pair(s) of %s, torch CPU fp32, %d threads of %d host cores, at %dx%d bs=1 after %d warm-up" % (
                    steps, what, cores, os.cpu_count() or 1, size, size, done_warm),
                s_per_step_pair=dt)


def workload_config(args, world):
    B, S = args.batch, args.size
    return dict(workload="%s %dx%d synthetic RGB, bs=%d per GPU, step-pair = dis_update + gen_update" % (args.config, S, S, B),
                global_batch=B * world, parallelism="dp%d" % world, cuda_graphs=not args.no_graphs,
                l2="per-step working set (GBs of activations) far exceeds the 126 MB L2; no explicit flush")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = load_cfg(args.config)
    steps = max(1, args.steps)
    cb = cpu_baseline(cfg, args.size, steps=steps, warmup=max(1, args.warmup))
    line = dict(metric="training images/sec @256x256 (dis_update + gen_update step-pair)", value=cb["value"],
                unit="images/s", n_gpus=args.gpus, steps=cb["steps"], warmup=cb["warmup"],
                ms_per_step=cb["s_per_step_pair"] * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=workload_config(args, int(os.environ.get("WORLD_SIZE", "1"))),
                cpu_baseline=cb, e2e=dict(value=cb["value"], unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def library_bar(cfg, batch, size, steps=3, warmup=2):
    """extra leg (not the reference arm): the UNMODIFIED reference (oracle/_ref) as eager PyTorch -> cuDNN / ATen on the same
    B200 - "the Blackwell kernel to beat" of SURVEY.md 2.2.  cudnn.benchmark = True as train.py:29; (a) stock settings
    (conv TF32 allowed, the torch default) and (b) the same under torch.autocast(bfloat16)."""
    import torch
    if not _reference_available():
        return dict(unavailable="oracle/_ref not present on this box (run __graft_entry__.build() where /root/reference exists)")
    import ref_shim
    _, trainer_mod, _ = ref_shim.import_reference()
    torch.backends.cudnn.benchmark = True
    out = dict(what="unmodified reference trainer.py dis_update + gen_update, eager PyTorch %s + cuDNN on cuda:0, bs=%d %dx%d, "
                    "%d step-pairs after %d warm-up" % (torch.__version__, batch, size, size, steps, warmup))
    g = torch.Generator().manual_seed(1234)
    x_a = (torch.rand(batch, 3, size, size, generator=g) * 2 - 1).cuda()
    x_b = (torch.rand(batch, 3, size, size, generator=g) * 2 - 1).cuda()
    import contextlib
    for tag, ctx in (("tf32_default", contextlib.nullcontext), ("bf16_autocast", lambda: torch.autocast("cuda", dtype=torch.bfloat16))):
        try:
            torch.manual_seed(0)
            tr = trainer_mod.aclgan_Trainer(copy.deepcopy(cfg)).cuda()
            with ctx():
                for _ in range(warmup):
                    tr.dis_update(x_a, x_b, cfg)
                    tr.gen_update(x_a, x_b, cfg)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    tr.dis_update(x_a, x_b, cfg)
                    tr.gen_update(x_a, x_b, cfg)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[tag] = dict(images_per_s=batch / (ms * 1e-3), ms_per_step_pair=ms)
            del tr
            torch.cuda.empty_cache()
        except Exception as ex:      # noqa: BLE001 - a measurement leg must never take the bench line down
            out[tag] = dict(error="%s: %s" % (type(ex).__name__, str(ex)[:200]))
    return out


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the roofline kernel from the committed ncu --set full capture
# (profiles/r2_top_kernels.md): 19.19 MB read (input plane 17.84 MB + packed weights 1.18 MB), the 16.8 MB output stays in L2
ROOFLINE_TRAFFIC_BYTES = 19.19e6


def kernel_roofline(eng_precision, batch, pk):
    """dominant kernel = the 3x3 256->256 reflect-pad conv of the residual blocks (forward implicit GEMM,
    igemm_seg_pair_kernel: 16 of the 28 generator convs per encode+decode), timed alone with CUDA events, L2 flushed
    between launches.  Algorithmic work per launch = 2 * (batch*64*64) * 256 * (9*256) FLOP."""
    import ctypes as C
    import torch
    import aclgan_native as N
    import engine as E
    eng = E.Engine(eng_precision)
    c, h = 256, 64
    w = torch.nn.Parameter(torch.randn(c, c, 3, 3, device="cuda") * 0.02)
    b = torch.nn.Parameter(torch.zeros(c, device="cuda"))
    arena = E.GradArena(eng.device)
    layer = E.ConvLayer(eng, arena, w, b, 1, 1)
    arena.finalize()
    x = E.ActT(eng, batch, h, h, c, 1, zero=True)
    x.buf.normal_()
    y = eng.new_dense(batch, h, h, c)                   # as in the residual blocks: raw conv output feeding the norm
    o = eng._out_dense(y, b)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()
    for _ in range(3):
        eng.conv_fwd_launch(layer, x, o)
    times = []
    for _ in range(10):
        flush.zero_()                                   # evict L2 between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        eng.conv_fwd_launch(layer, x, o)
        e1.record(stream)
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = sum(times) / len(times)
    segs = 3 if eng_precision == "fp32x3" else 1
    flops = 2.0 * batch * h * h * c * c * 9
    achieved = flops / (ms * 1e-3) / 1e12
    return dict(bound="tensor", kernel="igemm_seg_pair_kernel 3x3 256->256 s1 reflect, %dx64x64 (M=%d N=256 K=2304)" % (batch, batch * h * h),
                achieved=achieved, peak=pk["burst"], unit="TFLOP/s", frac=achieved / pk["burst"],
                traffic=ROOFLINE_TRAFFIC_BYTES if (batch == 8 and eng_precision == "bf16") else None, traffic_unit="bytes of DRAM per launch (ncu)",
                peak_source=pk["src"] + " bf16_tflops (burst: kernel timed alone)", ms=ms, tensor_passes=segs,
                note="achieved counts algorithmic conv FLOPs once; fp32x3 executes 3 bf16 tensor-core passes per product")


def run_b200(args):
    import torch
    import torch.distributed as dist
    import aclgan_native as N
    import trainer as T
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise N.NativeError("bench.py --impl b200 needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl")
    cfg = load_cfg(args.config)
    cfg["precision"] = args.precision
    cfg["cuda_graphs"] = 0 if args.no_graphs else 1
    torch.manual_seed(0)                                    # identical initial weights on every rank
    tr = T.aclgan_Trainer(copy.deepcopy(cfg)).cuda()
    g = torch.Generator().manual_seed(1234 + rank)
    B, S = args.batch, args.size
    host_a = (torch.rand(B, 3, S, S, generator=g) * 2 - 1).pin_memory()
    host_b = (torch.rand(B, 3, S, S, generator=g) * 2 - 1).pin_memory()
    dev_a, dev_b = host_a.cuda(), host_b.cuda()
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        tr.dis_update(dev_a, dev_b, cfg)
        tr.gen_update(dev_a, dev_b, cfg)

    stage_a, stage_b = torch.empty_like(dev_a), torch.empty_like(dev_b)
    loss_host = torch.zeros(2).pin_memory()

    def step_e2e():
        stage_a.copy_(host_a, non_blocking=True)
        stage_b.copy_(host_b, non_blocking=True)
        tr.dis_update(stage_a, stage_b, cfg)
        tr.gen_update(stage_a, stage_b, cfg)
        loss_host[0:1].copy_(tr.loss_dis_total.detach().reshape(1), non_blocking=True)
        loss_host[1:2].copy_(tr.loss_gen_total.detach().reshape(1), non_blocking=True)
        stream.synchronize()                                # the user-visible read of the step's result

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(k):
            fn()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / k

    for _ in range(max(3, args.warmup)):
        step_resident()
    barrier()
    if args.profile_step:
        torch.cuda.profiler.start()
        step_resident()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print("profiled one step-pair (%s kernels of libaclgan_b200.so per step-pair)" % tr.launches_per_step_pair)
        return
    n0 = N.launch_count
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_res = timed(step_resident, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    loss_val = [float(v) for v in loss_host]

    def step_2to1():        # the shipped schedule (D_update 1, G_update 2: train.py:71-74): two iterations = 2 x dis + 1 x gen
        tr.dis_update(dev_a, dev_b, cfg)
        tr.gen_update(dev_a, dev_b, cfg)
        tr.dis_update(dev_a, dev_b, cfg)
    step_2to1()
    ms_2to1 = timed(step_2to1, max(2, args.steps // 2))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    out_dim = cfg["gen"]["output_dim"]
    f_alg = F_ALG_TFLOP_PER_IMG.get(out_dim, 2.623) * (S / 256.0) ** 2
    f_exec = executed_tflop_per_image(cfg, S, f_alg, os.environ.get("ACLGAN_SUBPIXEL", "1") != "0")
    img_s = B * world / (ms_res * 1e-3)
    img_s_e2e = B * world / (ms_e2e * 1e-3)
    launches = getattr(tr, "launches_per_step_pair", None)
    line = dict(
        metric="training images/sec @256x256 (dis_update + gen_update step-pair)", value=img_s, unit="images/s",
        n_gpus=world, steps=args.steps, warmup=max(3, args.warmup), ms_per_step=ms_res, higher_is_better=True,
        scaling="weak", vs_baseline=None, dtype="bf16" if args.precision == "bf16" else "bf16x3(~f32)", data="synthetic",
        config=workload_config(args, world),
        e2e=dict(value=img_s_e2e, unit="images/s", h2d_bytes_per_step=2 * host_a.numel() * 4 + 3 * 2 * B * 8 * 4,
                 d2h_bytes_per_step=8, ms_per_step=ms_e2e, losses=loss_val),
        gpu_launches=launches if launches is not None else -1,
        clocks=clocks,
        step_tensor_util=dict(f_alg_tflop_per_image=f_alg, achieved_tflops=f_alg * img_s / world,
                              peak=pk["sustained"], frac=f_alg * img_s / world / pk["sustained"],
                              executed_tflop_per_image=f_exec, executed_tflops=f_exec * img_s / world,
                              executed_frac=f_exec * img_s / world / pk["sustained"],
                              peak_source=pk["src"] + " bf16_tflops_sustained (kernel inside a long step)",
                              note="frac counts the algorithmic FLOPs of the reference's layer shapes (SURVEY.md 8d F_alg); "
                                   "executed_* counts what the kernels execute: the sub-pixel up-blocks run 9 instead of 25 taps "
                                   "per output pixel (+ ring strips)"),
    )
    line["schedule_2to1"] = dict(value=2 * B * world / (ms_2to1 * 1e-3), unit="images/s (iterations x batch; D_update 1, G_update 2 "
                                 "as shipped in the YAML: 2 dis_update + 1 gen_update per 2 iterations)", ms_per_2_iterations=ms_2to1)
    line["roofline"] = kernel_roofline(args.precision, B, pk)
    del tr
    torch.cuda.empty_cache()
    if not args.no_library_bar and world == 1:
        line["library_bar"] = library_bar(load_cfg(args.config), B, S)
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(load_cfg(args.config), S, steps=3, warmup=1)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
