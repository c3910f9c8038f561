#!/usr/bin/env python
"""Benchmark of the CycleGAN 256x256 training step (BASELINE.json metric: train images/sec).

    python bench.py --gpus N --steps K --warmup W            # B200 path (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the stand-in on the host CPU cores

One "step" = one full CycleGAN optimisation step (6 generator forwards, G backward + Adam, D forward/
backward + Adam) on one batch of synthetic image pairs; one "image" = one (A, B) pair.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import threading
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cyclegan_256_train_images_per_sec"
UNIT = "images/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(source="measured (MEASURED_PEAKS.json)", hbm_gbs=p["hbm_gbs"], tf_burst=p["bf16_tflops"],
                    tf_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]))
    return dict(source="fallback (B200_PROFILING.md)", hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            with open(self.path) as f:
                for line in f:
                    parts = [p.strip() for p in line.split(",")]
                    if len(parts) < 9:
                        continue
                    try:
                        sm.append(float(parts[1]))
                        mx.append(float(parts[2]))
                    except ValueError:
                        continue
                    for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                         parts[5:9]):
                        if val.lower().startswith("active"):
                            reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the stand-in oracle on the host cores
# ----------------------------------------------------------------------------------------------
def time_standin(batch: int, size: int, steps: int, warmup: int):
    import torch
    from oracle import cyclegan_standin as ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    G_AB, G_BA, D_A, D_B = ref.build_models(seed=0)
    tr = ref.CycleGANTrainer(G_AB, G_BA, D_A, D_B)
    real_A, real_B = ref.synthetic_pair(batch, size, seed=1234)
    for _ in range(warmup):
        tr.train_step(real_A, real_B)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        tr.train_step(real_A, real_B)
        times.append(time.perf_counter() - t0)
    return times, cores


def run_reference(args):
    """the stand-in's own CPU train step on all host cores, with the SAME --steps / --warmup as the B200 arm; only when
    that would exceed ~4 minutes (slow hosts: ~5 s per step on 8 cores) is the number of timed steps cut, and the
    line says so (`steps_requested`)"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    budget_s = 240.0
    t0 = time.perf_counter()
    probe, cores = time_standin(args.batch, args.size, 1, 1)  # 1 warm-up (oneDNN primitive creation) + 1 probe step
    est = probe[0]
    first = time.perf_counter() - t0
    warmup = max(args.warmup, 1)
    extra_warm = max(0, min(warmup - 1, int((budget_s * 0.25) / est)))
    steps = max(1, min(args.steps, int((budget_s - first - extra_warm * est) / est)))
    times, cores = time_standin(args.batch, args.size, steps, extra_warm)
    warmup = 1 + extra_warm
    sec = sum(times) / len(times)
    value = args.batch / sec
    sample = (f"{steps} timed + {warmup} warm-up full train steps of oracle/cyclegan_standin.py (fp32, torch CPU, "
              f"{cores} threads), batch {args.batch}, {args.size}x{args.size}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "steps_requested": args.steps, "warmup_requested": args.warmup,
        "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"CycleGAN {args.size}x{args.size} ResNet-9blk + 70x70 PatchGAN, batch {args.batch}, "
                               "one full train step (stand-in, CPU)", "batch_per_gpu": args.batch, "size": args.size},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------
class StallWatchdog:
    """A device-side stall must not hold the caller until its own time limit: if no phase of the bench completes for
    `limit_s` seconds, say so on stdout / stderr and leave the process (the step graph is replayed asynchronously, so
    a stalled kernel would otherwise block the next synchronisation forever)."""

    def __init__(self, limit_s: float):
        self.limit_s, self.t, self.what = limit_s, time.time(), "start"
        threading.Thread(target=self._run, daemon=True).start()

    def beat(self, what: str):
        self.t, self.what = time.time(), what

    def _run(self):
        while True:
            time.sleep(5.0)
            if time.time() - self.t > self.limit_s:
                msg = f"bench.py: no progress for {self.limit_s:.0f} s after phase '{self.what}' (device stall?)"
                print(json.dumps({"error": msg}), flush=True)
                print(msg, file=sys.stderr, flush=True)
                os._exit(3)


def run_b200(args):
    import torch
    import torch.distributed as dist

    import unpaired_image_generation_b200 as cgb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (B200 arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # rank 0 prints ONE JSON line on stdout: NCCL writes its version banner to stdout when NCCL_DEBUG is set
        # (by us or by the environment), so stdout points at stderr while the communicator is created
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)  # forces communicator creation (and the banner) now
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    peaks = load_peaks()
    batch, size = args.batch, args.size
    dog = StallWatchdog(float(os.environ.get("CGB_BENCH_STALL_S", "300")))

    G_AB, G_BA = cgb.Generator(seed=1), cgb.Generator(seed=2)
    D_A, D_B = cgb.Discriminator(seed=3), cgb.Discriminator(seed=4)
    tr = cgb.CycleGANTrainer(G_AB, G_BA, D_A, D_B)
    g = torch.Generator().manual_seed(1234 + rank)
    host_A = (torch.rand(batch, 3, size, size, generator=g) * 2 - 1).pin_memory()
    host_B = (torch.rand(batch, 3, size, size, generator=g) * 2 - 1).pin_memory()
    dev_A, dev_B = host_A.cuda(), host_B.cuda()
    eng = tr._ensure_engine(dev_A)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        # inputs already resident in HBM; no host round trip inside the step
        if world == 1:
            with torch.cuda.stream(tr.stream):
                eng.stage_inputs(dev_A, dev_B)  # device-to-device copy into the staging buffers the step graph reads
                eng.train_step()
        else:
            tr._train_step_dp_nosync(eng, lambda: eng.stage_inputs(dev_A, dev_B))

    # ---- value: whole-job throughput, device-resident inputs.  The timed region is EXACTLY --steps steps, bracketed by
    # barrier + synchronize; it is repeated back to back until at least ~1 s of device time has been measured (20 steps
    # are only 0.1 s) and the value is the mean over all repetitions, each reduced with MAX over the ranks.
    dog.beat("engine created")
    for _ in range(max(args.warmup, 3)):
        device_step()
    barrier()
    dog.beat("warm-up")
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    rep_ms = []
    reps = 1
    while len(rep_ms) < reps:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(tr.stream)
        for _ in range(args.steps):
            device_step()
        ev1.record(tr.stream)
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rep_ms.append(float(t.item()))
        dog.beat(f"timed repetition {len(rep_ms)}")
        if len(rep_ms) == 1:
            reps = int(min(50, max(1, -(-1000.0 // rep_ms[0]))))  # same on every rank: derived from the reduced time
    clocks = sampler.stop() if rank == 0 else None
    ms_step = sum(rep_ms) / len(rep_ms) / args.steps
    value = world * batch / (ms_step * 1e-3)

    # ---- e2e: public API, pinned HOST inputs, losses read back every step
    for _ in range(2):
        tr.train_step(host_A, host_B)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = args.steps
    for _ in range(e2e_steps):
        losses = tr.train_step(host_A, host_B)
        dog.beat("e2e step")
    barrier()
    e2e_sec = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_sec], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * batch / float(te.item())
    h2d = 2 * batch * 3 * size * size * 4
    d2h = len(cgb.LOSS_KEYS) * 4

    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel class (tcgen05 implicit-GEMM fprop/dgrad), timed live with CUDA events
        with torch.cuda.stream(tr.stream):
            ms_ig, n_ig, fl_ig = eng.profile_kind(1, reps=5)
            ms_wg, n_wg, fl_wg = eng.profile_kind(2, reps=5)
            ms_wd, n_wd, fl_wd = eng.profile_kind(3, reps=3)
            ms_pw, n_pw, _ = eng.profile_kind(4, reps=5)
            # graph-replayed segments timed separately (forward only / whole G phase / D phase / optimisers)
            seg_ms = {}
            for name, seg in (("forward_6_generators", 5), ("G_phase_incl_forward", 1), ("D_phase", 2), ("adam_G", 3), ("adam_D", 4)):
                for _ in range(3):
                    eng.run_segment(seg)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(tr.stream)
                for _ in range(5):
                    eng.run_segment(seg)
                e1.record(tr.stream)
                torch.cuda.synchronize()
                seg_ms[name] = e0.elapsed_time(e1) / 5
            # the FLOP-dominant instantiation: residual-block 3x3 convs (108 of the fprop launches, 83 % of FLOPs)
            res_fprop = res_dgrad = res_wgrad = None
            try:
                os.environ["CGB_PROFILE_OPS"] = "1"
                txt = eng.timeline()
                os.environ.pop("CGB_PROFILE_OPS", None)
                def pick(prefix):
                    vals = [float(l.split()[-2]) for l in txt.splitlines() if l.startswith(prefix)]
                    return sum(vals) / len(vals) if vals else None
                res_fprop, res_dgrad, res_wgrad = pick("fprop res."), pick("dgrad res."), pick("wgrad res.")
            except Exception:
                os.environ.pop("CGB_PROFILE_OPS", None)
        # Dominant kernel: igemm_patch_kernel on the residual-block conv (256->256, 3x3 reflect, 64x64 map at 256x256):
        # 216 of the 330 fprop/dgrad launches of a step and 83 % of its conv FLOPs.  `achieved` = algorithmic FLOPs of
        # one launch (2 * pixels * 256 * 256 * 9; dgrad: the same taps on the 66x66 padded domain, counted on the
        # 64x64 map) / the average launch duration measured live with CUDA events on the launching stream.
        res_flops = 2.0 * batch * (size // 4) ** 2 * 256 * 256 * 9
        agg = fl_ig / (ms_ig * 1e-3) / 1e12
        if res_fprop and res_dgrad:
            us = 0.5 * (res_flops / (res_fprop * 1e12) + res_flops / (res_dgrad * 1e12)) * 1e6
            achieved = res_flops / (us * 1e-6) / 1e12
        else:
            us, achieved = None, agg
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch from the tracked `ncu --set full` capture of this
        # kernel (never measured in this run: ncu replays kernels); profiles/ncu_traffic.json names the source file
        ncu_traffic, traffic_source = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                ent = json.load(f).get(f"res_conv_fprop_b{batch}_{size}")
            if ent:
                ncu_traffic, traffic_source = ent["dram_bytes_per_launch"], ent["source"]
        except Exception:
            pass
        roofline = {
            "bound": "tensor",
            "kernel": "igemm_patch_kernel<BN,MT,KPS,KA,CG=2> (persistent tcgen05 patch-resident implicit GEMM on CTA pairs), residual-block conv "
                      "256->256 3x3 reflect, fprop + dgrad",
            "achieved": achieved, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["tf_sustained"],
            "traffic": ncu_traffic, "traffic_source": traffic_source,
            "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
            "us_per_launch": us, "flops_per_launch": res_flops, "launches_per_step": 216,
            "algorithmic_bytes_per_launch": 2.0 * batch * ((size // 4 + 2) ** 2 + (size // 4) ** 2) * 256 + 2.0 * 256 * 256 * 9,
            "res_block_conv_tflops": {"fprop": res_fprop, "dgrad": res_dgrad, "wgrad": res_wgrad},
            "all_igemm_launches": {"achieved": agg, "frac": agg / peaks["tf_sustained"], "launches_per_step": n_ig,
                                   "ms_per_step_serial": ms_ig, "flops_per_step": fl_ig},
            "other_kernels": {
                "wgrad (wgrad_pair_kernel on the stride-1 layers, wgrad_kernel on the rest; tcgen05)": {"ms_per_step": ms_wg, "launches": n_wg,
                                          "tflops": fl_wg / (ms_wg * 1e-3) / 1e12 if ms_wg > 0 else None},
                "wgrad_small(im2col4 + tcgen05 GEMM, 3-/1-channel layers)": {"ms_per_step": ms_wd, "launches": n_wd,
                                                   "tflops": fl_wd / (ms_wd * 1e-3) / 1e12 if ms_wd > 0 else None},
                "instnorm_pointwise": {"ms_per_step": ms_pw, "launches": n_pw},
            },
            "segments_ms": seg_ms,
            "step_conv_tflops": eng.conv_flops_per_step / (ms_step * 1e-3) / 1e12,
            "step_conv_frac_of_peak": eng.conv_flops_per_step / (ms_step * 1e-3) / 1e12 / peaks["tf_sustained"],
        }
        # ---- cpu_baseline: the stand-in on this box's host cores, bounded sample
        cpu = None
        dog.beat("roofline profiling")
        if not args.no_cpu_baseline and world == 1:  # rank 0 at N = 1 only (the other ranks would idle in the barrier)
            times, cores = time_standin(1, size, 2, 1)
            dog.beat("cpu baseline")
            sec = min(times)
            cpu = {"value": 1.0 / sec, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"best of 2 full train steps after 1 warm-up of oracle/cyclegan_standin.py (fp32 torch CPU, "
                             f"{cores} threads), batch 1, {size}x{size}"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "timed_repetitions": len(rep_ms), "timed_steps_total": len(rep_ms) * args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"CycleGAN {size}x{size} ResNet-9blk + 70x70 PatchGAN, batch {batch} per GPU, bf16 "
                                   "storage / fp32 accumulate, full train step (BASELINE.json "
                                   + ("configs[1])" if (batch, size) == (1, 256) else
                                      "configs[3] geometry)" if size == 512 else "configs[2] geometry)"),
                       "batch_per_gpu": batch, "global_batch": batch * world, "size": size,
                       "parallelism": f"dp{world}",
                       "l2": f"per-step working set {eng.workspace_bytes / 2**20:.0f} MiB >> 126 MB L2 (no explicit flush)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": eng.launches_per_step * args.steps,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "losses_last_step": losses,
        }
    # ---- the other training configurations of BASELINE.json at THIS number of GPUs (every rank takes part):
    # configs[2] = 256x256, batch 8 per GPU; configs[3] = 512x512, batch 4 per GPU.  Same timing rules as the headline.
    del tr, eng
    torch.cuda.empty_cache()
    extras = {}
    if not args.no_extra_configs:
        for key, (xb, xs) in (("configs2_256_b8", (8, 256)), ("configs3_512_b4", (4, 512))):
            if (xb, xs) == (batch, size):
                continue
            dog.beat("before " + key)
            try:
                extras[key] = measure_extra(cgb, torch, dist, world, rank, xb, xs, max(3, args.steps // 2), peaks)
            except Exception as ex:  # never lose the headline line
                extras[key] = {"error": str(ex)[:200]}
            torch.cuda.empty_cache()
    if rank == 0:
        line.update(extras)
        if "configs2_256_b8" in extras:
            line["extra_batch"] = extras["configs2_256_b8"]  # (round-1 name of the same datum)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def measure_extra(cgb, torch, dist, world, rank, batch, size, steps, peaks):
    """device-resident whole-job throughput (+ conv roofline on rank 0) of another per-GPU batch / image size on the
    same GPUs: data parallel over `world` ranks, barrier + synchronize on both sides, MAX over the ranks"""
    mods = (cgb.Generator(seed=1), cgb.Generator(seed=2), cgb.Discriminator(seed=3), cgb.Discriminator(seed=4))
    tr = cgb.CycleGANTrainer(*mods)
    g = torch.Generator().manual_seed(99 + rank)
    dev_A = (torch.rand(batch, 3, size, size, generator=g) * 2 - 1).cuda()
    dev_B = (torch.rand(batch, 3, size, size, generator=g) * 2 - 1).cuda()
    eng = tr._ensure_engine(dev_A)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        if world == 1:
            with torch.cuda.stream(tr.stream):
                eng.stage_inputs(dev_A, dev_B)
                eng.train_step()
        else:
            tr._train_step_dp_nosync(eng, lambda: eng.stage_inputs(dev_A, dev_B))

    for _ in range(3):
        step()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(tr.stream)
    for _ in range(steps):
        step()
    ev1.record(tr.stream)
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    out = {"workload": f"CycleGAN {size}x{size}, batch {batch} per GPU, bf16, full train step, dp{world}",
           "batch_per_gpu": batch, "size": size, "n_gpus": world, "value": world * batch / (ms * 1e-3), "unit": UNIT,
           "ms_per_step": ms, "steps": steps, "schedule": "paired" if batch >= 2 else "unpaired",
           "step_conv_tflops_per_gpu": eng.conv_flops_per_step / (ms * 1e-3) / 1e12,
           "step_conv_frac_of_peak": eng.conv_flops_per_step / (ms * 1e-3) / 1e12 / peaks["tf_sustained"]}
    if rank == 0 and world == 1:
        with torch.cuda.stream(tr.stream):
            ms_ig, n_ig, fl_ig = eng.profile_kind(1, reps=3)
            ms_wg, n_wg, fl_wg = eng.profile_kind(2, reps=3)
            ms_pw, n_pw, _ = eng.profile_kind(4, reps=3)
            res = {}
            try:
                os.environ["CGB_PROFILE_OPS"] = "1"
                txt = eng.timeline()
                for key, prefix in (("fprop", "fprop res."), ("dgrad", "dgrad res."), ("wgrad", "wgrad res.")):
                    vals = [float(l.split()[-2]) for l in txt.splitlines() if l.startswith(prefix)]
                    res[key] = sum(vals) / len(vals) if vals else None
            finally:
                os.environ.pop("CGB_PROFILE_OPS", None)
        out.update({"res_block_conv_tflops": res,
                    "res_block_fprop_frac_of_peak": (res.get("fprop") / peaks["tf_sustained"]) if res.get("fprop") else None,
                    "igemm_tflops": fl_ig / (ms_ig * 1e-3) / 1e12,
                    "igemm_frac_of_peak": fl_ig / (ms_ig * 1e-3) / 1e12 / peaks["tf_sustained"],
                    "wgrad_tflops": fl_wg / (ms_wg * 1e-3) / 1e12,
                    "instnorm_pointwise_ms_serial": ms_pw, "instnorm_pointwise_launches": n_pw})
    del tr, eng
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1, help="image pairs per GPU per step (configs[1]: 1)")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extra-batch", type=int, default=8, help="(kept for old scripts; 0 = same as --no-extra-configs)")
    ap.add_argument("--no-extra-configs", action="store_true",
                    help="skip the configs[2] (256x256 batch 8/GPU) and configs[3] (512x512 batch 4/GPU) measurements")
    args = ap.parse_args()
    if args.extra_batch == 0:
        args.no_extra_configs = True
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
