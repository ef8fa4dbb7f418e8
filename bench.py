#!/usr/bin/env python
"""bench.py — audio-seconds synthesised per wall-second on the BASELINE.json workload.

  python bench.py --gpus 1 --steps 5 --warmup 3            # our arm (C2: medium, 32 x 256 phonemes)
  python bench.py --impl reference --steps 2 --warmup 1    # reference arm: CPU path on host cores
  torchrun ... bench.py --gpus N ...                        # one rank per GPU, weak scaling (32 utts / GPU)

A "step" is one pass of the phoneme-id -> waveform hot path over one batch of synthetic ids.
`value`   : whole-job audio-s/s with ids already in the job (device-resident result, no D2H).
`e2e`     : same metric through the public call (`speak_batch_ids`): host ids in, host waveforms out.
`roofline`: dominant kernel class (HiFi-GAN ResBlock convolutions), CUDA-event timed in the same run.
`cpu_baseline`: the oracle (a port of the reference's onnxruntime graph) on this box's host cores.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# stdout carries exactly one JSON line: NCCL's own banner / debug output ("NCCL version ...") goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 22050
HOP = 256


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json), or null."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """samples from here on belong to the timed region"""
        self.first = len(self.lines)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        first = getattr(self, "first", 0)
        window = "timed region"
        lines = self.lines[first:]
        if not lines:                       # region shorter than one sampling period: use the warm-up samples too
            lines, window = self.lines, "warm-up + timed region"
        for ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def tune_cpu_threads(quality: str, cores: int) -> int:
    """onnxruntime's default (all cores, piper/src/lib.rs:79-86) oversubscribes small B=1 ops on many-core
    hosts; give the CPU arm its best shot: probe a short utterance at a few thread counts and keep the best."""
    best, best_v = min(cores, 8), 0.0
    for t in sorted({min(cores, x) for x in (4, 8, 16, 32, 64)}):
        cpu_reference(quality, 24, 1, t)
        a_, w_ = cpu_reference(quality, 48, 1, t)
        if a_ / w_ > best_v:
            best, best_v = t, a_ / w_
    return best


def cpu_reference(quality: str, n_phonemes: int, n_utts: int, threads: int):
    """Times the oracle (CPU port of the reference's ort graph) B=1 sequentially, like speak_batch
    (piper/src/lib.rs:433-435).  Returns (audio_seconds, wall_seconds)."""
    import torch
    from oracle import vits_oracle as vo
    from sonata_b200 import voicegen, workload
    torch.set_num_threads(threads)
    W = vo.to_torch(voicegen.make_tensors(quality))
    a = vo.arch_of(W)
    g = torch.Generator().manual_seed(5)
    scales = [0.667, 1.0, 0.8]
    audio = 0.0
    t0 = time.perf_counter()
    for u in range(n_utts):
        ids = workload.synthetic_ids(n_phonemes, utt=u)
        ew = torch.randn(1, 2, len(ids), generator=g)
        st = {}
        with torch.inference_mode():
            z_ = None
            # noise for z_p needs y_len: draw after the duration predictor like the graph does
            x, m_p, logs_p = vo.text_encoder(W, torch.as_tensor(ids).view(1, -1), a)
            logw = vo.sdp_reverse(W, x, ew, scales[2], a)
            _, w_ceil, y_len = vo.durations(logw, scales[1])
            ez = torch.randn(1, a["inter"], y_len, generator=g)
            z_p, _ = vo.expand(m_p, logs_p, w_ceil, y_len, ez, scales[0])
            z = vo.flow_reverse(W, z_p, a)
            wav = vo.decoder(W, z, a)
        audio += wav.numel() / SR
    return audio, time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C1", "C2", "C3"])
    ap.add_argument("--backend", type=int, default=int(os.environ.get("SB200_BACKEND", "1")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    from sonata_b200 import workload
    quality, B, NPH = workload.CONFIGS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg_desc = {"workload": f"{args.workload}: synthetic-{quality} (en_US-lessac-{quality} architecture), "
                            f"{B} x {NPH}-phoneme utterances per GPU, scales [0.667,1,0.8]",
                "quality": quality, "batch_per_gpu": B, "phonemes": NPH, "ids_per_utt": 2 * NPH + 2,
                "l2": "working set (>5 GB of activations per step) far exceeds the 126 MB L2",
                "parallelism": f"dp{world}: independent utterance shards, one process per GPU, no data-path collective"}
    cores = os.cpu_count() or 1

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        n_per_step = 1
        threads = tune_cpu_threads(quality, cores)
        for _ in range(max(args.warmup, 0)):
            cpu_reference(quality, NPH, 1, threads)
        audio, wall = 0.0, 0.0
        for s in range(args.steps):
            a_, w_ = cpu_reference(quality, NPH, n_per_step, threads)
            audio += a_; wall += w_
        v = audio / wall
        print(json.dumps({
            "impl": "reference", "metric": "audio-sec/sec", "value": v, "unit": "audio-s/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg_desc,
            "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": threads, "kind": "port", "host_cpus": cores,
                             "sample": f"{n_per_step} utterance(s) of the workload per step, B=1 sequential like speak_batch, "
                                       f"torch threads auto-tuned to {threads} of {cores} host CPUs; "
                                       "PyTorch-CPU restatement of the reference's onnxruntime graph (ort itself is absent offline)"},
            "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return

    # ------------------------------------------------------------------ our arm
    # stdout must carry exactly ONE line (the JSON): native libraries (NCCL's version banner, for one) write to file
    # descriptor 1 directly, so fd 1 points at stderr until the line is printed
    sys.stdout.flush()
    _stdout_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import sonata_b200
    from sonata_b200 import voicegen, shard
    from sonata_b200.job import SynthesisJob
    from sonata_b200 import _native

    if not os.path.exists(_native.LIB_PATH):
        from sonata_b200 import build as _b
        _b.build()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        cfg_path = voicegen.write_voice(voicegen.default_voice_dir(), quality)
    if world > 1:
        dist.barrier()
    cfg_path = voicegen.write_voice(voicegen.default_voice_dir(), quality)
    model = sonata_b200.from_config_path(cfg_path, device=local_rank)
    model.set_backend(args.backend)

    total_utts = B * world
    all_batches = [workload.synthetic_ids(NPH, utt=u) for u in range(total_utts)]
    ids_per_step = sum(len(b) for b in all_batches)
    # this rank's shard: the same LPT partition `shard.scatter_ids` computes, evaluated locally (it is a pure function
    # of the id counts), so the timed passes need no collective -- utterances are independent
    my_idx = shard.lpt_partition([len(b) for b in all_batches], world)[rank] if world > 1 else list(range(total_utts))
    local_batches = [all_batches[i] for i in my_idx]
    lib = _native.lib()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out_cap = int(B * (2 * NPH + 2) * 8 * HOP)     # generous: 8 frames per id
    d_out = torch.empty(out_cap, dtype=torch.float32, device="cuda") if world > 1 else None

    def step_device(utt_batches):
        """device-resident pass; returns (audio_seconds_local, device_ms, job)"""
        job = SynthesisJob(model, utt_batches)
        ms = job.run()
        frames, samples, _ = job.lengths()
        return sum(samples) / SR, ms, job

    # warm-up (the clock sampler starts here so that nvidia-smi is already streaming when the timed region begins)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        _, _, j = step_device(local_batches)
        j.close()

    prof_acc = {}
    barrier()
    launches0 = int(lib.sb200_launch_count())
    sampler.mark()
    t0 = time.perf_counter()
    audio_local, dev_ms = 0.0, 0.0
    for s in range(args.steps):
        a_, ms, j = step_device(local_batches)
        audio_local += a_; dev_ms += ms
        for r in j.profile():
            acc = prof_acc.setdefault(r["name"], {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
            for k in ("ms", "flops", "bytes", "launches"):
                acc[k] += r[k]
        j.close()
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = int(lib.sb200_launch_count()) - launches0

    t = torch.tensor([wall, audio_local, dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        wall_max, audio_total, dev_ms_max = float(tmax[0]), float(tsum[1]), float(tmax[2])
    else:
        wall_max, audio_total, dev_ms_max = wall, audio_local, dev_ms
    value = audio_total / wall_max

    # ---------------- e2e: public call, host ids in -> host waveforms out ----------------
    # Every rank serves its own batch through the public API (`infer_batch_with_values` = speak_batch on ids: ids from
    # host memory, waveforms into pinned host memory).  The shards are independent -- no data-path collective -- so the
    # whole-job number is the sum over ranks over the slowest rank's wall time.  The "one frontend on rank 0" variant
    # (NCCL scatter of the ids, NCCL gather of the waveforms, one device->host copy on rank 0: `shard.py`) is timed
    # separately below and reported as `e2e.frontend_rank0`; it funnels every GPU's audio through one PCIe link.
    def step_e2e():
        auds = model.infer_batch_with_values(local_batches)
        return sum(len(a) for a in auds) / SR

    pinned = {}

    def step_frontend():
        mine = shard.scatter_ids(all_batches if rank == 0 else None)
        job = SynthesisJob(model, mine)
        job.run(d_out.data_ptr(), out_cap)
        _, samples, _ = job.lengths()
        tot = int(sum(samples))
        res = shard.gather_waveforms(d_out[:tot], samples, to_host=False)
        job.close()
        if rank != 0:
            return 0.0
        gl, owner, lens_all = res
        n_audio = 0
        for r_, g_ in enumerate(gl):
            cnt = int(lens_all[owner == r_].sum())
            if r_ not in pinned or pinned[r_].numel() < g_.numel():
                pinned[r_] = torch.empty(g_.numel(), dtype=torch.float32, pin_memory=True)
            pinned[r_][:cnt].copy_(g_[:cnt], non_blocking=True)
            n_audio += cnt
        torch.cuda.synchronize()
        return n_audio / SR

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_audio = 0.0
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(e2e_steps):
        e2e_audio += step_e2e()
    barrier()
    e2e_wall = time.perf_counter() - t0
    te = torch.tensor([e2e_wall, e2e_audio], dtype=torch.float64, device="cuda")
    if world > 1:
        tm = te.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        tsm = te.clone(); dist.all_reduce(tsm, op=dist.ReduceOp.SUM)
        e2e_wall, e2e_audio = float(tm[0]), float(tsm[1])
    e2e_value = e2e_audio / e2e_wall
    d2h_bytes = int(4 * e2e_audio * SR / e2e_steps)
    frontend_value = None
    if world > 1:
        step_frontend()
        barrier()
        t0 = time.perf_counter()
        fa = 0.0
        for _ in range(2):
            fa += step_frontend()
        barrier()
        fw = time.perf_counter() - t0
        tf_ = torch.tensor([fw, fa], dtype=torch.float64, device="cuda")
        tm = tf_.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        frontend_value = float(tm[1]) / float(tm[0])

    if rank == 0:
        peaks = read_peaks()
        # dominant kernel class: HiFi-GAN ResBlock convolutions (dec.mrf*)
        mrf = {k: v for k, v in prof_acc.items() if k.startswith("dec.mrf")}
        mrf_ms = sum(v["ms"] for v in mrf.values()); mrf_l = sum(v["launches"] for v in mrf.values())
        mrf_bytes = sum(v["bytes"] for v in mrf.values()); mrf_flops = sum(v["flops"] for v in mrf.values())
        all_ms = sum(v["ms"] for v in prof_acc.values())
        ach_gbs = mrf_bytes / (mrf_ms * 1e-3) / 1e9 if mrf_ms else 0.0
        roofline = {
            "bound": "hbm", "kernel": ("conv_tc_kernel" if args.backend == 1 else "conv_simt_kernel") + " on dec.mrf* (HiFi-GAN ResBlock dilated Conv1d + residual; largest share of the step)",
            "achieved": ach_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach_gbs / peaks["hbm_gbs"],
            "peak_source": f"{peaks['source']} (MEASURED_PEAKS.json hbm_gbs)" if peaks["source"] == "measured" else "fallback 6.65 TB/s",
            "traffic": ncu_traffic(),
            "launches": mrf_l, "avg_launch_ms": mrf_ms / mrf_l if mrf_l else None,
            "bytes_per_launch": mrf_bytes / mrf_l if mrf_l else None,
            "share_of_step": mrf_ms / all_ms if all_ms else None,
            "achieved_tflops": mrf_flops / (mrf_ms * 1e-3) / 1e12 if mrf_ms else 0.0,
            "tensor_peak_tflops": peaks["bf16_tflops_sustained"],
            "note": "algorithmic bytes = each conv's input + output (+ the residual when it is NOT the conv input, + the accumulated buffer when read-modify-written) once, fp32, + weights; "
                    "FLOPs at 2/MAC over valid rows",
        }
        regions = {k: {"ms_per_step": v["ms"] / args.steps, "tflops": v["flops"] / max(v["ms"], 1e-9) / 1e9,
                       "gbs": v["bytes"] / max(v["ms"], 1e-9) / 1e6, "launches_per_step": v["launches"] / args.steps}
                   for k, v in prof_acc.items()}
        cpu_base = None
        if not args.no_cpu_baseline and world == 1:
            n_s = 4
            threads = tune_cpu_threads(quality, cores)
            a_, w_ = cpu_reference(quality, NPH, n_s, threads)
            cpu_base = {"value": a_ / w_, "unit": "audio-s/s", "cores": threads, "kind": "port", "host_cpus": cores,
                        "sample": f"{n_s} utterances of the workload ({NPH} phonemes each), B=1 sequential like speak_batch, "
                                  f"PyTorch-CPU port of the reference graph, torch threads auto-tuned to {threads} of {cores}"}
        line = {
            "metric": "audio-sec/sec", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * wall_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg_desc,
            "device_ms_per_step": dev_ms_max / args.steps, "audio_s_per_step": audio_total / args.steps,
            "backend": "tcgen05-bf16x2 (flow+decoder), fp32 CUDA cores (encoder, duration predictor)" if args.backend == 1 else "fp32-simt",
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": int(8 * ids_per_step),
                    "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps,
                    "path": "per-rank public call (host ids -> pinned host waveforms), summed over ranks",
                    "frontend_rank0": frontend_value},
            "roofline": roofline, "regions": regions, "cpu_baseline": cpu_base,
        }
        sys.stdout.flush()
        os.dup2(_stdout_fd, 1)
        print(json.dumps(line), flush=True)
        sys.stdout.flush()
        os.dup2(2, 1)
    model.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
