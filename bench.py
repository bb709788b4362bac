#!/usr/bin/env python
"""bench.py - headline measurement of the RAG diffusion sampling path.

Metric (BASELINE.json): denoising-steps/sec at TED shape, B=512 clips per GPU, T=1000
ancestral schedule, classifier-free guidance on (one step = the whole batch advanced one
timestep = 2 denoiser passes per clip + guidance + posterior update).

    python bench.py --gpus 1 --steps 200 --warmup 20
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29511 bench.py --gpus 8 --steps 200 --warmup 20
    python bench.py --impl reference --steps 3 --warmup 1      # CPU arm (oracle port)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_SAMPLE_STEP = {"ted": 337207296, "beat": 400027648}   # BASELINE.md section 3 (algorithmic, CFG = 2 passes)
T_FULL = 1000


def model_args(dims):
    return types.SimpleNamespace(mdm_condm='text', latent_dim=512, ff_size=1024, layers=dims.layers,
                                 cond_mask_prob=0.1, arch='trans_enc', emb_trans_dec=False, dataset='humanml',
                                 lang_model=None, mlpact='silu', diffusion_steps=T_FULL, noise_schedule='cosine',
                                 sigma_small=True, lambda_vel=1.0, lambda_rcxyz=0.0, lambda_fc=0.0,
                                 njoints=dims.njoints)


def ncu_traffic(kernel_impl, dataset, batch, steps_per_launch):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/), or None
    when the capture is for another kernel / workload.  Never measured live: a run under ncu is not a bench."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if kernel_impl == "tc_bf16x3" and dataset == "ted" and batch == 512 and os.path.exists(p):
        d = json.load(open(p))
        if d.get("steps_per_launch", 1) == steps_per_launch:
            return d.get("dram_bytes_per_launch")
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1400.0, "fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md's clocks line), read through NVML
    from a thread of this process (two light queries every 50 ms).  Spawning / polling `nvidia-smi` instead
    measurably stalls the launching thread of a 150 ms timed region (one run: 878 instead of ~1400 steps/s); it
    remains the fallback when pynvml is missing.  Started before the warm-up, marked around the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index      # rows: (t, sm_mhz, max_mhz, [reason names])
        self.t0 = self.t1 = None
        self.nvml, self.stop_flag, self.how = None, False, None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.idx])
            except (ValueError, IndexError):
                pass
        return self.idx

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            bits = [("hw_slowdown", pynvml.nvmlClocksEventReasonHwSlowdown),
                    ("hw_thermal_slowdown", pynvml.nvmlClocksEventReasonHwThermalSlowdown),
                    ("sw_thermal_slowdown", pynvml.nvmlClocksEventReasonSwThermalSlowdown),
                    ("sw_power_cap", pynvml.nvmlClocksEventReasonSwPowerCap)]

            def loop():
                while not self.stop_flag:
                    try:
                        sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                        r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        self.rows.append((time.perf_counter(), float(sm), float(mx), [n for n, b in bits if r & b]))
                    except Exception:
                        pass
                    time.sleep(0.05)

            self.nvml, self.how = pynvml, "nvml"
            threading.Thread(target=loop, daemon=True).start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self._physical_index()), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.how = "nvidia-smi"
            threading.Thread(target=self._read_smi, daemon=True).start()
        except OSError:
            self.proc = None

    def _read_smi(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                self.rows.append((time.perf_counter(), float(r[1]), float(r[2]),
                                  [n for n, v in zip(self.NAMES, r[5:9]) if v.lower().startswith("active")]))
            except (ValueError, IndexError):
                continue

    def wait_first_sample(self, timeout=5.0):
        t = time.perf_counter()
        while self.how is not None and not self.rows and time.perf_counter() - t < timeout:
            time.sleep(0.01)

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.how is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        rows = [r for r in self.rows if self.t0 is not None and self.t0 - 0.06 <= r[0] <= (self.t1 or r[0]) + 0.12]
        window = "timed region"
        if not rows:
            rows, window = list(self.rows), "whole run (no sample fell into the timed region)"
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        reasons = sorted({n for r in rows for n in r[3]})
        return {"sm_mhz": statistics.median(r[1] for r in rows), "sm_max_mhz": max(r[2] for r in rows),
                "reasons": reasons, "samples": len(rows), "window": window, "via": self.how}


def cpu_port_steps_per_s(dims, batch_equiv, n_steps, warmup, sample_batch, threads):
    """The oracle (a port of the reference's CPU path) timed on the host cores.
    One oracle step at `sample_batch` clips, scaled to steps/s at `batch_equiv` clips
    (steps are iso-cost per clip on CPU at these sizes: BASELINE.md section 4)."""
    import torch
    from livelyspeaker_b200 import synthetic
    from oracle import sampler_oracle, schedule_oracle
    torch.set_num_threads(threads)
    sd = synthetic.synth_state_dict(dims, seed=1)
    tab, tmap = schedule_oracle.build("cosine", T_FULL, "")
    y = synthetic.synth_cond(dims, sample_batch)
    tape = sampler_oracle.NoiseTape(seed=0)
    x = tape.draw(sample_batch, dims.njoints, dims.nfeats, 34)
    times = []
    with torch.no_grad():
        for k in range(warmup + n_steps):
            t0 = time.perf_counter()
            x, _ = sampler_oracle.p_sample_step(sd, tab, tmap, x, T_FULL - 1 - k, y, tape, dims.njoints, dims.nfeats)
            if k >= warmup:
                times.append(time.perf_counter() - t0)
    per_step = sum(times) / len(times)
    return (sample_batch / batch_equiv) / per_step, per_step


def gpu_eager_port_steps_per_s(dims, batch, n_steps, dev):
    """The same oracle port run as eager PyTorch on the GPU (fp32, TF32 off): the honest "reference code on a B200"
    comparator of SURVEY.md 8d (the reference tree itself is not on the GPU box).  A reported baseline only."""
    import torch
    from livelyspeaker_b200 import synthetic
    from oracle import sampler_oracle, schedule_oracle
    sd = {k: v.to(dev) for k, v in synthetic.synth_state_dict(dims, seed=1).items()}
    tab, tmap = schedule_oracle.build("cosine", T_FULL, "")
    y = synthetic.synth_cond(dims, batch, device=dev)

    class DevTape(sampler_oracle.NoiseTape):
        def __init__(self):
            self.replay, self.record, self.gen, self.device = None, [], None, dev

        def draw(self, *s_):
            return torch.randn(*s_, device=dev)

        def draw_like(self, x_):
            return torch.randn_like(x_)

    pick, tf32 = sampler_oracle._pick, (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    sampler_oracle._pick = lambda table, i: pick(table, i).to(dev)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        tape = DevTape()
        x = tape.draw(batch, dims.njoints, dims.nfeats, 34)
        with torch.no_grad():
            for k in range(2 + n_steps):
                if k == 2:
                    torch.cuda.synchronize(dev)
                    t0 = time.perf_counter()
                x, _ = sampler_oracle.p_sample_step(sd, tab, tmap, x, T_FULL - 1 - k, y, tape, dims.njoints, dims.nfeats)
            torch.cuda.synchronize(dev)
        return n_steps / (time.perf_counter() - t0)
    finally:
        sampler_oracle._pick = pick
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32


def run_reference_arm(a, dims, rank):
    """--impl reference: the reference's own CPU implementation of the path.  The reference
    tree is not present on the GPU box, so this times oracle/ (kind "port"), which is
    pinned to the reference draw-by-draw (tests/golden/PIN_REPORT.json)."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_steps = max(1, min(a.steps, 8))
    sample_batch = min(256, a.batch)
    value, per_step = cpu_port_steps_per_s(dims, a.batch, n_steps, max(1, min(a.warmup, 1)), sample_batch, threads)
    sample = "%d oracle p_sample steps at B=%d clips (CFG on), %.3f s each, scaled to B=%d" % (
        n_steps, sample_batch, per_step, a.batch)
    # the unit (steps of 512 clips) does not depend on N: the host cores do not scale with --gpus
    line = {"impl": "reference", "metric": "denoising-steps/sec", "value": value,
            "unit": "steps/s (1 step = %d clips advanced one timestep, CFG on)" % a.batch, "n_gpus": a.gpus,
            "steps": n_steps, "warmup": 1, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(a, dims),
            "cpu_baseline": {"value": value, "unit": "steps/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "clips_per_s": value * a.batch / T_FULL}
    print(json.dumps(line), flush=True)


def workload_config(a, dims):
    return {"workload": "TED RAG sampling B=%d/GPU, F=34, J*D=%d, T=%d ancestral, 8-layer/512-d, CFG scale 1.5"
                        % (a.batch, dims.jd, T_FULL) if dims.dataset == "ted" else
                        "BEAT RAG sampling B=%d/GPU, F=34, J*D=%d, T=%d ancestral" % (a.batch, dims.jd, T_FULL),
            "global_batch": a.batch * a.gpus, "timesteps": T_FULL, "sampler": a.sampler,
            "l2": "flushed between timed launches of %d steps (256 MiB memset outside the event bracket)" % a.chunk,
            "steps_per_launch": a.chunk,
            "parallelism": "batch shard x%d, no per-step collective, 1 all_gather at loop end" % a.gpus}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kernel", default="auto", choices=["auto", "simt", "tc_bf16x3", "tc_bf16"])
    ap.add_argument("--dataset", default="ted", choices=["ted", "beat"])
    ap.add_argument("--batch", type=int, default=512, help="clips per GPU")
    ap.add_argument("--sampler", default="ancestral", choices=["ancestral", "ddim"])
    ap.add_argument("--chunk", type=int, default=16, help="loop iterations per launch (ls_step_multi), 1..16")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from livelyspeaker_b200 import synthetic
    dims = synthetic.dims_for(a.dataset)
    if a.impl == "reference":
        run_reference_arm(a, dims, rank)
        return

    import torch
    import torch.distributed as dist
    import livelyspeaker_b200 as ls
    from livelyspeaker_b200 import beat_model_util, sharding
    from livelyspeaker_b200 import gaussian_diffusion as gd

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == a.gpus, "WORLD_SIZE %d != --gpus %d (launch with torch.distributed.run)" % (world, a.gpus)

    B, K, W = a.batch, a.steps, a.warmup
    ddim = a.sampler == "ddim"
    mk = beat_model_util.create_model_and_diffusion if a.dataset == "beat" else ls.create_model_and_diffusion
    model, diffusion = mk(model_args(dims), "")
    sd = synthetic.synth_state_dict(dims, seed=1)
    model.load_state_dict(sd, strict=True)
    model.set_impl(a.kernel)
    cfg = ls.ClassifierFreeSampleModel(model).to(dev).eval()
    eng = model.engine(B)
    shape = (B, dims.njoints, dims.nfeats, 34)

    # conditioning of this rank's shard of the global batch (seed differs per rank)
    y_host = synthetic.synth_cond(dims, B, seed=233 + rank)
    y_pinned = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in y_host.items()}
    y_dev = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in y_host.items()}

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    torch.manual_seed(1000 + rank)

    # ------------------------------------------------------------------ device-resident steps
    eng.set_cond(y_dev, force=True)
    scale = y_dev["scale"].float().contiguous()
    x = torch.randn(*shape, device=dev)
    perm_like = torch.empty(34, B, dims.njoints, dims.nfeats, device=dev).permute(1, 2, 3, 0)
    # Warm-up in whole launches: the second one is the first to take the steady-state route (one-launch draws, cached
    # allocator blocks), and at least ~0.3 s of them: the timed region of 200 steps is only 0.14 s long, and runs
    # timed right after a 2-launch warm-up came out up to 20 % low while kernel and e2e (timed later) did not move.
    Cw = max(1, min(a.chunk, ls.MAX_FUSED_STEPS))
    W = max(-(-W // Cw), -(-384 // Cw)) * Cw
    idx = [T_FULL - 1 - (k % T_FULL) for k in range(W + K)]
    params = [diffusion.step_params(i, ddim=ddim, eta=0.0, clip_denoised=False) for i in idx]

    C = max(1, min(a.chunk, ls.MAX_FUSED_STEPS))

    def run_chunk(k0, n, x_in):
        """n consecutive loop iterations as p_sample_loop runs them on the fused route: the reference's
        per-step draws (for full chunks after the first: one torch-compatible launch, or a CUDA-graph replay of the
        3n torch kernels when that kernel does not verify against torch), then ONE
        ls_step_multi launch."""
        graphed = None
        if n == C and n > 1 and k0 > 0 and diffusion.graph_draws:
            # like p_sample_loop: full chunks after the first replay their 3n draws from a CUDA graph
            graphed = gd.chunk_draws(eng, n, B, 512, perm_like)
            if graphed is None:
                diffusion.graph_draws = False
        if graphed is not None:
            e_c, e_u, nz = graphed.draw()
        else:
            e_c = [torch.randn(B, 1, 512, device=dev) for _ in range(n)]
            e_u = [torch.randn(B, 1, 512, device=dev) for _ in range(n)]
            nz = [torch.randn_like(perm_like) for _ in range(n)]
        xs = torch.empty((n,) + tuple(x_in.shape), device=dev)
        if n == 1:
            eng.step(params[k0], x_in, e_c[0], e_u[0], nz[0], scale, xs[0], None)
        else:
            eng.step_multi(params[k0:k0 + n], x_in, e_c, e_u, nz, scale, xs, None)
        return xs[n - 1]

    def chunks(k0, n):
        return [(k0 + i, min(C, n - i)) for i in range(0, n, C)]

    clocks = ClockSampler(local_rank)
    clocks.start()
    for k0, n in chunks(0, W):
        x = run_chunk(k0, n, x)
    sync_all()
    clocks.wait_first_sample()
    launches0 = eng.launch_count()
    timed = chunks(W, K)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in timed]
    import gc
    gc.collect()
    gc.disable()                               # no collector pause on the launching thread inside the timed region
    sync_all()
    clocks.mark_begin()
    for j, (k0, n) in enumerate(timed):
        flush_buf.zero_()                      # evict L2 (126 MB) - outside the event bracket
        ev[j][0].record()
        x = run_chunk(k0, n, x)
        ev[j][1].record()
    sync_all()
    clocks.mark_end()
    gc.enable()
    launches = eng.launch_count() - launches0
    clk = clocks.stop()
    t_ms = sum(s.elapsed_time(e) for s, e in ev)
    assert torch.isfinite(x).all(), "sampler diverged"

    # dominant kernel alone (the ls_step_multi launch without the torch RNG launches), for the roofline
    g = torch.Generator(device=dev).manual_seed(5)
    nk = min(C, K)
    e_c = [torch.randn(B, 1, 512, device=dev, generator=g) for _ in range(nk)]
    e_u = [torch.randn(B, 1, 512, device=dev, generator=g) for _ in range(nk)]
    nz = [torch.randn(*shape, device=dev, generator=g) for _ in range(nk)]
    xs = torch.empty((nk,) + tuple(x.shape), device=dev)
    kk = max(3, min(K, 50) // nk)
    evk = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(kk)]
    l0 = eng.launch_count()
    for k in range(kk):
        flush_buf.zero_()
        evk[k][0].record()
        if nk == 1:
            eng.step(params[W], x, e_c[0], e_u[0], nz[0], scale, xs[0], None)
        else:
            eng.step_multi(params[W:W + nk], x, e_c, e_u, nz, scale, xs, None)
        evk[k][1].record()
    torch.cuda.synchronize()
    kern_ms = sum(s.elapsed_time(e) for s, e in evk) / kk          # per launch of nk steps
    launches_per_step = (eng.launch_count() - l0) / (kk * nk)

    # ------------------------------------------------------------------ end to end (host buffers)
    # The whole T=1000 loop when the run is long enough to afford it (0.8 s on a B200), else K steps.
    n_e2e = T_FULL if K >= 100 else min(K, T_FULL)
    h2d = sum(v.numel() * v.element_size() for k_, v in y_pinned.items()
              if torch.is_tensor(v) and k_ in ("audio_input", "origin_x", "vid_indices", "scale", "emo"))
    sample_fn = diffusion.ddim_sample_loop if ddim else diffusion.p_sample_loop
    diffusion.fused_chunk = C

    def e2e_once(n):
        yk = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in y_pinned.items()}
        if world > 1:
            local = sample_fn(cfg, shape, clip_denoised=False, model_kwargs={"y": yk}, skip_timesteps=T_FULL - n)
            out = sharding.all_gather_samples(local, B * world)[rank * B:(rank + 1) * B]
        else:
            out = sample_fn(cfg, shape, clip_denoised=False, model_kwargs={"y": yk}, skip_timesteps=T_FULL - n)
        return out.to("cpu", non_blocking=False)

    for _ in range(3):                        # warm-up of the e2e path: allocator pools, pinned staging (>= 3 loops of 2 launches)
        e2e_once(min(2 * C, n_e2e))
    sync_all()
    s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s_ev.record()
    res = e2e_once(n_e2e)
    e_ev.record()
    sync_all()
    e2e_ms = max(s_ev.elapsed_time(e_ev), (time.perf_counter() - t0) * 1e3 if world == 1 else 0.0)
    d2h = res.numel() * res.element_size()

    # ------------------------------------------------------------------ reduce over ranks
    stats = torch.tensor([t_ms, e2e_ms, kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    t_ms, e2e_ms, kern_ms = [float(v) for v in stats.tolist()]

    if rank == 0:
        peak, peak_src = measured_peaks()
        steps_per_s = world * K / (t_ms / 1e3)
        flop_step = B * FLOP_PER_SAMPLE_STEP[a.dataset]
        flop_launch = nk * flop_step
        achieved = flop_launch / (kern_ms / 1e3) / 1e12
        impl = eng.get_impl()
        line = {
            "metric": "denoising-steps/sec", "value": steps_per_s,
            "unit": "steps/s (1 step = %d clips advanced one timestep, CFG on)" % B,
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": t_ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {"simt": "f32", "tc_bf16x3": "bf16x3 split operands, f32 accumulate",
                      "tc_bf16": "bf16 operands, f32 accumulate"}[impl],
            "data": "synthetic", "config": workload_config(a, dims), "impl": "ours", "kernel": impl,
            "clips_per_s": steps_per_s * B / T_FULL,
            "e2e": {"value": world * n_e2e / (e2e_ms / 1e3), "unit": "steps/s", "steps_in_loop": n_e2e,
                    "h2d_bytes_per_step": h2d / n_e2e, "d2h_bytes_per_step": d2h / n_e2e,
                    "what": "p_sample_loop(model, shape, model_kwargs=pinned host cond) -> .cpu(): H2D of the cond, "
                            "WavEncoder + cond precompute, %d steps, D2H of the samples; timed after 3 short "
                            "warm-up loops, max(CUDA events, host wall clock)" % n_e2e},
            "gpu_launches": launches,
            "launches_per_step": launches_per_step,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(impl, a.dataset, B, nk), "peak_source": peak_src,
                         "kernel_ms": kern_ms, "steps_per_launch": nk, "flop_per_launch": flop_launch,
                         "note": "algorithmic flops (BASELINE.md section 3): x3 of the bf16x3 split and padding not counted"},
            "clocks": clk,
        }
        if not a.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            sb = min(256, B)
            v, per = cpu_port_steps_per_s(dims, B, 4, 1, sb, threads)
            line["cpu_baseline"] = {"value": v, "unit": "steps/s", "cores": threads, "kind": "port",
                                    "sample": "4 oracle p_sample steps at B=%d clips (CFG on, WavEncoder recomputed "
                                              "per step as in the reference), %.3f s each, scaled to B=%d" % (sb, per, B)}
            try:
                line["gpu_eager_baseline"] = {
                    "value": gpu_eager_port_steps_per_s(dims, B, 5, dev), "unit": "steps/s", "kind": "port",
                    "what": "the oracle port as eager PyTorch on this GPU (fp32, TF32 off, %d clips, WavEncoder "
                            "recomputed per pass like the reference): reference-style code on a B200, for scale" % B}
            except Exception as e:      # a reported extra, never a reason to lose the bench line
                line["gpu_eager_baseline"] = {"error": "%s: %s" % (type(e).__name__, e)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
