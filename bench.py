#!/usr/bin/env python
"""bench.py - headline measurement of the RAG diffusion sampling path.

Metric (BASELINE.json): denoising-steps/sec (one step = the whole batch advanced one timestep = 2 denoiser passes per
clip + guidance + posterior update) at T=1000 ancestral, classifier-free guidance on.  Workloads (--config):
  2 (default)  TED RAG sampling, B=512 clips per GPU                               (BASELINE configs[1], the headline)
  3            TED LivelySpeaker sampling, B=256: SAG decoder -> init_image, skip_timesteps=800 of T=1000
  4            BEAT RAG sampling, B=256, 47 joints x 6
(config 5 = config 2 under `torch.distributed.run --nproc-per-node 8`: 512 clips per GPU, weak scaling.)

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
        --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5
    python bench.py --impl reference --steps 3 --warmup 1      # CPU arm: the unmodified reference (oracle/_ref)

Whatever --steps / --warmup say, the timed region is made of WHOLE 16-step launches and lasts at least --min-seconds
(0.5 s): `steps` / `warmup` in the JSON line are the numbers of steps actually run (>= the requested ones), `e2e` is
always the full loop of the workload.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
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
CONFIGS = {     # --config: dataset, clips per GPU, skipped timesteps, SAG init_image
    2: dict(dataset="ted", batch=512, skip=0, sag=False,
            name="TED RAG sampling B=%d/GPU, F=34, J*D=27, T=1000 ancestral, 8-layer/512-d, CFG scale 1.5"),
    3: dict(dataset="ted", batch=256, skip=800, sag=True,
            name="TED LivelySpeaker sampling B=%d/GPU: SAG decoder -> init_image, skip_timesteps=800 of T=1000 ancestral "
                 "(200 RAG steps), CFG scale 1.5"),
    4: dict(dataset="beat", batch=256, skip=0, sag=False,
            name="BEAT RAG sampling B=%d/GPU, F=34, J*D=282 (47 joints x 6), T=1000 ancestral, CFG scale 1.5"),
}


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
    measurably stalls the launching thread; it remains the fallback when pynvml is missing.  Started before the
    warm-up, marked around the timed region."""
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
                        try:        # board power and its limit (mW): direct evidence of a power-capped kernel
                            pw = (pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1000.0)
                        except Exception:
                            pw = None
                        self.rows.append((time.perf_counter(), float(sm), float(mx), [n for n, b in bits if r & b], pw))
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
        out = {"sm_mhz": statistics.median(r[1] for r in rows), "sm_max_mhz": max(r[2] for r in rows),
               "reasons": reasons, "samples": len(rows), "window": window, "via": self.how}
        pw = [r[4] for r in rows if len(r) > 4 and r[4] is not None]
        if pw:
            out["power_w"] = round(statistics.median(p[0] for p in pw), 1)
            out["power_w_max"] = round(max(p[0] for p in pw), 1)
            out["power_limit_w"] = round(max(p[1] for p in pw), 1)
        return out


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own CPU path on the host cores
# ---------------------------------------------------------------------------------------------------------------
def ref_tree(dataset):
    """oracle/_ref/<dataset>: the unmodified reference files placed there by oracle/make_ref.sh at build time."""
    p = os.path.join(ROOT, "oracle", "_ref", dataset)
    return p if os.path.exists(os.path.join(p, "mdm_utils", "model_util.py")) else None


def cpu_reference_steps_per_s(dims, batch_equiv, n_steps, warmup, sample_batch, threads):
    """The UNMODIFIED reference (its own create_model_and_diffusion / ClassifierFreeSampleModel / diffusion.p_sample,
    scripts/test_RAG_ted.py:152-182 wiring) on the host cores: `n_steps` p_sample steps at `sample_batch` clips,
    scaled to steps/s at `batch_equiv` clips.  Returns None when oracle/_ref is absent."""
    tree = ref_tree(dims.dataset)
    if tree is None:
        return None
    import torch
    from livelyspeaker_b200 import synthetic
    saved_path, saved_mods = list(sys.path), {k: sys.modules.get(k) for k in ("model", "diffusion", "mdm_utils", "clip")}
    sys.dont_write_bytecode = True
    sys.path.insert(0, tree)
    try:
        for k in list(sys.modules):
            if k.split(".")[0] in ("model", "diffusion", "mdm_utils", "clip"):
                del sys.modules[k]
        import warnings
        warnings.filterwarnings("ignore")
        from mdm_utils.model_util import create_model_and_diffusion, load_model_wo_clip
        from model.cfg_sampler import ClassifierFreeSampleModel
        torch.set_num_threads(threads)
        import contextlib
        import io
        model, diffusion = create_model_and_diffusion(model_args(dims), "")
        with contextlib.redirect_stdout(io.StringIO()):
            load_model_wo_clip(model, synthetic.synth_state_dict(dims, seed=1))
        model.eval()
        cfg = ClassifierFreeSampleModel(model).eval()
        y = synthetic.synth_cond(dims, sample_batch)
        torch.manual_seed(0)
        x = torch.randn(sample_batch, dims.njoints, dims.nfeats, 34)
        times = []
        with torch.no_grad():
            for k in range(warmup + n_steps):
                t0 = time.perf_counter()
                t = torch.tensor([T_FULL - 1 - k] * sample_batch)
                x = diffusion.p_sample(cfg, x, t, clip_denoised=False, model_kwargs={"y": y})["sample"]
                if k >= warmup:
                    times.append(time.perf_counter() - t0)
        per_step = sum(times) / len(times)
        return (sample_batch / batch_equiv) / per_step, per_step
    finally:
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k.split(".")[0] in ("model", "diffusion", "mdm_utils", "clip"):
                del sys.modules[k]
        for k, v in saved_mods.items():
            if v is not None:
                sys.modules[k] = v


def cpu_port_steps_per_s(dims, batch_equiv, n_steps, warmup, sample_batch, threads):
    """Fallback when oracle/_ref is absent: the oracle (a port of the reference's CPU path, pinned to it draw by draw)."""
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


def cpu_arm(dims, batch, n_steps, warmup):
    """(value steps/s at `batch` clips, cpu_baseline dict): the real reference when oracle/_ref is there, else the port."""
    threads = os.cpu_count() or 1
    sb = min(256, batch)
    r = cpu_reference_steps_per_s(dims, batch, n_steps, warmup, sb, threads)
    kind = "reference"
    what = ("%d diffusion.p_sample steps of the UNMODIFIED reference (oracle/_ref/%s, files copied by oracle/make_ref.sh; "
            "CFG on, WavEncoder and deepcopy(y) per pass as shipped) at B=%d clips" % (n_steps, dims.dataset, sb))
    if r is None:
        r = cpu_port_steps_per_s(dims, batch, n_steps, warmup, sb, threads)
        kind = "port"
        what = "%d oracle p_sample steps (port of the reference; oracle/_ref absent on this box) at B=%d clips" % (n_steps, sb)
    value, per_step = r
    return value, {"value": value, "unit": "steps/s", "cores": threads, "kind": kind,
                   "sample": "%s, %.3f s each, scaled to B=%d" % (what, per_step, batch)}


def gpu_eager_port_steps_per_s(dims, batch, n_steps, dev):
    """The oracle port run as eager PyTorch on the GPU (fp32, TF32 off): "reference-style code on a B200"
    (SURVEY.md 8d).  A reported extra only."""
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


def workload_config(a, dims, cfgd):
    return {"workload": cfgd["name"] % a.batch, "baseline_config": a.config,
            "global_batch": a.batch * a.gpus, "timesteps": T_FULL, "steps_in_loop": T_FULL - cfgd["skip"],
            "sampler": a.sampler,
            "l2": "flushed between timed launches of %d steps (256 MiB memset outside the event bracket)" % a.chunk,
            "steps_per_launch": a.chunk,
            "parallelism": "batch shard x%d, no per-step collective, 1 all_gather at loop end" % a.gpus}


def run_reference_arm(a, dims, cfgd, rank):
    """--impl reference: the reference's own CPU implementation of the path on all host cores, bounded sample."""
    if rank != 0:
        return
    n_steps = max(1, min(a.steps, 8))
    value, base = cpu_arm(dims, a.batch, n_steps, 1)
    # the unit (steps of `batch` clips) does not depend on N: the host cores do not scale with --gpus
    line = {"impl": "reference", "metric": "denoising-steps/sec", "value": value,
            "unit": "steps/s (1 step = %d clips advanced one timestep, CFG on)" % a.batch, "n_gpus": a.gpus,
            "steps": n_steps, "warmup": 1, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(a, dims, cfgd), "cpu_baseline": base,
            "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "clips_per_s": value * a.batch / (T_FULL - cfgd["skip"])}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kernel", default="auto", choices=["auto", "simt", "tc_bf16x3", "tc_bf16"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json configs index")
    ap.add_argument("--batch", type=int, default=None, help="clips per GPU (default: the config's)")
    ap.add_argument("--sampler", default="ancestral", choices=["ancestral", "ddim"])
    ap.add_argument("--chunk", type=int, default=16, help="loop iterations per launch (ls_step_multi), 1..16")
    ap.add_argument("--min-seconds", type=float, default=0.5, help="lower bound of the timed region")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    cfgd = CONFIGS[a.config]
    a.dataset = cfgd["dataset"]
    if a.batch is None:
        a.batch = cfgd["batch"]
    a.warmup = max(a.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from livelyspeaker_b200 import synthetic
    dims = synthetic.dims_for(a.dataset)
    if a.impl == "reference":
        run_reference_arm(a, dims, cfgd, rank)
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

    B = a.batch
    ddim = a.sampler == "ddim"
    n_loop = T_FULL - cfgd["skip"]                       # steps of one whole loop of this workload
    mk = beat_model_util.create_model_and_diffusion if a.dataset == "beat" else ls.create_model_and_diffusion
    model, diffusion = mk(model_args(dims), "")
    sd = synthetic.synth_state_dict(dims, seed=1)
    model.load_state_dict(sd, strict=True)
    model.set_impl(a.kernel)
    cfg = ls.ClassifierFreeSampleModel(model).to(dev).eval()
    eng = model.engine(B)
    shape = (B, dims.njoints, dims.nfeats, 34)
    sag = None
    if cfgd["sag"]:
        sag = ls.Decoder_TRANSFORMER(latent_dim=512, n_pre_poses=4, use_style=False)
        sag.load_state_dict(synthetic.synth_sag_state_dict(seed=3), strict=True)
        sag = sag.to(dev).eval()

    # Conditioning of the GLOBAL batch, identical on every rank (seed 233 = the reference's fixseed): the end-to-end
    # leg hands it to sample_sharded, which slices this rank's shard; the device-resident leg uses the shard directly.
    y_glob = synthetic.synth_cond(dims, B * world, seed=233)
    y_glob_pinned = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in y_glob.items()}
    lo, hi = sharding.shard_bounds(B * world, world, rank)
    y_dev = {k: (v[lo:hi].to(dev) if torch.is_tensor(v) else v) for k, v in y_glob.items()}
    z_glob = torch.randn(B * world, 512, generator=torch.Generator().manual_seed(7)).pin_memory() if sag else None

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
    C = max(1, min(a.chunk, ls.MAX_FUSED_STEPS))
    params_all = [diffusion.step_params(i, ddim=ddim, eta=0.0, clip_denoised=False) for i in range(n_loop)]

    def params_at(k0, n):                 # loop iterations k0 .. k0+n-1 (the loop counts i = n_loop-1 .. 0, wrapping)
        return [params_all[n_loop - 1 - ((k0 + j) % n_loop)] for j in range(n)]

    draw_launches = [0]

    def run_chunk(k0, x_in, steady):
        """C consecutive loop iterations as p_sample_loop runs them on the fused route: the reference's per-step draws
        (full chunks after the first: ONE torch-compatible launch, or a CUDA-graph replay of the 3C torch kernels when
        that kernel does not verify against torch on this box), then ONE ls_step_multi launch."""
        graphed = None
        if steady and C > 1 and diffusion.graph_draws:
            graphed = gd.chunk_draws(eng, C, B, 512, perm_like)
            if graphed is None:
                diffusion.graph_draws = False
        if graphed is not None:
            e_c, e_u, nz = graphed.draw()
            if isinstance(graphed, gd._FusedDraws):
                draw_launches[0] += 1         # ls_randn_torch_compat: one kernel of ours per chunk
        else:
            e_c = [torch.randn(B, 1, 512, device=dev) for _ in range(C)]
            e_u = [torch.randn(B, 1, 512, device=dev) for _ in range(C)]
            nz = [torch.randn_like(perm_like) for _ in range(C)]
        xs = torch.empty((C,) + tuple(x_in.shape), device=dev)
        if C == 1:
            eng.step(params_at(k0, 1)[0], x_in, e_c[0], e_u[0], nz[0], scale, xs[0], None)
        else:
            eng.step_multi(params_at(k0, C), x_in, e_c, e_u, nz, scale, xs, None)
        return xs[C - 1]

    clocks = ClockSampler(local_rank)
    clocks.start()
    # Warm-up in whole launches: >= the requested steps, >= 3 launches and >= 0.3 s (the second launch is the first to
    # take the steady-state route: one-launch draws, cached allocator blocks; clocks settle within ~0.3 s).
    n_warm = max(3, -(-a.warmup // C))
    t_w0 = time.perf_counter()
    k = 0
    while k < n_warm or time.perf_counter() - t_w0 < 0.3:
        x = run_chunk(k * C, x, steady=k > 0)
        k += 1
        if k % 4 == 0:
            torch.cuda.synchronize()
    n_warm = k
    sync_all()
    # one probe launch sizes the timed region: whole launches, >= the requested steps, >= --min-seconds
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    x = run_chunk(n_warm * C, x, steady=True)
    p1.record()
    torch.cuda.synchronize()
    probe = torch.tensor([p0.elapsed_time(p1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(probe, op=dist.ReduceOp.MAX)
    n_timed = max(-(-a.steps // C), int(a.min_seconds * 1e3 / max(float(probe), 1e-3)) + 1)
    n_warm += 1
    K, W = n_timed * C, n_warm * C
    clocks.wait_first_sample()
    launches0 = eng.launch_count() + draw_launches[0]
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_timed)]
    import gc
    gc.collect()
    gc.disable()                               # no collector pause on the launching thread inside the timed region
    sync_all()
    clocks.mark_begin()
    for j in range(n_timed):
        flush_buf.zero_()                      # evict L2 (126 MB) - outside the event bracket
        ev[j][0].record()
        x = run_chunk((n_warm + j) * C, x, steady=True)
        ev[j][1].record()
    sync_all()
    clocks.mark_end()
    gc.enable()
    launches = eng.launch_count() + draw_launches[0] - launches0      # fused launches + the one-launch draws
    clk = clocks.stop()
    per_launch = [s.elapsed_time(e) for s, e in ev]
    t_ms = sum(per_launch)
    assert torch.isfinite(x).all(), "sampler diverged"

    # dominant kernel alone (the ls_step_multi launch without the RNG launch), for the roofline: >= 0.3 s of launches
    g = torch.Generator(device=dev).manual_seed(5)
    e_c = [torch.randn(B, 1, 512, device=dev, generator=g) for _ in range(C)]
    e_u = [torch.randn(B, 1, 512, device=dev, generator=g) for _ in range(C)]
    nz = [torch.randn(*shape, device=dev, generator=g) for _ in range(C)]
    xs = torch.empty((C,) + tuple(x.shape), device=dev)
    kk = max(5, int(300.0 / max(float(probe), 1e-3)) + 1)
    evk = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(kk)]
    l0 = eng.launch_count()
    pk = params_at(0, C)
    for k in range(kk):
        flush_buf.zero_()
        evk[k][0].record()
        if C == 1:
            eng.step(pk[0], x, e_c[0], e_u[0], nz[0], scale, xs[0], None)
        else:
            eng.step_multi(pk, x, e_c, e_u, nz, scale, xs, None)
        evk[k][1].record()
    torch.cuda.synchronize()
    kern = [s.elapsed_time(e) for s, e in evk]
    kern_ms = sum(kern) / kk                                        # per launch of C steps
    launches_per_step = (eng.launch_count() - l0) / (kk * C)

    # ------------------------------------------------------------------ end to end (host buffers), ALWAYS the whole loop
    keys = ("audio_input", "origin_x", "vid_indices", "scale", "emo")
    h2d = sum(v[lo:hi].numel() * v.element_size() for k_, v in y_glob_pinned.items() if torch.is_tensor(v) and k_ in keys)
    if sag:
        h2d += (hi - lo) * 512 * 4
    sample_fn = diffusion.ddim_sample_loop if ddim else diffusion.p_sample_loop
    diffusion.fused_chunk = C
    gshape = (B * world,) + tuple(shape[1:])

    def e2e_once(n):
        """The call a user makes (scripts/test_RAG_ted.py:71-82 / test_LivelySpeaker_ted.py:85-113): pinned host cond in,
        host samples out.  n = loop length (skip_timesteps = T - n)."""
        kw = dict(clip_denoised=False, skip_timesteps=T_FULL - n)
        if sag is not None:
            zl = z_glob[lo:hi].to(dev, non_blocking=True)
            ox = y_glob_pinned["origin_x"][lo:hi].to(dev, non_blocking=True)
            dec = sag({"x": ox, "z": zl, "mask": torch.ones(hi - lo, 34, dtype=torch.bool, device=dev)})["output"]
            if world > 1:
                full = torch.zeros(gshape, device=dev)
                full[lo:hi] = dec
                kw["init_image"] = full                 # sample_sharded slices [lo:hi] back out
            else:
                kw["init_image"] = dec
        if world > 1:
            out = sharding.sample_sharded(sample_fn, cfg, gshape, {"y": y_glob_pinned}, diffusion=diffusion,
                                          rng="per_rank", fork_seed=False, **kw)
        else:
            out = sample_fn(cfg, shape, model_kwargs={"y": dict(y_glob_pinned)}, **kw)
        return out.to("cpu", non_blocking=False)

    for _ in range(3):                        # warm-up of the e2e path: allocator pools, pinned staging
        e2e_once(min(2 * C, n_loop))
    sync_all()
    s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s_ev.record()
    res = e2e_once(n_loop)
    e_ev.record()
    torch.cuda.synchronize()
    e2e_ms = max(s_ev.elapsed_time(e_ev), (time.perf_counter() - t0) * 1e3)
    sync_all()
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
        flop_launch = C * flop_step
        achieved = flop_launch / (kern_ms / 1e3) / 1e12
        impl = eng.get_impl()
        line = {
            "metric": "denoising-steps/sec", "value": steps_per_s,
            "unit": "steps/s (1 step = %d clips advanced one timestep, CFG on)" % B,
            "n_gpus": world, "steps": K, "warmup": W, "requested_steps": a.steps, "requested_warmup": a.warmup,
            "ms_per_step": t_ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {"simt": "f32", "tc_bf16x3": "bf16x3 split operands, f32 accumulate",
                      "tc_bf16": "bf16 operands, f32 accumulate"}[impl],
            "data": "synthetic", "config": workload_config(a, dims, cfgd), "impl": "ours", "kernel": impl,
            "clips_per_s": steps_per_s * B / n_loop,
            "timed_region": {"seconds": t_ms / 1e3, "launches": n_timed, "steps_per_launch": C,
                             "launch_ms_median": statistics.median(per_launch), "launch_ms_min": min(per_launch),
                             "launch_ms_max": max(per_launch)},
            "e2e": {"value": world * n_loop / (e2e_ms / 1e3), "unit": "steps/s", "steps_in_loop": n_loop,
                    "seconds": e2e_ms / 1e3, "clips_per_s": world * B / (e2e_ms / 1e3),
                    "h2d_bytes_per_step": h2d / n_loop, "d2h_bytes_per_step": d2h / n_loop,
                    "what": ("%sp_sample_loop(model, shape, model_kwargs=pinned host cond)%s -> .cpu(): H2D of the cond, "
                             "WavEncoder + cond precompute, the WHOLE loop of %d steps, D2H of the samples; timed once after "
                             "3 short warm-up loops, max(CUDA events, host wall clock), max over ranks"
                             % ("SAG decoder -> init_image -> " if sag else "",
                                " through sharding.sample_sharded (1 all_gather)" if world > 1 else "", n_loop))},
            "gpu_launches": launches,
            "launches_per_step": launches_per_step,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(impl, a.dataset, B, C), "peak_source": peak_src,
                         "kernel_ms": kern_ms, "kernel_ms_min": min(kern), "kernel_ms_max": max(kern), "kernel_launches": kk,
                         "steps_per_launch": C, "flop_per_launch": flop_launch,
                         "note": "algorithmic flops (BASELINE.md section 3): x3 of the bf16x3 split and padding not counted"},
            "clocks": clk,
        }
        if not a.no_cpu_baseline and world == 1:
            _, line["cpu_baseline"] = cpu_arm(dims, B, 4, 1)
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
